"""SURVEY.md 8f #3, first half: Powell's three-sequence affine Ukkonen aligner (src/ukk.checkp.c + src/ukkCommon.c, the
aligner behind Sequence.Align.readjust_3d, src/sequence.ml:1075-1139) compiled UNMODIFIED into oracle/_ref/libpoyref.so and
driven through its own OCaml stub powell_3D_align (src/ukkCommon.c:110-145) on blocks built by oracle/caml_runtime.c.
These tests pin the oracle (it runs here, its answers have the properties an exact aligner's must have); the CUDA
implementation of this aligner is not built yet."""
import os

import numpy as np
import pytest

import test_stubs as TS


@pytest.fixture(scope="module")
def ref():
    from oracle import oracle

    oracle.build(ref=True)
    if not os.path.exists(TS.REF_SO):
        pytest.skip("oracle/_ref/libpoyref.so not built")
    r = TS.Side(TS.REF_SO)
    if not hasattr(r.L, "powell_3D_align"):
        pytest.skip("libpoyref.so predates the Powell recipe")
    return r


def powell(ref, x, y, z, mm=1, go=3, ge=2):
    cap = len(x) + len(y) + len(z)
    s = [ref.seq(v) for v in (x, y, z)]
    o = [ref.empty(cap) for _ in range(3)]
    cost = TS.int_val(ref.call("powell_3D_align", s[0], s[1], s[2], o[0], o[1], o[2], TS.val_int(mm), TS.val_int(go), TS.val_int(ge)))
    return cost, [ref.read(v) for v in o]


def _sp_affine(rows, mm, go, ge):
    """Cost of an alignment as Powell's state machine charges it: per column the mismatch cost among the non-gap characters
    (whichCharCost, src/ukkCommon.c:155-184: 0 / 1 / 2 x mm) plus, per sequence, go when a gap run starts and ge per gap."""
    a, b, c = rows
    cost = 0
    prev = [False, False, False]
    for col in zip(a, b, c):
        chars = [x for x in col if x != 16]
        k = len(set(chars))
        cost += mm * (0 if k <= 1 else (1 if k == 2 else 2))
        for q, x in enumerate(col):
            g = x == 16
            if g:
                cost += ge + (0 if prev[q] else go)
            prev[q] = g
    return cost


def test_powell_oracle_runs_and_is_consistent(ref):
    rng = np.random.default_rng(31)
    for n, p in ((30, 0.1), (60, 0.15), (90, 0.05)):
        a = TS._dna(rng, n)
        b, c = TS._mutate(rng, a, p), TS._mutate(rng, a, p)
        cost, rows = powell(ref, a, b, c)
        assert len(rows[0]) == len(rows[1]) == len(rows[2])
        # removing the gaps gives the operands back (codes collapse to their lowest base, copySequence :84-105), behind a
        # leading all-gap column
        for src, row in zip((a, b, c), rows):
            assert row[0] == 16
            body = row[1:][row[1:] != 16]
            assert np.array_equal(body, src[1:]), "alignment does not spell its operand"
        assert cost > 0
        assert cost <= _sp_affine([r[1:] for r in rows], 1, 3, 2)  # never worse than what its own alignment costs naively
    a = TS._dna(rng, 50)
    assert powell(ref, a, a, a)[0] == 0
    # a single base deleted from one sequence: one gap opening + one extension
    b = np.delete(a, 20)
    cost, rows = powell(ref, a, a, b, mm=1, go=3, ge=2)
    assert cost == 3 + 2 or cost == 2 * (3 + 2), cost
