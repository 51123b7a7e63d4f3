"""CPU tests of the checkers themselves (no GPU, no CUDA calls).

* the plain-C port (oracle/poy_oracle.c) replays every committed golden fixture bit-for-bit -- the fixtures are
  outputs of the unmodified reference algn.c (tests/golden/make_golden.py);
* where the compiled reference is present (this container / shipped oracle/_ref), the port is also compared with
  it directly on fresh seeded inputs, including direction matrices, empty sequences and the cost-only variant
  whose block-diagonal rule differs from the traceback variant;
* the host-side cost-matrix builder reproduces the tables the fixtures were generated with, and the deltaw rule of
  Sequence.Align.cost_2 shows the protein quirk of SURVEY.md A15.
"""
import os

import numpy as np
import pytest

from golden_util import golden_files, load


@pytest.fixture(scope="module")
def O():
    from oracle import oracle

    oracle.build(ref=True)
    return oracle


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        x, y = a[k], b[k]
        if x.ndim == 2 and x.shape != y.shape:
            w = min(x.shape[1], y.shape[1])
            assert not x[:, w:].any() and not y[:, w:].any()
            x, y = x[:, :w], y[:, :w]
        assert np.array_equal(x, y), k


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_port_replays_reference_golden(O, path):
    cm, pool, pairs, mode, deltaw, ref = load(path)
    got = O.Port(cm).batch(mode, pool.pool, pool.off, pool.len, pairs, deltaw=deltaw, nthreads=2)
    _same(got, ref)


def test_there_are_golden_fixtures():
    assert len(golden_files()) >= 8


def _need_ref(O):
    if not O.Reference.available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")


def test_port_vs_compiled_reference_linear(O):
    _need_ref(O)
    from poyd_b200 import cost_matrix as CM, synth

    for cm, alph in ((CM.default_nucleotides(), "dna"), (CM.nucleotides(3, 1), "dna"), (CM.default_aminoacids(), "protein")):
        P, R = O.Port(cm), O.Reference(cm)
        pool, pairs = synth.ragged_batch(250, max_len=200, seed=101, alphabet=alph, gap_ambiguity=0.04 if alph == "dna" else 0)
        rng = np.random.default_rng(5)
        for k, (a, b) in enumerate(pairs):
            s1, s2 = pool.seq(a), pool.seq(b)
            if len(s1) < len(s2):
                s1, s2 = s2, s1
            dw = int(rng.integers(0, 70))
            cp, dp = P.cost_2(s1, s2, dw, want_dir=True)
            cr, dr = R.cost_2(s1, s2, dw, want_dir=True)
            assert cp == cr and np.array_equal(dp, dr), (k, len(s1), len(s2), dw)
        dw = rng.integers(0, 60, size=len(pairs)).astype(np.int32)
        _same(P.batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=dw), R.batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=dw))


def test_port_vs_compiled_reference_affine(O):
    _need_ref(O)
    from poyd_b200 import cost_matrix as CM, synth

    ndiff = 0
    for cm in (CM.nucleotides(1, 2, 3), CM.nucleotides(2, 1, 1), CM.nucleotides(3, 1, 5)):
        P, R = O.Port(cm), O.Reference(cm)
        for gap_amb in (0.0, 0.15):
            pool, pairs = synth.ragged_batch(200, max_len=220, seed=77, gap_ambiguity=gap_amb)
            a, b = P.batch(3, pool.pool, pool.off, pool.len, pairs), R.batch(3, pool.pool, pool.off, pool.len, pairs)
            _same(a, b)
            ca, cb = P.batch(2, pool.pool, pool.off, pool.len, pairs)["cost"], R.batch(2, pool.pool, pool.off, pool.len, pairs)["cost"]
            assert np.array_equal(ca, cb)
            ndiff += int((ca != a["cost"]).sum())
    assert ndiff > 0, "the cost-only and traceback recurrences should disagree somewhere (SURVEY.md A.2)"


def test_port_vs_compiled_reference_medians_and_empty(O):
    _need_ref(O)
    from poyd_b200 import cost_matrix as CM

    rng = np.random.default_rng(3)
    for cm in (CM.default_nucleotides(), CM.nucleotides(1, 2, 3), CM.default_aminoacids()):
        P, R = O.Port(cm), O.Reference(cm)
        hi = 32 if cm.combinations else 23
        for n in (0, 1, 2, 17, 90):
            a = rng.integers(1, hi, size=n).astype(np.uint8)
            b = rng.integers(1, hi, size=n).astype(np.uint8)
            for which in (0, 1, 2):
                assert np.array_equal(P.median_2(which, a, b), R.median_2(which, a, b)), (which, n)
    cm = CM.nucleotides(1, 2, 3)
    P, R = O.Port(cm), O.Reference(cm)
    empty, one = np.array([16], np.uint8), np.array([16, 4], np.uint8)
    for x, y in ((empty, empty), (empty, one), (one, empty), (one, one)):
        rp, rr = P.align_affine_3(x, y), R.align_affine_3(x, y)
        assert rp[0] == rr[0] and all(np.array_equal(u, v) for u, v in zip(rp[1:], rr[1:]))


def test_cost_matrix_builder_matches_fixture_tables():
    from poyd_b200 import cost_matrix as CM

    want = {"affine_cfg2_medianlike": CM.nucleotides(1, 2, 3), "affine_ragged": CM.nucleotides(2, 1, 1),
            "linear_ragged": CM.default_nucleotides(), "protein_cfg3a_full": CM.default_aminoacids()}
    for path in golden_files():
        name = os.path.basename(path)[:-4]
        if name in want:
            cm = load(path)[0]
            for f in ("cost", "median", "worst", "prepend_cost", "tail_cost"):
                assert np.array_equal(getattr(cm, f), getattr(want[name], f)), (name, f)
            assert (cm.lcm, cm.gap, cm.a_sz, cm.gap_open) == (want[name].lcm, want[name].gap, want[name].a_sz, want[name].gap_open)


def test_cost_matrix_known_values():
    """Spot values that follow from cost_matrix.ml by hand: DNA 1/2 (default, metric) and its affine clone."""
    from poyd_b200 import cost_matrix as CM

    m = CM.default_nucleotides()
    assert (m.lcm, m.gap, m.a_sz, m.combinations) == (5, 16, 31, 1)
    assert m.cost[1, 2] == 1 and m.cost[1, 16] == 2 and m.cost[1, 1] == 0
    assert m.cost[3, 1] == 0 and m.median[3, 1] == 1        # {A,C} vs A: shared state, cost 0
    assert m.cost[3, 12] == 1 and m.median[3, 12] == 15      # {A,C} vs {G,T}: any of the four at cost 1
    assert m.cost[1, 17] == 0 and m.median[1, 17] == 1       # A vs {A,gap}
    assert m.tail_cost[4] == 2 and m.prepend_cost[8] == 2
    a = CM.nucleotides(1, 2, 3)
    assert a.cost_model_type == 1 and a.gap_open == 3
    assert a.cost[1, 2] == 1 and a.median[1, 2] == 3         # bitwise rule keeps the union of the closest pair
    p = CM.default_aminoacids()
    assert (p.lcm, p.gap, p.a_sz, p.combinations) == (6, 22, 22, 0)
    assert p.cost[3, 7] == 1 and p.cost[3, 21] == 0 and p.cost[3, 22] == 2


def test_deltaw_rule_and_protein_quirk():
    """Sequence.Align.cost_2's deltaw (src/sequence.ml:691-714) needs no GPU: gaps + f(lengths)."""
    from poyd_b200 import sequence as S, synth

    l1, l2 = np.array([501, 501, 301]), np.array([501, 440, 301])
    assert list(S.deltaw_calc(l1, l2, None)) == [25, 2, 15]
    assert list(S.deltaw_calc(l1, l2, np.array([8, 61, 8]))) == [50, 61, 30]
    pool, pairs = synth.pair_batch(8, 300, seed=3, alphabet="protein", subst=0.15, indel=0.02)
    cnt = pool.count(22)
    # 18 of the 21 residue codes share a bit with the gap code 22: the "gap count" is ~ 0.86 len
    assert (cnt > 0.7 * pool.len).all()
    pool, pairs = synth.pair_batch(8, 500, seed=3)
    assert (pool.count(16) == 1).all()  # DNA: only the leading gap


def test_port_vs_compiled_reference_cube(O):
    """3-D: the port reproduces the reference's executed (defective) cube fill, its walk and its constant median."""
    _need_ref(O)
    from poyd_b200 import cost_matrix as CM

    cm3 = CM.of_two_dim(CM.default_nucleotides())
    assert cm3.cost[1, 2, 4] == 2 and cm3.cost[1, 1, 1] == 0 and cm3.cost[1, 1, 16] == 2 and cm3.median[1, 1, 16] == 1
    P, R = O.Port3(cm3), O.Reference3(cm3)
    rng = np.random.default_rng(4)

    def rnd(n):
        return np.concatenate([[16], rng.choice(np.array([1, 2, 4, 8], np.uint8), size=n)]).astype(np.uint8)

    # the two probes of SURVEY.md Appendix B: the cube is NOT an optimal 3-way alignment
    x = rnd(10)
    assert R.align_3(x, x, x)[0] > 0
    for _ in range(120):
        a = rnd(int(rng.integers(0, 45)))
        b = a.copy() if rng.random() < 0.5 else rnd(int(rng.integers(0, 45)))
        c = rnd(int(rng.integers(0, 45)))
        rp, rr = P.align_3(a, b, c, want_dir=True), R.align_3(a, b, c, want_dir=True)
        assert rp[0] == rr[0] and rp[1] == rr[1]
        assert np.array_equal(rp[6], rr[6])
        for u, v in zip(rp[2:6], rr[2:6]):
            assert np.array_equal(u, v)
