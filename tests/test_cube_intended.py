"""SURVEY.md "3-D policy" (1): the recurrence algn_fill_cube INTENDS (oracle/poy_oracle.c po_cost_3_intended -- the reference's
seven candidates, order and strictness, src/algn.c:2812-2857, 3039-3046, read from the rows a three-sequence Needleman-Wunsch
needs) checked against brute force: exhaustive enumeration of every alignment for tiny triples, an independent memoised
recursion for larger ones.  The product kernel reproduces the EXECUTED (lagging-row) form; these tests pin what a corrected
reference would return and show the distance between the two."""
import functools
import itertools

import numpy as np
import pytest

from oracle import oracle
from poyd_b200 import cost_matrix as CM

GAP = 16


@pytest.fixture(scope="module")
def checker():
    oracle.build(ref=False)
    cm3 = CM.of_two_dim(CM.default_nucleotides())
    return oracle.Port3(cm3), np.asarray(cm3.cost).reshape(32, 32, 32), np.asarray(cm3.median).reshape(32, 32, 32)


def seq(rng, n, ambiguity=0.0):
    s = rng.choice(np.array([1, 2, 4, 8], np.uint8), n)
    if ambiguity:
        amb = rng.random(n) < ambiguity
        s[amb] = rng.integers(1, 16, int(amb.sum()), dtype=np.uint8)
    return np.concatenate([[GAP], s]).astype(np.uint8)


MOVES = [m for m in itertools.product((0, 1), repeat=3) if any(m)]


def enumerate_all(cost3, a, b, c):
    """Minimum column-cost sum over EVERY alignment of a[1:], b[1:], c[1:] (depth-first over the seven column types)."""
    best = [None]

    def go(i, j, k, acc):
        if best[0] is not None and acc >= best[0]:
            return
        if i == len(a) - 1 and j == len(b) - 1 and k == len(c) - 1:
            best[0] = acc
            return
        for di, dj, dk in MOVES:
            if i + di < len(a) and j + dj < len(b) and k + dk < len(c):
                x = a[i + di] if di else GAP
                y = b[j + dj] if dj else GAP
                z = c[k + dk] if dk else GAP
                go(i + di, j + dj, k + dk, acc + int(cost3[x, y, z]))

    go(0, 0, 0, 0)
    return best[0]


def memoised(cost3, a, b, c):
    @functools.lru_cache(maxsize=None)
    def f(i, j, k):
        if i == 0 and j == 0 and k == 0:
            return 0
        best = None
        for di, dj, dk in MOVES:
            if i - di >= 0 and j - dj >= 0 and k - dk >= 0:
                x = a[i] if di else GAP
                y = b[j] if dj else GAP
                z = c[k] if dk else GAP
                v = f(i - di, j - dj, k - dk) + int(cost3[x, y, z])
                best = v if best is None or v < best else best
        return best

    return f(len(a) - 1, len(b) - 1, len(c) - 1)


def check_alignment(cost3, med3, a, b, c, res):
    cost, status, r1, r2, r3, med = res
    assert status == 0
    assert len(r1) == len(r2) == len(r3) == len(med)
    for r, s in ((r1, a), (r2, b), (r3, c)):
        assert np.array_equal(r[r != GAP], s[1:])  # the rows spell the sequences (no leading all-gap column, A13)
    assert not np.any((r1 == GAP) & (r2 == GAP) & (r3 == GAP))
    assert int(cost3[r1, r2, r3].sum()) == cost      # the traceback realises the cost of the fill
    assert np.array_equal(med, med3[r1, r2, r3])     # one cm_get_median_3d per column


def test_intended_cube_equals_exhaustive_enumeration(checker):
    port, cost3, med3 = checker
    rng = np.random.default_rng(1)
    for _ in range(60):
        a, b, c = (seq(rng, int(rng.integers(0, 4)), ambiguity=0.2) for _ in range(3))
        res = port.align_3_intended(a, b, c)
        assert res[0] == enumerate_all(cost3, a, b, c), (a, b, c)
        if len(a) + len(b) + len(c) > 3:
            check_alignment(cost3, med3, a, b, c, res)


def test_intended_cube_equals_memoised_recursion(checker):
    port, cost3, med3 = checker
    rng = np.random.default_rng(2)
    for _ in range(40):
        a = seq(rng, int(rng.integers(1, 14)), ambiguity=0.1)
        b = a.copy() if rng.random() < 0.5 else seq(rng, int(rng.integers(1, 14)))
        c = seq(rng, int(rng.integers(1, 14)), ambiguity=0.1)
        if rng.random() < 0.5 and len(b) > 3:
            b = np.delete(b, 2)
        res = port.align_3_intended(a, b, c)
        assert res[0] == memoised(cost3, tuple(a), tuple(b), tuple(c))
        check_alignment(cost3, med3, a, b, c, res)


def test_executed_form_is_not_an_optimal_alignment(checker):
    """SURVEY.md A12's probe: the compiled reference (and its port, and the CUDA kernel) return 13 where the optimum is 0 on
    three identical 10-mers; the intended recurrence returns the optimum, and never more than the executed form."""
    port, cost3, _ = checker
    rng = np.random.default_rng(3)
    a = seq(rng, 10)
    assert port.align_3_intended(a, a, a)[0] == 0
    assert port.align_3(a, a, a)[0] > 0
    worse = 0
    for _ in range(30):
        x, y, z = (seq(rng, int(rng.integers(3, 12))) for _ in range(3))
        ci, ce = port.align_3_intended(x, y, z)[0], port.align_3(x, y, z)[0]
        assert ci == memoised(cost3, tuple(x), tuple(y), tuple(z))
        worse += ce > ci
    assert worse > 0
