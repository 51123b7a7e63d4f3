"""Powell's three-sequence affine Ukkonen aligner (src/ukk.checkp.c, the aligner behind
Sequence.Align.readjust_3d).  The CUDA kernel executes poyd_b200/csrc/powell_core.h; here the same source runs on one host
thread (tests/native/powell_host.cpp, a test harness -- the product library has no CPU path) and must reproduce the compiled
reference bit for bit: cost and the three aligned rows, i.e. every tie the reference's check-pointed recursion breaks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import powell_util as PU

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host():
    out = os.path.join(HERE, "..", "build", "libpwhost.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(HERE, "native", "powell_host.cpp")
    dep = os.path.join(HERE, "..", "poyd_b200", "csrc", "powell_core.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out, src])
    return C.CDLL(out)


@pytest.fixture(scope="module")
def ref():
    lib = PU.reference()
    if lib is None:
        pytest.skip("oracle/_ref/libpoyref.so with the Powell recipe not built")
    return lib


def host_powell(host, a, b, c, mm, go, ge):
    cap = len(a) + len(b) + len(c) + 2
    R = 16
    while True:
        rows = [np.zeros(cap, np.uint8) for _ in range(3)]
        n, st = C.c_int(0), C.c_int(0)
        cost = host.pw_host_align(a.ctypes.data_as(PU.u8), len(a), b.ctypes.data_as(PU.u8), len(b), c.ctypes.data_as(PU.u8), len(c),
                                  mm, go, ge, R, 0, *[r.ctypes.data_as(PU.u8) for r in rows], C.byref(n), C.byref(st), None)
        if st.value == 1 and R < 512:  # PW_EBOX: the diagonal box was too small, like the library's next round
            R *= 2
            continue
        return cost, [r[: n.value] for r in rows], st.value


def test_tables_match_setup(host, ref):
    """make_tables() against the globals the reference's own setup() leaves behind (src/ukkCommon.c:46-59, 247-350): the 16
    states in its order, their neighbour steps, continuation / second costs, the transition matrix, maxSingleStep."""
    a = np.array([16, 1, 2, 4], np.uint8)
    for mm, go, ge in PU.COSTS:
        PU.ref_powell(ref, a, a, a, mm, go, ge)  # runs setup() for these costs
        I27 = C.c_int * 27
        got = {k: I27() for k in ("neighbours", "contCost", "secondCost")}
        trans = (C.c_int * (27 * 27))()
        ns, mss = C.c_int(0), C.c_int(0)
        host.pw_host_tables(mm, go, ge, got["neighbours"], got["contCost"], got["secondCost"], trans, C.byref(ns), C.byref(mss))
        assert ns.value == C.c_int.in_dll(ref, "numStates").value == 16
        assert mss.value == C.c_int.in_dll(ref, "maxSingleStep").value, (mm, go, ge)
        for k, arr in got.items():
            want = I27.in_dll(ref, k)
            assert list(arr)[:16] == list(want)[:16], (k, (mm, go, ge))
        want_t = (C.c_int * (27 * 27)).in_dll(ref, "transCost")
        for s1 in range(16):
            assert list(trans)[s1 * 27: s1 * 27 + 16] == list(want_t)[s1 * 27: s1 * 27 + 16], (s1, (mm, go, ge))


def test_restatement_matches_the_reference(host, ref):
    bad = []
    for k, (a, b, c) in enumerate(PU.triples(seed=5, count=120, max_len=70)):
        mm, go, ge = PU.COSTS[k % len(PU.COSTS)]
        rc, rr = PU.ref_powell(ref, a, b, c, mm, go, ge)
        hc, hr, st = host_powell(host, a, b, c, mm, go, ge)
        if st != 0 or rc != hc or not all(np.array_equal(x, y) for x, y in zip(rr, hr)):
            bad.append((k, len(a), len(b), len(c), (mm, go, ge), rc, hc, st))
    assert not bad, bad[:5]


def test_restatement_on_longer_and_unequal_triples(host, ref):
    rng = np.random.default_rng(17)
    cases = []
    for n, p in ((150, 0.08), (220, 0.04), (120, 0.2)):
        a = PU.dna(rng, n)
        cases.append((a, PU.mutate(rng, a, p), PU.mutate(rng, a, p)))
    a = PU.dna(rng, 90)
    cases.append((a, a[:40].copy(), PU.mutate(rng, a, 0.05)))     # very different lengths: the final diagonal is far out
    cases.append((a, a.copy(), a.copy()))                          # identical: cost 0, one run of matches
    cases.append((PU.dna(rng, 1), PU.dna(rng, 1), PU.dna(rng, 1)))
    e, s4 = np.array([16], np.uint8), np.array([16, 1, 2, 4], np.uint8)  # empty operands (readjust_3d never passes them; same anyway)
    cases += [(e, s4, s4), (s4, e, s4), (s4, s4, e), (e, e, e), (e, e, s4)]
    for k, (a, b, c) in enumerate(cases):
        mm, go, ge = PU.COSTS[k % len(PU.COSTS)]
        rc, rr = PU.ref_powell(ref, a, b, c, mm, go, ge)
        hc, hr, st = host_powell(host, a, b, c, mm, go, ge)
        assert st == 0 and rc == hc, (k, rc, hc, st)
        for x, y in zip(rr, hr):
            assert np.array_equal(x, y), f"case {k}: aligned rows differ"


def test_element_without_a_base_is_an_error(host):
    a = np.array([16, 1, 2, 16, 4], np.uint8)  # a bare gap inside: copySequence raises "This is impossible!"
    b = np.array([16, 1, 2, 4], np.uint8)
    _, _, st = host_powell(host, a, b, b, 1, 3, 2)
    assert st == 5


def test_readjust_3d_branches():
    """The host part of Sequence.Align.readjust_3d (src/sequence.ml:1117-1129): which rows return a copy and which align."""
    from poyd_b200 import sequence as S

    e = np.array([16], np.uint8)
    a, b, c = np.array([16, 1, 2, 4], np.uint8), np.array([16, 1, 2], np.uint8), np.array([16, 1, 2, 4, 8], np.uint8)
    gg = np.array([16, 16], np.uint8)  # two gaps: empty as well (is_empty looks at every element)
    pool = S.SeqPool([e, a, b, c, a.copy(), gg])
    quads = np.array([[1, 4, 1, 4],   # four equal lengths: keep m
                      [0, 1, 2, 3],   # s1 empty: s2
                      [1, 5, 2, 3],   # s2 empty: s1
                      [1, 2, 3, 0],   # p empty: p
                      [1, 2, 3, 3],   # nothing empty, lengths differ: align
                      [0, 5, 2, 3]],  # s1 and s2 both empty, p not: the reference falls through to the aligner
                     np.int32)
    assert list(S.readjust_3d_classify(pool, quads, 16)) == [0, 1, 2, 3, 4, 4]
