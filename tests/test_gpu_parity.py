"""GPU parity tests: every result of the CUDA path, called through the C ABI, against the CPU checker
(the compiled reference algn.c when oracle/_ref is present, else the plain-C port) on the same seeded inputs.
Bar: bit-exact costs, aligned sequences and medians, including traceback tie-breaking."""
import numpy as np
import pytest

from helpers import assert_aligned_equal, right_rows_equal

pytestmark = pytest.mark.gpu

ALL = 7  # WANT_MEDIAN | WANT_MEDIANWG | WANT_ALIGNED


@pytest.fixture(scope="module")
def S():
    from poyd_b200 import sequence

    return sequence


def _affine_cases():
    from poyd_b200 import cost_matrix as CM

    return [("sub1_indel2_go3", CM.nucleotides(1, 2, 3)), ("sub2_indel1_go1", CM.nucleotides(2, 1, 1)),
            ("sub3_indel1_go5", CM.nucleotides(3, 1, 5)), ("sub1_indel2_go0", CM.nucleotides(1, 2, 0))]


def _linear_cases():
    from poyd_b200 import cost_matrix as CM

    return [("dna_1_2", CM.default_nucleotides(), "dna"), ("dna_1_1", CM.nucleotides(1, 1), "dna"),
            ("dna_3_1", CM.nucleotides(3, 1), "dna"), ("protein_1_2", CM.default_aminoacids(), "protein")]


@pytest.mark.parametrize("gap_amb", [0.0, 0.12])
def test_affine_align_ragged(S, checker_factory, gap_amb):
    from poyd_b200 import synth

    for name, cm in _affine_cases():
        pool, pairs = synth.ragged_batch(600, max_len=260, seed=11, gap_ambiguity=gap_amb)
        al = S.Align(cm)
        g = al.align_affine_3(pool, pairs, ALL)
        o = checker_factory(cm).batch(3, pool.pool, pool.off, pool.len, pairs)
        assert_aligned_equal(g, o, label=f"affine {name} gapamb={gap_amb}")
        gc = al.cost_2(pool, pairs)
        oc = checker_factory(cm).batch(2, pool.pool, pool.off, pool.len, pairs)["cost"]
        assert np.array_equal(gc, oc), f"cost_affine_3 {name}"
        al.close()


def test_affine_headline_shape(S, checker_factory):
    """configs[1] shape: 500 bp parents, 10 % substitutions, 2 % indels, gap opening 3."""
    from poyd_b200 import cost_matrix as CM, synth

    cm = CM.nucleotides(1, 2, 3)
    for gap_amb, seed in ((0.0, 2), (0.10, 3)):
        pool, pairs = synth.pair_batch(1500, 500, seed=seed, min_len=450, ambiguity=0.005 if gap_amb else 0.0,
                                       gap_ambiguity=gap_amb)
        al = S.Align(cm)
        g = al.align_affine_3(pool, pairs, ALL)
        o = checker_factory(cm).batch(3, pool.pool, pool.off, pool.len, pairs, nthreads=8)
        assert_aligned_equal(g, o, label=f"cfg2 gapamb={gap_amb}")
        oc = checker_factory(cm).batch(2, pool.pool, pool.off, pool.len, pairs, nthreads=8)["cost"]
        assert np.array_equal(al.cost_2(pool, pairs), oc)
        al.close()


def test_affine_empty_and_tiny(S, checker_factory):
    from poyd_b200 import cost_matrix as CM

    cm = CM.nucleotides(1, 2, 3)
    seqs = [np.array([16], np.uint8), np.array([16, 1], np.uint8), np.array([16, 17], np.uint8),
            np.array([16, 1, 2, 4, 8], np.uint8), np.array([16, 8, 4, 18, 1, 1, 1], np.uint8),
            np.array([16] + [1] * 45, np.uint8), np.array([16] + [2] * 41, np.uint8)]
    pool = S.SeqPool(seqs)
    pairs = np.array([(a, b) for a in range(len(seqs)) for b in range(len(seqs))], np.int32)
    al = S.Align(cm)
    g = al.align_affine_3(pool, pairs, ALL)
    o = checker_factory(cm).batch(3, pool.pool, pool.off, pool.len, pairs)
    assert_aligned_equal(g, o, label="affine tiny")
    al.close()


def test_linear_align_ragged(S, checker_factory):
    from poyd_b200 import synth

    for name, cm, alph in _linear_cases():
        pool, pairs = synth.ragged_batch(500, max_len=240, seed=5, alphabet=alph, gap_ambiguity=0.03 if alph == "dna" else 0)
        al = S.Align(cm)
        dw = al.deltaw_for(pool, pairs)
        g = al.align_2(pool, pairs, ALL)
        o = checker_factory(cm).batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=dw)
        assert_aligned_equal(g, o, label=f"linear {name}")
        gc = al.cost_2(pool, pairs)
        assert np.array_equal(gc, o["cost"]), f"cost_2 {name}"
        # the candidate-edge sweep hint of SeqCS.DOS.distance: deltaw = max 8 |la - lb| (src/seqCS.ml:856-866)
        la, lb = pool.len[pairs[:, 0]], pool.len[pairs[:, 1]]
        hint = np.maximum(8, np.abs(la - lb)).astype(np.int32)
        dw2 = al.deltaw_for(pool, pairs, hint)
        oc2 = checker_factory(cm).batch(0, pool.pool, pool.off, pool.len, pairs, deltaw=dw2)["cost"]
        assert np.array_equal(al.cost_2(pool, pairs, deltaw=hint), oc2), f"cost_2 with deltaw hint {name}"
        al.close()


def test_linear_explicit_bands(S, checker_factory):
    """deltaw is an input of the external (algn_CAML_simple_2): sweep it directly, narrow to full."""
    from poyd_b200 import cost_matrix as CM, synth

    cm = CM.default_nucleotides()
    pool, pairs = synth.pair_batch(300, 300, seed=9, min_len=240)
    al = S.Align(cm)
    for dwv in (0, 3, 16, 60, 271):
        dw = np.full(len(pairs), dwv, np.int32)
        g = al.align_2(pool, pairs, ALL, deltaw=dw, raw_deltaw=True)
        o = checker_factory(cm).batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=dw)
        assert_aligned_equal(g, o, label=f"linear deltaw={dwv}")
    al.close()


def test_protein_config_shapes(S, checker_factory):
    """configs[2]: 300 aa pairs, 22x22 matrix; (3a) deltaw as the product computes it (full matrix, SURVEY.md
    A15 quirk) and (3b) an explicit band of 16."""
    from poyd_b200 import cost_matrix as CM, synth

    cm = CM.default_aminoacids()
    pool, pairs = synth.pair_batch(400, 300, seed=3, alphabet="protein", subst=0.15, indel=0.02)
    al = S.Align(cm)
    dwa = al.deltaw_for(pool, pairs)
    assert dwa.min() > 200  # the quirk: 18 of the 21 residue codes share a bit with gap code 22
    for dw in (dwa, np.full(len(pairs), 16, np.int32)):
        g = al.align_2(pool, pairs, ALL, deltaw=dw, raw_deltaw=True)
        o = checker_factory(cm).batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=dw, nthreads=8)
        assert_aligned_equal(g, o, label="protein")
    al.close()


def test_median_2_entry_points(S, checker_factory):
    from poyd_b200 import cost_matrix as CM, synth

    for cm in (CM.default_nucleotides(), CM.nucleotides(1, 2, 3), CM.default_aminoacids()):
        alph = "dna" if cm.combinations else "protein"
        pool, pairs = synth.ragged_batch(200, max_len=120, seed=21, alphabet=alph, gap_ambiguity=0.1 if alph == "dna" else 0)
        chk = checker_factory(cm)
        lin = cm if cm.cost_model_type == 0 else CM.nucleotides(1, 2)
        src = checker_factory(lin).batch(1, pool.pool, pool.off, pool.len, pairs,
                                         deltaw=np.full(len(pairs), 30, np.int32))
        a, b, ln = src["ra"], src["rb"], src["lens"][:, 2].copy()
        stride = (a.shape[1] + 15) // 16 * 16
        a = np.pad(a, ((0, 0), (0, stride - a.shape[1])))
        b = np.pad(b, ((0, 0), (0, stride - b.shape[1])))
        al = S.Align(cm)
        for which, fn in ((0, al.ancestor_2), (1, al.median_2_with_gaps), (2, al.median_2)):
            out, olen = fn(a, b, ln)
            for p in range(len(pairs)):
                exp = chk.median_2(which, a[p, :ln[p]], b[p, :ln[p]])
                got = out[p, out.shape[1] - olen[p]:]
                assert np.array_equal(got, exp), f"median_2 which={which} pair {p}"
        al.close()


def test_shared_operand_sweep(S, checker_factory):
    """Candidate-edge sweep: one clade sequence against many edges -- the pool holds each sequence once."""
    from poyd_b200 import cost_matrix as CM, synth

    cm = CM.nucleotides(1, 2, 3)
    pool, _ = synth.pair_batch(64, 400, seed=17, min_len=350)
    pairs = np.array([(0, k) for k in range(1, len(pool))], np.int32)
    al = S.Align(cm)
    oc = checker_factory(cm).batch(2, pool.pool, pool.off, pool.len, pairs)["cost"]
    assert np.array_equal(al.cost_2(pool, pairs), oc)
    al.close()


def test_error_paths(S):
    from poyd_b200 import cost_matrix as CM

    al = S.Align(CM.nucleotides(1, 2, 3))
    pool = S.SeqPool([np.array([16, 1, 2], np.uint8), np.array([16, 1], np.uint8)])
    with pytest.raises(S.PoyB200Error):
        al.cost_2(pool, np.array([[0, 5]], np.int32))  # pair index out of range
    lin = S.Align(CM.default_nucleotides())
    b, _ = lin.make_batch(pool, np.array([[0, 1]], np.int32))
    with pytest.raises(S.PoyB200Error):
        lin._check(lin.L.poyb200_batch_cost_2(lin.h, __import__("ctypes").byref(b)))  # deltaw missing
    al.close()
    lin.close()


def _pairs_with_length_gap(rng, n, diffs, base_lo=30, base_hi=260, gap_amb=0.08):
    seqs = []
    for k in range(n):
        la = int(rng.integers(base_lo, base_hi))
        lb = la + int(diffs[k % len(diffs)])
        a = rng.choice(np.array([1, 2, 4, 8], np.uint8), size=la)
        b = rng.choice(np.array([1, 2, 4, 8], np.uint8), size=lb)
        m = min(la, lb)
        keep = rng.random(m) < 0.8
        b[:m][keep] = a[:m][keep]  # related prefix so that the band matters
        for s in (a, b):
            s[rng.random(len(s)) < gap_amb] |= 16
        a = np.concatenate([[16], a]).astype(np.uint8)
        b = np.concatenate([[16], b]).astype(np.uint8)
        seqs += ([a, b] if k % 2 == 0 else [b, a])
    return seqs


def test_affine_every_stripe_shape(S, checker_factory, monkeypatch):
    """Length differences that select each register-stripe shape (W = 39 + max(40, |dl| + 8)), the spare-diagonal
    (LOW) variants, and the generic kernel beyond the widest shape; then the same batch forced through the generic
    kernels only."""
    from poyd_b200 import cost_matrix as CM

    cm = CM.nucleotides(1, 2, 3)
    rng = np.random.default_rng(31)
    diffs = [0, 5, 32, 33, 40, 49, 50, 80, 81, 120, 145, 146, 200, 209, 210, 300, 337, 338, 400, 465, 466, 520, 700]
    pool = S.SeqPool(_pairs_with_length_gap(rng, 4 * len(diffs), diffs))
    pairs = np.arange(2 * 4 * len(diffs), dtype=np.int32).reshape(-1, 2)
    chk = checker_factory(cm)
    o = chk.batch(3, pool.pool, pool.off, pool.len, pairs, nthreads=8)
    oc = chk.batch(2, pool.pool, pool.off, pool.len, pairs, nthreads=8)["cost"]
    # ring kernels (default), the legacy fast / stripe kernels + separate traceback, the generic kernels
    for config in ({}, {"use_ring": 1}, {"use_ring": 0}, {"use_ring": 0, "allow_fast": 0}, {"force_generic": 1}):
        al = S.Align(cm, config=config)
        g = al.align_affine_3(pool, pairs, ALL)
        assert_aligned_equal(g, o, label=f"affine shapes {config}")
        assert np.array_equal(al.cost_2(pool, pairs), oc), f"affine cost shapes {config}"
        al.close()


def test_affine_generic_matches_on_headline_shape(S, checker_factory, monkeypatch):
    from poyd_b200 import cost_matrix as CM, synth

    cm = CM.nucleotides(1, 2, 3)
    pool, pairs = synth.pair_batch(300, 500, seed=8, min_len=450, gap_ambiguity=0.05)
    o = checker_factory(cm).batch(3, pool.pool, pool.off, pool.len, pairs, nthreads=8)
    al = S.Align(cm, config={"force_generic": 1})
    assert_aligned_equal(al.align_affine_3(pool, pairs, ALL), o, label="generic kernel, cfg2 shape")
    al.close()


def test_gpu_replays_reference_golden(S):
    """The committed fixtures (outputs of the compiled reference) through the C ABI: needs no checker at all."""
    import os

    from golden_util import golden_files, load

    for path in golden_files():
        cm, pool, pairs, mode, deltaw, ref = load(path)
        al = S.Align(cm)
        name = os.path.basename(path)
        if mode == 3:
            assert_aligned_equal(al.align_affine_3(pool, pairs, ALL), ref, label=name)
        elif mode == 2:
            assert np.array_equal(al.cost_2(pool, pairs), ref["cost"]), name
        elif mode == 1:
            assert_aligned_equal(al.align_2(pool, pairs, ALL, deltaw=deltaw, raw_deltaw=True), ref, label=name)
            assert np.array_equal(al.cost_2(pool, pairs, deltaw=deltaw, raw_deltaw=True), ref["cost"]), name
        al.close()


def test_large_batch_properties(S):
    """BASELINE.json size (1M pairs is run by bench.py; here 120k): size-independent properties.
    * symmetry of the cost-only kernel in its operands (the C stub orders them by length, src/algn.c:2661);
    * the traceback cost never exceeds ... equals the recomputed cost of its own alignment under the linear
      model is not defined for affine_3, so the checked invariants are structural: aligned rows have equal
      length, removing gaps from them gives back the inputs, and medianwg has that same length."""
    from poyd_b200 import cost_matrix as CM, synth

    cm = CM.nucleotides(1, 2, 3)
    pool, pairs = synth.pair_batch(120_000, 500, seed=41, min_len=450)
    al = S.Align(cm)
    c1 = al.cost_2(pool, pairs)
    c2 = al.cost_2(pool, pairs[:, ::-1].copy())
    assert np.array_equal(c1, c2)
    r = al.align_affine_3(pool, pairs, ALL)
    assert (r.lens[:, 2] == r.lens[:, 3]).all() and (r.lens[:, 1] == r.lens[:, 2]).all()
    assert (r.lens[:, 0] <= r.lens[:, 1]).all()
    stride = r.aligned_a.shape[1]
    rng = np.random.default_rng(0)
    for p in rng.integers(0, len(pairs), size=400):
        for buf, k, s in ((r.aligned_a, 2, pairs[p, 0]), (r.aligned_b, 3, pairs[p, 1])):
            row = buf[p, stride - r.lens[p, k]:]
            body = row[1:]
            assert row[0] == 16
            assert np.array_equal(body[body != 16], pool.seq(s)[1:])
    al.close()


def test_linear_custom_tail_and_prepend_costs(S, checker_factory, monkeypatch):
    """Cost_matrix.fill_tail / fill_prepend (src/cost_matrix.ml:418-460) make tail_cost[a] differ from cost(a, gap): the
    last-column rule (algn_fill_last_column) and the first-column / first-row costs then really matter."""
    from poyd_b200 import cost_matrix as CM, synth

    cm = CM.default_nucleotides().clone()
    rng = np.random.default_rng(12)
    cm.tail_cost[1:32] = rng.integers(0, 4, size=31)
    cm.prepend_cost[1:32] = rng.integers(0, 4, size=31)
    pool, pairs = synth.ragged_batch(400, max_len=200, seed=15, gap_ambiguity=0.02)
    chk = checker_factory(cm)
    for force in (0, 1):
        al = S.Align(cm, config={"force_generic": force})
        for dwv in (2, 20, 200):
            dw = np.full(len(pairs), dwv, np.int32)
            g = al.align_2(pool, pairs, ALL, deltaw=dw, raw_deltaw=True)
            o = chk.batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=dw, nthreads=4)
            assert_aligned_equal(g, o, label=f"custom tail deltaw={dwv} force_generic={force}")
        al.close()


def test_linear_full_matrix_rows_kernel(S, checker_factory):
    """Full matrices (algn_fill_plane: cases 1 and 3a of algn_fill_plane_2, src/algn.c:893, :936) take the column-striped
    lin_rows_kernel: every shape of it (columns 1 .. 512), both tie orders, custom tail / prepend costs, against the
    checker and against the diagonal-stripe kernels (allow_rows = 0)."""
    from poyd_b200 import cost_matrix as CM, synth

    tail_cm = CM.default_nucleotides().clone()
    rng = np.random.default_rng(21)
    tail_cm.tail_cost[1:32] = rng.integers(0, 4, size=31)
    tail_cm.prepend_cost[1:32] = rng.integers(0, 4, size=31)
    cases = [("dna", CM.default_nucleotides(), "dna"), ("dna_3_1", CM.nucleotides(3, 1), "dna"),
             ("protein", CM.default_aminoacids(), "protein"), ("dna_custom_tail", tail_cm, "dna")]
    for name, cm, alph in cases:
        chk = checker_factory(cm)
        batches = [synth.ragged_batch(300, max_len=130, seed=31, alphabet=alph, gap_ambiguity=0.02 if alph == "dna" else 0),
                   synth.pair_batch(64, 500, seed=32, alphabet=alph, min_len=380, indel=0.05),
                   synth.pair_batch(96, 250, seed=33, alphabet=alph, min_len=100, indel=0.1)]
        al, al0 = S.Align(cm), S.Align(cm, config={"allow_rows": 0})
        for bi, (pool, pairs) in enumerate(batches):
            dw = np.full(len(pairs), 600, np.int32)  # 8 >= l1 - height: the full matrix whatever the lengths
            o = chk.batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=dw, nthreads=8)
            g = al.align_2(pool, pairs, ALL, deltaw=dw, raw_deltaw=True)
            assert_aligned_equal(g, o, label=f"rows kernel {name} batch {bi}")
            g0 = al0.align_2(pool, pairs, ALL, deltaw=dw, raw_deltaw=True)
            assert_aligned_equal(g0, o, label=f"stripe kernels on full matrices {name} batch {bi}")
            assert np.array_equal(al.cost_2(pool, pairs, deltaw=dw, raw_deltaw=True), o["cost"]), f"cost-only rows kernel {name} {bi}"
            for flag in (0, 1):
                sw = np.full(len(pairs), flag, np.uint8)
                ga = al.align_2(pool, pairs, ALL, deltaw=dw, raw_deltaw=True, swaped=sw)
                gb = al0.align_2(pool, pairs, ALL, deltaw=dw, raw_deltaw=True, swaped=sw)
                assert np.array_equal(ga.cost, gb.cost) and np.array_equal(ga.lens, gb.lens), f"swaped={flag} {name} batch {bi}"
                for k, attr in enumerate(("median", "medianwg", "aligned_a", "aligned_b")):
                    x, y = getattr(ga, attr), getattr(gb, attr)
                    live = np.arange(x.shape[1])[None, :] >= x.shape[1] - ga.lens[:, k][:, None]
                    assert np.array_equal(np.where(live, x, 0), np.where(live, y, 0)), f"swaped={flag} {name} batch {bi}: {attr}"
        al.close()
        al0.close()


def test_linear_generic_matches_stripe(S, checker_factory, monkeypatch):
    from poyd_b200 import cost_matrix as CM, synth

    cm = CM.default_nucleotides()
    pool, pairs = synth.pair_batch(300, 500, seed=6, min_len=450)
    chk = checker_factory(cm)
    al = S.Align(cm, config={"force_generic": 1})
    dw = al.deltaw_for(pool, pairs)
    o = chk.batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=dw, nthreads=8)
    assert_aligned_equal(al.align_2(pool, pairs, ALL), o, label="linear generic cfg2-lin")
    al.close()
    al = S.Align(cm, config={"force_generic": 0})
    assert_aligned_equal(al.align_2(pool, pairs, ALL), o, label="linear stripe cfg2-lin")
    # explicit swaped flag of the external (algn_CAML_backtrack_2d): both values, on equal-length operands
    same = np.nonzero(pool.len[pairs[:, 0]] == pool.len[pairs[:, 1]])[0][:64]
    if len(same):
        sub = pairs[same]
        for flag in (0, 1):
            g = al.align_2(pool, sub, 4, deltaw=dw[same], raw_deltaw=True, swaped=np.full(len(sub), flag, np.uint8))
            ref = checker_factory(cm)
            for k, (a, b) in enumerate(sub):
                if hasattr(ref, "L") and ref.kind == "reference":
                    import ctypes as C

                    s1, s2 = np.ascontiguousarray(pool.seq(a)), np.ascontiguousarray(pool.seq(b))
                    cap = len(s1) + len(s2)
                    r1, r2, rl = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8), C.c_int(0)
                    ref.L.ref_align_2(ref.h, ref.ws, s1.ctypes.data_as(C.POINTER(C.c_uint8)), len(s1),
                                      s2.ctypes.data_as(C.POINTER(C.c_uint8)), len(s2), int(dw[same][k]), flag,
                                      r1.ctypes.data_as(C.POINTER(C.c_uint8)), r2.ctypes.data_as(C.POINTER(C.c_uint8)),
                                      C.byref(rl))
                    assert np.array_equal(g.get("aligned_a", k), r1[: rl.value]), (flag, k)
                    assert np.array_equal(g.get("aligned_b", k), r2[: rl.value]), (flag, k)
    al.close()


def test_maximum_lengths_take_the_generic_kernels(S, checker_factory):
    """Sequences near the reference's 16384-element cap (src/seq.c:370) and pairs wider than any register shape:
    they must run (generic CUDA kernels) and still match."""
    from poyd_b200 import cost_matrix as CM

    rng = np.random.default_rng(99)

    def rnd(n, gapamb=0.0):
        s = rng.choice(np.array([1, 2, 4, 8], np.uint8), size=n)
        if gapamb:
            s[rng.random(n) < gapamb] |= 16
        return np.concatenate([[16], s]).astype(np.uint8)

    a = rnd(6000, 0.02)
    b = a.copy()
    b[1:][rng.random(6000) < 0.1] = 4
    seqs = [a, b, rnd(16383), rnd(2500), rnd(900), rnd(5200, 0.02), rnd(400)]
    pool = S.SeqPool(seqs)
    pairs = np.array([[0, 1], [3, 4], [4, 0], [5, 1]], np.int32)
    aff = CM.nucleotides(1, 2, 3)
    al = S.Align(aff)
    o = checker_factory(aff).batch(3, pool.pool, pool.off, pool.len, pairs, nthreads=4)
    assert_aligned_equal(al.align_affine_3(pool, pairs, ALL), o, label="affine long")
    al.close()
    lin = CM.default_nucleotides()
    al = S.Align(lin)
    # [2, 3]: 16384 x 2501, full matrix (l1 >= 1.5 l2), too wide for any register kernel; [2, 6]: 16384 x 401, full matrix
    # in the column-striped kernel at the staging cap
    pairs = np.array([[0, 1], [3, 4], [2, 3], [1, 5], [2, 6], [6, 2]], np.int32)
    dw = al.deltaw_for(pool, pairs)
    o = checker_factory(lin).batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=dw, nthreads=4)
    assert_aligned_equal(al.align_2(pool, pairs, ALL), o, label="linear long")
    al.close()


def test_three_sequence_cube_as_the_reference_executes_it(S):
    """configs[3] path (algn_nw_3d + backtrack_3d + algn_get_median_3d).  The reference's cube fill is defective
    (SURVEY.md A12-A14); parity means the same costs, walks and medians as the compiled reference, triple by triple."""
    from oracle import oracle
    from poyd_b200 import cost_matrix as CM

    oracle.build(ref=True)
    cm = CM.default_nucleotides()
    cm3 = CM.of_two_dim(cm)
    chk = oracle.best_checker_3(cm3)
    rng = np.random.default_rng(8)

    def rnd(n):
        return np.concatenate([[16], rng.choice(np.array([1, 2, 4, 8], np.uint8), size=n)]).astype(np.uint8)

    seqs, triples = [], []
    shapes = [(0, 0, 0), (1, 0, 2), (5, 5, 5), (12, 30, 7), (40, 40, 40), (33, 20, 70), (60, 61, 59), (1, 50, 1),
              (90, 80, 100), (100, 100, 100), (20, 20, 600), (150, 20, 20)]
    for (n1, n2, n3) in shapes * 2:
        a = rnd(n1)
        b = a.copy() if (n2 == n1 and rng.random() < 0.7) else rnd(n2)
        if len(b) > 1:
            b[1:][rng.random(len(b) - 1) < 0.1] = 2
        c = a.copy() if (n3 == n1 and rng.random() < 0.7) else rnd(n3)
        if len(c) > 1:
            c[1:][rng.random(len(c) - 1) < 0.1] = 8
        k = len(seqs)
        seqs += [a, b, c]
        triples.append((k, k + 1, k + 2))
    pool = S.SeqPool(seqs)
    triples = np.array(triples, np.int32)
    al = S.Align3(cm, cm3)
    g = al.align_3(pool, triples, want=3)
    assert np.array_equal(al.cost_3(pool, triples), g.cost)
    for t, (i1, i2, i3) in enumerate(triples):
        cost, status, r1, r2, r3, med = chk.align_3(pool.seq(i1), pool.seq(i2), pool.seq(i3))
        assert g.cost[t] == cost, (t, g.cost[t], cost)
        assert g.status[t] == status, t
        if status == 0:
            assert g.lens[t] == len(r1), t
            assert np.array_equal(g.get("aligned_1", t), r1), t
            assert np.array_equal(g.get("aligned_2", t), r2), t
            assert np.array_equal(g.get("aligned_3", t), r3), t
            assert np.array_equal(g.get("median", t), med), t
    al.close()


def test_cube_at_the_configured_size_300(S):
    """configs[3] shape: 300 bp triples (301 x 301 x 301 cells).  A handful of triples against the compiled reference,
    every output (cost, status, the three aligned sequences, the median)."""
    from oracle import oracle
    from poyd_b200 import cost_matrix as CM, synth

    oracle.build(ref=True)
    cm = CM.default_nucleotides()
    cm3 = CM.of_two_dim(cm)
    chk = oracle.best_checker_3(cm3)
    pool, triples = synth.triple_batch(5, 300, seed=77)
    assert int(pool.len.max()) == 301
    al = S.Align3(cm, cm3)
    g = al.align_3(pool, triples, want=3)
    assert np.array_equal(al.cost_3(pool, triples), g.cost)
    for t, (i1, i2, i3) in enumerate(triples):
        cost, status, r1, r2, r3, med = chk.align_3(pool.seq(int(i1)), pool.seq(int(i2)), pool.seq(int(i3)))
        assert g.cost[t] == cost, (t, g.cost[t], cost)
        assert g.status[t] == status, t
        if status == 0:
            assert g.lens[t] == len(r1), t
            for name, want in (("aligned_1", r1), ("aligned_2", r2), ("aligned_3", r3), ("median", med)):
                assert np.array_equal(g.get(name, t), want), (t, name)
    al.close()


def test_one_million_pairs_bit_exact(S, checker_factory):
    """BASELINE.json target: bit-exact costs and medians against algn.c on >= 1 M synthetic pairs (configs[1] shape:
    500 bp DNA, affine).  Compared in slices against the compiled reference on all host threads; half of the slices carry
    gap-ambiguous codes (median-like operands, full block-diagonal cell), half do not (NOEB fast path)."""
    import os

    from poyd_b200 import cost_matrix as CM, synth

    cm = CM.nucleotides(1, 2, 3)
    chk = checker_factory(cm)
    threads = len(os.sched_getaffinity(0))
    al = S.Align(cm)
    total, slice_pairs = 0, 125_000
    for k in range(8):
        extra = dict(ambiguity=0.005, gap_ambiguity=0.10) if k % 2 else {}
        pool, pairs = synth.pair_batch(slice_pairs, 500, seed=1000 + k, min_len=450, stride=512, **extra)
        g = al.align_affine_3(pool, pairs, S.WANT_MEDIAN)
        o = chk.batch(3, pool.pool, pool.off, pool.len, pairs, nthreads=threads)
        assert np.array_equal(g.cost, o["cost"]), f"slice {k}: cost mismatch"
        assert np.array_equal(g.lens[:, 0], o["lens"][:, 0]), f"slice {k}: median length mismatch"
        # medians: compare as right-aligned blocks in one vectorised pass
        L = o["lens"][:, 0]
        w = int(L.max())
        gm = g.median[:, g.median.shape[1] - w:]
        om = np.zeros((slice_pairs, w), np.uint8)
        cols = np.arange(w)[None, :]
        src = cols - (w - L[:, None])
        valid = src >= 0
        om[valid] = o["median"][np.nonzero(valid)[0], src[valid]]
        assert np.array_equal(np.where(valid, gm, 0), om), f"slice {k}: median mismatch"
        cg = al.cost_2(pool, pairs)
        oc = chk.batch(2, pool.pool, pool.off, pool.len, pairs, nthreads=threads)["cost"]
        assert np.array_equal(cg, oc), f"slice {k}: cost-only mismatch"
        if k < 2:
            # one leaf-like and one median-like slice: every output of align_affine_3 (median, medianwg, both aligned
            # sequences), bytewise, and the DOS.median payload built from the same walk
            assert_aligned_equal(al.align_affine_3(pool, pairs, ALL), o, label=f"slice {k}, all outputs")
            gb = al.align_affine_3(pool, pairs, S.WANT_MEDIAN | S.WANT_BITSETS)
            n2 = o["lens"][:, 2]
            for name, key in (("a", "ra"), ("b", "rb"), ("wg", "medianwg")):
                bits = np.unpackbits(getattr(gb, "bits_" + name), axis=1)
                assert right_rows_equal(bits, (o[key] != cm.gap).astype(np.uint8), n2), f"slice {k}: bitset {name}"
        total += slice_pairs
    assert total >= 1_000_000
    al.close()


@pytest.mark.parametrize("affine", [True, False])
def test_dos_median_payload_bitsets(S, checker_factory, affine):
    """POYB200_WANT_BITSETS: cost + median + the three gap bitsets SeqCS.DOS.median keeps (src/seqCS.ml:769-771) --
    without transferring the aligned sequences -- against seq_to_bitset of the checker's aligned sequences, for
    ragged lengths (so operands swap roles) and for operands that carry gap bits."""
    from poyd_b200 import cost_matrix as CM, synth

    cm = CM.nucleotides(1, 2, 3) if affine else CM.default_nucleotides()
    chk = checker_factory(cm)
    for pool, pairs in (synth.ragged_batch(1500, max_len=260, seed=41),
                        synth.pair_batch(700, 500, seed=42, min_len=430, gap_ambiguity=0.08)):
        al = S.Align(cm)
        if affine:
            g = al.align_affine_3(pool, pairs, want=S.WANT_MEDIAN | S.WANT_BITSETS)
            o = chk.batch(3, pool.pool, pool.off, pool.len, pairs)
        else:
            dw = al.deltaw_for(pool, pairs)
            g = al.align_2(pool, pairs, want=S.WANT_MEDIAN | S.WANT_BITSETS)
            o = chk.batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=dw)
        assert g.aligned_a is None and g.medianwg is None
        assert np.array_equal(g.cost, o["cost"]) and np.array_equal(g.lens, o["lens"])
        for p in range(len(pairs)):
            n, nm = int(o["lens"][p, 2]), int(o["lens"][p, 0])
            assert np.array_equal(g.get("median", p), o["median"][p, :nm])
            assert np.array_equal(g.bitset("a", p), (o["ra"][p, :n] != cm.gap).astype(np.uint8)), p
            assert np.array_equal(g.bitset("b", p), (o["rb"][p, :n] != cm.gap).astype(np.uint8)), p
            assert np.array_equal(g.bitset("wg", p), (o["medianwg"][p, :n] != cm.gap).astype(np.uint8)), p
        # the full payload in one call gives the same bitsets
        g2 = al.align_affine_3(pool, pairs, want=15) if affine else al.align_2(pool, pairs, want=15)
        for p in range(0, len(pairs), 7):
            assert np.array_equal(g2.bitset("a", p), g.bitset("a", p)) and np.array_equal(g2.bitset("wg", p), g.bitset("wg", p))
        al.close()


@pytest.mark.parametrize("kind", ["affine", "linear", "protein"])
def test_closest_fused_kernel(S, checker_factory, kind):
    """Sequence.Align.closest (src/sequence.ml:967-1033) fused into the traceback (POYB200_WANT_CLOSEST) against the
    column rule applied to the checker's aligned pair: get_closest as a table for DNA (tree.closest_table, a restatement
    of src/cost_matrix.ml:681-700), the `all`-code rule for protein (:986-997)."""
    from poyd_b200 import cost_matrix as CM, synth, tree as T

    cm = {"affine": CM.nucleotides(1, 2, 3), "linear": CM.default_nucleotides(), "protein": CM.default_aminoacids()}[kind]
    alphabet = "protein" if kind == "protein" else "dna"
    pool, pairs = synth.ragged_batch(1200, max_len=240, seed=51, alphabet=alphabet, gap_ambiguity=0.0 if kind == "protein" else 0.06)
    pairs = np.concatenate([pairs, np.stack([pairs[:40, 0], pairs[:40, 0]], axis=1)])  # s1 = s2 early exit
    al = S.Align(cm)
    got = al.closest(pool, pairs)
    chk = checker_factory(cm)
    mode = 3 if kind == "affine" else 1
    o = chk.batch(mode, pool.pool, pool.off, pool.len, pairs, deltaw=None if kind == "affine" else al.deltaw_for(pool, pairs))
    tab = T.closest_table(cm) if cm.combine() else None
    for p, (i, j) in enumerate(pairs):
        s1, s2 = pool.seq(int(i)), pool.seq(int(j))
        if np.all(s2 == cm.gap):
            want = s2
        else:
            n = int(o["lens"][p, 2])
            a1, b1 = o["ra"][p, :n], o["rb"][p, :n]
            if tab is not None and len(s1) == len(s2) and np.array_equal(s1, s2):
                a1 = s1.copy()
                a1[1:] &= np.uint8(~cm.gap & 0xFF)
                b1 = a1
            if tab is not None:
                sel = tab[a1, b1]
            else:
                sel = np.where(b1 == cm.all_elements, np.where(a1 == cm.all_elements, 1, a1), b1).astype(np.uint8)
            want = np.concatenate([[cm.gap], sel[sel != cm.gap]]).astype(np.uint8)
        assert np.array_equal(got[p], want), f"{kind} pair {p}"
    al.close()


def test_multi_device_call_matches_single_device(S, checker_factory):
    """poyb200_multi_batch (csrc/multi.cu): one batch cut into contiguous shards of about equal work, one context and one
    host thread per shard, every shard uploading only its window of the pool and writing into the caller's rows.  On a
    one-GPU box the three contexts share device 0 -- the host logic (cuts, pool windows, row offsets) is what is tested;
    bench.py --gpus N runs the same call over N devices."""
    import torch

    from poyd_b200 import cost_matrix as CM, synth

    ndev = torch.cuda.device_count()
    devices = [k % ndev for k in range(3)]
    for cm, mode in ((CM.nucleotides(1, 2, 3), 3), (CM.default_nucleotides(), 1)):
        pool, pairs = synth.ragged_batch(900, max_len=300, seed=61, gap_ambiguity=0.05 if mode == 3 else 0.0)
        # a candidate-edge sweep at the end: one operand shared by many pairs, far away in the pool
        sweep = np.stack([np.zeros(200, np.int32), np.arange(1, 401, 2, dtype=np.int32)], axis=1)
        pairs = np.concatenate([pairs, sweep])
        ma = S.MultiAlign(cm, devices)
        al = S.Align(cm)
        if mode == 3:
            g, s = ma.align_affine_3(pool, pairs, ALL | S.WANT_BITSETS), al.align_affine_3(pool, pairs, ALL | S.WANT_BITSETS)
            o = checker_factory(cm).batch(3, pool.pool, pool.off, pool.len, pairs, nthreads=8)
        else:
            g, s = ma.align_2(pool, pairs, ALL | S.WANT_BITSETS), al.align_2(pool, pairs, ALL | S.WANT_BITSETS)
            o = checker_factory(cm).batch(1, pool.pool, pool.off, pool.len, pairs, deltaw=al.deltaw_for(pool, pairs), nthreads=8)
        assert_aligned_equal(g, o, label=f"multi-device mode {mode}")
        for p in range(0, len(pairs), 5):  # the bitsets of the sharded call = those of the one-device call
            for name in ("a", "b", "wg"):
                assert np.array_equal(g.bitset(name, p), s.bitset(name, p)), (mode, p, name)
        assert np.array_equal(ma.cost_2(pool, pairs), al.cost_2(pool, pairs))
        cuts = ma.shards()
        assert cuts[0] == 0 and cuts[-1] == len(pairs) and np.all(np.diff(cuts) > 0), cuts
        ma.close()
        al.close()


@pytest.mark.parametrize("affine", [False, True])
def test_uppass_three_medians_and_union_distance(S, affine):
    """SeqCS.DOS.median_3_no_union / median_3_union / distance and SeqCS.Union.distance_union (src/seqCS.ml:778-867,
    1569-1616) as batches, against the same compositions of the compiled reference's own functions (align_2 /
    align_affine_3, median_2, worst_2), vertex by vertex."""
    from oracle import oracle
    from poyd_b200 import cost_matrix as CM, seqcs, synth

    oracle.build(ref=True)
    cm = CM.nucleotides(1, 2, 3) if affine else CM.default_nucleotides()
    chk = oracle.Reference(cm)
    gap = cm.gap
    pool, pairs = synth.ragged_batch(150, max_len=220, seed=71, gap_ambiguity=0.04 if affine else 0.0, min_len=5)
    n = 100
    parent, c1, c2 = pairs[:n, 0], pairs[:n, 1], pairs[50:50 + n, 1]
    al = S.Align(cm)
    dos = seqcs.DOS(al)

    def ref_align(a, b):  # Sequence.Align.align_2: affine matrices are diverted to align_affine_3 (src/sequence.ml:851-858)
        if affine:
            c, _, _, ra, rb = chk.align_affine_3(a, b)
            return ra, rb, c
        c, ra, rb = chk.align_2(a, b, int(al.deltaw_for(S.SeqPool([a, b]), np.array([[0, 1]], np.int32))[0]))
        return ra, rb, c

    def ref_with_parent(p, c):
        ra, rb, cost = ref_align(p, c)
        return chk.median_2(2, ra, rb), cost, chk.worst_2(ra, rb)

    seqs, cmin, cmax = dos.median_3_no_union(pool, parent, c1, c2)
    for k in range(n):
        m1, k1, w1 = ref_with_parent(pool.seq(int(parent[k])), pool.seq(int(c1[k])))
        m2, k2, w2 = ref_with_parent(pool.seq(int(parent[k])), pool.seq(int(c2[k])))
        m, c, w = (m1, k1, w1) if k1 < k2 else (m2, k2, w2)
        if len(m) == 0 or m[0] != gap:
            m = np.concatenate([[gap], m]).astype(np.uint8)
        assert np.array_equal(seqs[k], m) and cmin[k] == c and cmax[k] == w, k
    # median_3_union: the aligned children of a vertex = the aligned pair of (c1, c2)
    al_a, al_b = [], []
    for k in range(n):
        ra, rb, _ = ref_align(pool.seq(int(c1[k])), pool.seq(int(c2[k])))
        al_a.append(ra)
        al_b.append(rb)
    useqs, ucost, uworst = dos.median_3_union(pool, parent, al_a, al_b)
    for k in range(n):
        u = np.bitwise_or(al_a[k], al_b[k])
        ra, rb, c = ref_align(pool.seq(int(parent[k])), u)
        m = chk.median_2(2, ra, rb)
        if len(m) == 0 or m[0] != gap:
            m = np.concatenate([[gap], m]).astype(np.uint8)
        assert np.array_equal(useqs[k], m) and ucost[k] == c and uworst[k] == chk.worst_2(ra, rb), k
    # distances with the DOS.distance hint; Union.distance_union scales them
    d = dos.distance(pool, pairs[:n])
    la, lb = pool.len[pairs[:n, 0]].astype(np.int64), pool.len[pairs[:n, 1]].astype(np.int64)
    want = al.cost_2(pool, pairs[:n], deltaw=np.maximum(np.abs(la - lb), 8))
    assert np.array_equal(d, want)
    assert np.allclose(seqcs.Union(al).distance_union(pool, pairs[:n]), (0.8 if affine else 1.0) * want)
    al.close()


def test_powell_3d_aligner_matches_the_reference(S):
    """poyb200_batch_powell_3 (the CUDA port of src/ukk.checkp.c) against the compiled reference's
    powell_3D_align, triple by triple: cost, the three aligned rows (every tie), and the median of align_3_powell_inter
    (src/sequence.ml:1103-1114) recomputed here from the reference's rows through the 3-D matrix."""
    import powell_util as PU
    from poyd_b200 import cost_matrix as CM

    ref = PU.reference()
    if ref is None:
        pytest.skip("oracle/_ref/libpoyref.so with the Powell recipe not built")
    cm = CM.nucleotides(1, 2, 3)
    cm3 = CM.of_two_dim(cm)
    al = S.Align3(cm, cm3)
    cases = PU.triples(seed=23, count=60, max_len=60) + PU.triples(seed=24, count=12, max_len=160, rates=(0.03, 0.08))
    rng = np.random.default_rng(3)
    a = PU.dna(rng, 90)
    cases += [(a, a[:40].copy(), PU.mutate(rng, a, 0.05)), (a, a.copy(), a.copy()), (PU.dna(rng, 1), PU.dna(rng, 1), PU.dna(rng, 1))]
    seqs = [s for t in cases for s in t]
    pool = S.SeqPool(seqs)
    triples = np.arange(3 * len(cases), dtype=np.int32).reshape(-1, 3)
    med3 = np.asarray(cm3.median).reshape(32, 32, 32)
    for mm, go, ge in PU.COSTS[:3]:
        g = al.align_3_powell(pool, triples, mm, go, ge, want=3)
        assert not g.status.any(), g.status
        for t, (x, y, z) in enumerate(cases):
            rc, rows = PU.ref_powell(ref, x, y, z, mm, go, ge)
            assert g.cost[t] == rc, (t, (mm, go, ge), g.cost[t], rc)
            for k, name in enumerate(("aligned_1", "aligned_2", "aligned_3")):
                assert np.array_equal(g.get(name, t), rows[k]), f"triple {t} costs {(mm, go, ge)}: {name}"
            med = med3[rows[0], rows[1], rows[2]]
            want_med = np.concatenate([[16], med[med != 16]]).astype(np.uint8)
            assert np.array_equal(g.get("median", t), want_med), f"triple {t}: median"
    # long operands (the sequences no longer fit the shared-memory staging: the kernel reads them from its HBM copy)
    big = PU.dna(rng, 14000)
    b2, c2 = big.copy(), big.copy()
    b2[[700, 5000, 9000]] = [1, 2, 4]
    c2 = np.delete(c2, [3000, 3001, 12000])
    lp = S.SeqPool([big, b2, c2])
    gl = al.align_3_powell(lp, np.array([[0, 1, 2]], np.int32), 1, 3, 2, want=1)
    rc, rows = PU.ref_powell(ref, big, b2, c2, 1, 3, 2)
    assert gl.status[0] == 0 and gl.cost[0] == rc
    for k, name in enumerate(("aligned_1", "aligned_2", "aligned_3")):
        assert np.array_equal(gl.get(name, 0), rows[k]), f"long triple: {name}"
    # Sequence.Align.readjust_3d (src/sequence.ml:1116-1139) and SeqCS.DOS.readjust in `ThreeD mode (src/seqCS.ml:680-727):
    # copies for the trivial rows, Powell + the 3-D median for the others, both values of first_gap
    from poyd_b200 import seqcs

    e = np.array([16], np.uint8)
    x, y, z = cases[0]
    rp = S.SeqPool([e, x, y, z, x.copy()])
    quads = np.array([[1, 4, 1, 4], [0, 2, 3, 1], [1, 0, 3, 2], [1, 2, 3, 0], [1, 2, 1, 3], [2, 3, 1, 1]], np.int32)
    what = S.readjust_3d_classify(rp, quads, 16)
    assert list(what[:4]) == [0, 1, 2, 3]
    rcost, rseq, rch = al.readjust_3d(rp, quads)
    assert list(rcost[:4]) == [0, 0, 0, 0]
    assert np.array_equal(rseq[0], x) and np.array_equal(rseq[1], y) and np.array_equal(rseq[2], x) and np.array_equal(rseq[3], e)
    assert list(rch[:4]) == [False, not np.array_equal(z, y), not np.array_equal(z, x), not np.array_equal(z, e)]
    for k in (4, 5):
        if what[k] != 4:  # (all four lengths equal by chance)
            continue
        s1, s2, mm_, pp = (rp.seq(int(i)) for i in quads[k])
        rc, rows = PU.ref_powell(ref, s1, s2, pp, 1, 3, 2)
        med = med3[rows[0], rows[1], rows[2]]
        want = np.concatenate([[16], med[med != 16]]).astype(np.uint8)
        assert rcost[k] == rc and np.array_equal(rseq[k], want), k
        assert rch[k] == (not np.array_equal(want, mm_))
    # first_gap = false is for sequences stored WITHOUT the leading gap: it is prepended for the aligner and dropped from
    # the result (src/sequence.ml:1132-1138); a second gap in front of a stored one is an element without a base
    bare = S.SeqPool([x[1:].copy(), y[1:].copy(), z[1:].copy(), x[2:].copy()])
    c0, s0, _ = al.readjust_3d(bare, np.array([[0, 1, 3, 2]], np.int32), first_gap=False)
    rc, rows = PU.ref_powell(ref, x, y, z, 1, 3, 2)
    med = med3[rows[0], rows[1], rows[2]]
    assert c0[0] == rc and np.array_equal(s0[0], med[med != 16].astype(np.uint8))
    dos = seqcs.DOS(al)
    ch, seqs, cst = seqcs.readjust_3d(dos, al, rp, [1, 0, 1, 0], [2, 2, 0, 0], [3, 3, 3, 3], [1, 1, 1, 1])
    rc, rows = PU.ref_powell(ref, x, y, z, 1, 3, 2)
    med = med3[rows[0], rows[1], rows[2]]
    if what[4] == 4:
        assert cst[0] == rc and np.array_equal(seqs[0], np.concatenate([[16], med[med != 16]]).astype(np.uint8))
    pm = al.align_affine_3(rp, np.array([[2, 3], [1, 3]], np.int32), 1)  # one empty child: the pairwise median with the parent
    assert cst[1] == pm.cost[0] and np.array_equal(seqs[1], seqcs.select_one(pm.get("median", 0), cm))
    assert cst[2] == pm.cost[1] and np.array_equal(seqs[2], seqcs.select_one(pm.get("median", 1), cm))
    assert cst[3] == 0 and np.array_equal(seqs[3], e) and ch[3]  # both children empty: ch1
    # align_3_powell_inter takes its three costs from the 2-D matrix: (1, 3, 2) here
    g2 = al.align_3_powell_inter(pool, triples[:8])
    g1 = al.align_3_powell(pool, triples[:8], 1, 3, 2, want=3)
    assert np.array_equal(g1.cost, g2.cost) and np.array_equal(g1.aligned_1, g2.aligned_1)
    # an element without a base: the reference raises, the batch marks the triple and goes on
    bad = S.SeqPool([np.array([16, 1, 2, 16, 4], np.uint8), np.array([16, 1, 2, 4], np.uint8), np.array([16, 1, 2, 4], np.uint8)] + list(cases[0]))
    gb = al.align_3_powell(bad, np.array([[0, 1, 2], [3, 4, 5]], np.int32), 1, 3, 2, want=1)
    assert gb.status[0] == 5 and gb.status[1] == 0
    assert gb.cost[1] == PU.ref_powell(ref, *cases[0], 1, 3, 2)[0]
    al.close()


def test_two_pairs_per_group_kernel(S, checker_factory):
    """aff_x2_kernel (csrc/aff_x2_kernels.cuh: two alignments in the 16-bit halves of every register) against the compiled
    reference, every output, and against the library with the kernel switched off (pair2 = 0):
    * leaf-like 500 bp pairs, pair counts that leave a half-filled group, a lone batch and an odd number of batches;
    * lengths from 80 to 500 in one launch (the two pairs of a lane group end at different steps);
    * batches mixed with pairs carrying gap bits (declined two batches at a time and taken by the next kernels);
    * a cost matrix too large for 16 bits (everything declined), one near the limit (long pairs declined, short ones taken)."""
    from poyd_b200 import cost_matrix as CM, synth

    def both(cm, pool, pairs, label, ref=True):
        on = S.Align(cm, config={"pair2": 1})
        off = S.Align(cm, config={"pair2": 0})
        g1 = on.align_affine_3(pool, pairs, ALL)
        g0 = off.align_affine_3(pool, pairs, ALL)
        assert np.array_equal(g1.cost, g0.cost), f"{label}: cost differs from pair2 = 0 at {np.nonzero(g1.cost != g0.cost)[0][:8]}"
        for name in ("median", "medianwg", "aligned_a", "aligned_b", "lens"):
            assert np.array_equal(getattr(g1, name), getattr(g0, name)), f"{label}: {name} differs from pair2 = 0"
        if ref:
            o = checker_factory(cm).batch(3, pool.pool, pool.off, pool.len, pairs, nthreads=8)
            assert_aligned_equal(g1, o, label=label)
        on.close()
        off.close()

    cm = CM.nucleotides(1, 2, 3)
    for n, seed in ((1, 5), (3, 6), (4, 7), (13, 8), (37, 9), (1000, 10)):
        pool, pairs = synth.pair_batch(n, 500, seed=seed, min_len=450)
        both(cm, pool, pairs, f"leaf-like, {n} pairs")
    # other costs (gap opening 1 and 5, substitutions dearer than indels)
    for name, cm2 in _affine_cases()[1:3]:
        pool, pairs = synth.pair_batch(600, 500, seed=21, min_len=450)
        both(cm2, pool, pairs, f"leaf-like, {name}")
    # ragged lengths in one launch: same stripe shape (5, 8) needs |lr - lc| small, so shorten both operands together
    rng = np.random.default_rng(4)
    seqs = []
    for k in range(800):
        L = int(rng.integers(80, 500))
        a = np.concatenate([[16], rng.choice(np.array([1, 2, 4, 8], np.uint8), L)]).astype(np.uint8)
        b = a.copy()
        sub = rng.random(L + 1) < 0.1
        sub[0] = False
        b[sub] = rng.choice(np.array([1, 2, 4, 8], np.uint8), int(sub.sum()))
        cut = rng.integers(1, L, size=int(rng.integers(0, 6)))
        b = np.delete(b, cut)
        seqs += [a, b]
    pool = S.SeqPool(seqs)
    pairs = np.arange(1600, dtype=np.int32).reshape(-1, 2)
    both(cm, pool, pairs, "ragged lengths")
    # every eleventh pair carries gap bits: its double batch goes to aff_fast_kernel, its batch on to the ring kernel
    pool, pairs = synth.pair_batch(900, 500, seed=31, min_len=450)
    for p in range(5, 900, 11):
        r = pool.seq(int(pairs[p, 1]))
        r[7::13] |= 16
    both(cm, pool, pairs, "mixed with gap bits")
    # costs: 16-bit range exceeded (4 * 40 * 1000 > 20000) -> all declined; near the limit -> decided pair by pair
    pool, pairs = synth.pair_batch(300, 500, seed=41, min_len=450)
    both(CM.nucleotides(20, 40, 30), pool, pairs, "costs too large for 16 bits")
    seqs = []
    for k in range(300):
        L = 300 if (k // 24) % 2 == 0 else 500
        a = np.concatenate([[16], rng.choice(np.array([1, 2, 4, 8], np.uint8), L)]).astype(np.uint8)
        b = a.copy()
        b[3::9] = 1
        seqs += [a, np.delete(b, [L // 2])]
    pool = S.SeqPool(seqs)
    both(CM.nucleotides(3, 6, 2), pool, np.arange(600, dtype=np.int32).reshape(-1, 2), "costs near the 16-bit limit")
