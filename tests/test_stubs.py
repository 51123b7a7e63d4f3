"""The drop-in boundary, executed: every algn_CAML_* symbol of stubs/poyb200_stubs.c (oracle/_ref/libpoystubs.so, on the
GPU) next to the reference's OWN algn_CAML_* symbol of the same name (oracle/_ref/libpoyref.so = unmodified src/algn.c),
called the way the OCaml runtime calls them: tagged ints and custom blocks (`struct seq`, `struct cm`, `struct cm_3d`,
`struct matrices`) built by oracle/caml_runtime.c.  Same blocks in, same blocks out, byte for byte.

Reference externals: src/sequence.ml:453-762, 919; C side src/algn.c:2551-2680, 3382-3475, 3908-4020, 4198-4300."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpoyref.so")
STUBS_SO = os.path.join(ROOT, "oracle", "_ref", "libpoystubs.so")

DROP_IN = ["algn_CAML_simple_2", "algn_CAML_backtrack_2d", "algn_CAML_backtrack_2d_bc", "algn_CAML_align_2d",
           "algn_CAML_align_2d_bc", "algn_CAML_cost_affine_3", "algn_CAML_align_affine_3", "algn_CAML_align_affine_3_bc",
           "algn_CAML_median_2_no_gaps", "algn_CAML_median_2_with_gaps", "algn_CAML_ancestor_2", "algn_CAML_worst_2",
           "algn_CAML_verify_2", "algn_CAML_simple_3", "algn_CAML_simple_3_bc", "algn_CAML_backtrack_3d",
           "algn_CAML_backtrack_3d_bc", "algn_CAML_align_3d", "algn_CAML_align_3d_bc", "algn_CAML_median_3", "powell_3D_align",
           "powell_3D_align_bc"]
BATCHED = ["poyb200_CAML_batch_align_affine_3", "poyb200_CAML_batch_align_affine_3_bc", "poyb200_CAML_batch_cost_2",
           "poyb200_CAML_batch_median", "poyb200_CAML_batch_closest"]

V = C.c_ssize_t  # OCaml `value`


def val_int(x: int) -> int:
    return (int(x) << 1) + 1


def int_val(v: int) -> int:
    return int(v) >> 1


class Side:
    """One of the two libraries, with its own stand-in runtime (failwith is armed per call through camlrt_call)."""

    def __init__(self, path):
        L = self.L = C.CDLL(path)
        for f in ("camlrt_seq", "camlrt_cm", "camlrt_cm3", "camlrt_matrices", "camlrt_array", "camlrt_array_get", "camlrt_call"):
            getattr(L, f).restype = V
        L.camlrt_last_failure.restype = C.c_char_p
        L.camlrt_seq.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.camlrt_cm.argtypes = [C.c_void_p]
        L.camlrt_cm3.argtypes = [C.c_void_p]
        L.camlrt_call.argtypes = [C.c_void_p, C.c_int, C.POINTER(V), C.POINTER(C.c_int)]
        for f in ("camlrt_seq_scramble", "camlrt_seq_clear"):
            getattr(L, f).argtypes = [V]
            getattr(L, f).restype = None
        L.camlrt_seq_len.argtypes = [V]
        L.camlrt_seq_read.argtypes = [V, C.c_void_p]
        L.camlrt_array.argtypes = [C.c_int]
        L.camlrt_array_set.argtypes = [V, C.c_int, V]
        L.camlrt_array_set.restype = None
        L.camlrt_array_get.argtypes = [V, C.c_int]
        L.camlrt_array_len.argtypes = [V]
        L.camlrt_bytes_read.argtypes = [V, C.c_void_p, C.c_int]
        L.camlrt_bytes_read.restype = None
        L.camlrt_string_length.argtypes = [V]
        L.camlrt_string_length.restype = C.c_size_t

    def call(self, name, *args, expect_failure=False):
        fn = C.cast(getattr(self.L, name), C.c_void_p)
        arr = (V * len(args))(*args)
        failed = C.c_int(0)
        res = self.L.camlrt_call(fn, len(args), arr, C.byref(failed))
        if expect_failure:
            assert failed.value == 1, f"{name}: expected an OCaml Failure"
            return self.L.camlrt_last_failure().decode()
        assert failed.value == 0, f"{name} raised Failure: {self.L.camlrt_last_failure().decode()}"
        return res

    def seq(self, data, cap=None) -> int:
        data = np.ascontiguousarray(data, np.uint8)
        return self.L.camlrt_seq(data.ctypes.data, len(data), len(data) if cap is None else cap)

    def empty(self, cap) -> int:
        return self.L.camlrt_seq(None, 0, cap)

    def read(self, v) -> np.ndarray:
        out = np.zeros(self.L.camlrt_seq_len(v) + 1, np.uint8)
        n = self.L.camlrt_seq_read(v, out.ctypes.data)
        return out[:n].copy()


@pytest.fixture(scope="module")
def sides():
    from oracle import oracle

    oracle.build(ref=True)
    if not (os.path.exists(REF_SO) and os.path.exists(STUBS_SO)):
        pytest.skip("oracle/_ref/libpoyref.so / libpoystubs.so not built (needs /root/reference once)")
    return Side(REF_SO), Side(STUBS_SO)


def test_stub_library_exports_every_drop_in_symbol():
    """CPU check: the stub library loads and exports the reference's own symbol names (link-time replacement) plus the
    batched externals."""
    from oracle import oracle

    oracle.build(ref=True)
    if not os.path.exists(STUBS_SO):
        pytest.skip("libpoystubs.so not built")
    L = C.CDLL(STUBS_SO)
    ref = C.CDLL(REF_SO)
    for name in DROP_IN:
        assert hasattr(L, name), name
        assert hasattr(ref, name), f"the reference does not define {name}?"
    for name in BATCHED:
        assert hasattr(L, name), name


def test_drop_in_link_has_one_definition_of_every_external():
    """stubs/algn_b200.c + stubs/poyb200_stubs.c link into one library (what libpoycside would contain): the externals'
    original names resolve (to the GPU stubs), the reference's renamed CPU versions and its untouched externals are there."""
    import subprocess

    from oracle import oracle

    oracle.build(ref=True)
    so = os.path.join(ROOT, "oracle", "_ref", "libpoydropin.so")
    if not os.path.exists(so):
        pytest.skip("libpoydropin.so not built")
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout.split("\n")
    names = [ln.split()[-1] for ln in syms if ln.strip()]
    for n in DROP_IN:
        assert names.count(n) == 1, n
        if n.startswith("algn_"):  # the externals of algn.c are renamed by algn_b200.c; powell_3D_align lives in ukkCommon.c,
            assert names.count(n + "_cpu") == 1, n + "_cpu"  # which simply leaves the link
    for n in ("algn_CAML_union", "algn_CAML_myers", "algn_CAML_limit_2", "algn_CAML_create_backtrack", "cm_CAML_create", "seq_CAML_create"):
        assert names.count(n) == 1, n


def _ref_cm(ref: Side, cm):
    """(handle keep-alive, value block) of a cost matrix made by the reference's own cm_set_val (ref_driver.c)."""
    from oracle import oracle

    r = oracle.Reference(cm)
    r.L.ref_cm_struct.restype = C.c_void_p
    return r, ref.L.camlrt_cm(r.L.ref_cm_struct(r.h))


def _ref_cm3(ref: Side, cm3):
    from oracle import oracle

    r = oracle.Reference3(cm3)
    r.L.ref_cm3_struct.restype = C.c_void_p
    return r, ref.L.camlrt_cm3(r.L.ref_cm3_struct(r.h))


def _dna(rng, n, gapamb=0.0, amb=0.0):
    s = rng.choice(np.array([1, 2, 4, 8], np.uint8), size=n)
    if amb:
        m = rng.random(n) < amb
        s[m] = rng.integers(1, 16, size=int(m.sum()), dtype=np.uint8)
    if gapamb:
        s[rng.random(n) < gapamb] |= 16
    return np.concatenate([[16], s]).astype(np.uint8)


def _mutate(rng, a, p=0.1):
    b = list(a[1:])
    out = []
    for x in b:
        r = rng.random()
        if r < p / 4:
            continue
        if r < p / 2:
            out.append(int(rng.choice([1, 2, 4, 8])))
        out.append(int(rng.choice([1, 2, 4, 8])) if rng.random() < p else int(x))
    return np.array([16] + out, np.uint8)


@pytest.mark.gpu
def test_linear_externals_side_by_side(sides):
    from poyd_b200 import cost_matrix as CM

    ref, gpu = sides
    rng = np.random.default_rng(5)
    for cm in (CM.default_nucleotides(), CM.nucleotides(3, 1)):
        keep, vcm = _ref_cm(ref, cm)
        mat_r, mat_g = ref.L.camlrt_matrices(), gpu.L.camlrt_matrices()
        for n1, n2, dw in [(60, 60, 4), (200, 150, 12), (300, 299, 30), (40, 7, 2), (1, 1, 2), (0, 0, 2), (500, 470, 26), (130, 20, 0)]:
            a = _dna(rng, n1, amb=0.05)
            b = _mutate(rng, a)[: n2 + 1] if n2 <= n1 and n2 > 10 else _dna(rng, n2, amb=0.05)
            if len(b) > len(a):
                a, b = b, a
            s1, s2 = ref.seq(a), ref.seq(b)
            cr = int_val(ref.call("algn_CAML_simple_2", s1, s2, vcm, mat_r, val_int(dw)))
            gpu.L.camlrt_seq_scramble(s1)
            gpu.L.camlrt_seq_scramble(s2)
            cg = int_val(gpu.call("algn_CAML_simple_2", s1, s2, vcm, mat_g, val_int(dw)))
            assert cr == cg, (n1, n2, dw, cr, cg)
            cap = len(a) + len(b)
            for sw in (0, 1):
                o = [ref.empty(cap) for _ in range(4)]
                ref.call("algn_CAML_backtrack_2d", s1, s2, o[0], o[1], mat_r, vcm, val_int(sw))
                gpu.call("algn_CAML_backtrack_2d", s1, s2, o[2], o[3], mat_g, vcm, val_int(sw))
                assert np.array_equal(ref.read(o[0]), ref.read(o[2])) and np.array_equal(ref.read(o[1]), ref.read(o[3])), (n1, n2, sw)
                # the combined external, native and bytecode entry
                o2 = [ref.empty(cap) for _ in range(4)]
                c1 = int_val(ref.call("algn_CAML_align_2d", s1, s2, vcm, mat_r, o2[0], o2[1], val_int(dw), val_int(sw)))
                c2 = int_val(gpu.call("algn_CAML_align_2d", s1, s2, vcm, mat_g, o2[2], o2[3], val_int(dw), val_int(sw)))
                assert c1 == c2 == cr
                assert np.array_equal(ref.read(o2[0]), ref.read(o2[2])) and np.array_equal(ref.read(o2[1]), ref.read(o2[3]))
                # functions of the aligned pair
                ra, rb = o2[0], o2[1]
                n = ref.L.camlrt_seq_len(ra)
                for name in ("algn_CAML_ancestor_2", "algn_CAML_median_2_with_gaps", "algn_CAML_median_2_no_gaps"):
                    m1, m2 = ref.empty(n + 1), ref.empty(n + 1)
                    ref.call(name, ra, rb, vcm, m1)
                    gpu.call(name, ra, rb, vcm, m2)
                    assert np.array_equal(ref.read(m1), ref.read(m2)), (name, n1, n2)
                for name in ("algn_CAML_worst_2", "algn_CAML_verify_2"):
                    assert ref.call(name, ra, rb, vcm) == gpu.call(name, ra, rb, vcm), (name, n1, n2)
        # a backtrack that does not follow a simple_2 on the same operands is refused, not answered from stale state
        x, y = ref.seq(_dna(rng, 30)), ref.seq(_dna(rng, 30))
        msg = gpu.call("algn_CAML_backtrack_2d", x, y, ref.empty(64), ref.empty(64), mat_g, vcm, val_int(0), expect_failure=True)
        assert "simple_2" in msg


@pytest.mark.gpu
def test_affine_externals_side_by_side(sides):
    from poyd_b200 import cost_matrix as CM

    ref, gpu = sides
    rng = np.random.default_rng(6)
    for cm in (CM.nucleotides(1, 2, 3), CM.nucleotides(2, 1, 1)):
        keep, vcm = _ref_cm(ref, cm)
        mat_r, mat_g = ref.L.camlrt_matrices(), gpu.L.camlrt_matrices()
        for n1, n2, ga in [(80, 80, 0.0), (200, 230, 0.1), (500, 480, 0.0), (500, 500, 0.1), (7, 60, 0.05), (0, 5, 0.0), (0, 0, 0.0)]:
            a = _dna(rng, n1, gapamb=ga, amb=0.02)
            b = _dna(rng, n2, gapamb=ga, amb=0.02) if abs(n1 - n2) > 40 or n1 < 10 else np.concatenate([_mutate(rng, a)[: n2 + 1]])
            s1, s2 = ref.seq(a), ref.seq(b)
            assert ref.call("algn_CAML_cost_affine_3", s1, s2, vcm, mat_r) == gpu.call("algn_CAML_cost_affine_3", s1, s2, vcm, mat_g)
            cap = len(a) + len(b) + 2
            o = [ref.empty(cap) for _ in range(8)]
            gpu.L.camlrt_seq_scramble(s1)
            c1 = ref.call("algn_CAML_align_affine_3", s1, s2, vcm, mat_r, o[0], o[1], o[2], o[3])
            gpu.L.camlrt_seq_scramble(s2)
            c2 = gpu.call("algn_CAML_align_affine_3", s1, s2, vcm, mat_g, o[4], o[5], o[6], o[7])
            assert c1 == c2, (n1, n2, int_val(c1), int_val(c2))
            for k in range(4):
                assert np.array_equal(ref.read(o[k]), ref.read(o[4 + k])), (n1, n2, k)


@pytest.mark.gpu
def test_cost_matrix_is_tracked_by_content(sides):
    """cm_CAML_set_cost mutates a matrix in place; a new matrix can land on an old address.  The stubs must follow."""
    from poyd_b200 import cost_matrix as CM

    ref, gpu = sides
    rng = np.random.default_rng(7)
    cm = CM.default_nucleotides()
    keep, vcm = _ref_cm(ref, cm)
    mat_r, mat_g = ref.L.camlrt_matrices(), gpu.L.camlrt_matrices()
    a = _dna(rng, 120)
    b = _mutate(rng, a, 0.3)
    if len(b) > len(a):
        a, b = b, a
    s1, s2 = ref.seq(a), ref.seq(b)
    c0 = gpu.call("algn_CAML_simple_2", s1, s2, vcm, mat_g, val_int(10))
    assert c0 == ref.call("algn_CAML_simple_2", s1, s2, vcm, mat_r, val_int(10))
    # in-place mutation of the SAME tables (same addresses): triple every substitution cost
    keep.L.ref_cm_struct.restype = C.c_void_p
    base = keep.L.ref_cm_struct(keep.h)  # struct cm: eight ints, then `int *cost` (src/cm.h:32-41)
    cost_ptr = C.cast(C.c_void_p.from_address(base + 32).value, C.POINTER(C.c_int32))
    dim = 1 << cm.lcm
    for x in range(dim * dim):
        cost_ptr[x] *= 3
    c1r = ref.call("algn_CAML_simple_2", s1, s2, vcm, mat_r, val_int(10))
    c1g = gpu.call("algn_CAML_simple_2", s1, s2, vcm, mat_g, val_int(10))
    assert c1r == c1g and c1g != c0


@pytest.mark.gpu
def test_three_sequence_externals_side_by_side(sides):
    from poyd_b200 import cost_matrix as CM

    ref, gpu = sides
    rng = np.random.default_rng(8)
    cm = CM.default_nucleotides()
    cm3 = CM.of_two_dim(cm)
    keep, vcm3 = _ref_cm3(ref, cm3)
    mat_r, mat_g = ref.L.camlrt_matrices(), gpu.L.camlrt_matrices()
    walked = 0
    for n1, n2, n3 in [(5, 5, 5), (12, 30, 7), (40, 40, 40), (33, 20, 70), (60, 61, 59), (1, 50, 1), (90, 80, 100), (0, 0, 0)]:
        a, b, c = _dna(rng, n1), _dna(rng, n2), _dna(rng, n3)
        if n1 == n2 == n3 and n1 > 1:
            b, c = a.copy(), a.copy()
            b[1:][rng.random(n1) < 0.1] = 2
            c[1:][rng.random(n1) < 0.1] = 8
        s = [ref.seq(x) for x in (a, b, c)]
        c1 = ref.call("algn_CAML_simple_3", s[0], s[1], s[2], vcm3, mat_r, val_int(0))
        c2 = gpu.call("algn_CAML_simple_3", s[0], s[1], s[2], vcm3, mat_g, val_int(0))
        assert c1 == c2, (n1, n2, n3)
        _, status, r1, r2, r3, med = keep.align_3(a, b, c)
        cap = len(a) + len(b) + len(c)
        if status != 0:
            # the reference's unchecked walk would leave the sequences: the drop-in raises instead of corrupting memory
            gpu.call("algn_CAML_backtrack_3d", s[0], s[1], s[2], ref.empty(cap), ref.empty(cap), ref.empty(cap), mat_g, vcm3,
                     expect_failure=True)
            continue
        walked += 1
        o = [ref.empty(cap) for _ in range(6)]
        ref.call("algn_CAML_backtrack_3d", s[0], s[1], s[2], o[0], o[1], o[2], mat_r, vcm3)
        gpu.call("algn_CAML_backtrack_3d", s[0], s[1], s[2], o[3], o[4], o[5], mat_g, vcm3)
        for k in range(3):
            assert np.array_equal(ref.read(o[k]), ref.read(o[3 + k])), (n1, n2, n3, k)
        o2 = [ref.empty(cap) for _ in range(3)]
        c3 = gpu.call("algn_CAML_align_3d", s[0], s[1], s[2], vcm3, mat_g, o2[0], o2[1], o2[2], val_int(0))
        assert c3 == c1
        for k in range(3):
            assert np.array_equal(ref.read(o2[k]), ref.read(o[k]))
        m1, m2 = ref.empty(cap + 1), ref.empty(cap + 1)
        ref.call("algn_CAML_median_3", o[0], o[1], o[2], vcm3, m1)
        gpu.call("algn_CAML_median_3", o[0], o[1], o[2], vcm3, m2)
        assert np.array_equal(ref.read(m1), ref.read(m2)), (n1, n2, n3)
    assert walked >= 3


@pytest.mark.gpu
def test_powell_external_side_by_side(sides):
    """powell_3D_align (src/ukkCommon.c:110-145): the drop-in symbol next to the reference's own, same blocks."""
    ref, gpu = sides
    if not hasattr(ref.L, "powell_3D_align"):
        pytest.skip("libpoyref.so predates the Powell recipe")
    rng = np.random.default_rng(41)
    for n, p, costs in [(12, 0.2, (1, 3, 2)), (40, 0.1, (1, 0, 1)), (70, 0.08, (2, 1, 1)), (25, 0.0, (1, 3, 2)), (55, 0.3, (1, 2, 1))]:
        a = _dna(rng, n)
        b, c = _mutate(rng, a, p), _mutate(rng, a, p)
        cap = len(a) + len(b) + len(c)
        s = [ref.seq(x) for x in (a, b, c)]
        o = [ref.empty(cap) for _ in range(6)]
        args = [val_int(v) for v in costs]
        c1 = ref.call("powell_3D_align", s[0], s[1], s[2], o[0], o[1], o[2], *args)
        c2 = gpu.call("powell_3D_align", s[0], s[1], s[2], o[3], o[4], o[5], *args)
        assert c1 == c2, (n, p, costs)
        for k in range(3):
            assert np.array_equal(ref.read(o[k]), ref.read(o[3 + k])), (n, p, costs, k)
    # an element without a base: both raise Failure "This is impossible!"
    bad = ref.seq(np.array([16, 1, 16, 2], np.uint8))
    good = ref.seq(np.array([16, 1, 2], np.uint8))
    for side in (ref, gpu):
        side.call("powell_3D_align", bad, good, good, ref.empty(12), ref.empty(12), ref.empty(12), val_int(1), val_int(3), val_int(2),
                  expect_failure=True)


@pytest.mark.gpu
def test_batched_externals_against_the_single_calls(sides):
    """poyb200_CAML_batch_* (the externals of the batching layer) return, pair by pair, what the reference's single-call
    externals return -- through OCaml arrays, preallocated result sequences and the (costs, lens, bitsets) tuple."""
    from poyd_b200 import cost_matrix as CM

    ref, gpu = sides
    rng = np.random.default_rng(9)
    for cm in (CM.nucleotides(1, 2, 3), CM.default_nucleotides()):
        affine = cm.cost_model_type == 1
        keep, vcm = _ref_cm(ref, cm)
        mat_r = ref.L.camlrt_matrices()
        seqs = [_dna(rng, int(rng.integers(20, 260)), gapamb=0.05 if affine else 0.0, amb=0.03) for _ in range(24)]
        seqs += [_mutate(rng, s) for s in seqs[:12]]
        ns = len(seqs)
        pairs = [(i, 24 + i) for i in range(12)] + [(int(rng.integers(ns)), int(rng.integers(ns))) for _ in range(20)]
        n = len(pairs)
        vseqs = gpu.L.camlrt_array(ns)
        blocks = [ref.seq(s) for s in seqs]
        for i, b in enumerate(blocks):
            gpu.L.camlrt_array_set(vseqs, i, b)
        vpairs = gpu.L.camlrt_array(2 * n)
        vdw = gpu.L.camlrt_array(n)
        dws = []
        for p, (i, j) in enumerate(pairs):
            gpu.L.camlrt_array_set(vpairs, 2 * p, val_int(i))
            gpu.L.camlrt_array_set(vpairs, 2 * p + 1, val_int(j))
            dws.append(int(rng.integers(2, 40)))
            gpu.L.camlrt_array_set(vdw, p, val_int(dws[-1]))
        # ---- batch_cost_2
        vc = gpu.call("poyb200_CAML_batch_cost_2", vseqs, vpairs, vdw, vcm)
        assert gpu.L.camlrt_array_len(vc) == n
        single = []
        for p, (i, j) in enumerate(pairs):
            if affine:
                single.append(int_val(ref.call("algn_CAML_cost_affine_3", blocks[i], blocks[j], vcm, mat_r)))
            else:
                x, y = (i, j) if len(seqs[i]) >= len(seqs[j]) else (j, i)
                single.append(int_val(ref.call("algn_CAML_simple_2", blocks[x], blocks[y], vcm, mat_r, val_int(dws[p]))))
            assert int_val(gpu.L.camlrt_array_get(vc, p)) == single[-1], p
        # ---- batch_median: cost, median, three bitsets
        vmed = gpu.L.camlrt_array(n)
        for p, (i, j) in enumerate(pairs):
            gpu.L.camlrt_array_set(vmed, p, ref.empty(len(seqs[i]) + len(seqs[j]) + 2))
        res = gpu.call("poyb200_CAML_batch_median", vseqs, vpairs, vdw, vcm, vmed)
        vcosts, vlens, vba, vbb, vbm = (gpu.L.camlrt_array_get(res, k) for k in range(5))
        for p, (i, j) in enumerate(pairs):
            cap = len(seqs[i]) + len(seqs[j]) + 2
            if affine:
                o = [ref.empty(cap) for _ in range(4)]
                c = int_val(ref.call("algn_CAML_align_affine_3", blocks[i], blocks[j], vcm, mat_r, o[0], o[1], o[2], o[3]))
                ra, rb, med, wg = ref.read(o[0]), ref.read(o[1]), ref.read(o[2]), ref.read(o[3])
            else:
                swapped = len(seqs[i]) < len(seqs[j])
                x, y = (j, i) if swapped else (i, j)
                o = [ref.empty(cap) for _ in range(4)]
                c = int_val(ref.call("algn_CAML_align_2d", blocks[x], blocks[y], vcm, mat_r, o[0], o[1], val_int(dws[p]),
                                     val_int(0 if swapped else 1)))  # create_edited_2: swaped = sz1 >= sz2 (sequence.ml:818)
                ra_, rb_ = (o[1], o[0]) if swapped else (o[0], o[1])
                ref.call("algn_CAML_ancestor_2", ra_, rb_, vcm, o[2])
                ref.call("algn_CAML_median_2_with_gaps", ra_, rb_, vcm, o[3])
                ra, rb, med, wg = ref.read(ra_), ref.read(rb_), ref.read(o[2]), ref.read(o[3])
            assert int_val(gpu.L.camlrt_array_get(vcosts, p)) == c, p
            assert int_val(gpu.L.camlrt_array_get(vlens, p)) == len(ra), p
            assert np.array_equal(ref.read(gpu.L.camlrt_array_get(vmed, p)), med), p
            nb = (len(ra) + 7) // 8
            for arr, want in ((vba, ra), (vbb, rb), (vbm, wg)):
                s = gpu.L.camlrt_array_get(arr, p)
                assert gpu.L.camlrt_string_length(s) == nb
                raw = np.zeros(nb + 1, np.uint8)
                gpu.L.camlrt_bytes_read(s, raw.ctypes.data, nb)
                bits = np.unpackbits(raw[:nb], bitorder="little")[: len(want)]  # extlib BitSet: bit i of byte i / 8
                assert np.array_equal(bits, (want != cm.gap).astype(np.uint8)), p
        # ---- batch_align_affine_3 (affine only)
        if affine:
            arrs = [gpu.L.camlrt_array(n) for _ in range(4)]
            for p, (i, j) in enumerate(pairs):
                for a in arrs:
                    gpu.L.camlrt_array_set(a, p, ref.empty(len(seqs[i]) + len(seqs[j]) + 2))
            vc2 = gpu.call("poyb200_CAML_batch_align_affine_3", vseqs, vpairs, vcm, arrs[0], arrs[1], arrs[2], arrs[3])
            for p, (i, j) in enumerate(pairs):
                cap = len(seqs[i]) + len(seqs[j]) + 2
                o = [ref.empty(cap) for _ in range(4)]
                c = ref.call("algn_CAML_align_affine_3", blocks[i], blocks[j], vcm, mat_r, o[0], o[1], o[2], o[3])
                assert gpu.L.camlrt_array_get(vc2, p) == c
                for k in range(4):  # resi, resj, median, medianwg
                    assert np.array_equal(ref.read(gpu.L.camlrt_array_get(arrs[k], p)), ref.read(o[k])), (p, k)
