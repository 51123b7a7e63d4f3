"""TEST INFRASTRUCTURE ONLY -- ctypes front-ends for the two CPU checkers.

* ``Port``      : oracle/_build/libpoyoracle.so, the repo's plain-C restatement (poy_oracle.c).
* ``Reference`` : oracle/_ref/libpoyref.so, the unmodified reference ``src/algn.c`` compiled in the
                  container (oracle/Makefile `ref`) behind ref_driver.c.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The product package (``poyd_b200``) never does.
Both classes expose the same methods so tests can run one against the other.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_build", "libpoyoracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libpoyref.so")

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_longlong)
_u16p = C.POINTER(C.c_uint16)


def build(ref: bool = True) -> None:
    """Compile the checkers (building the checker is not using it)."""
    import sys

    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"], stdout=sys.stderr)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"], stdout=sys.stderr)


def _p(a: Optional[np.ndarray], t):
    return None if a is None else a.ctypes.data_as(t)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint8))


class _PoCm(C.Structure):
    _fields_ = [("a_sz", C.c_int32), ("lcm", C.c_int32), ("gap", C.c_int32), ("cost_model_type", C.c_int32),
                ("combinations", C.c_int32), ("gap_open", C.c_int32), ("cost", _i32p), ("median", _u8p),
                ("prepend", _i32p), ("tail", _i32p)]


class _PoCm3(C.Structure):
    _fields_ = [("lcm", C.c_int32), ("gap", C.c_int32), ("cost3", _i32p), ("median3", _u8p)]


class _PoBand(C.Structure):
    _fields_ = [("full", C.c_int), ("dlo", C.c_int), ("dhi", C.c_int)]


class _Base:
    """Shared batch plumbing."""

    def batch(self, mode: int, pool, off, length, pairs, deltaw=None, nthreads: int = 1, want_seqs: bool = True):
        pool = _u8(pool)
        off = np.ascontiguousarray(off, dtype=np.int64)
        length = np.ascontiguousarray(length, dtype=np.int32)
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pairs.shape[0]
        if deltaw is None:
            deltaw = np.zeros(n, np.int32)
        deltaw = np.ascontiguousarray(deltaw, dtype=np.int32)
        cost = np.zeros(n, np.int32)
        out = {"cost": cost}
        med = medwg = ra = rb = lens = None
        stride = 0
        if mode in (1, 3) and want_seqs and n:
            stride = int((length[pairs[:, 0]].astype(np.int64) + length[pairs[:, 1]]).max()) + 2
            med, medwg, ra, rb = (np.zeros((n, stride), np.uint8) for _ in range(4))
            lens = np.zeros((n, 4), np.int32)
            out.update(median=med, medianwg=medwg, ra=ra, rb=rb, lens=lens)
        self._batch(mode, pool, off, length, pairs, deltaw, n, nthreads, cost, med, medwg, ra, rb, lens, stride)
        return out


class Port(_Base):
    kind = "port"

    def __init__(self, cm):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        L = self.L = C.CDLL(PORT_SO)
        self.cm = cm
        self._cost = np.ascontiguousarray(cm.cost, np.int32)
        self._median = np.ascontiguousarray(cm.median, np.uint8)
        self._prep = np.ascontiguousarray(cm.prepend_cost, np.int32)
        self._tail = np.ascontiguousarray(cm.tail_cost, np.int32)
        self.c = _PoCm(cm.a_sz, cm.lcm, cm.gap, cm.cost_model_type, cm.combinations, cm.gap_open,
                       _p(self._cost, _i32p), _p(self._median, _u8p), _p(self._prep, _i32p), _p(self._tail, _i32p))
        L.po_linear_band.restype = _PoBand
        L.po_cells_linear.restype = C.c_longlong
        L.po_cells_affine.restype = C.c_longlong

    def linear_band(self, l1, l2, deltaw):
        b = self.L.po_linear_band(l1, l2, deltaw)
        return b.full, b.dlo, b.dhi

    def cells_linear(self, l1, l2, deltaw):
        return int(self.L.po_cells_linear(l1, l2, deltaw))

    def cells_affine(self, li, lj):
        return int(self.L.po_cells_affine(li, lj))

    def cost_2(self, s1, s2, deltaw, want_dir=False):
        s1, s2 = _u8(s1), _u8(s2)
        assert len(s1) >= len(s2)
        d = np.zeros((len(s1), len(s2)), np.uint16) if want_dir else None
        r = self.L.po_cost_2(C.byref(self.c), _p(s1, _u8p), len(s1), _p(s2, _u8p), len(s2), int(deltaw), _p(d, _u16p))
        return (r, d) if want_dir else r

    def align_2(self, sa, sb, deltaw):
        sa, sb = _u8(sa), _u8(sb)
        cap = len(sa) + len(sb)
        ra, rb = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
        rl = C.c_int(0)
        r = self.L.po_align_2(C.byref(self.c), _p(sa, _u8p), len(sa), _p(sb, _u8p), len(sb), int(deltaw),
                              _p(ra, _u8p), _p(rb, _u8p), C.byref(rl))
        return r, ra[: rl.value].copy(), rb[: rl.value].copy()

    def median_2(self, which, a, b):
        a, b = _u8(a), _u8(b)
        out = np.zeros(len(a) + 2, np.uint8)
        n = self.L.po_median_2(C.byref(self.c), which, _p(a, _u8p), _p(b, _u8p), len(a), _p(out, _u8p))
        return out[:n].copy()

    def align_affine_3(self, sa, sb, want_dir=False):
        sa, sb = _u8(sa), _u8(sb)
        cap = len(sa) + len(sb) + 2
        o = [np.zeros(cap, np.uint8) for _ in range(4)]
        lens = np.zeros(4, np.int32)
        d = np.zeros(len(sa) * len(sb), np.uint16) if want_dir else None
        r = self.L.po_align_affine_3(C.byref(self.c), _p(sa, _u8p), len(sa), _p(sb, _u8p), len(sb),
                                     *[_p(x, _u8p) for x in o], _p(lens, _i32p), _p(d, _u16p))
        res = (r, o[0][: lens[0]].copy(), o[1][: lens[1]].copy(), o[2][: lens[2]].copy(), o[3][: lens[3]].copy())
        return res + (d,) if want_dir else res

    def cost_affine_3(self, sa, sb):
        sa, sb = _u8(sa), _u8(sb)
        return self.L.po_cost_affine_3(C.byref(self.c), _p(sa, _u8p), len(sa), _p(sb, _u8p), len(sb))

    def _batch(self, mode, pool, off, length, pairs, deltaw, n, nthreads, cost, med, medwg, ra, rb, lens, stride):
        self.L.po_batch(C.byref(self.c), mode, _p(pool, _u8p), _p(off, _i64p), _p(length, _i32p), _p(pairs, _i32p),
                        _p(deltaw, _i32p), n, nthreads, _p(cost, _i32p), _p(med, _u8p), _p(medwg, _u8p),
                        _p(ra, _u8p), _p(rb, _u8p), _p(lens, _i32p), C.c_longlong(stride))


class Port3:
    """3-D cube in the reference's executed (defective) form, plain-C port."""

    kind = "port"

    def __init__(self, cm3):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        self.L = C.CDLL(PORT_SO)
        self._cost = np.ascontiguousarray(cm3.cost, np.int32)
        self._med = np.ascontiguousarray(cm3.median, np.uint8)
        self.c = _PoCm3(cm3.lcm, cm3.gap, _p(self._cost, _i32p), _p(self._med, _u8p))

    def align_3(self, s1, s2, s3, want_dir=False):
        """Returns (cost, status, r1, r2, r3, median[, dir])."""
        s1, s2, s3 = _u8(s1), _u8(s2), _u8(s3)
        l1, l2, l3 = len(s1), len(s2), len(s3)
        d = np.zeros(l1 * l2 * l3, np.uint8)
        cost = self.L.po_cost_3(C.byref(self.c), _p(s1, _u8p), l1, _p(s2, _u8p), l2, _p(s3, _u8p), l3, _p(d, _u8p))
        cap = l1 + l2 + l3
        r = [np.zeros(cap + 1, np.uint8) for _ in range(4)]
        ml, st = C.c_int(0), C.c_int(0)
        n = self.L.po_backtrack_3(C.byref(self.c), _p(d, _u8p), _p(s1, _u8p), l1, _p(s2, _u8p), l2, _p(s3, _u8p), l3,
                                  _p(r[0], _u8p), _p(r[1], _u8p), _p(r[2], _u8p), _p(r[3], _u8p), C.byref(ml), C.byref(st))
        res = (cost, st.value, r[0][:n].copy(), r[1][:n].copy(), r[2][:n].copy(), r[3][: ml.value].copy())
        return res + (d,) if want_dir else res


    def align_3_intended(self, s1, s2, s3):
        """The recurrence algn_fill_cube intends (po_cost_3_intended, SURVEY.md "3-D policy" (1)): a correct three-sequence
        Needleman-Wunsch with the reference's candidate order.  Returns (cost, status, r1, r2, r3, median)."""
        s1, s2, s3 = _u8(s1), _u8(s2), _u8(s3)
        l1, l2, l3 = len(s1), len(s2), len(s3)
        d = np.zeros(l1 * l2 * l3, np.uint8)
        cost = self.L.po_cost_3_intended(C.byref(self.c), _p(s1, _u8p), l1, _p(s2, _u8p), l2, _p(s3, _u8p), l3, _p(d, _u8p))
        cap = l1 + l2 + l3
        r = [np.zeros(cap + 1, np.uint8) for _ in range(4)]
        st = C.c_int(0)
        n = self.L.po_backtrack_3_intended(C.byref(self.c), _p(d, _u8p), _p(s1, _u8p), l1, _p(s2, _u8p), l2, _p(s3, _u8p), l3,
                                           _p(r[0], _u8p), _p(r[1], _u8p), _p(r[2], _u8p), _p(r[3], _u8p), C.byref(st))
        return (cost, st.value, r[0][:n].copy(), r[1][:n].copy(), r[2][:n].copy(), r[3][:n].copy())


class Reference3:
    """3-D cube through the compiled reference (algn_nw_3d + backtrack_3d + algn_get_median_3d)."""

    kind = "reference"

    def __init__(self, cm3):
        L = self.L = C.CDLL(REF_SO)
        L.ref_cm3_create.restype = C.c_void_p
        L.ref_ws_create.restype = C.c_void_p
        cost = np.ascontiguousarray(cm3.cost, np.int32)
        med = np.ascontiguousarray(cm3.median, np.uint8)
        self.h = C.c_void_p(L.ref_cm3_create(cm3.a_sz_in, cm3.combinations, cm3.cost_model_type, cm3.gap_open,
                                             cm3.all_elements, _p(cost, _i32p), _p(med, _u8p)))
        self.ws = C.c_void_p(L.ref_ws_create())

    def align_3(self, s1, s2, s3, want_dir=False):
        s1, s2, s3 = _u8(s1), _u8(s2), _u8(s3)
        l1, l2, l3 = len(s1), len(s2), len(s3)
        cap = l1 + l2 + l3
        r = [np.zeros(cap + 1, np.uint8) for _ in range(4)]
        rl, ml, st = C.c_int(0), C.c_int(0), C.c_int(0)
        d = np.zeros(l1 * l2 * l3, np.uint16) if want_dir else None
        cost = self.L.ref_align_3(self.h, self.ws, _p(s1, _u8p), l1, _p(s2, _u8p), l2, _p(s3, _u8p), l3, _p(r[0], _u8p),
                                  _p(r[1], _u8p), _p(r[2], _u8p), C.byref(rl), _p(r[3], _u8p), C.byref(ml), C.byref(st),
                                  _p(d, _u16p))
        n = rl.value
        res = (cost, st.value, r[0][:n].copy(), r[1][:n].copy(), r[2][:n].copy(), r[3][: ml.value].copy())
        return res + (d,) if want_dir else res


def best_checker_3(cm3):
    return Reference3(cm3) if Reference.available() else Port3(cm3)


class Reference(_Base):
    kind = "reference"

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self, cm):
        L = self.L = C.CDLL(REF_SO)
        L.ref_cm_create.restype = C.c_void_p
        L.ref_ws_create.restype = C.c_void_p
        self.cm = cm
        cost = np.ascontiguousarray(cm.cost, np.int32)
        median = np.ascontiguousarray(cm.median, np.uint8)
        worst = np.ascontiguousarray(cm.worst, np.int32)
        prep = np.ascontiguousarray(cm.prepend_cost, np.int32)
        tail = np.ascontiguousarray(cm.tail_cost, np.int32)
        self.h = C.c_void_p(L.ref_cm_create(cm.a_sz_in, cm.combinations, cm.cost_model_type, cm.gap_open,
                                            cm.is_metric, cm.all_elements, _p(cost, _i32p), _p(median, _u8p),
                                            _p(worst, _i32p), _p(prep, _i32p), _p(tail, _i32p)))
        assert L.ref_cm_lcm(self.h) == cm.lcm and L.ref_cm_gap(self.h) == cm.gap and L.ref_cm_a_sz(self.h) == cm.a_sz
        self.ws = C.c_void_p(L.ref_ws_create())

    def cost_2(self, s1, s2, deltaw, want_dir=False):
        s1, s2 = _u8(s1), _u8(s2)
        assert len(s1) >= len(s2)
        if want_dir:
            # the reference never clears its direction matrix (src/matrices.c:77-119); size and zero it first
            # so that cells outside the visited region read 0
            self.L.ref_cost_2(self.h, self.ws, _p(s1, _u8p), len(s1), _p(s2, _u8p), len(s2), int(deltaw))
            self.L.ref_clear_dir(self.ws)
        r = self.L.ref_cost_2(self.h, self.ws, _p(s1, _u8p), len(s1), _p(s2, _u8p), len(s2), int(deltaw))
        if want_dir:
            d = np.zeros((len(s1), len(s2)), np.uint16)
            self.L.ref_get_dir_2(self.ws, len(s1), len(s2), _p(d, _u16p))
            return r, d
        return r

    def align_2(self, sa, sb, deltaw):
        """Sequence.Align.align_2, linear branch (src/sequence.ml:813-823, 849-861)."""
        sa, sb = _u8(sa), _u8(sb)
        cap = len(sa) + len(sb)
        ra, rb = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
        rl = C.c_int(0)
        if len(sa) >= len(sb):
            r = self.L.ref_align_2(self.h, self.ws, _p(sa, _u8p), len(sa), _p(sb, _u8p), len(sb), int(deltaw), 1,
                                   _p(ra, _u8p), _p(rb, _u8p), C.byref(rl))
        else:
            r = self.L.ref_align_2(self.h, self.ws, _p(sb, _u8p), len(sb), _p(sa, _u8p), len(sa), int(deltaw), 0,
                                   _p(rb, _u8p), _p(ra, _u8p), C.byref(rl))
        return r, ra[: rl.value].copy(), rb[: rl.value].copy()

    def median_2(self, which, a, b):
        a, b = _u8(a), _u8(b)
        out = np.zeros(len(a) + 2, np.uint8)
        n = self.L.ref_median_2(self.h, which, _p(a, _u8p), _p(b, _u8p), len(a), _p(out, _u8p))
        return out[:n].copy()

    def worst_2(self, a, b):
        """algn_worst_2 (src/algn.c:3373) of an aligned pair."""
        a, b = _u8(a), _u8(b)
        return int(self.L.ref_worst_2(self.h, _p(a, _u8p), _p(b, _u8p), len(a)))

    def align_affine_3(self, sa, sb, want_dir=False):
        sa, sb = _u8(sa), _u8(sb)
        cap = len(sa) + len(sb) + 2
        o = [np.zeros(cap, np.uint8) for _ in range(4)]
        lens = np.zeros(4, np.int32)
        d = np.zeros(len(sa) * len(sb), np.uint16) if want_dir else None
        r = self.L.ref_align_affine_3(self.h, self.ws, _p(sa, _u8p), len(sa), _p(sb, _u8p), len(sb),
                                      *[_p(x, _u8p) for x in o], _p(lens, _i32p), _p(d, _u16p))
        res = (r, o[0][: lens[0]].copy(), o[1][: lens[1]].copy(), o[2][: lens[2]].copy(), o[3][: lens[3]].copy())
        return res + (d,) if want_dir else res

    def cost_affine_3(self, sa, sb):
        sa, sb = _u8(sa), _u8(sb)
        return self.L.ref_cost_affine_3(self.h, self.ws, _p(sa, _u8p), len(sa), _p(sb, _u8p), len(sb))

    def _batch(self, mode, pool, off, length, pairs, deltaw, n, nthreads, cost, med, medwg, ra, rb, lens, stride):
        self.L.ref_batch(self.h, mode, _p(pool, _u8p), _p(off, _i64p), _p(length, _i32p), _p(pairs, _i32p),
                         _p(deltaw, _i32p), n, nthreads, _p(cost, _i32p), _p(med, _u8p), _p(medwg, _u8p),
                         _p(ra, _u8p), _p(rb, _u8p), _p(lens, _i32p), C.c_longlong(stride))


def best_checker(cm):
    """The compiled reference when present (this container, or shipped in oracle/_ref), else the port."""
    return Reference(cm) if Reference.available() else Port(cm)
