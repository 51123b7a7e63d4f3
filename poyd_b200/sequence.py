"""Host-side mirror of the reference's ``Sequence`` / ``Sequence.Align`` (src/sequence.ml:453-1141) over batches.

Where the reference aligns one pair per call through the OCaml externals ``algn_CAML_*``, a caller here hands
over a :class:`SeqPool` (the distinct sequences) and a pair list, and every pair gets exactly what the
reference's function of the same name returns for it.  The functions keep the reference's names, argument
meaning and selection logic:

* :meth:`Align.cost_2`   -- ``Sequence.Align.cost_2`` (:691-723): affine matrices go to ``cost_2_affine``
  (``algn_CAML_cost_affine_3``), others to ``algn_CAML_simple_2`` with the Ukkonen ``deltaw`` computed from
  the gap counts and lengths exactly as :691-714 does.
* :meth:`Align.align_2`  -- ``Sequence.Align.align_2`` (:849-869): ``align_affine_3`` or
  ``cost_2`` + ``create_edited_2`` (:813-823).
* :meth:`Align.align_affine_3` (:469-478), :meth:`Align.median_2` (:907-918),
  :meth:`Align.median_2_with_gaps` (:895-905), :meth:`Align.ancestor_2` (:922-932),
  :meth:`Align.full_median_2` (:949-957).

All compute runs in the CUDA library behind include/poyb200.h; nothing here computes an alignment on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence as _Seq

import numpy as np

from . import _lib
from .cost_matrix import CostMatrix

WANT_MEDIAN, WANT_MEDIANWG, WANT_ALIGNED, WANT_BITSETS, WANT_CLOSEST = 1, 2, 4, 8, 16
MODE_COST_2, MODE_ALIGN_2, MODE_COST_AFFINE_3, MODE_ALIGN_AFFINE_3 = 0, 1, 2, 3


class PoyB200Error(RuntimeError):
    """Raised where the reference would raise ``Failure`` (failwith) -- plus CUDA errors."""


class SeqPool:
    """The distinct sequences of a batch: one ``uint8`` per element, leading gap included (src/seq.h:48-58).

    Sequence starts are 16-byte aligned so the kernels can stage them with bulk copies."""

    ALIGN = 16

    def __init__(self, seqs: Iterable[np.ndarray]):
        seqs = [np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]
        lens = np.array([len(s) for s in seqs], dtype=np.int32)
        padded = (lens.astype(np.int64) + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        off = np.zeros(len(seqs), dtype=np.int64)
        if len(seqs):
            off[1:] = np.cumsum(padded)[:-1]
        pool = np.zeros(int(padded.sum()) + self.ALIGN, dtype=np.uint8)
        for s, o in zip(seqs, off):
            pool[o:o + len(s)] = s
        self.pool, self.off, self.len = pool, off, lens
        self._zero_padded = True  # built here: the bytes between sequences are 0

    @classmethod
    def from_matrix(cls, mat: np.ndarray, lens: np.ndarray) -> "SeqPool":
        """Rows of a [n, stride] uint8 matrix, row r holding lens[r] elements; stride must be a multiple of 16."""
        self = cls.__new__(cls)
        mat = np.ascontiguousarray(mat, dtype=np.uint8)
        assert mat.shape[1] % cls.ALIGN == 0
        self.pool = mat.reshape(-1)
        self.off = np.arange(mat.shape[0], dtype=np.int64) * mat.shape[1]
        self.len = np.ascontiguousarray(lens, dtype=np.int32)
        self._stride = mat.shape[1]  # the caller's rows may hold anything past lens[r]
        return self

    def __len__(self) -> int:
        return len(self.len)

    def seq(self, i: int) -> np.ndarray:
        return self.pool[self.off[i]:self.off[i] + self.len[i]]

    def count(self, gap: int) -> np.ndarray:
        """``Sequence.count gap s`` = seq_CAML_count (src/seq.c:570-582) for every sequence: the number of
        elements with ``code land gap <> 0`` (a bitwise test even for sequential alphabets)."""
        cache = self.__dict__.setdefault("_count_cache", {})
        if gap not in cache:
            off = np.asarray(self.off, np.int64)
            if getattr(self, "_stride", 0):
                # matrix-shaped pool: count inside each row's first len elements only (the padding is the caller's), by blocks
                st, n = self._stride, len(self.len)
                mat = self.pool[:n * st].reshape(n, st)
                out = np.zeros(n, np.int32)
                cols = np.arange(st, dtype=np.int32)[None, :]
                for lo in range(0, n, 65536):
                    hi = min(n, lo + 65536)
                    hit = ((mat[lo:hi] & np.uint8(gap)) != 0) & (cols < self.len[lo:hi, None])
                    out[lo:hi] = hit.sum(axis=1, dtype=np.int32)
                cache[gap] = out
            elif getattr(self, "_zero_padded", False) and len(off):
                # pool built by __init__: ordered, padding bytes are 0 and never hit -- one pass
                hit = (self.pool & np.uint8(gap)) != 0
                cache[gap] = np.add.reduceat(hit, off, dtype=np.int32).astype(np.int32) * (self.len > 0)
            else:
                hit = (self.pool & np.uint8(gap)) != 0
                cs = np.concatenate([[0], np.cumsum(hit, dtype=np.int64)])
                cache[gap] = (cs[off + self.len] - cs[off]).astype(np.int32)
        return cache[gap]


@dataclass
class Aligned:
    """Result of an align call.  Rows are right aligned in [n, stride] arrays (the layout of the reference's
    ``struct seq``: begin = end - len + 1); use :meth:`get` for one trimmed sequence."""

    cost: np.ndarray
    lens: Optional[np.ndarray] = None  # [n, 4]: median, medianwg, aligned a, aligned b
    median: Optional[np.ndarray] = None
    medianwg: Optional[np.ndarray] = None
    aligned_a: Optional[np.ndarray] = None
    aligned_b: Optional[np.ndarray] = None
    bits_a: Optional[np.ndarray] = None   # WANT_BITSETS: right-aligned bit rows (numpy.unpackbits order)
    bits_b: Optional[np.ndarray] = None
    bits_wg: Optional[np.ndarray] = None

    _COL = {"median": 0, "medianwg": 1, "aligned_a": 2, "aligned_b": 3}

    def get(self, what: str, p: int) -> np.ndarray:
        buf = getattr(self, what)
        n = int(self.lens[p, self._COL[what]])
        return buf[p, buf.shape[1] - n:]

    def bitset(self, what: str, p: int) -> np.ndarray:
        """``seq_to_bitset gap`` of the aligned operand a / b or of medianwg (src/seqCS.ml:649-655) as 0/1 flags in
        column order: 1 where the element is not the gap."""
        n = int(self.lens[p, 2])
        return np.unpackbits(getattr(self, "bits_" + what)[p])[-n:] if n else np.zeros(0, np.uint8)


def deltaw_calc(s1len: np.ndarray, s2len: np.ndarray, deltaw: Optional[np.ndarray]) -> np.ndarray:
    """``deltaw_calc`` of Sequence.Align.cost_2 (src/sequence.ml:692-702), s1len >= s2len."""
    dif = s1len - s2len
    lower = (s1len.astype(np.float64) * 0.10).astype(np.int64)  # int_of_float truncates
    if deltaw is None:
        return np.where(dif < lower, lower // 2, 2).astype(np.int32)
    return np.where(dif < lower, lower, deltaw).astype(np.int32)


class Align:
    """``Sequence.Align`` bound to one cost matrix and one GPU."""

    def __init__(self, cm: CostMatrix, device: int = -1, config: Optional[dict] = None):
        """config: overrides of poyb200_config fields (include/poyb200.h), e.g. {"force_generic": 1}."""
        self.L = _lib.lib()
        self.cm = cm
        h = C.c_void_p()
        cfg = _lib.make_config(config)
        rc = self.L.poyb200_create_ex(device, C.byref(cfg), C.byref(h))
        if rc != 0:
            raise PoyB200Error(f"poyb200_create_ex failed ({rc}): bad config or no usable CUDA device; there is no CPU fallback")
        self.h = h
        self._set_cm(cm)
        self._keep = None

    def close(self) -> None:
        if getattr(self, "h", None):
            self.L.poyb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise PoyB200Error(f"poyb200 error {rc}: {self.L.poyb200_last_error(self.h).decode()}")

    def _set_cm(self, cm: CostMatrix) -> None:
        self._tabs = [np.ascontiguousarray(cm.cost, np.int32), np.ascontiguousarray(cm.median, np.uint8),
                      np.ascontiguousarray(cm.worst, np.int32), np.ascontiguousarray(cm.prepend_cost, np.int32),
                      np.ascontiguousarray(cm.tail_cost, np.int32)]
        c = _lib.CM(cm.a_sz, cm.lcm, cm.gap, cm.cost_model_type, cm.combinations, cm.gap_open, cm.is_metric,
                    cm.all_elements, self._tabs[0].ctypes.data_as(_lib.i32p), self._tabs[1].ctypes.data_as(_lib.u8p),
                    self._tabs[2].ctypes.data_as(_lib.i32p), self._tabs[3].ctypes.data_as(_lib.i32p),
                    self._tabs[4].ctypes.data_as(_lib.i32p))
        self._check(self.L.poyb200_set_cm(self.h, C.byref(c)))

    @property
    def is_affine(self) -> bool:
        return self.cm.cost_model_type == 1

    _ONE_SHOT = ("poyb200_batch_cost_2", "poyb200_batch_align_2", "poyb200_batch_cost_affine_3", "poyb200_batch_align_affine_3")

    def _one_shot(self, mode: int, batch) -> None:
        """The one-shot C ABI call of `mode` (H2D, kernels, D2H) on this object's device(s)."""
        self._check(getattr(self.L, self._ONE_SHOT[mode])(self.h, C.byref(batch)))

    # ---- deltaw exactly as Sequence.Align.cost_2 computes it -------------------------------------------
    def deltaw_for(self, pool: SeqPool, pairs: np.ndarray, deltaw: Optional[np.ndarray] = None) -> np.ndarray:
        cnt = pool.count(self.cm.gap)
        la, lb = pool.len[pairs[:, 0]].astype(np.int64), pool.len[pairs[:, 1]].astype(np.int64)
        gaps = np.maximum(cnt[pairs[:, 0]], cnt[pairs[:, 1]])
        return (gaps + deltaw_calc(np.maximum(la, lb), np.minimum(la, lb), deltaw)).astype(np.int32)

    # ---- batch plumbing -----------------------------------------------------------------------------
    def make_batch(self, pool: SeqPool, pairs, deltaw=None, swaped=None, want: int = 0, outputs: bool = True):
        """Builds the C batch descriptor (and output arrays).  Returns (batch, Aligned)."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = len(pairs)
        keep = [pool.pool, pool.off, pool.len, pairs]
        b = _lib.Batch()
        b.pool, b.pool_bytes = pool.pool.ctypes.data, pool.pool.nbytes
        b.seq_off, b.seq_len, b.n_seqs = pool.off.ctypes.data, pool.len.ctypes.data, len(pool)
        b.pairs, b.n_pairs = pairs.ctypes.data, n
        if deltaw is not None:
            deltaw = np.ascontiguousarray(deltaw, dtype=np.int32)
            keep.append(deltaw)
            b.deltaw = deltaw.ctypes.data
        if swaped is not None:
            swaped = np.ascontiguousarray(swaped, dtype=np.uint8)
            keep.append(swaped)
            b.swaped = swaped.ctypes.data
        res = Aligned(cost=np.zeros(n, np.int32))
        b.cost = res.cost.ctypes.data
        b.want = want
        if (want & WANT_BITSETS) and outputs:
            cap = int((pool.len[pairs[:, 0]].astype(np.int64) + pool.len[pairs[:, 1]]).max()) + 2 if n else 16
            bstride = ((cap + 7) // 8 + 3) // 4 * 4
            res.bits_a, res.bits_b, res.bits_wg = (np.zeros((n, bstride), np.uint8) for _ in range(3))
            b.bits_a, b.bits_b, b.bits_wg = res.bits_a.ctypes.data, res.bits_b.ctypes.data, res.bits_wg.ctypes.data
            b.bits_stride = bstride
        if want and outputs:
            stride = 16
            if n:
                stride = int((pool.len[pairs[:, 0]].astype(np.int64) + pool.len[pairs[:, 1]]).max()) + 2
            stride = (stride + 15) // 16 * 16
            res.lens = np.zeros((n, 4), np.int32)
            b.out_len, b.out_stride = res.lens.ctypes.data, stride
            if want & (WANT_MEDIAN | WANT_CLOSEST):
                res.median = np.zeros((n, stride), np.uint8)
                b.median = res.median.ctypes.data
            if want & WANT_MEDIANWG:
                res.medianwg = np.zeros((n, stride), np.uint8)
                b.medianwg = res.medianwg.ctypes.data
            if want & WANT_ALIGNED:
                res.aligned_a = np.zeros((n, stride), np.uint8)
                res.aligned_b = np.zeros((n, stride), np.uint8)
                b.aligned_a, b.aligned_b = res.aligned_a.ctypes.data, res.aligned_b.ctypes.data
        self._keep = keep
        return b, res

    # ---- Sequence.Align ------------------------------------------------------------------------------
    def cost_2(self, pool: SeqPool, pairs, deltaw: Optional[np.ndarray] = None, raw_deltaw: bool = False) -> np.ndarray:
        """``Sequence.Align.cost_2 ?deltaw s1 s2 m1 m2`` for every pair (src/sequence.ml:716-723).

        raw_deltaw=True passes ``deltaw`` straight to ``algn_CAML_simple_2`` (the external's own argument)."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        if self.is_affine:
            b, res = self.make_batch(pool, pairs)
            self._one_shot(MODE_COST_AFFINE_3, b)
        else:
            dw = np.asarray(deltaw, np.int32) if raw_deltaw else self.deltaw_for(pool, pairs, deltaw)
            b, res = self.make_batch(pool, pairs, deltaw=dw)
            self._one_shot(MODE_COST_2, b)
        return res.cost

    def align_affine_3(self, pool: SeqPool, pairs, want: int = WANT_MEDIAN | WANT_MEDIANWG | WANT_ALIGNED) -> Aligned:
        """``Sequence.Align.align_affine_3 si sj cm`` (src/sequence.ml:469-478): median, resi, resj, cost,
        medianwg for every pair."""
        b, res = self.make_batch(pool, pairs, want=want)
        self._one_shot(MODE_ALIGN_AFFINE_3, b)
        return res

    def align_2(self, pool: SeqPool, pairs, want: int = WANT_ALIGNED, deltaw: Optional[np.ndarray] = None,
                raw_deltaw: bool = False, swaped: Optional[np.ndarray] = None) -> Aligned:
        """``Sequence.Align.align_2 s1 s2 c m`` (src/sequence.ml:849-861) for every pair; with WANT_MEDIAN /
        WANT_MEDIANWG also ``ancestor_2`` / ``median_2_with_gaps`` of the aligned pair, which is what
        ``SeqCS.DOS.median`` computes next (src/seqCS.ml:757-766)."""
        if self.is_affine:
            return self.align_affine_3(pool, pairs, want)
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        dw = np.asarray(deltaw, np.int32) if raw_deltaw else self.deltaw_for(pool, pairs, deltaw)
        b, res = self.make_batch(pool, pairs, deltaw=dw, swaped=swaped, want=want)
        self._one_shot(MODE_ALIGN_2, b)
        return res

    def closest(self, pool: SeqPool, pairs, prechecked: bool = False) -> List[np.ndarray]:
        """``Sequence.Align.closest s1 s2 cm m`` (src/sequence.ml:967-1033) for every pair (s1 = pair[0], s2 = pair[1]):
        the sequence of s2's elements closest to s1's along their alignment, gaps removed.  The alignment, the
        column rule (``Cost_matrix.Two_D.get_closest``) and the compaction run in the traceback kernel
        (POYB200_WANT_CLOSEST); the two early exits -- empty s2, s1 = s2 -- are decided here."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        gap = self.cm.gap
        out: List[Optional[np.ndarray]] = [None] * len(pairs)
        todo = []
        if prechecked:  # the caller has taken both early exits already (poyd_b200/tree.py does)
            r = self.align_2(pool, pairs, WANT_CLOSEST)
            return [r.get("median", q).copy() for q in range(len(pairs))]
        for k, (i, j) in enumerate(pairs):
            s1, s2 = pool.seq(int(i)), pool.seq(int(j))
            if np.all(s2 == gap):  # is_empty s2: (s2, 0)
                out[k] = s2.copy()
            elif self.cm.combine() and len(s1) == len(s2) and np.array_equal(s1, s2):
                v = s2.copy()  # :1000-1009: gap bits cleared past the first element, get_closest of x with itself
                v[1:] &= np.uint8(~gap & 0xFF)
                one = np.array([self._closest_same(int(x)) for x in v], np.uint8)
                out[k] = np.concatenate([[gap], one[one != gap]]).astype(np.uint8)
            else:
                todo.append(k)
        if todo:
            r = self.align_2(pool, pairs[todo], WANT_CLOSEST)
            for q, k in enumerate(todo):
                out[k] = r.get("median", q).copy()
        return out  # type: ignore[return-value]

    def _closest_same(self, x: int) -> int:
        """get_closest cm x x for one element (src/cost_matrix.ml:681-700)."""
        gap, b = self.cm.gap, x
        if x != gap:
            b = gap if (x & gap) else (x & ~gap)
        best, cur = x, None
        for bit in range(self.cm.lcm):
            if b & (1 << bit):
                nc = int(self.cm.cost[x, 1 << bit])
                if cur is None or nc < cur:
                    best, cur = 1 << bit, nc
        return best

    def _median(self, which: int, a: np.ndarray, b: np.ndarray, lens: np.ndarray):
        a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        if a.shape != b.shape:
            raise ValueError("The size of the sequences is not the same.")  # Invalid_Argument, sequence.ml:903
        n, stride = a.shape
        ostride = (stride + 1 + 15) // 16 * 16
        out, olen = np.zeros((n, ostride), np.uint8), np.zeros(n, np.int32)
        self._check(self.L.poyb200_batch_median_2(self.h, which, a.ctypes.data, b.ctypes.data, stride, lens.ctypes.data, n,
                                                  out.ctypes.data, ostride, olen.ctypes.data))
        return out, olen

    def ancestor_2(self, a, b, lens):
        """``Sequence.Align.ancestor_2`` (src/sequence.ml:922-932) on rows of LEFT-aligned aligned sequences."""
        return self._median(0, a, b, lens)

    def median_2_with_gaps(self, a, b, lens):
        """``Sequence.Align.median_2_with_gaps`` (src/sequence.ml:895-905)."""
        return self._median(1, a, b, lens)

    def median_2(self, a, b, lens):
        """``Sequence.Align.median_2`` (src/sequence.ml:907-918)."""
        return self._median(2, a, b, lens)

    def worst_2(self, a, b, lens, verify: bool = False) -> np.ndarray:
        """``Sequence.Align.max_cost_2`` (``algn_CAML_worst_2``, src/algn.c:3382; verify=True: ``algn_CAML_verify_2``, :3395) on
        rows of LEFT-aligned aligned sequences."""
        a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        if a.shape != b.shape:
            raise ValueError("The size of the sequences is not the same.")
        n, stride = a.shape
        out = np.zeros(n, np.int32)
        self._check(self.L.poyb200_batch_worst_2(self.h, 1 if verify else 0, a.ctypes.data, b.ctypes.data, stride, lens.ctypes.data, n,
                                                 out.ctypes.data))
        return out

    def full_median_2(self, pool: SeqPool, pairs):
        """``Sequence.Align.full_median_2`` (src/sequence.ml:949-957): the affine median, or align_2 followed
        by median_2 (no gaps).  Returns (rows, lens), rows right aligned."""
        if self.is_affine:
            r = self.align_affine_3(pool, pairs, WANT_MEDIAN)
            return r.median, r.lens[:, 0].copy()
        r = self.align_2(pool, pairs, WANT_ALIGNED)
        n, stride = r.aligned_a.shape
        la = r.lens[:, 2]
        a, b = np.zeros_like(r.aligned_a), np.zeros_like(r.aligned_b)
        for p in range(n):  # left align for the median call (host-side plumbing only)
            a[p, :la[p]] = r.aligned_a[p, stride - la[p]:]
            b[p, :la[p]] = r.aligned_b[p, stride - la[p]:]
        return self.median_2(a, b, la)

    # ---- split calls for benchmarks ----------------------------------------------------------------------
    def stage(self, mode: int, batch) -> None:
        self._check(self.L.poyb200_stage(self.h, mode, C.byref(batch)))

    def run(self) -> None:
        self._check(self.L.poyb200_run(self.h))

    def sync(self) -> None:
        self._check(self.L.poyb200_sync(self.h))

    def fetch(self) -> None:
        self._check(self.L.poyb200_fetch(self.h))

    def launch_count(self) -> int:
        return int(self.L.poyb200_launch_count(self.h))

    def last_run_ms(self):
        ms = (C.c_float * 2)()
        self._check(self.L.poyb200_last_run_ms(self.h, ms))
        return float(ms[0]), float(ms[1])

    def int32_peak(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._check(self.L.poyb200_int32_peak(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value


class MultiAlign(Align):
    """``Sequence.Align`` over several GPUs of the node in ONE process (``poyb200_multi_*``): every batch call is cut into
    contiguous pair ranges of about equal work, one per device, each on its own host thread; results land in the same
    arrays a single-device call fills.  `devices` = list of CUDA ordinals."""

    def __init__(self, cm: CostMatrix, devices, config: Optional[dict] = None):
        self.L = _lib.lib()
        self.cm = cm
        devices = [int(d) for d in devices]
        arr = (C.c_int * len(devices))(*devices)
        cfg = _lib.make_config(config)
        m = C.c_void_p()
        rc = self.L.poyb200_multi_create(arr, len(devices), C.byref(cfg), C.byref(m))
        if rc != 0:
            raise PoyB200Error(f"poyb200_multi_create failed ({rc}): bad config or a device is not usable; there is no CPU fallback")
        self.m = m
        self.h = C.c_void_p(self.L.poyb200_multi_ctx(m, 0))  # median_2 / worst_2 / probes run on the first device
        self.devices = devices
        self._tabs = [np.ascontiguousarray(cm.cost, np.int32), np.ascontiguousarray(cm.median, np.uint8),
                      np.ascontiguousarray(cm.worst, np.int32), np.ascontiguousarray(cm.prepend_cost, np.int32),
                      np.ascontiguousarray(cm.tail_cost, np.int32)]
        c = _lib.CM(cm.a_sz, cm.lcm, cm.gap, cm.cost_model_type, cm.combinations, cm.gap_open, cm.is_metric,
                    cm.all_elements, self._tabs[0].ctypes.data_as(_lib.i32p), self._tabs[1].ctypes.data_as(_lib.u8p),
                    self._tabs[2].ctypes.data_as(_lib.i32p), self._tabs[3].ctypes.data_as(_lib.i32p),
                    self._tabs[4].ctypes.data_as(_lib.i32p))
        if self.L.poyb200_multi_set_cm(m, C.byref(c)) != 0:
            raise PoyB200Error(self.L.poyb200_multi_last_error(m).decode())
        self._keep = None

    def close(self) -> None:
        if getattr(self, "m", None):
            self.L.poyb200_multi_destroy(self.m)
            self.m = None
            self.h = None

    def _one_shot(self, mode: int, batch) -> None:
        rc = self.L.poyb200_multi_batch(self.m, mode, C.byref(batch))
        if rc != 0:
            raise PoyB200Error(f"poyb200 error {rc}: {self.L.poyb200_multi_last_error(self.m).decode()}")

    def launch_count(self) -> int:
        return int(self.L.poyb200_multi_launch_count(self.m))

    def shards(self) -> np.ndarray:
        """Pair index at which each device's shard of the last call began (len(devices) + 1 entries)."""
        out = (C.c_int64 * (len(self.devices) + 1))()
        n = self.L.poyb200_multi_shards(self.m, out, len(self.devices) + 1)
        return np.array(out[:n], dtype=np.int64)


def cells_linear(l1: int, l2: int, deltaw: int) -> int:
    """DP cells ``algn_fill_plane_2`` visits for stored lengths l1, l2."""
    return int(_lib.lib().poyb200_cells_linear(int(l1), int(l2), int(deltaw)))


def cells_affine(la: int, lb: int) -> int:
    """DP cells ``algn_fill_plane_3_aff`` visits (band + left-edge cells + the initialised row)."""
    return int(_lib.lib().poyb200_cells_affine(int(la), int(lb)))


@dataclass
class Aligned3:
    """Result of :meth:`Align3.align_3`: rows right aligned, ``lens[t]`` elements each; ``status[t] = 1`` marks triples
    whose traceback the reference would run off the start of a sequence (undefined there; nothing is returned)."""

    cost: np.ndarray
    lens: Optional[np.ndarray] = None
    status: Optional[np.ndarray] = None
    aligned_1: Optional[np.ndarray] = None
    aligned_2: Optional[np.ndarray] = None
    aligned_3: Optional[np.ndarray] = None
    median: Optional[np.ndarray] = None
    median_lens: Optional[np.ndarray] = None  # align_3_powell only: the median is shorter than the rows

    def get(self, what: str, t: int) -> np.ndarray:
        buf = getattr(self, what)
        n = int(self.median_lens[t]) if (what == "median" and self.median_lens is not None) else int(self.lens[t])
        return buf[t, buf.shape[1] - n:]


def readjust_3d_classify(pool: "SeqPool", quads, gap: int) -> np.ndarray:
    """The branch ``Sequence.Align.readjust_3d`` takes for every row (s1, s2, m, p) (src/sequence.ml:1117-1129): 0 keep m
    (all four lengths say nothing can change), 1 return s2 (s1 empty), 2 return s1 (s2 empty), 3 return p (p empty),
    4 align (s1, s2, p).  `is_empty` = every element is the gap (src/sequence.ml:442-449)."""
    quads = np.asarray(quads, np.int32).reshape(-1, 4)
    out = np.full(len(quads), 4, np.int8)
    for k, (s1, s2, m, p) in enumerate(quads):
        l1, l2, lm, lp = (int(pool.len[i]) for i in (s1, s2, m, p))
        if l1 == l2 and lm == lp and l1 == lm:
            out[k] = 0
            continue
        e1, e2, ep = (bool(np.all(pool.seq(int(i)) == gap)) for i in (s1, s2, p))
        if e1 and not e2:
            out[k] = 1
        elif e2 and not e1:
            out[k] = 2
        elif ep:
            out[k] = 3
    return out


class Align3(Align):
    """``Sequence.Align.align_3`` / ``cost_3`` / ``median_3`` (src/sequence.ml:727-762, 871-893, 934-947) over a batch
    of triples, with the reference's cube semantics as executed."""

    def __init__(self, cm: CostMatrix, cm3, device: int = -1, config: Optional[dict] = None):
        super().__init__(cm, device, config)
        self._t3 = [np.ascontiguousarray(cm3.cost, np.int32), np.ascontiguousarray(cm3.median, np.uint8)]
        c = _lib.CM3(cm3.lcm, cm3.gap, self._t3[0].ctypes.data_as(_lib.i32p), self._t3[1].ctypes.data_as(_lib.u8p))
        self._check(self.L.poyb200_set_cm_3d(self.h, C.byref(c)))

    def align_3(self, pool: SeqPool, triples, want: int = 3) -> Aligned3:
        triples = np.ascontiguousarray(triples, dtype=np.int32).reshape(-1, 3)
        n = len(triples)
        b = _lib.Batch3()
        b.pool, b.pool_bytes = pool.pool.ctypes.data, pool.pool.nbytes
        b.seq_off, b.seq_len, b.n_seqs = pool.off.ctypes.data, pool.len.ctypes.data, len(pool)
        b.triples, b.n_triples, b.want = triples.ctypes.data, n, want
        res = Aligned3(cost=np.zeros(n, np.int32))
        b.cost = res.cost.ctypes.data
        if want:
            stride = 16
            if n:
                stride = int(pool.len[triples].astype(np.int64).sum(axis=1).max())
            stride = (stride + 15) // 16 * 16
            res.lens, res.status = np.zeros(n, np.int32), np.zeros(n, np.int32)
            b.out_len, b.status, b.out_stride = res.lens.ctypes.data, res.status.ctypes.data, stride
            if want & 1:
                res.aligned_1, res.aligned_2, res.aligned_3 = (np.zeros((n, stride), np.uint8) for _ in range(3))
                b.aligned_1, b.aligned_2, b.aligned_3 = (x.ctypes.data for x in (res.aligned_1, res.aligned_2, res.aligned_3))
            if want & 2:
                res.median = np.zeros((n, stride), np.uint8)
                b.median = res.median.ctypes.data
        self._check(self.L.poyb200_batch_align_3(self.h, C.byref(b)))
        return res

    def cost_3(self, pool: SeqPool, triples) -> np.ndarray:
        return self.align_3(pool, triples, want=0).cost

    def align_3_powell(self, pool: SeqPool, triples, mm: int, go: int, ge: int, want: int = 1) -> Aligned3:
        """``Sequence.Align.align_3_powell s1 s2 s3 mm go ge`` (src/sequence.ml:1078-1087) for every triple: Powell's
        three-sequence aligner under affine gap costs (src/ukk.checkp.c).  want & 1: the three aligned rows; want & 2: the
        median of ``align_3_powell_inter`` (:1103-1114).  ``lens`` = aligned lengths, ``median_lens`` = median lengths;
        ``status[t] = 5`` marks a triple holding an element without a base (the reference raises there)."""
        triples = np.ascontiguousarray(triples, dtype=np.int32).reshape(-1, 3)
        n = len(triples)
        b = _lib.Batch3()
        b.pool, b.pool_bytes = pool.pool.ctypes.data, pool.pool.nbytes
        b.seq_off, b.seq_len, b.n_seqs = pool.off.ctypes.data, pool.len.ctypes.data, len(pool)
        b.triples, b.n_triples, b.want = triples.ctypes.data, n, want
        res = Aligned3(cost=np.zeros(n, np.int32))
        b.cost = res.cost.ctypes.data
        stride = 16
        if n:
            stride = int(pool.len[triples].astype(np.int64).sum(axis=1).max())
        stride = (stride + 15) // 16 * 16
        lens2 = np.zeros((n, 2), np.int32)
        res.status = np.zeros(n, np.int32)
        b.out_len, b.status, b.out_stride = lens2.ctypes.data, res.status.ctypes.data, stride
        if want & 1:
            res.aligned_1, res.aligned_2, res.aligned_3 = (np.zeros((n, stride), np.uint8) for _ in range(3))
            b.aligned_1, b.aligned_2, b.aligned_3 = (x.ctypes.data for x in (res.aligned_1, res.aligned_2, res.aligned_3))
        if want & 2:
            res.median = np.zeros((n, stride), np.uint8)
            b.median = res.median.ctypes.data
        self._check(self.L.poyb200_batch_powell_3(self.h, C.byref(b), int(mm), int(go), int(ge)))
        res.lens = np.ascontiguousarray(lens2[:, 0])
        res.median_lens = np.ascontiguousarray(lens2[:, 1])
        return res

    def readjust_3d(self, pool: SeqPool, quads, first_gap: bool = True):
        """``Sequence.Align.readjust_3d ?first_gap s1 s2 m cm cm3 p`` (src/sequence.ml:1116-1139) for every row
        (s1, s2, m, p) of `quads`: the early exits on the host, everything else as ONE batch of Powell alignments of
        (s1, s2, p).  Returns (cost[n], list of the n new sequences, changed[n])."""
        quads = np.ascontiguousarray(quads, dtype=np.int32).reshape(-1, 4)
        n = len(quads)
        what = readjust_3d_classify(pool, quads, self.cm.gap)
        cost = np.zeros(n, np.int64)
        res: List[Optional[np.ndarray]] = [None] * n
        for k in np.nonzero(what != 4)[0]:
            res[k] = pool.seq(int(quads[k, (2, 1, 0, 3)[what[k]]])).copy()  # 0: m, 1: s2, 2: s1, 3: p
        todo = np.nonzero(what == 4)[0]
        if len(todo):
            tri = quads[todo][:, (0, 1, 3)]
            work = pool
            if not first_gap:  # prepend_char s gap on all three, del_first_char on the result (:1132-1138)
                used = np.unique(tri)
                seqs = [np.concatenate([[self.cm.gap], pool.seq(int(i))]).astype(np.uint8) for i in used]
                work = SeqPool(seqs)
                tri = np.searchsorted(used, tri).astype(np.int32)
            g = self.align_3_powell_inter(work, tri, want=2)
            if g.status.any():
                raise PoyB200Error(f"powell_3D_align: status {g.status[g.status != 0][:4]} (5 = an element without a base)")
            for q, k in enumerate(todo):
                m = g.get("median", q).copy()
                res[k] = m if first_gap else m[1:]
                cost[k] = int(g.cost[q])
        changed = np.array([not np.array_equal(res[k], pool.seq(int(quads[k, 2]))) for k in range(n)], bool)
        return cost, res, changed

    def align_3_powell_inter(self, pool: SeqPool, triples, want: int = 3) -> Aligned3:
        """``Sequence.Align.align_3_powell_inter`` (src/sequence.ml:1089-1114): mismatch, gap opening and gap extension
        taken from the 2-D matrix (cost 1 2, the affine opening or 0, cost 1 16), medians through the 3-D matrix."""
        model = self.cm.affine()
        go = int(model[1]) if model[0] == "Affine" else 0
        return self.align_3_powell(pool, triples, int(self.cm.cost[1, 2]), go, int(self.cm.cost[1, 16]), want)
