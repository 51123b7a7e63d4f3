"""ctypes binding of include/poyb200_tree.h: the NATIVE (C++) tree driver on the device-resident sequence store.

Same results as :class:`poyd_b200.tree.Evaluator` (which stays the readable, line-by-line cited statement of the algorithm
and drives the CPU checker in the tests), but the level-order batching, the hash-consed median cache, the Wagner sweep and
the SPR search run in C++ and the sequences never leave the GPU: per batch only task descriptors go up and costs, lengths
and gap counts come back."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib, sequence as S
from .cost_matrix import CostMatrix
from .tree import Topology


class _Topo(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("ids", C.c_void_p), ("nbr", C.c_void_p), ("handle", C.c_int32)]


class _Cost(C.Structure):
    _fields_ = [("adjusted", C.c_int64), ("unadjusted", C.c_int64), ("root_a", C.c_int32), ("root_b", C.c_int32)]


class _Spr(C.Structure):
    _fields_ = [("start_cost", C.c_int64), ("final_cost", C.c_int64), ("rounds", C.c_int32), ("breaks", C.c_int64),
                ("joins_swept", C.c_int64), ("exact_evaluated", C.c_int64)]


@dataclass
class NativeCost:
    adjusted: int
    unadjusted: int
    root: Tuple[int, int]


def _bind(L):
    if getattr(L, "_tree_bound", False):
        return
    L.poyb200_tree_create.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
    L.poyb200_tree_destroy.argtypes = [C.c_void_p]
    L.poyb200_tree_destroy.restype = None
    L.poyb200_tree_last_error.argtypes = [C.c_void_p]
    L.poyb200_tree_last_error.restype = C.c_char_p
    L.poyb200_tree_set_leaf.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
    L.poyb200_tree_reset.argtypes = [C.c_void_p]
    L.poyb200_tree_reset.restype = None
    L.poyb200_tree_evaluate.argtypes = [C.c_void_p, C.POINTER(_Topo), C.c_int32, C.c_int32, C.POINTER(_Cost)]
    L.poyb200_tree_single.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
    L.poyb200_tree_wagner.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.c_void_p]
    L.poyb200_tree_spr.argtypes = [C.c_void_p, C.POINTER(_Topo), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                   C.POINTER(_Spr)]
    L.poyb200_tree_spr_round.argtypes = [C.c_void_p, C.POINTER(_Topo), C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                         C.POINTER(_Spr)]
    L.poyb200_tree_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 5
    L.poyb200_tree_stats.restype = None
    L._tree_bound = True


def _pack(topo: Topology):
    ids = np.array(sorted(topo.nodes), np.int32)
    nbr = np.full((len(ids), 3), -1, np.int32)
    for k, v in enumerate(ids):
        n = topo.nodes[int(v)]
        nbr[k, :len(n)] = n
    return ids, nbr


def _unpack(ids, nbr, handle, n_taxa) -> Topology:
    nodes = {}
    for k, v in enumerate(ids):
        n = tuple(int(x) for x in nbr[k] if x >= 0)
        nodes[int(v)] = n
    return Topology(nodes, int(handle), n_taxa)


class NativeEvaluator:
    """The C++ driver bound to one cost matrix and one GPU."""

    def __init__(self, cm: CostMatrix, leaves: Dict[int, List[np.ndarray]], device: int = 0, config: Optional[dict] = None):
        self.al = S.Align(cm, device=device, config=config)  # owns the context and its cost matrix
        self.L = self.al.L
        _bind(self.L)
        self.n_loci = len(next(iter(leaves.values())))
        self.n_taxa = len(leaves)
        h = C.c_void_p()
        self._ck(self.L.poyb200_tree_create(self.al.h, self.n_loci, C.byref(h)), None)
        self.h = h
        for code, seqs in leaves.items():
            for l, q in enumerate(seqs):
                q = np.ascontiguousarray(q, np.uint8)
                self._ck(self.L.poyb200_tree_set_leaf(self.h, int(code), l, q.ctypes.data, len(q)))

    def _ck(self, rc, h="self"):
        if rc != 0:
            msg = self.L.poyb200_tree_last_error(self.h).decode() if h == "self" and getattr(self, "h", None) else ""
            raise S.PoyB200Error(f"poyb200_tree error {rc}: {msg}")

    def close(self):
        if getattr(self, "h", None):
            self.L.poyb200_tree_destroy(self.h)
            self.h = None
        self.al.close()

    def evaluate_many(self, topos: List[Topology], keep: bool = False) -> List[NativeCost]:
        packed = [_pack(t) for t in topos]
        arr = (_Topo * len(topos))()
        for k, (t, (ids, nbr)) in enumerate(zip(topos, packed)):
            arr[k] = _Topo(len(ids), ids.ctypes.data, nbr.ctypes.data, t.handle)
        out = (_Cost * len(topos))()
        self._ck(self.L.poyb200_tree_evaluate(self.h, arr, len(topos), int(keep), out))
        return [NativeCost(int(o.adjusted), int(o.unadjusted), (int(o.root_a), int(o.root_b))) for o in out]

    def evaluate(self, topo: Topology, keep: bool = False) -> NativeCost:
        return self.evaluate_many([topo], keep)[0]

    def single(self, which: int, vertex: int, locus: int) -> np.ndarray:
        buf = np.zeros(16384 + 16, np.uint8)
        n = C.c_int32(0)
        self._ck(self.L.poyb200_tree_single(self.h, which, vertex, locus, buf.ctypes.data, len(buf), C.byref(n)))
        return buf[: n.value].copy()

    def wagner(self, order: Optional[List[int]] = None):
        order = np.array(order if order is not None else list(range(1, self.n_taxa + 1)), np.int32)
        n = len(order)
        ids, nbr = np.zeros(2 * n, np.int32), np.zeros((2 * n, 3), np.int32)
        nn, handle = C.c_int32(0), C.c_int32(0)
        steps = np.zeros((max(n - 2, 0), 4), np.int64)
        self._ck(self.L.poyb200_tree_wagner(self.h, order.ctypes.data, n, ids.ctypes.data, nbr.ctypes.data, C.byref(nn), C.byref(handle),
                                            steps.ctypes.data))
        topo = _unpack(ids[: nn.value], nbr[: nn.value], handle.value, self.n_taxa)
        return topo, [{"taxon": int(s[0]), "edge": (int(s[1]), int(s[2])), "delta": int(s[3])} for s in steps]

    def spr(self, start: Topology, max_rounds: int = 0, window: int = 32):
        ids, nbr = _pack(start)
        st = _Topo(len(ids), ids.ctypes.data, nbr.ctypes.data, start.handle)
        oid, onbr = np.zeros(len(ids), np.int32), np.zeros((len(ids), 3), np.int32)
        handle = C.c_int32(0)
        res = _Spr()
        self._ck(self.L.poyb200_tree_spr(self.h, C.byref(st), max_rounds, window, oid.ctypes.data, onbr.ctypes.data, C.byref(handle),
                                         C.byref(res)))
        stats = {k: int(getattr(res, k)) for k, _ in _Spr._fields_}
        return _unpack(oid, onbr, handle.value, start.n_taxa), stats

    def spr_round(self, cur: Topology, best_cost: int, shard: int, nshards: int, window: int = 32):
        """One sharded round: (found, key, cost, joined topology or None, counters)."""
        ids, nbr = _pack(cur)
        st = _Topo(len(ids), ids.ctypes.data, nbr.ctypes.data, cur.handle)
        oid, onbr = np.zeros(len(ids), np.int32), np.zeros((len(ids), 3), np.int32)
        handle, found = C.c_int32(0), C.c_int32(0)
        key, cost = C.c_int64(0), C.c_int64(0)
        res = _Spr()
        self._ck(self.L.poyb200_tree_spr_round(self.h, C.byref(st), int(best_cost), shard, nshards, window, C.byref(found), C.byref(key),
                                               C.byref(cost), oid.ctypes.data, onbr.ctypes.data, C.byref(handle), C.byref(res)))
        stats = {k: int(getattr(res, k)) for k, _ in _Spr._fields_}
        topo = _unpack(oid, onbr, handle.value, cur.n_taxa) if found.value else None
        return bool(found.value), int(key.value), int(cost.value), topo, stats

    def stats(self) -> Dict[str, int]:
        v = [C.c_int64(0) for _ in range(5)]
        self.L.poyb200_tree_stats(self.h, *[C.byref(x) for x in v])
        return dict(zip(("calls", "pairs", "cells", "medians", "sequences"), (int(x.value) for x in v)))


def tree_cost_native(cm: CostMatrix, fasta: str, tree_file: str, which: int = 0, device: int = 0) -> NativeCost:
    """``read (fasta)  read (tree)`` -> ``Ptree.get_cost `Adjusted`` through the C++ driver (cf. tree.tree_cost)."""
    from .tree import GAP, parse_trees, read_fasta

    taxa = read_fasta(fasta)
    codes = {name: i + 1 for i, (name, _) in enumerate(taxa)}
    n_loci = max(len(fr) for _, fr in taxa)
    leaves = {codes[name]: fr + [np.array([GAP], np.uint8)] * (n_loci - len(fr)) for name, fr in taxa}
    with open(tree_file) as f:
        tree = parse_trees(f.read())[which]
    topo = Topology.convert_to(tree, codes)
    ev = NativeEvaluator(cm, leaves, device=device)
    try:
        return ev.evaluate(topo)
    finally:
        ev.close()
