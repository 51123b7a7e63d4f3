"""``SeqCS.DOS`` / ``SeqCS.Union`` operations of the uppass as BATCHES over the C ABI.

The reference runs these one vertex at a time (src/seqCS.ml); here every function takes lists of vertices and issues a
fixed number of library calls whatever the number of vertices:

* :meth:`DOS.median_3_no_union`  (src/seqCS.ml:778-796)  align_2 of the parent with either child, median_2 (no gaps) and
  max_cost_2 (``algn_CAML_worst_2``) of both aligned pairs, the cheaper one wins, a gap is prepended if missing;
* :meth:`DOS.median_3_union`     (src/seqCS.ml:798-817)  ``Sequence.Align.union`` of the vertex's two aligned children
  (``algn_union``, src/algn.c:4176-4184: element-wise OR), align_2 of the parent with it, median_2, max_cost_2;
* :meth:`DOS.distance`           (src/seqCS.ml:819-867)  cost_2 with the ``max 8 |la - lb|`` hint, empty operands cost 0;
* :meth:`Union.distance_union`   (src/seqCS.ml:1569-1616) the same cost per locus, times 0.8 under affine gaps.

All alignment and all per-column work is the library's (``poyb200_batch_align_2 / align_affine_3 / median_2 /
worst_2``); this module only orders operands and moves rows, like the OCaml code it mirrors."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import sequence as S


def _left(rows: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """Right-aligned rows (the library's output layout) -> left-aligned rows of the same width (its aligned-pair input)."""
    n, w = rows.shape
    out = np.zeros_like(rows)
    cols = np.arange(w)[None, :]
    src = cols + (w - lens[:, None])
    ok = src < w
    out[ok] = rows[np.nonzero(ok)[0], src[ok]]
    return out


class DOS:
    def __init__(self, al: S.Align):
        self.al = al
        self.gap = al.cm.gap

    def _empty(self, pool: S.SeqPool, i: int) -> bool:
        return bool(np.all(pool.seq(int(i)) == self.gap))

    def distance(self, pool: S.SeqPool, pairs) -> np.ndarray:
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        out = np.zeros(len(pairs), np.int64)
        todo = [k for k, (a, b) in enumerate(pairs) if not (self._empty(pool, a) or self._empty(pool, b))]
        if todo:
            pp = pairs[todo]
            la, lb = pool.len[pp[:, 0]].astype(np.int64), pool.len[pp[:, 1]].astype(np.int64)
            out[todo] = self.al.cost_2(pool, pp, deltaw=np.maximum(np.abs(la - lb), 8))
        return out

    def _with_parent(self, pool: S.SeqPool, pairs: np.ndarray):
        """align_2 p c, median_2 s1 s2, max_cost_2 s1 s2 for every (p, c): (median rows, lens, cost, worst)."""
        r = self.al.align_2(pool, pairs, S.WANT_ALIGNED)
        n2 = r.lens[:, 2].astype(np.int32)
        a, b = _left(r.aligned_a, n2), _left(r.aligned_b, n2)
        med, mlen = self.al.median_2(a, b, n2)
        return med, mlen, r.cost.astype(np.int64), self.al.worst_2(a, b, n2)

    def median_3_no_union(self, pool: S.SeqPool, parent, child1, child2) -> Tuple[List[np.ndarray], np.ndarray, np.ndarray]:
        """For every vertex k: (sequence, min cost, max cost) of ``median_3_no_union h p n c1 c2``."""
        parent, child1, child2 = (np.asarray(x, np.int32) for x in (parent, child1, child2))
        n = len(parent)
        pairs = np.concatenate([np.stack([parent, child1], 1), np.stack([parent, child2], 1)])
        med, mlen, cost, worst = self._with_parent(pool, pairs)
        seqs, cmin, cmax = [], np.zeros(n, np.int64), np.zeros(n, np.int64)
        for k in range(n):
            q = k if cost[k] < cost[n + k] else n + k  # `if cost1 < cost2 then res1 else res2`
            s = med[q, med.shape[1] - mlen[q]:]
            if len(s) == 0 or s[0] != self.gap:
                s = np.concatenate([[self.gap], s]).astype(np.uint8)
            seqs.append(s.copy())
            cmin[k], cmax[k] = cost[q], worst[q]
        return seqs, cmin, cmax

    def median_3_union(self, pool: S.SeqPool, parent, aligned_a: List[np.ndarray], aligned_b: List[np.ndarray]):
        """``median_3_union``: aligned_a[k] / aligned_b[k] = the vertex's aligned children (bitset_to_seq of its
        aligned_children, equal lengths)."""
        parent = np.asarray(parent, np.int32)
        n = len(parent)
        unions = [np.bitwise_or(a, b).astype(np.uint8) for a, b in zip(aligned_a, aligned_b)]  # Sequence.Align.union
        seqs = [pool.seq(int(p)) for p in parent] + unions
        sub = S.SeqPool(seqs)
        pairs = np.stack([np.arange(n, dtype=np.int32), np.arange(n, 2 * n, dtype=np.int32)], 1)
        med, mlen, cost, worst = self._with_parent(sub, pairs)
        out = []
        for k in range(n):
            s = med[k, med.shape[1] - mlen[k]:]
            if len(s) == 0 or s[0] != self.gap:
                s = np.concatenate([[self.gap], s]).astype(np.uint8)
            out.append(s.copy())
        return out, cost, worst


def select_one(seq: np.ndarray, cm) -> np.ndarray:
    """``Sequence.select_one`` (src/sequence.ml:1149-1156): for combination alphabets every element collapses to its lowest
    bit; other alphabets are returned unchanged."""
    if not cm.combine():
        return seq
    v = seq.astype(np.int32)
    return (v & -v).astype(np.uint8)


def readjust_3d(dos: "DOS", al3: "S.Align3", pool: S.SeqPool, ch1, ch2, parent, mine):
    """``SeqCS.DOS.readjust (`ThreeD _) h ch1 ch2 parent mine`` (src/seqCS.ml:680-727) for every vertex k, batched: the vertices
    with three non-empty neighbours go through ``Sequence.Align.readjust_3d`` (Powell's aligner, one GPU batch), the ones
    with an empty neighbour through the pairwise median of the other two (one batch per pairing), the rest are copies.
    Returns (changed[n], list of sequences, cost[n])."""
    ch1, ch2, parent, mine = (np.asarray(x, np.int32) for x in (ch1, ch2, parent, mine))
    n = len(mine)
    e1 = np.array([dos._empty(pool, i) for i in ch1])
    e2 = np.array([dos._empty(pool, i) for i in ch2])
    ep = np.array([dos._empty(pool, i) for i in parent])
    res = [None] * n
    cost = np.zeros(n, np.int64)
    full = np.nonzero(~e1 & ~e2 & ~ep)[0]
    if len(full):
        c, seqs, _ = al3.readjust_3d(pool, np.stack([ch1[full], ch2[full], mine[full], parent[full]], axis=1))
        for q, k in enumerate(full):
            res[k], cost[k] = seqs[q], c[q]
    for k in range(n):  # | true, true, _ -> ch1 | true, _, true -> ch1 | _, true, true -> ch2
        if e1[k] and (e2[k] or ep[k]):
            res[k] = pool.seq(int(ch1[k])).copy()
        elif e2[k] and ep[k]:
            res[k] = pool.seq(int(ch2[k])).copy()
    # one empty neighbour: `algn` of the other two (the affine median, or align_2 + median_2), then select_one
    for mask, x, y in ((~e1 & ~e2 & ep, ch1, ch2), (~e1 & e2 & ~ep, ch1, parent), (e1 & ~e2 & ~ep, ch2, parent)):
        idx = np.nonzero(mask)[0]
        if not len(idx):
            continue
        pairs = np.stack([x[idx], y[idx]], axis=1)
        rows, lens = al3.full_median_2(pool, pairs)
        c = al3.cost_2(pool, pairs)
        for q, k in enumerate(idx):
            res[k] = select_one(rows[q, rows.shape[1] - int(lens[q]):].copy(), al3.cm)
            cost[k] = int(c[q])
    changed = np.array([not np.array_equal(res[k], pool.seq(int(mine[k]))) for k in range(n)], bool)
    return changed, res, cost


class Union:
    def __init__(self, al: S.Align):
        self.dos = DOS(al)
        self.sub_factor = 0.8 if al.is_affine else 1.0

    def distance_union(self, pool: S.SeqPool, pairs) -> np.ndarray:
        """``distance_union`` per locus pair (the caller sums over the loci of a vertex)."""
        return self.sub_factor * self.dos.distance(pool, pairs).astype(np.float64)
