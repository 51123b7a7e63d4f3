"""Batched host driver for the tree-level callers of the alignment path.

The reference evaluates a tree by forcing one lazy median at a time (``AllDirChar.internal_downpass``,
src/allDirChar.ml:722-786), every force ending in one ``algn_CAML_*`` call.  Here the same quantities are computed in
**level-order batches**: all directional medians of one dependency level form one call into the CUDA library, then
all edge (root) medians, then the single assignments level by level from the root, then all edge distances.  The
numbers are the reference's:

* tree construction with the reference's vertex codes -- ``Tree.convert_to`` / ``add_tree_to`` (src/tree.ml:724-838);
* the three directional medians of every interior vertex -- ``create_lazy_interior_down/up``
  (src/allDirChar.ml:49-97), operand order by ``min_child_code`` as ``Node.cs_median`` does (src/node.ml:343-348),
  each median = ``SeqCS.DOS.median`` (src/seqCS.ml:747-776);
* edge medians and root selection -- ``refresh_all_edges`` (:672-700), ``create_root`` (:99-125),
  ``general_pick_best_root`` + ``blindly_trust_downpass`` (:787-869): strict improvement over the handle's root,
  edges visited by descending (a, b);
* single assignment -- ``assign_single`` (:283-399) through ``SeqCS.DOS.to_single`` (:730-745) and
  ``Sequence.Align.closest`` (src/sequence.ml:967-1033) with ``Cost_matrix.Two_D.get_closest``
  (src/cost_matrix.ml:681-700);
* the adjusted tree cost -- ``check_cost`` (:179-211): the sum over the edges, oriented away from the handle, of
  ``SeqCS.DOS.distance`` (src/seqCS.ml:819-867) between the single assignments.  This is the number
  ``Ptree.get_cost `Adjusted`` returns and the reference's ``test/cost_tests`` pin (``test/cc*.costs``).

All alignment work goes through an *engine* (three batch calls: median, align, distance).  The product engine is
:class:`GpuEngine` over :class:`poyd_b200.sequence.Align`; there is no CPU engine in this package.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence as _Seq, Tuple

import numpy as np

from .cost_matrix import CostMatrix
from . import sequence as S

# ---- Alphabet.nucleotides (src/alphabet.ml:144-187) ---------------------------------------------------------------
NUCLEOTIDES: Dict[str, int] = {
    "A": 1, "C": 2, "G": 4, "T": 8, "U": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10, "K": 12, "V": 7, "H": 11, "D": 13,
    "B": 14, "N": 15, "X": 15, "-": 16, "1": 17, "2": 18, "3": 19, "4": 20, "5": 21, "6": 22, "7": 23, "8": 24, "9": 25,
    "0": 26, "!": 27, "^": 28, "$": 29, "#": 30, "*": 31, "?": 31,
}
GAP = 16


def read_fasta(path: str) -> List[Tuple[str, List[np.ndarray]]]:
    """``Parser.Fasta.of_file Nucleic_Acids`` (src/parser.ml:220-392) for DNA: taxa in file order, names trimmed
    (``Data.process_taxon_code``, src/data.ml:860-893), fragments split at ``#`` / ``|`` / ``@`` and flattened
    (src/data.ml:1047-1059), gaps removed, a leading gap prepended (``process_sequence``, :272-288)."""
    taxa: List[Tuple[str, List[np.ndarray]]] = []
    name: Optional[str] = None
    chunks: List[str] = []

    def flush():
        if name is None or name == "":
            return
        text = "".join(chunks)
        frags = []
        for part in re.split(r"[#|@]", text):
            codes = [NUCLEOTIDES[ch] for ch in part.upper() if ch != " " and ch != "\t" and ch != "\r"]
            codes = [c for c in codes if c != GAP]
            frags.append(np.array([GAP] + codes, dtype=np.uint8))
        taxa.append((name.strip(), frags))

    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                flush()
                name, chunks = line[1:], []
            elif line.strip():
                chunks.append(line.strip())
    flush()
    return taxa


def parse_trees(text: str) -> List[list]:
    """Parenthetical trees as ``Parser.Tree`` reads them: names separated by blanks or commas, ``[...]`` annotations
    ignored, one tree per top-level parenthesis.  A tree is a nested list of names."""
    text = re.sub(r"\[[^\]]*\]", " ", text)
    toks = re.findall(r"[()]|[^\s(),;*]+", text)
    trees, stack = [], []
    for t in toks:
        if t == "(":
            stack.append([])
        elif t == ")":
            node = stack.pop()
            if stack:
                stack[-1].append(node)
            else:
                trees.append(node)
        else:
            name = t.split(":")[0]
            if stack:
                stack[-1].append(name)
    return trees


# ---- Tree.u_tree as Tree.convert_to builds it (src/tree.ml:724-860) ------------------------------------------------
@dataclass
class Topology:
    """``nodes[id]`` = (parent,) for a leaf, (parent, c1, c2) for an interior vertex -- the ``Leaf`` / ``Interior``
    constructors of src/tree.ml with the neighbour order the reference stores."""

    nodes: Dict[int, Tuple[int, ...]]
    handle: int
    n_taxa: int

    @classmethod
    def convert_to(cls, tree: list, taxon_code: Dict[str, int], n_taxa: Optional[int] = None) -> "Topology":
        total = n_taxa if n_taxa is not None else len(taxon_code)
        avail = iter(range(total + 1, 2 * total + 1))  # add_available, src/tree.ml:841-846
        nodes: Dict[int, Tuple[int, ...]] = {}

        def resolve(children: list) -> list:  # resolve_more_children, :752-773
            children = list(children)
            if len(children) < 2:
                raise ValueError("Illegal Tree file format")
            while len(children) > 2:
                children = [[children[0], children[1]]] + children[2:]
            return children

        def assign(parent: int, t) -> int:  # assign_codes, :735-785
            if isinstance(t, str):
                tc = taxon_code[t]
                nodes[tc] = (parent,)
                return tc
            sc = next(avail)
            a, b = resolve(t)
            ca = assign(sc, a)
            cb = assign(sc, b)
            nodes[sc] = (parent, ca, cb)
            return sc

        if isinstance(tree, list) and len(tree) == 1:
            tree = tree[0]
        if isinstance(tree, str):
            raise ValueError("We need trees with more than two taxa")
        sc = assign(-1, tree)
        _, ca, cb = nodes.pop(sc)
        nodes[ca] = (cb,) + nodes[ca][1:]  # replace_parent
        nodes[cb] = (ca,) + nodes[cb][1:]
        return cls(nodes, ca, total)

    def is_leaf(self, v: int) -> bool:
        return len(self.nodes[v]) == 1

    def other_two_nbrs(self, nbr: int, v: int) -> Tuple[int, int]:  # src/tree.ml:989-1010
        n1, n2, n3 = self.nodes[v]
        if nbr == n1:
            return n2, n3
        if nbr == n2:
            return n1, n3
        assert nbr == n3
        return n1, n2

    def pre_order_edges(self) -> List[Tuple[int, int]]:
        """``Tree.get_pre_order_edges handle`` (src/tree.ml:1356-1423, 1634-1637): every edge once, oriented away
        from the handle.  (The reference flips a coin for the child order; only the orientation is used.)"""
        out: List[Tuple[int, int]] = []
        h = self.handle
        first = self.nodes[h][0]
        stack = [(h, first)]
        if not self.is_leaf(h):
            x, y = self.other_two_nbrs(first, h)
            stack = [(h, y), (h, x), (h, first)]
        while stack:
            pred, v = stack.pop()
            out.append((pred, v))
            if not self.is_leaf(v):
                x, y = self.other_two_nbrs(pred, v)
                stack.append((v, y))
                stack.append((v, x))
        return out


def spr_neighbours(topo: Topology, limit: Optional[int] = None, seed: int = 0) -> List[Topology]:
    """The SPR neighbourhood of a tree (what ``Ptree.single_spr_round`` walks, src/ptree.ml:1141-1168): every subtree is
    pruned with its attachment vertex and regrafted on every other edge.  Vertex codes are kept (the attachment vertex
    moves with the subtree), the handle stays in the part that is not pruned.  ``limit`` draws a seeded sample."""
    def repl(tup, old, new):
        return tuple(new if z == old else z for z in tup)

    out: List[Topology] = []
    moves = []
    for p, nb in topo.nodes.items():
        if len(nb) != 3 or p == topo.handle:
            continue
        for c in nb:
            x, y = [z for z in nb if z != c]
            pruned, stack = {p, c}, [c]
            while stack:
                v = stack.pop()
                for w in topo.nodes[v]:
                    if w not in pruned:
                        pruned.add(w)
                        stack.append(w)
            if topo.handle in pruned:
                continue
            for u in topo.nodes:
                if u in pruned:
                    continue
                for v in topo.nodes[u]:
                    if v in pruned or v < u or {u, v} == {x, y}:
                        continue
                    moves.append((p, c, x, y, u, v))
    if limit is not None and len(moves) > limit:
        rng = np.random.default_rng(seed)
        moves = [moves[i] for i in sorted(rng.choice(len(moves), size=limit, replace=False))]
    for (p, c, x, y, u, v) in moves:
        n2 = dict(topo.nodes)
        n2[x] = repl(n2[x], p, y)
        n2[y] = repl(n2[y], p, x)
        n2[u] = repl(n2[u], v, p)
        n2[v] = repl(n2[v], u, p)
        n2[p] = (u, v, c)
        out.append(Topology(n2, topo.handle, topo.n_taxa))
    return out


# ---- engines --------------------------------------------------------------------------------------------------
class GpuEngine:
    """The three batch calls of the driver on :class:`poyd_b200.sequence.Align` (CUDA; no CPU path)."""

    def __init__(self, cm: CostMatrix, device: int = 0):
        self.cm = cm
        self.al = S.Align(cm, device=device)
        self.calls = 0
        self.pairs = 0
        self.log: List[Tuple[str, np.ndarray, np.ndarray, Optional[np.ndarray]]] = []  # (kind, len a, len b, deltaw)

    def close(self) -> None:
        self.al.close()

    def _note(self, kind: str, pool, pp, dw=None) -> None:
        self.calls += 1
        self.pairs += len(pp)
        self.log.append((kind, pool.len[pp[:, 0]].copy(), pool.len[pp[:, 1]].copy(), None if dw is None else np.asarray(dw).copy()))

    @staticmethod
    def _pool(store: _Seq[np.ndarray], pairs: np.ndarray):
        used, inv = np.unique(pairs.reshape(-1), return_inverse=True)
        return S.SeqPool([store[i] for i in used]), inv.reshape(-1, 2).astype(np.int32)

    def median(self, store, pairs):
        """``SeqCS.DOS.median`` for every pair: (cost, median)."""
        pool, pp = self._pool(store, pairs)
        r = self.al.align_2(pool, pp, S.WANT_MEDIAN)
        self._note("median", pool, pp, None if self.al.is_affine else self.al.deltaw_for(pool, pp))
        return r.cost, [r.get("median", p).copy() for p in range(len(pp))]

    def align(self, store, pairs):
        """``Sequence.Align.align_2`` for every pair: the two aligned sequences."""
        pool, pp = self._pool(store, pairs)
        r = self.al.align_2(pool, pp, S.WANT_ALIGNED)
        self._note("align", pool, pp, None if self.al.is_affine else self.al.deltaw_for(pool, pp))
        return [(r.get("aligned_a", p).copy(), r.get("aligned_b", p).copy()) for p in range(len(pp))]

    def closest(self, store, pairs):
        """``Sequence.Align.closest parent mine`` for every pair, fused into the traceback kernel."""
        pool, pp = self._pool(store, pairs)
        self._note("closest", pool, pp, None if self.al.is_affine else self.al.deltaw_for(pool, pp))
        return self.al.closest(pool, pp, prechecked=True)

    def distance(self, store, pairs):
        """``SeqCS.DOS.distance`` (src/seqCS.ml:856-866): ``cost_2 ~deltaw:(max 8 |la - lb|)``."""
        pool, pp = self._pool(store, pairs)
        la, lb = pool.len[pp[:, 0]].astype(np.int64), pool.len[pp[:, 1]].astype(np.int64)
        hint = np.maximum(np.abs(la - lb), 8)
        self._note("distance", pool, pp, None if self.al.is_affine else self.al.deltaw_for(pool, pp, hint))
        return self.al.cost_2(pool, pp, deltaw=hint)


def logged_cells(log, affine: bool) -> int:
    """DP cells the reference visits for the calls an engine logged, counted after the fact."""
    total = 0
    for _, la, lb, dw in log:
        la, lb = la.astype(np.int64), lb.astype(np.int64)
        if affine:
            key = la * 65536 + lb
            u, cnt = np.unique(key, return_counts=True)
            total += sum(S.cells_affine(int(k) >> 16, int(k) & 65535) * int(c) for k, c in zip(u, cnt))
        else:
            l1, l2 = np.maximum(la, lb), np.minimum(la, lb)
            key = (l1 * 65536 + l2) * 65536 + dw.astype(np.int64)
            u, cnt = np.unique(key, return_counts=True)
            total += sum(S.cells_linear(int(k) >> 32, (int(k) >> 16) & 65535, int(k) & 65535) * int(c) for k, c in zip(u, cnt))
    return int(total)


# ---- Cost_matrix.Two_D.get_closest as a table ------------------------------------------------------------------------
def closest_table(cm: CostMatrix) -> np.ndarray:
    """``T[a, b] = Cost_matrix.Two_D.get_closest cm a b`` (src/cost_matrix.ml:681-700) for combination alphabets:
    ``b`` loses its gap bit (or collapses to the gap when both carry it), then the lowest single bit of ``b`` with
    the strictly smallest ``cost a x`` wins."""
    n = 1 << cm.lcm
    gap = cm.gap
    T = np.zeros((n, n), np.uint8)
    for a in range(1, n):
        for b0 in range(1, n):
            b = b0
            if a == gap or b == gap:
                pass
            elif (a & gap) and (b & gap):
                b = gap
            else:
                b &= ~gap
            best, cur = a, None
            for bit in range(cm.lcm):  # list_of_bits b a_sz, ascending after List.rev; b < 2^lcm
                x = 1 << bit
                if b & x:
                    nc = int(cm.cost[a, x])
                    if cur is None or nc < cur:
                        best, cur = x, nc
            T[a, b0] = best
    return T


# ---- the evaluation ---------------------------------------------------------------------------------------------
@dataclass
class TreeCost:
    adjusted: int                     # Ptree.get_cost `Adjusted
    unadjusted: int                   # component_cost: root_cost of the chosen root median
    root: Tuple[int, int]
    singles: Dict[int, List[np.ndarray]]
    root_costs: Dict[Tuple[int, int], int]
    batches: int = 0
    medians: int = 0
    stats: Dict[str, int] = field(default_factory=dict)


def _is_empty(s: np.ndarray, gap: int) -> bool:
    return bool(np.all(s == gap))  # Sequence.is_empty, src/sequence.ml:442-450


class Evaluator:
    """Downpass + uppass of one tree over ``n_loci`` independent sequence characters."""

    def __init__(self, engine, cm: CostMatrix):
        self.e = engine
        self.cm = cm
        self.gap = cm.gap
        self._closest = closest_table(cm) if cm.combine() else None

    # -- SeqCS.DOS.median over a batch, with the empty-operand rule (src/seqCS.ml:748-752)
    def _medians(self, store: List[np.ndarray], jobs: List[Tuple[int, int]]) -> List[Tuple[int, int]]:
        """jobs: (a, b) store indices.  Returns (store index of the median, cost) per job; new medians are appended."""
        out: List[Optional[Tuple[int, int]]] = [None] * len(jobs)
        todo, where = [], []
        for k, (a, b) in enumerate(jobs):
            if self._emp(a):
                out[k] = (b, 0)
            elif self._emp(b):
                out[k] = (a, 0)
            else:
                todo.append((a, b))
                where.append(k)
        if todo:
            cost, med = self.e.median(store, np.array(todo, np.int32))
            for k, c, m in zip(where, cost, med):
                store.append(m)
                out[k] = (len(store) - 1, int(c))
        return out  # type: ignore[return-value]

    # -- Sequence.Align.closest over a batch (src/sequence.ml:967-1033); the cost it returns is not used by check_cost
    def _closest_batch(self, store: List[np.ndarray], jobs: List[Tuple[int, int]]) -> List[int]:
        gap = self.gap
        out: List[Optional[int]] = [None] * len(jobs)
        todo, where = [], []
        pre: Dict[int, Tuple[np.ndarray, np.ndarray]] = {}
        for k, (p, m) in enumerate(jobs):
            s1, s2 = store[p], store[m]
            if self._emp(m):
                out[k] = m
            elif self._closest is not None and len(s1) == len(s2) and np.array_equal(s1, s2):
                mask = np.uint8(~gap & 0xFF)
                a, b = s1.copy(), s2.copy()
                a[1:] &= mask
                b[1:] &= mask
                pre[k] = (a, b)
            else:
                todo.append((p, m))
                where.append(k)
        if todo and hasattr(self.e, "closest"):  # the fused kernel (GpuEngine)
            for k, res in zip(where, self.e.closest(store, np.array(todo, np.int32))):
                store.append(res)
                out[k] = len(store) - 1
        elif todo:  # engines that only align: the column rule is applied here
            al = self.e.align(store, np.array(todo, np.int32))
            for k, ab in zip(where, al):
                pre[k] = ab
        for k, (a1, b1) in pre.items():
            if self._closest is not None:
                sel = self._closest[a1, b1]
            else:  # no combinations: sequence.ml:986-997
                allv = self.cm.all_elements
                sel = np.where(b1 == allv, np.where(a1 == allv, 1, a1), b1).astype(np.uint8)
            res = np.concatenate([[gap], sel[sel != gap]]).astype(np.uint8)  # remove_gaps + prepend gap
            store.append(res)
            out[k] = len(store) - 1
        return out  # type: ignore[return-value]

    # -- persistent state: every sequence ever produced, and the medians keyed by WHAT they are medians of.  A
    # directional node depends only on the rooted subtree behind it, so subtrees are interned (hash-consed) and their
    # medians survive from one tree to the next -- a Wagner step or an SPR move recomputes only what it changed.
    def _reset(self) -> None:
        self.store: List[np.ndarray] = []
        self._sig: Dict[Tuple[int, int], int] = {}          # (sig x, sig y) ordered -> sig id of the median node
        self._node: Dict[int, Tuple[List[int], int, int]] = {}  # sig id -> (store index per locus, total cost, min_child_code)
        self._edge: Dict[Tuple[int, int], Optional[Tuple[List[int], int]]] = {}
        self._leafsig: Dict[int, int] = {}
        self.n_medians = 0
        self._next_sig = 0
        self._empty: Dict[int, bool] = {}
        self._clo: Dict[Tuple[int, int], int] = {}    # closest by (parent, mine) store index
        self._dist: Dict[Tuple[int, int], int] = {}   # DOS.distance by store index pair

    def _emp(self, i: int) -> bool:
        """Sequence.is_empty of store[i], remembered (the store only grows)."""
        e = self._empty.get(i)
        if e is None:
            e = self._empty[i] = _is_empty(self.store[i], self.gap)
        return e

    def _new_sig(self) -> int:
        self._next_sig += 1
        return self._next_sig - 1

    def _leaf(self, code: int, seqs: List[np.ndarray]) -> int:
        sig = self._leafsig.get(code)
        if sig is None:
            idx = []
            for q in seqs:
                self.store.append(np.ascontiguousarray(q, np.uint8))
                idx.append(len(self.store) - 1)
            sig = self._leafsig[code] = self._new_sig()
            self._node[sig] = (idx, 0, code)
        return sig

    def _collect(self, topo: Topology, leaves: Dict[int, List[np.ndarray]]) -> Dict[Tuple[int, int], int]:
        """Signature of every directional node ``(u, v)`` = u looking away from v (``AllDirNode.not_with v``).  Medians
        that are not in the cache yet are only *registered* (``self._fresh``); :meth:`_flush` computes them."""
        sig: Dict[Tuple[int, int], int] = {}
        for u in topo.nodes:
            for v in topo.nodes[u]:
                stack = [(u, v)]
                while stack:
                    a, b = stack[-1]
                    if (a, b) in sig:
                        stack.pop()
                        continue
                    if topo.is_leaf(a):
                        sig[(a, b)] = self._leaf(a, leaves[a])
                        stack.pop()
                        continue
                    x, y = topo.other_two_nbrs(b, a)
                    miss = [k for k in ((x, a), (y, a)) if k not in sig]
                    if miss:
                        stack.extend(miss)
                        continue
                    sx, sy = sig[(x, a)], sig[(y, a)]
                    # Node.cs_median: the operand with the smaller min_child_code first (src/node.ml:343-348)
                    if not self._minc(sx) < self._minc(sy):
                        sx, sy = sy, sx
                    s_ = self._sig.get((sx, sy))
                    if s_ is None:
                        s_ = self._sig[(sx, sy)] = self._new_sig()
                        lv = 1 + max(self._fresh.get(sx, (0, 0, 0, 0))[2], self._fresh.get(sy, (0, 0, 0, 0))[2])
                        self._fresh[s_] = (sx, sy, lv, min(self._minc(sx), self._minc(sy)))
                    sig[(a, b)] = s_
                    stack.pop()
        return sig

    def _flush(self, n_loci: int) -> None:
        """Computes the registered medians level by level: one batch per dependency level, whatever number of trees
        registered them."""
        if not self._fresh:
            return
        top = max(v[2] for v in self._fresh.values())
        by_level: List[List[int]] = [[] for _ in range(top + 1)]
        for k, v in self._fresh.items():
            by_level[v[2]].append(k)
        for lv in range(1, top + 1):
            keys = by_level[lv]
            jobs = []
            for k in keys:
                sx, sy = self._fresh[k][0], self._fresh[k][1]
                for l in range(n_loci):
                    jobs.append((self._node[sx][0][l], self._node[sy][0][l]))
            res = self._medians(self.store, jobs)
            self.n_medians += len(jobs)
            for i, k in enumerate(keys):
                sx, sy = self._fresh[k][0], self._fresh[k][1]
                r = res[i * n_loci:(i + 1) * n_loci]
                nx, ny = self._node[sx], self._node[sy]
                self._node[k] = ([j for j, _ in r], nx[1] + ny[1] + sum(c for _, c in r), min(nx[2], ny[2]))
        self._fresh.clear()

    def directional(self, topo: Topology, leaves: Dict[int, List[np.ndarray]]) -> Dict[Tuple[int, int], int]:
        """Signatures of all directional nodes of one tree, their medians computed."""
        sig = self._collect(topo, leaves)
        self._flush(len(next(iter(leaves.values()))))
        return sig

    def _minc(self, s_: int) -> int:
        return self._node[s_][2] if s_ in self._node else self._fresh[s_][3]

    def edge_medians(self, topo: Topology, sig: Dict[Tuple[int, int], int], edges: List[Tuple[int, int]],
                     n_loci: int) -> Dict[Tuple[int, int], Tuple[List[int], int]]:
        """``refresh_all_edges`` (src/allDirChar.ml:672-700): the median across every edge and its root cost."""
        return self._edge_medians_many([(sig, edges)], n_loci)[0]

    def _edge_medians_many(self, trees, n_loci: int):
        jobs, todo, keys = [], [], []
        for sig, edges in trees:
            ks = []
            for (a, b) in edges:
                sa, sb = sig[(a, b)], sig[(b, a)]
                if not self._node[sa][2] < self._node[sb][2]:
                    sa, sb = sb, sa
                ks.append((sa, sb))
                if (sa, sb) not in self._edge:
                    self._edge[(sa, sb)] = None  # claimed: computed once for all the trees that share the edge
                    todo.append((sa, sb))
                    for l in range(n_loci):
                        jobs.append((self._node[sa][0][l], self._node[sb][0][l]))
            keys.append(ks)
        res = self._medians(self.store, jobs)
        self.n_medians += len(jobs)
        for i, (sa, sb) in enumerate(todo):
            r = res[i * n_loci:(i + 1) * n_loci]
            self._edge[(sa, sb)] = ([j for j, _ in r], self._node[sa][1] + self._node[sb][1] + sum(c for _, c in r))
        return [{e: self._edge[k] for e, k in zip(edges, ks)} for (_, edges), ks in zip(trees, keys)]

    def _closest_cached(self, jobs: List[Tuple[int, int]]) -> List[int]:
        """``closest parent mine`` by store index, each distinct (parent, mine) computed once."""
        new = []
        for j in jobs:
            if j not in self._clo:
                self._clo[j] = -1
                new.append(j)
        if new:
            for j, r in zip(new, self._closest_batch(self.store, new)):
                self._clo[j] = r
        return [self._clo[j] for j in jobs]

    def evaluate(self, topo: Topology, leaves: Dict[int, List[np.ndarray]], keep: bool = False) -> TreeCost:
        """Downpass + uppass of one tree.  ``keep=True`` keeps the median cache for the next tree over the same leaves."""
        return self.evaluate_many([topo], leaves, keep=keep)[0]

    def evaluate_many(self, topos: List[Topology], leaves: Dict[int, List[np.ndarray]], keep: bool = False) -> List[TreeCost]:
        """Downpass + uppass of several trees over the same leaves -- an SPR/TBR neighbourhood, the candidates of an
        exhaustive join -- in lockstep: the batches of every phase hold the work of ALL trees, and whatever the trees
        share (subtree medians, edge medians, single assignments below a common root, edge distances) is computed once."""
        if not keep or not hasattr(self, "store"):
            self._reset()
        self._fresh = {}
        n_loci = len(next(iter(leaves.values())))
        store = self.store
        nb0, nm0 = getattr(self.e, "calls", 0), self.n_medians
        sigs = [self._collect(t, leaves) for t in topos]
        self._flush(n_loci)
        edges_all = [t.pre_order_edges() for t in topos]
        Es = self._edge_medians_many(list(zip(sigs, edges_all)), n_loci)
        # general_pick_best_root with blindly_trust_downpass, tree by tree
        roots, bests = [], []
        for topo, edges, E in zip(topos, edges_all, Es):
            h = topo.handle
            root = (h, topo.nodes[h][0])  # create_root: the handle and its parent
            best = E[root][1]
            for e in sorted(edges, key=lambda e: (-e[0], -e[1])):
                c = E[e][1]
                if abs(best) > abs(c):
                    best, root = c, e
            roots.append(root)
            bests.append(best)
        # assign_single (uppass): all trees advance one depth per batch
        D = [{k: self._node[v][0] for k, v in sig.items()} for sig in sigs]
        singles: List[Dict[int, List[int]]] = [dict() for _ in topos]
        jobs = []
        for ti, (root, E) in enumerate(zip(roots, Es)):
            a, b = root
            for l in range(n_loci):
                mine = D[ti][(a, b)][l]
                jobs.append((self._nonempty_parent(store, E[root][0][l], mine), mine))
        rs = self._closest_cached(jobs)
        frontier = []  # (tree, parent vertex, current vertex, parent's singles)
        for ti, (a, b) in enumerate(roots):
            r = rs[ti * n_loci:(ti + 1) * n_loci]
            frontier += [(ti, b, a, r), (ti, a, b, r)]
        while frontier:
            jobs = []
            for (ti, p, cur, ps) in frontier:
                mine = D[ti][(cur, p)]
                for l in range(n_loci):
                    jobs.append((self._nonempty_parent(store, ps[l], mine[l]), mine[l]))
            res = self._closest_cached(jobs)
            nxt = []
            for k, (ti, p, cur, ps) in enumerate(frontier):
                sg = res[k * n_loci:(k + 1) * n_loci]
                singles[ti][cur] = sg
                if not topos[ti].is_leaf(cur):
                    x, y = topos[ti].other_two_nbrs(p, cur)
                    nxt.append((ti, cur, x, sg))
                    nxt.append((ti, cur, y, sg))
            frontier = nxt
        # check_cost: distances between single assignments along the edges, oriented away from the handle
        new = []
        for ti, edges in enumerate(edges_all):
            for (p, v) in edges:
                for l in range(n_loci):
                    j = (singles[ti][p][l], singles[ti][v][l])
                    if j not in self._dist:
                        if self._emp(j[0]) or self._emp(j[1]):
                            self._dist[j] = 0  # missing_distance
                        else:
                            self._dist[j] = -1
                            new.append(j)
        if new:
            for j, c in zip(new, self.e.distance(store, np.array(new, np.int32))):
                self._dist[j] = int(c)
        out = []
        for ti, (topo, edges, E) in enumerate(zip(topos, edges_all, Es)):
            adjusted = sum(self._dist[(singles[ti][p][l], singles[ti][v][l])] for (p, v) in edges for l in range(n_loci))
            out.append(TreeCost(adjusted=int(adjusted), unadjusted=int(bests[ti]), root=roots[ti],
                                singles={v: [store[i] for i in ix] for v, ix in singles[ti].items()},
                                root_costs={e: c for e, (_, c) in E.items()}, batches=getattr(self.e, "calls", 0) - nb0,
                                medians=self.n_medians - nm0,
                                stats={"edges": len(edges), "loci": n_loci, "sequences": len(store), "trees": len(topos)}))
        return out

    # -- Wagner build with a batched candidate-edge sweep (Ptree.make_wagner_tree, src/ptree.ml:948-1060) --------------
    def wagner(self, leaves: Dict[int, List[np.ndarray]], order: Optional[List[int]] = None):
        """Adds the taxa one at a time.  For each, ``AllDirChar.cost_fn`` (src/allDirChar.ml:1279-1317) --
        ``Node.distance clade (median across the edge)``, i.e. ``DOS.distance`` -- is evaluated on EVERY edge of the
        current tree in one batch (the Wagner manager really visits all edges, src/queues.ml:367-407), the taxon joins
        the edge of smallest cost, and only the medians the join invalidated are recomputed.  Deterministic variant:
        taxa in the given order, ties to the first edge in pre-order (the reference randomises both).
        Returns (topology, per-step records)."""
        self._reset()
        self._fresh = {}
        order = list(order) if order is not None else sorted(leaves)
        n_loci = len(next(iter(leaves.values())))
        t1, t2 = order[0], order[1]
        nodes: Dict[int, Tuple[int, ...]] = {t1: (t2,), t2: (t1,)}
        topo = Topology(nodes, t1, len(leaves))
        next_id = max(leaves) + 1
        steps = []
        for c in order[2:]:
            sig = self.directional(topo, leaves)
            edges = topo.pre_order_edges()
            E = self.edge_medians(topo, sig, edges, n_loci)
            cl = self._node[self._leaf(c, leaves[c])][0]
            jobs, owner = [], []
            for k, e in enumerate(edges):
                for l in range(n_loci):
                    m = E[e][0][l]
                    if not (self._emp(cl[l]) or self._emp(m)):
                        jobs.append((cl[l], m))
                        owner.append(k)
            delta = np.zeros(len(edges), np.int64)
            if jobs:
                np.add.at(delta, np.array(owner), self.e.distance(self.store, np.array(jobs, np.int32)).astype(np.int64))
            k = int(np.argmin(delta))  # first minimum
            a, b = edges[k]
            v = next_id
            next_id += 1
            nodes[v] = (a, b, c)
            nodes[a] = tuple(v if x == b else x for x in nodes[a])
            nodes[b] = tuple(v if x == a else x for x in nodes[b])
            nodes[c] = (v,)
            steps.append({"taxon": c, "edge": (a, b), "delta": int(delta[k]), "edges": len(edges)})
        return topo, steps

    def _nonempty_parent(self, store, parent: int, mine: int) -> int:
        # DOS.to_single (src/seqCS.ml:734-739): an empty parent is replaced by the vertex's own sequence
        return mine if self._emp(parent) else parent


def tree_cost(engine, cm: CostMatrix, fasta: str, tree_file: str, which: int = 0) -> TreeCost:
    """What ``read (fasta)  read (tree)`` leaves as ``Ptree.get_cost `Adjusted`` (src/poy_test.ml:43)."""
    taxa = read_fasta(fasta)
    codes = {name: i + 1 for i, (name, _) in enumerate(taxa)}
    n_loci = max(len(fr) for _, fr in taxa)
    leaves = {}
    for name, fr in taxa:  # a taxon short of fragments gets empty ones (src/data.ml:1092-1106)
        leaves[codes[name]] = fr + [np.array([GAP], np.uint8)] * (n_loci - len(fr))
    with open(tree_file) as f:
        tree = parse_trees(f.read())[which]
    topo = Topology.convert_to(tree, codes)
    return Evaluator(engine, cm).evaluate(topo, leaves)
