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


# ---- engines --------------------------------------------------------------------------------------------------
class GpuEngine:
    """The three batch calls of the driver on :class:`poyd_b200.sequence.Align` (CUDA; no CPU path)."""

    def __init__(self, cm: CostMatrix, device: int = 0):
        self.cm = cm
        self.al = S.Align(cm, device=device)
        self.calls = 0
        self.pairs = 0

    def close(self) -> None:
        self.al.close()

    @staticmethod
    def _pool(store: _Seq[np.ndarray], pairs: np.ndarray):
        used, inv = np.unique(pairs.reshape(-1), return_inverse=True)
        return S.SeqPool([store[i] for i in used]), inv.reshape(-1, 2).astype(np.int32)

    def median(self, store, pairs):
        """``SeqCS.DOS.median`` for every pair: (cost, median)."""
        pool, pp = self._pool(store, pairs)
        r = self.al.align_2(pool, pp, S.WANT_MEDIAN)
        self.calls += 1
        self.pairs += len(pp)
        return r.cost, [r.get("median", p).copy() for p in range(len(pp))]

    def align(self, store, pairs):
        """``Sequence.Align.align_2`` for every pair: the two aligned sequences."""
        pool, pp = self._pool(store, pairs)
        r = self.al.align_2(pool, pp, S.WANT_ALIGNED)
        self.calls += 1
        self.pairs += len(pp)
        return [(r.get("aligned_a", p).copy(), r.get("aligned_b", p).copy()) for p in range(len(pp))]

    def distance(self, store, pairs):
        """``SeqCS.DOS.distance`` (src/seqCS.ml:856-866): ``cost_2 ~deltaw:(max 8 |la - lb|)``."""
        pool, pp = self._pool(store, pairs)
        la, lb = pool.len[pp[:, 0]].astype(np.int64), pool.len[pp[:, 1]].astype(np.int64)
        hint = np.maximum(np.abs(la - lb), 8)
        self.calls += 1
        self.pairs += len(pp)
        return self.al.cost_2(pool, pp, deltaw=hint)


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
            if _is_empty(store[a], self.gap):
                out[k] = (b, 0)
            elif _is_empty(store[b], self.gap):
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
            if _is_empty(s2, gap):
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
        if todo:
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

    def evaluate(self, topo: Topology, leaves: Dict[int, List[np.ndarray]]) -> TreeCost:
        n_loci = len(next(iter(leaves.values())))
        store: List[np.ndarray] = []
        nb0 = getattr(self.e, "calls", 0)
        # directional nodes D[(u, v)] = u looking away from v: (store index per locus, total cost, min_child_code)
        D: Dict[Tuple[int, int], Tuple[List[int], int, int]] = {}
        leaf_entry: Dict[int, Tuple[List[int], int, int]] = {}
        for code, seqs in leaves.items():
            idx = []
            for s in seqs:
                store.append(np.ascontiguousarray(s, np.uint8))
                idx.append(len(store) - 1)
            leaf_entry[code] = (idx, 0, code)

        def nbrs(v):
            return topo.nodes[v]

        # dependency level of every directed pair
        level: Dict[Tuple[int, int], int] = {}
        order: List[Tuple[int, int]] = []
        for u in topo.nodes:
            for v in nbrs(u):
                stack = [(u, v)]
                while stack:
                    a, b = stack[-1]
                    if (a, b) in level:
                        stack.pop()
                        continue
                    if topo.is_leaf(a):
                        level[(a, b)] = 0
                        stack.pop()
                        continue
                    x, y = topo.other_two_nbrs(b, a)
                    need = [(x, a), (y, a)]
                    miss = [k for k in need if k not in level]
                    if miss:
                        stack.extend(miss)
                    else:
                        level[(a, b)] = 1 + max(level[need[0]], level[need[1]])
                        order.append((a, b))
                        stack.pop()
        for (a, b), lv in level.items():
            if lv == 0:
                D[(a, b)] = leaf_entry[a]
        n_medians = 0
        for lv in range(1, max(level.values()) + 1 if level else 1):
            keys = [k for k in order if level[k] == lv]
            jobs, spec = [], []
            for (a, b) in keys:
                x, y = topo.other_two_nbrs(b, a)
                dx, dy = D[(x, a)], D[(y, a)]
                if not dx[2] < dy[2]:  # Node.cs_median: the operand with the smaller min_child_code first
                    dx, dy = dy, dx
                spec.append((dx, dy))
                for l in range(n_loci):
                    jobs.append((dx[0][l], dy[0][l]))
            res = self._medians(store, jobs)
            n_medians += len(jobs)
            for k, (key, (dx, dy)) in enumerate(zip(keys, spec)):
                r = res[k * n_loci:(k + 1) * n_loci]
                D[key] = ([i for i, _ in r], dx[1] + dy[1] + sum(c for _, c in r), min(dx[2], dy[2]))
        # edge medians (refresh_all_edges) and their root costs
        edges = topo.pre_order_edges()
        jobs, spec = [], []
        for (a, b) in edges:
            da, db = D[(a, b)], D[(b, a)]
            if not da[2] < db[2]:
                da, db = db, da
            spec.append((da, db))
            for l in range(n_loci):
                jobs.append((da[0][l], db[0][l]))
        res = self._medians(store, jobs)
        n_medians += len(jobs)
        E: Dict[Tuple[int, int], Tuple[List[int], int]] = {}
        for k, (e, (da, db)) in enumerate(zip(edges, spec)):
            r = res[k * n_loci:(k + 1) * n_loci]
            E[e] = ([i for i, _ in r], da[1] + db[1] + sum(c for _, c in r))
        # general_pick_best_root with blindly_trust_downpass
        h = topo.handle
        root = (h, nbrs(h)[0])  # create_root: the handle and its parent
        best = E[root][1]
        for e in sorted(edges, key=lambda e: (-e[0], -e[1])):
            c = E[e][1]
            if abs(best) > abs(c):
                best, root = c, e
        # assign_single (uppass)
        a, b = root
        singles: Dict[int, List[int]] = {}
        rs = self._closest_batch(store, [(self._nonempty_parent(store, E[root][0][l], D[(a, b)][0][l]), D[(a, b)][0][l])
                                         for l in range(n_loci)])
        frontier = [(b, a, rs), (a, b, rs)]  # (parent vertex, current vertex, parent's singles)
        while frontier:
            jobs = []
            for (p, cur, ps) in frontier:
                mine = D[(cur, p)][0]
                for l in range(n_loci):
                    jobs.append((self._nonempty_parent(store, ps[l], mine[l]), mine[l]))
            res = self._closest_batch(store, jobs)
            nxt = []
            for k, (p, cur, ps) in enumerate(frontier):
                sg = res[k * n_loci:(k + 1) * n_loci]
                singles[cur] = sg
                if not topo.is_leaf(cur):
                    x, y = topo.other_two_nbrs(p, cur)
                    nxt.append((cur, x, sg))
                    nxt.append((cur, y, sg))
            frontier = nxt
        # check_cost: distances between single assignments along the edges, oriented away from the handle
        jobs, zero = [], 0
        for (p, v) in edges:
            for l in range(n_loci):
                s1, s2 = store[singles[p][l]], store[singles[v][l]]
                if _is_empty(s1, self.gap) or _is_empty(s2, self.gap):
                    zero += 1  # missing_distance = 0
                else:
                    jobs.append((singles[p][l], singles[v][l]))
        adjusted = int(self.e.distance(store, np.array(jobs, np.int32)).astype(np.int64).sum()) if jobs else 0
        return TreeCost(adjusted=adjusted, unadjusted=int(best), root=root,
                        singles={v: [store[i] for i in ix] for v, ix in singles.items()},
                        root_costs={e: c for e, (_, c) in E.items()}, batches=getattr(self.e, "calls", 0) - nb0,
                        medians=n_medians, stats={"edges": len(edges), "loci": n_loci, "sequences": len(store)})

    def _nonempty_parent(self, store, parent: int, mine: int) -> int:
        # DOS.to_single (src/seqCS.ml:734-739): an empty parent is replaced by the vertex's own sequence
        return mine if _is_empty(store[parent], self.gap) else parent


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
