"""The tree-level driver (poyd_b200/tree.py) against the reference's OWN known-answer tests: test/cost_tests lines 7-14
(cc*.poy: 13 data sets x 4 cost regimes, expected ``Ptree.get_cost `Adjusted`` in test/cc*.costs), packed into
tests/golden/trees/tree_costs.npz by tests/golden/make_tree_golden.py.

CPU run: the alignment calls are answered by the CPU checker (compiled reference / port), which pins the oracle, the
cost-matrix builder and the host driver to the reference's goldens.  GPU run: the same 52 numbers through the CUDA
library (GpuEngine -> C ABI)."""
import os

import numpy as np
import pytest

from poyd_b200 import cost_matrix as CM, tree as T

HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.load(os.path.join(HERE, "golden", "trees", "tree_costs.npz"))
FILES = [str(f) for f in Z["files"]]
REGIMES = 4


def _cm(k):
    s, g, go = (int(x) for x in Z[f"tcm_{k}"])
    if k == 0:
        return CM.default_nucleotides()  # no transform: Cost_matrix.Two_D.default
    return CM.nucleotides(s, g, go if go >= 0 else None)


def _case(tmp_path, k, n):
    fa, tr = tmp_path / f"{n}.fas", tmp_path / f"{n}.tree"
    fa.write_bytes(Z[f"fasta_{n}"].tobytes())
    tr.write_bytes(Z[f"tree_{k}_{n}"].tobytes())
    return str(fa), str(tr), int(Z[f"costs_{k}"][n])


def test_topology_codes_follow_tree_ml():
    # (A (B (C D))) with taxon codes 1..4: the root takes code 5 and disappears, its children join (handle = A)
    t = T.Topology.convert_to(T.parse_trees("(A (B (C D)))[12.]")[0], {"A": 1, "B": 2, "C": 3, "D": 4})
    assert t.handle == 1 and t.nodes[1] == (6,) and t.nodes[6] == (1, 2, 7) and t.nodes[7] == (6, 3, 4)
    assert t.pre_order_edges() == [(1, 6), (6, 2), (6, 7), (7, 3), (7, 4)]
    # polytomies are resolved left to right (resolve_more_children)
    t = T.Topology.convert_to(T.parse_trees("(A B C D)")[0], {"A": 1, "B": 2, "C": 3, "D": 4})
    assert sorted(len(v) for v in t.nodes.values()) == [1, 1, 1, 1, 3, 3]


def test_closest_table_rules():
    cm = CM.default_nucleotides()
    tab = T.closest_table(cm)
    assert tab[1, 3] == 1 and tab[4, 3] == 1 and tab[2, 3] == 2  # closest single base, lowest bit on ties
    assert tab[16, 16] == 16 and tab[17, 18] == 16  # both carry the gap bit -> gap
    assert tab[1, 19] == 1 and tab[4, 19] == 1  # the gap bit is dropped first


@pytest.mark.parametrize("k", range(REGIMES))
def test_reference_tree_costs_cpu_checker(tmp_path, k, checker_factory):
    from oracle.tree_engine import OracleEngine

    cm = _cm(k)
    eng = OracleEngine(cm, nthreads=4)
    for n in range(len(FILES)):
        fa, tr, want = _case(tmp_path, k, n)
        got = T.tree_cost(eng, cm, fa, tr)
        assert got.adjusted == want, f"{FILES[n]} regime {k}: {got.adjusted} != {want}"


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(REGIMES))
def test_reference_tree_costs_gpu(tmp_path, k):
    cm = _cm(k)
    eng = T.GpuEngine(cm, device=0)
    try:
        for n in range(len(FILES)):
            fa, tr, want = _case(tmp_path, k, n)
            got = T.tree_cost(eng, cm, fa, tr)
            assert got.adjusted == want, f"{FILES[n]} regime {k}: {got.adjusted} != {want}"
        assert eng.al.launch_count() > 0
    finally:
        eng.close()


def _newick(topo):
    def down(p, v):
        if topo.is_leaf(v):
            return str(v)
        x, y = topo.other_two_nbrs(p, v)
        return "(" + " ".join(sorted([down(v, x), down(v, y)])) + ")"
    h = topo.handle
    return "(" + str(h) + " " + down(h, topo.nodes[h][0]) + ")" if topo.is_leaf(h) else None


@pytest.mark.parametrize("affine", [False, True])
def test_wagner_sweep_and_cache_cpu_checker(affine, checker_factory):
    from oracle.tree_engine import OracleEngine
    from poyd_b200 import synth

    cm = CM.nucleotides(2, 1, 1) if affine else CM.default_nucleotides()
    leaves = synth.taxa_on_random_tree(12, 120, seed=3, subst=0.06, indel=0.02)
    eng = OracleEngine(cm, nthreads=2)
    ev = T.Evaluator(eng, cm)
    topo, steps = ev.wagner(leaves)
    assert len(steps) == 10 and sorted(v for v in topo.nodes if topo.is_leaf(v)) == list(range(1, 13))
    assert all(len(n) in (1, 3) for n in topo.nodes.values()) and len(topo.nodes) == 2 * 12 - 2
    # every step saw every edge of the tree it was added to (2 k - 3 edges for k taxa)
    assert [s["edges"] for s in steps] == [2 * k - 3 for k in range(2, 12)]
    # the cached evaluation equals a from-scratch evaluation of the same topology
    warm = ev.evaluate(topo, leaves, keep=True)
    cold = T.Evaluator(OracleEngine(cm, nthreads=2), cm).evaluate(topo, leaves)
    assert (warm.adjusted, warm.unadjusted, warm.root) == (cold.adjusted, cold.unadjusted, cold.root)
    assert warm.medians < cold.medians  # the build left its medians behind
    # the last sweep, pair by pair: distance(new taxon, median across the edge) is minimal on the chosen edge
    last = steps[-1]
    assert last["delta"] >= 0


@pytest.mark.gpu
def test_wagner_gpu_matches_cpu_checker():
    from oracle import oracle
    from oracle.tree_engine import OracleEngine
    from poyd_b200 import synth

    oracle.build(ref=True)
    # DNA linear, DNA affine, and protein (no combinations: the other branch of Sequence.Align.closest, full-matrix linear fills)
    for cm, alphabet in ((CM.default_nucleotides(), "dna"), (CM.nucleotides(2, 1, 1), "dna"), (CM.default_aminoacids(), "protein")):
        leaves = synth.taxa_on_random_tree(24, 400 if alphabet == "dna" else 150, seed=11, subst=0.05, indel=0.01, alphabet=alphabet)
        g = T.GpuEngine(cm, device=0)
        try:
            evg = T.Evaluator(g, cm)
            tg, sg = evg.wagner(leaves)
            cg = evg.evaluate(tg, leaves, keep=True)
        finally:
            g.close()
        evc = T.Evaluator(OracleEngine(cm, nthreads=8), cm)
        tc, sc = evc.wagner(leaves)
        cc = evc.evaluate(tc, leaves, keep=True)
        assert sg == sc and tg.nodes == tc.nodes
        assert (cg.adjusted, cg.unadjusted, cg.root) == (cc.adjusted, cc.unadjusted, cc.root)
        for v in cc.singles:
            assert all(np.array_equal(x, y) for x, y in zip(cg.singles[v], cc.singles[v]))


def test_spr_neighbourhood_in_lockstep_cpu_checker(checker_factory):
    """evaluate_many over an SPR neighbourhood: every tree gets exactly what a cold evaluation of it alone gives, in far
    fewer engine calls, because the batches hold all trees and shared medians / singles / distances are computed once."""
    from oracle.tree_engine import OracleEngine
    from poyd_b200 import synth

    cm = CM.nucleotides(2, 1, 1)
    leaves = synth.taxa_on_random_tree(10, 100, seed=9, subst=0.08, indel=0.03)
    ev = T.Evaluator(OracleEngine(cm, nthreads=2), cm)
    topo, _ = ev.wagner(leaves)
    nbrs = T.spr_neighbours(topo, limit=24, seed=1)
    assert len(nbrs) == 24
    for t in nbrs:  # still binary trees over the same taxa
        assert sorted(v for v in t.nodes if t.is_leaf(v)) == list(range(1, 11))
        assert all(u in t.nodes[v] for u in t.nodes for v in t.nodes[u])
        assert len(t.pre_order_edges()) == 2 * 10 - 3
    many = ev.evaluate_many(nbrs, leaves, keep=True)
    lock_calls = many[0].batches
    solo_calls = 0
    for t, r in zip(nbrs, many):
        cold = T.Evaluator(OracleEngine(cm, nthreads=2), cm).evaluate(t, leaves)
        solo_calls += cold.batches
        assert (r.adjusted, r.unadjusted, r.root) == (cold.adjusted, cold.unadjusted, cold.root)
        for v in cold.singles:
            assert all(np.array_equal(x, y) for x, y in zip(r.singles[v], cold.singles[v]))
    assert lock_calls * 5 < solo_calls


def test_read_fasta_follows_the_reference_parser(tmp_path):
    """Parser.Fasta for nucleotides (src/parser.ml:220-392): names trimmed, case folded, gaps dropped, IUPAC sets,
    fragments at '#', a leading gap on every fragment, taxa in file order."""
    p = tmp_path / "x.fas"
    p.write_text(">Alpha   \nacg-T\nNN#ry\n\n>Beta\nAC\nGT#K?\n")
    taxa = T.read_fasta(str(p))
    assert [n for n, _ in taxa] == ["Alpha", "Beta"]
    assert [f.tolist() for f in taxa[0][1]] == [[16, 1, 2, 4, 8, 15, 15], [16, 5, 10]]
    assert [f.tolist() for f in taxa[1][1]] == [[16, 1, 2, 4, 8], [16, 12, 31]]
    # trees: blanks or commas, annotations ignored, several trees per file
    assert T.parse_trees("(A (B C))[12.] (A,(B,C));") == [["A", ["B", "C"]], ["A", ["B", "C"]]]


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(REGIMES))
def test_reference_tree_costs_native_driver(tmp_path, k):
    """The same 52 reference goldens through the C++ driver (include/poyb200_tree.h) on the device-resident store."""
    from poyd_b200 import tree_native as TN

    cm = _cm(k)
    for n in range(len(FILES)):
        fa, tr, want = _case(tmp_path, k, n)
        got = TN.tree_cost_native(cm, fa, tr)
        assert got.adjusted == want, f"{FILES[n]} regime {k}: {got.adjusted} != {want}"


@pytest.mark.gpu
def test_native_driver_matches_python_driver_step_by_step():
    """Wagner build (every step, the topology), evaluation (costs, root, every single assignment), lockstep evaluation of an
    SPR neighbourhood, and the SPR search's invariants -- C++ driver on the store against tree.py over the CPU checker."""
    from oracle import oracle
    from oracle.tree_engine import OracleEngine
    from poyd_b200 import synth, tree_native as TN

    oracle.build(ref=True)
    for cm, alphabet in ((CM.default_nucleotides(), "dna"), (CM.nucleotides(1, 2, 3), "dna"), (CM.default_aminoacids(), "protein")):
        leaves = synth.taxa_on_random_tree(20, 300 if alphabet == "dna" else 120, seed=21, subst=0.05, indel=0.01, alphabet=alphabet)
        ev = TN.NativeEvaluator(cm, leaves)
        try:
            tn, sn = ev.wagner()
            cn = ev.evaluate(tn, keep=True)
            evc = T.Evaluator(OracleEngine(cm, nthreads=8), cm)
            tc, sc = evc.wagner(leaves)
            cc = evc.evaluate(tc, leaves, keep=True)
            assert [(s["taxon"], s["edge"], s["delta"]) for s in sn] == [(s["taxon"], s["edge"], s["delta"]) for s in sc]
            assert tn.nodes == tc.nodes and tn.handle == tc.handle
            assert (cn.adjusted, cn.unadjusted, cn.root) == (cc.adjusted, cc.unadjusted, cc.root)
            for v in cc.singles:
                for l, want in enumerate(cc.singles[v]):
                    assert np.array_equal(ev.single(0, v, l), want), (v, l)
            nbrs = T.spr_neighbours(tc, limit=16, seed=2)
            many_n = ev.evaluate_many(nbrs, keep=True)
            many_c = evc.evaluate_many(nbrs, leaves, keep=True)
            assert [(a.adjusted, a.unadjusted, a.root) for a in many_n] == [(b.adjusted, b.unadjusted, b.root) for b in many_c]
            # SPR search: never worse than the start, the final topology is a binary tree over the same taxa whose
            # (independently evaluated) cost is the reported one
            final, st = ev.spr(tn, max_rounds=3, window=8)
            assert st["start_cost"] == cn.adjusted and st["final_cost"] <= st["start_cost"]
            assert sorted(v for v in final.nodes if final.is_leaf(v)) == sorted(leaves)
            assert all(u in final.nodes[v] for u in final.nodes for v in final.nodes[u])
            cold = T.Evaluator(OracleEngine(cm, nthreads=8), cm).evaluate(final, leaves)
            assert cold.adjusted == st["final_cost"]
            assert st["joins_swept"] > 0 and st["breaks"] > 0
        finally:
            ev.close()
