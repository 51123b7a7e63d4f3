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
    from oracle_engine import OracleEngine

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
