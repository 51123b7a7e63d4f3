"""Packs the reference's own tree-cost known-answer tests into tests/golden/trees/tree_costs.npz.  Run here, where
/root/reference exists:

    python tests/golden/make_tree_golden.py

Source: test/cost_tests lines 7-14 of the reference -- scripts cc.poy / cc1.poy / cc2.poy / cc3.poy read N.fas
(N = 1..13, test/first_fasta) and the fixed tree N.fas[.k].tree under tcm (1,2) / (1,1) / (2,1)+gap_opening 1 /
(3,1)+gap_opening 5, and poy_test.ml compares ``Ptree.get_cost `Adjusted`` with the N-th line of test/cc[k].costs
(the cnc* scripts are the same with weightfactor:-1, i.e. the negated numbers).  The fixture stores the input text
(fasta + tree, zlib-compressed by numpy) and the expected costs; nothing is computed here.
"""
import os

import numpy as np

REF = "/root/reference/test"
HERE = os.path.dirname(os.path.abspath(__file__))
REGIMES = [("", "cc.costs", (1, 2, -1)), (".1", "cc1.costs", (1, 1, -1)), (".2", "cc2.costs", (2, 1, 1)), (".3", "cc3.costs", (3, 1, 5))]


def main():
    files = [l.strip() for l in open(os.path.join(REF, "first_fasta")) if l.strip()]
    out = {"files": np.array(files)}
    for n, fn in enumerate(files):
        out[f"fasta_{n}"] = np.frombuffer(open(os.path.join(REF, fn), "rb").read(), np.uint8)
    for k, (suf, costs, tcm) in enumerate(REGIMES):
        out[f"tcm_{k}"] = np.array(tcm, np.int32)  # substitution, indel, gap opening (-1 = none)
        out[f"costs_{k}"] = np.array([int(x) for x in open(os.path.join(REF, costs)).read().split()], np.int64)
        for n, fn in enumerate(files):
            out[f"tree_{k}_{n}"] = np.frombuffer(open(os.path.join(REF, fn + suf + ".tree"), "rb").read(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "trees", "tree_costs.npz"), **out)
    print("wrote tree_costs.npz", os.path.getsize(os.path.join(HERE, "trees", "tree_costs.npz")), "bytes")


if __name__ == "__main__":
    main()
