"""Loads a tests/golden/*.npz fixture (produced by tests/golden/make_golden.py from the compiled reference)."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_files():
    return sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load(path):
    from poyd_b200.cost_matrix import CostMatrix
    from poyd_b200.sequence import SeqPool

    z = np.load(path)
    s = z["cm_scalars"]
    cm = CostMatrix(a_sz_in=int(s[0]), a_sz=int(s[1]), lcm=int(s[2]), gap=int(s[3]), cost_model_type=int(s[4]),
                    combinations=int(s[5]), gap_open=int(s[6]), is_metric=int(s[7]), all_elements=int(s[8]),
                    cost=z["cm_cost"], median=z["cm_median"], worst=z["cm_worst"], prepend_cost=z["cm_prepend"],
                    tail_cost=z["cm_tail"])
    pool = SeqPool.__new__(SeqPool)
    pool.pool, pool.off, pool.len = z["pool"], z["off"], z["len"]
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    deltaw = z["deltaw"] if "deltaw" in z.files else None
    return cm, pool, z["pairs"], int(z["mode"]), deltaw, ref
