"""Small-batch latency of the one-shot C-ABI calls (what a tree-level caller sees): ms per call for n pairs of L bp."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poyd_b200 import cost_matrix as CM, sequence as S, synth  # noqa: E402

for cm_name, cm in (("affine", CM.nucleotides(1, 2, 3)), ("linear", CM.default_nucleotides())):
    al = S.Align(cm, device=0)
    for L in (500, 1500):
        for n in (1, 16, 128, 1024):
            pool, pairs = synth.pair_batch(n, L, seed=3, min_len=L - 50)
            for want, label in ((S.WANT_MEDIAN, "median"), (0, "cost")):
                f = (lambda: al.align_2(pool, pairs, want)) if want else (lambda: al.cost_2(pool, pairs))
                for _ in range(3):
                    f()
                reps = 30
                t0 = time.perf_counter()
                for _ in range(reps):
                    f()
                ms = (time.perf_counter() - t0) / reps * 1e3
                print(f"{cm_name} L={L} n={n} {label}: {ms:.3f} ms/call  ({ms / n * 1e3:.1f} us/pair)", flush=True)
    os.environ["POYB200_CONFIG"] = "trace=1,timing=1"
    al.close()
