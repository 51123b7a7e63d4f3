#!/usr/bin/env python
"""Latency of one small one-shot call (what a Wagner dependency level is): median time of poyb200_batch_align_affine_3 on
batches of 1 .. 2048 pairs of 1.5 kb, per kernel path (use_ring 0 / 1 / 2)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from poyd_b200 import cost_matrix as CM, sequence as S, synth  # noqa: E402


def main():
    cm = CM.nucleotides(1, 2, 3)
    bp = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    for gapamb in (0.0, 0.1):
        pool, pairs = synth.pair_batch(2048, bp, seed=5, min_len=bp - 100, gap_ambiguity=gapamb)
        for ring in (2, 1, 0):
            al = S.Align(cm, config={"use_ring": ring})  # use_ring = 2 here never switches (small_ring_pairs = 0)
            line = f"bp={bp} gap-bit operands={gapamb > 0} use_ring={ring}:"
            for n in (1, 8, 64, 512, 2048):
                sub = pairs[:n]
                al.align_affine_3(pool, sub, 9)  # DOS payload: median + bitsets
                ts = []
                for _ in range(9):
                    t0 = time.perf_counter()
                    al.align_affine_3(pool, sub, 9)
                    ts.append(time.perf_counter() - t0)
                line += f"  n={n}: {np.median(ts) * 1e3:.3f} ms"
            print(line, flush=True)
            al.close()


if __name__ == "__main__":
    main()
