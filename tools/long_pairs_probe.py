#!/usr/bin/env python
"""Operands longer than 2048 elements: register (stripe / ring) kernels against the generic kernels, same results, times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poyd_b200 import cost_matrix as CM, sequence as S, synth  # noqa: E402

for length, n in ((3000, 20000), (5000, 8000)):
    pool, pairs = synth.pair_batch(n, length, seed=3, min_len=length - 30, gap_ambiguity=0.05)
    cells = sum(S.cells_affine(int(pool.len[a]), int(pool.len[b])) for a, b in pairs[:200]) / 200 * n
    out = {}
    for name, cfg in (("register kernels", {}), ("generic kernels", {"force_generic": 1})):
        al = S.Align(CM.nucleotides(1, 2, 3), config=cfg)
        al.align_affine_3(pool, pairs[:256], S.WANT_MEDIAN)
        t0 = time.perf_counter()
        g = al.align_affine_3(pool, pairs, S.WANT_MEDIAN)
        dt = time.perf_counter() - t0
        out[name] = (g.cost.copy(), g.lens[:, 0].copy(), dt)
        al.close()
        print(f"{length} bp x {n} pairs, {name}: {dt * 1e3:.1f} ms end to end = {cells / dt * 1e-9:.1f} GCUPS", flush=True)
    a, b = out["register kernels"], out["generic kernels"]
    print("  identical costs and median lengths:", bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])))
