#!/usr/bin/env python
"""GCUPS of the three-sequence cube (BASELINE.json configs[3]: 300 bp DNA triples) -- a side benchmark; the headline
metric is bench.py.  Prints one JSON line: device time of poyb200_batch_align_3 (host buffers in and out) for N triples,
and the compiled reference (oracle/_ref) on a few triples, one thread."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--triples", type=int, default=1184)
    ap.add_argument("--length", type=int, default=300)
    ap.add_argument("--cpu-triples", type=int, default=4)
    args = ap.parse_args()
    from oracle import oracle
    from poyd_b200 import cost_matrix as CM, sequence as S, synth

    cm = CM.default_nucleotides()
    cm3 = CM.of_two_dim(cm)
    # parent + two children at 10 % (SURVEY.md 8d cfg 4): reuse the pair generator twice on the same parents
    pool_a, _ = synth.pair_batch(args.triples, args.length, seed=4, min_len=args.length - 30)
    pool_b, _ = synth.pair_batch(args.triples, args.length, seed=4 + 1000, min_len=args.length - 30)
    seqs = []
    for t in range(args.triples):
        seqs += [pool_a.seq(2 * t), pool_a.seq(2 * t + 1), pool_b.seq(2 * t + 1)]
    pool = S.SeqPool(seqs)
    triples = np.arange(3 * args.triples, dtype=np.int32).reshape(-1, 3)
    cells = int(np.prod(pool.len[triples].astype(np.int64), axis=1).sum())
    al = S.Align3(cm, cm3)
    al.align_3(pool, triples[: min(64, args.triples)], want=3)  # warm-up
    t0 = time.perf_counter()
    g = al.align_3(pool, triples, want=3)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    al.cost_3(pool, triples)
    dt_cost = time.perf_counter() - t0
    oracle.build(ref=True)
    chk = oracle.best_checker_3(cm3)
    t0 = time.perf_counter()
    ccells = 0
    for t in range(min(args.cpu_triples, args.triples)):
        i1, i2, i3 = triples[t]
        r = chk.align_3(pool.seq(i1), pool.seq(i2), pool.seq(i3))
        assert r[0] == g.cost[t]
        ccells += int(pool.len[i1]) * int(pool.len[i2]) * int(pool.len[i3])
    cdt = time.perf_counter() - t0
    print(json.dumps({"metric": "GCUPS, 3-D cube as the reference executes it (align_3 = fill + traceback + median)",
                      "triples": args.triples, "length": args.length, "cells": cells, "gpu_seconds": dt,
                      "gpu_gcups": cells / dt * 1e-9, "gpu_gcups_cost_only": cells / dt_cost * 1e-9,
                      "cpu_gcups_1_thread": ccells / cdt * 1e-9, "cpu_kind": chk.kind,
                      "walks_out_of_bounds": int(g.status.sum())}))
    al.close()


if __name__ == "__main__":
    main()
