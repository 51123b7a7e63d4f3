#!/usr/bin/env python
"""Quick GPU check of the ring kernels against the checker, with diagnostics (which output, which pair, where)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from poyd_b200 import cost_matrix as CM, sequence as S, synth  # noqa: E402


def compare(tag, g, o):
    bad = np.nonzero(g.cost != o["cost"])[0]
    msg = [f"{tag}: cost mismatches {len(bad)}/{len(o['cost'])}"]
    if len(bad):
        msg.append(f"  first {bad[:6]} gpu {g.cost[bad[:6]]} ref {o['cost'][bad[:6]]}")
    for name, k, buf in (("median", 0, g.median), ("medianwg", 1, g.medianwg), ("ra", 2, g.aligned_a), ("rb", 3, g.aligned_b)):
        if buf is None:
            continue
        nb = 0
        first = None
        for p in range(len(o["cost"])):
            L = int(o["lens"][p, k])
            if g.lens[p, k] != L or not np.array_equal(buf[p, buf.shape[1] - L:], o[name][p, :L]):
                nb += 1
                if first is None:
                    first = p
        msg.append(f"  {name}: {nb} rows differ" + (f" (first pair {first}, len gpu {g.lens[first, k]} ref {o['lens'][first, k]})" if nb else ""))
    print("\n".join(msg), flush=True)
    return len(bad) == 0


def main():
    cm = CM.nucleotides(1, 2, 3)
    chk = oracle.best_checker(cm)
    ok = True
    for name, kw in (("leaf-like 500", dict()), ("median-like 500", dict(ambiguity=0.005, gap_ambiguity=0.10))):
        pool, pairs = synth.pair_batch(3000, 500, seed=5, min_len=450, **kw)
        o = chk.batch(3, pool.pool, pool.off, pool.len, pairs, nthreads=8)
        oc = chk.batch(2, pool.pool, pool.off, pool.len, pairs, nthreads=8)["cost"]
        for cfg in ({}, {"use_ring": 1}, {"use_ring": 0}):
            al = S.Align(cm, config=cfg)
            g = al.align_affine_3(pool, pairs, 7)
            ok &= compare(f"{name} {cfg}", g, o)
            c = al.cost_2(pool, pairs)
            print(f"  cost-only mismatches {int((c != oc).sum())}", flush=True)
            ok &= bool((c == oc).all())
            al.close()
    pool, pairs = synth.ragged_batch(2000, max_len=400, seed=9, gap_ambiguity=0.05)
    o = chk.batch(3, pool.pool, pool.off, pool.len, pairs, nthreads=8)
    al = S.Align(cm)
    ok &= compare("ragged 400", al.align_affine_3(pool, pairs, 7), o)
    al.close()
    print("RING CHECK", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
