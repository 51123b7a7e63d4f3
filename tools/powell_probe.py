#!/usr/bin/env python
"""Times poyb200_batch_powell_3 on batches of triples and the compiled reference on a sample of the same triples."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import powell_util as PU  # noqa: E402
from poyd_b200 import cost_matrix as CM, sequence as S  # noqa: E402


def main():
    ref = PU.reference()
    cm = CM.nucleotides(1, 2, 3)
    al = S.Align3(cm, CM.of_two_dim(cm))
    configs = ((100, 0.05, 592), (300, 0.03, 592), (300, 0.10, 148), (500, 0.05, 148))
    if len(sys.argv) > 1:  # "n,p,count" triples on the command line
        configs = tuple((int(a.split(",")[0]), float(a.split(",")[1]), int(a.split(",")[2])) for a in sys.argv[1:])
    for n, p, count in configs:
        rng = np.random.default_rng(n)
        cases = []
        for _ in range(count):
            a = PU.dna(rng, n)
            cases.append((a, PU.mutate(rng, a, p), PU.mutate(rng, a, p)))
        pool = S.SeqPool([s for t in cases for s in t])
        triples = np.arange(3 * count, dtype=np.int32).reshape(-1, 3)
        al.align_3_powell(pool, triples, 1, 3, 2, want=1)  # warm-up at full size: the workspaces stay with the context
        t0 = time.perf_counter()
        g = al.align_3_powell(pool, triples, 1, 3, 2, want=3)
        dt = time.perf_counter() - t0
        line = f"n={n} p={p} triples={count}: {dt:.3f} s = {count / dt:.1f} triples/s, mean cost {g.cost.mean():.1f}, status {np.bincount(g.status)}"
        if ref is not None and not os.environ.get("PROBE_NO_REF"):
            k = 4 if n >= 300 else 12
            t0 = time.perf_counter()
            ok = True
            for t in range(k):
                rc, rows = PU.ref_powell(ref, *cases[t], 1, 3, 2)
                ok &= rc == g.cost[t] and np.array_equal(rows[0], g.get("aligned_1", t))
            dr = (time.perf_counter() - t0) / k
            line += f" | reference {1 / dr:.2f} triples/s on one core, same results on the sample: {ok}"
        print(line, flush=True)
    al.close()


if __name__ == "__main__":
    main()
