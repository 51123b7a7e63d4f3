import os, sys, time
os.environ["POYB200_CONFIG"] = "trace=1,timing=1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poyd_b200 import cost_matrix as CM, sequence as S, synth
for name, cm in (("linear", CM.default_nucleotides()), ("affine", CM.nucleotides(1, 2, 3))):
    al = S.Align(cm, device=0)
    for n in (128, 1024, 8192):
        pool, pairs = synth.pair_batch(n, 500, seed=3, min_len=450)
        dw = None if al.is_affine else al.deltaw_for(pool, pairs)
        for want in (0, S.WANT_MEDIAN):
            mode = (2 if al.is_affine else 0) + (1 if want else 0)
            b, res = al.make_batch(pool, pairs, deltaw=dw, want=want)
            al.stage(mode, b); al.sync()
            for _ in range(3):
                al.run()
            al.sync()
            t0 = time.perf_counter()
            al.run(); al.sync()
            dt = (time.perf_counter() - t0) * 1e3
            print(f"{name} n={n} want={want}: run+sync {dt:.3f} ms, kernel phases {al.last_run_ms()}, launches so far {al.launch_count()}", flush=True)
            sys.stderr.flush()
            t0 = time.perf_counter()
            if want:
                al.align_2(pool, pairs, want)
            else:
                al.cost_2(pool, pairs)
            print(f"   one-shot {(time.perf_counter() - t0) * 1e3:.3f} ms", flush=True)
    al.close()
