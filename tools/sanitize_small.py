"""Small batches through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poyd_b200 import cost_matrix as CM, sequence as S, synth  # noqa: E402

pool, pairs = synth.ragged_batch(48, max_len=160, seed=3, gap_ambiguity=0.05)
al = S.Align(CM.nucleotides(1, 2, 3))
al.align_affine_3(pool, pairs, 7)
al.cost_2(pool, pairs)
al.close()
pool2, pairs2 = synth.pair_batch(24, 500, seed=2, min_len=450)
al = S.Align(CM.nucleotides(1, 2, 3))
al.align_affine_3(pool2, pairs2, 7)
al.close()
al = S.Align(CM.default_nucleotides())
al.align_2(pool, pairs, 7)
al.align_2(pool2, pairs2, 7)
al.cost_2(pool, pairs)
al.close()
al = S.Align(CM.nucleotides(1, 2, 3), config={"force_generic": 1})
al.align_affine_3(pool, pairs[:16], 7)
al.close()
al = S.Align(CM.default_nucleotides(), config={"force_generic": 1})
al.align_2(pool, pairs[:16], 7)
al.close()
cm = CM.default_nucleotides()
a3 = S.Align3(cm, CM.of_two_dim(cm))
tri = np.arange(12, dtype=np.int32).reshape(-1, 3)
a3.align_3(pool, tri, 3)
a3.close()
print("sanitize_small: done")
