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
# round 2: full linear matrices (lin_rows_kernel), the 6-bit band of the default affine path, gap-bit operands (ring kernel),
# protein, and Powell's 3-D aligner
al = S.Align(CM.default_nucleotides())
full = np.full(len(pairs), 600, np.int32)
al.align_2(pool, pairs, 15 & ~1 | 1, deltaw=full, raw_deltaw=True)
al.align_2(pool2, pairs2, 7, deltaw=np.full(len(pairs2), 600, np.int32), raw_deltaw=True)
al.close()
poolp, pairsp = synth.pair_batch(16, 300, seed=3, alphabet="protein", subst=0.15, indel=0.02)
al = S.Align(CM.default_aminoacids())
al.align_2(poolp, pairsp, 7)
al.close()
pool3, pairs3 = synth.pair_batch(24, 500, seed=4, min_len=450, gap_ambiguity=0.1, ambiguity=0.005)
al = S.Align(CM.nucleotides(1, 2, 3))
al.align_affine_3(pool3, pairs3, 9)
al.close()
rng = np.random.default_rng(5)
trs = []
for n in (12, 40, 90):
    a = np.concatenate([[16], rng.choice(np.array([1, 2, 4, 8], np.uint8), size=n)]).astype(np.uint8)
    b, c = a.copy(), a.copy()
    b[1 + rng.integers(0, n, size=max(1, n // 10))] = 2
    c = np.delete(c, 1 + rng.integers(0, n, size=max(1, n // 15)))
    trs += [a, b, c]
pw = S.SeqPool(trs)
cm = CM.nucleotides(1, 2, 3)
a3 = S.Align3(cm, CM.of_two_dim(cm))
a3.align_3_powell_inter(pw, np.arange(9, dtype=np.int32).reshape(-1, 3))
a3.close()
print("sanitize_small: done")
