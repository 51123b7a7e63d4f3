"""One kernel family per run, for compute-sanitizer --tool synccheck: argv[1] in {aff, lin_stripe, lin_rows, powell}."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poyd_b200 import cost_matrix as CM, sequence as S, synth  # noqa: E402

what = sys.argv[1]
pool, pairs = synth.ragged_batch(48, max_len=160, seed=3, gap_ambiguity=0.05)
if what == "aff":
    al = S.Align(CM.nucleotides(1, 2, 3))
    al.align_affine_3(pool, pairs, 7)
elif what == "lin_stripe":
    al = S.Align(CM.default_nucleotides(), config={"allow_rows": 0})
    al.align_2(pool, pairs, 7)
elif what == "lin_rows":
    al = S.Align(CM.default_nucleotides())
    al.align_2(pool, pairs, 7, deltaw=np.full(len(pairs), 600, np.int32), raw_deltaw=True)
else:
    cm = CM.nucleotides(1, 2, 3)
    al = S.Align3(cm, CM.of_two_dim(cm))
    rng = np.random.default_rng(5)
    a = np.concatenate([[16], rng.choice(np.array([1, 2, 4, 8], np.uint8), size=40)]).astype(np.uint8)
    b = a.copy(); b[5] = 2
    al.align_3_powell_inter(S.SeqPool([a, b, a.copy()]), np.array([[0, 1, 2]], np.int32))
al.close()
print("sync_probe", what, "done")
