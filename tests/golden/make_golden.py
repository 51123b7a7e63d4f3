"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libpoyref.so = /root/reference/src/algn.c
compiled in this container, see oracle/Makefile).  Run here, where /root/reference exists:

    python tests/golden/make_golden.py

The reference repository holds no per-pair golden vectors for this path, so these fixtures pin
the plain-C port (oracle/poy_oracle.c) and the CUDA path to outputs of the reference itself.  Each file stores the
inputs (pool, offsets, lengths, pairs, deltaw, cost-matrix tables) next to the reference outputs, so the tests
need nothing but numpy to replay them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import oracle  # noqa: E402
from poyd_b200 import cost_matrix as CM, synth  # noqa: E402


def cm_arrays(cm):
    return dict(cm_scalars=np.array([cm.a_sz_in, cm.a_sz, cm.lcm, cm.gap, cm.cost_model_type, cm.combinations, cm.gap_open,
                                     cm.is_metric, cm.all_elements], np.int32),
                cm_cost=cm.cost, cm_median=cm.median, cm_worst=cm.worst, cm_prepend=cm.prepend_cost, cm_tail=cm.tail_cost)


def deltaw_like_sequence_ml(pool, pairs, gap):
    cnt = pool.count(gap)
    la, lb = pool.len[pairs[:, 0]].astype(np.int64), pool.len[pairs[:, 1]].astype(np.int64)
    l1, l2 = np.maximum(la, lb), np.minimum(la, lb)
    lower = (l1 * 0.10).astype(np.int64)
    return (np.maximum(cnt[pairs[:, 0]], cnt[pairs[:, 1]]) + np.where(l1 - l2 < lower, lower // 2, 2)).astype(np.int32)


def save(name, cm, pool, pairs, mode, deltaw=None):
    ref = oracle.Reference(cm)
    o = ref.batch(mode, pool.pool, pool.off, pool.len, pairs, deltaw=deltaw)
    d = dict(pool=pool.pool, off=pool.off, len=pool.len, pairs=pairs, mode=np.int32(mode), **cm_arrays(cm))
    if deltaw is not None:
        d["deltaw"] = deltaw
    for k, v in o.items():
        d["ref_" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, len(pairs), "pairs, cost sum", int(o["cost"].sum()))


ONLY = set(sys.argv[1:])  # fixture names to (re)generate; none given = all


def main():
    oracle.build(ref=True)
    assert oracle.Reference.available(), "the reference must be compiled here"
    global save
    save_all = save

    def save(name, *a, **k):  # noqa: F811
        if not ONLY or name in ONLY:
            save_all(name, *a, **k)

    aff = CM.nucleotides(1, 2, 3)
    # leaf-like pairs of the headline shape (no gap bits): the pairs aff_x2_kernel / aff_fast_kernel take; 44 pairs = eleven
    # batches of four = five double batches of aff_x2_kernel and one with an empty half
    pool, pairs = synth.pair_batch(44, 500, seed=7, min_len=450)
    save("affine_cfg2_leaflike", aff, pool, pairs, 3)
    save("affine_cfg2_leaflike_cost", aff, pool, pairs, 2)
    pool, pairs = synth.pair_batch(48, 500, seed=2, min_len=450, ambiguity=0.005, gap_ambiguity=0.10)
    save("affine_cfg2_medianlike", aff, pool, pairs, 3)
    save("affine_cfg2_medianlike_cost", aff, pool, pairs, 2)
    pool, pairs = synth.ragged_batch(160, max_len=200, seed=12, gap_ambiguity=0.1)
    save("affine_ragged", CM.nucleotides(2, 1, 1), pool, pairs, 3)
    save("affine_ragged_cost", CM.nucleotides(2, 1, 1), pool, pairs, 2)
    lin = CM.default_nucleotides()
    pool, pairs = synth.ragged_batch(160, max_len=220, seed=13, gap_ambiguity=0.03)
    save("linear_ragged", lin, pool, pairs, 1, deltaw_like_sequence_ml(pool, pairs, lin.gap))
    pool, pairs = synth.pair_batch(48, 500, seed=5, min_len=450)
    save("linear_cfg2lin", lin, pool, pairs, 1, deltaw_like_sequence_ml(pool, pairs, lin.gap))
    prot = CM.default_aminoacids()
    pool, pairs = synth.pair_batch(32, 300, seed=3, alphabet="protein", subst=0.15, indel=0.02)
    save("protein_cfg3a_full", prot, pool, pairs, 1, deltaw_like_sequence_ml(pool, pairs, prot.gap))
    save("protein_cfg3b_band16", prot, pool, pairs, 1, np.full(len(pairs), 16, np.int32))


if __name__ == "__main__":
    main()
