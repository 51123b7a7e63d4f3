"""world_size-2 gloo test of the pair-sharding path (CPU): shard, run, gather == unsharded run.
The per-rank worker is the CPU oracle here -- the point is the host logic; on the GPU box the worker is Align."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import oracle
    from poyd_b200 import cost_matrix as CM, sharding, synth

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cm = CM.nucleotides(1, 2, 3)
    pool, pairs = synth.ragged_batch(64, max_len=120, seed=4, gap_ambiguity=0.05)
    chk = oracle.Port(cm)
    work = np.array([chk.cells_affine(int(pool.len[a]), int(pool.len[b])) for a, b in pairs])

    def run(idx):
        o = chk.batch(3, pool.pool, pool.off, pool.len, pairs[idx])
        med = np.zeros((len(idx), 256), np.uint8)  # rows of one common width on every rank
        med[:, : o["median"].shape[1]] = o["median"]
        return {"cost": o["cost"], "median": med, "lens": o["lens"]}

    got, idx = sharding.run_sharded(run, work)
    total = sharding.cost_sum(run(idx)["cost"])
    full = chk.batch(3, pool.pool, pool.off, pool.len, pairs)
    assert total == int(full["cost"].astype(np.int64).sum())
    if rank == 0:
        assert np.array_equal(got["cost"], full["cost"])
        assert np.array_equal(got["lens"], full["lens"])
        w = min(got["median"].shape[1], full["median"].shape[1])
        for p in range(len(pairs)):
            L = full["lens"][p, 0]
            assert np.array_equal(got["median"][p, :L], full["median"][p, :L])
        open(os.path.join(tmpdir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_partition():
    sys.path.insert(0, ROOT)
    from poyd_b200 import sharding

    work = np.random.default_rng(0).integers(1, 1000, size=101)
    for world in (1, 2, 3, 8):
        parts = [sharding.shard_indices(work, world, r) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(101))
        loads = [work[p].sum() for p in parts]
        assert max(loads) - min(loads) <= work.max()


def test_two_rank_gloo_gather(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import oracle

    oracle.build(ref=False)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
