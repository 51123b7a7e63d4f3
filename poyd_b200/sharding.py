"""Multi-GPU execution of a pair batch: one process per GPU, pairs sharded by index, no data-path collective.

Every pair is a pure function of (two sequences, cost matrix, parameters) -- SURVEY.md 8(e) -- so a batch is
split by work, each rank runs its shard through its own :class:`poyd_b200.sequence.Align`, and results are
gathered on the host in the caller's pair order.  The only collective the design knows is the optional sum of
per-shard cost totals (what a downpass level needs when only the tree cost is wanted).

poyd itself distributes whole scripts on whole trees between servants (src/poyd/PoydParallel.ml:486-514), never
single alignments.  Two ways to use several GPUs exist in this repository:

* inside ONE process: ``poyb200_multi_batch`` (csrc/multi.cu, ``sequence.MultiAlign``) -- one host thread per device,
  results written straight into the caller's buffers; this is the north_star's "shard by pair index, gather on the host";
* across processes (one per GPU, ``torch.distributed``): this module.  ``bench.py --gpus N`` runs its
  ``run_sharded`` + ``cost_sum`` on the GPUs (the ``rank_sharded_gather`` block of the bench line, NCCL backend) and
  tests/test_sharding_gloo.py runs the same host logic with two gloo ranks on the CPU.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np


def shard_indices(work: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Indices of the pairs rank `rank` processes: pairs sorted by descending work (DP cells) and dealt
    round-robin, so every rank gets the same mix of long and short alignments.  Deterministic; the union over
    ranks is a partition of range(len(work))."""
    order = np.argsort(-np.asarray(work, dtype=np.int64), kind="stable")
    return np.sort(order[rank::world])


def _dist():
    import torch.distributed as dist

    return dist


def gather_rows(local: np.ndarray, idx: np.ndarray, n_total: int, dst: int = 0) -> Optional[np.ndarray]:
    """Host gather of per-pair results (first axis = pairs of this rank's shard, in `idx` order) into an array
    over all pairs on rank `dst`; returns None elsewhere.  Works with the gloo and the nccl backend."""
    import torch

    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = np.zeros((n_total,) + local.shape[1:], local.dtype)
        out[idx] = local
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(idx)], dtype=torch.int64, device=dev))
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    if local.ndim > 1:
        # trailing shapes may differ between ranks (the row width of an Align result follows the shard's longest pair):
        # agree on the widest and pad on the right
        shp = [torch.zeros(local.ndim - 1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(shp, torch.tensor(local.shape[1:], dtype=torch.int64, device=dev))
        widest = tuple(int(max(s[d].item() for s in shp)) for d in range(local.ndim - 1))
        if widest != tuple(local.shape[1:]):
            grown = np.zeros((local.shape[0],) + widest, local.dtype)
            grown[(slice(None),) + tuple(slice(0, k) for k in local.shape[1:])] = local
            local = grown
    pad_idx = torch.full((m,), -1, dtype=torch.int64, device=dev)
    pad_idx[: len(idx)] = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(dev)
    pad_val = torch.zeros((m,) + local.shape[1:], dtype=torch.from_numpy(local[:0].copy()).dtype, device=dev)
    pad_val[: len(idx)] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    idx_list = [torch.empty_like(pad_idx) for _ in range(world)] if rank == dst else None
    val_list = [torch.empty_like(pad_val) for _ in range(world)] if rank == dst else None
    dist.gather(pad_idx, idx_list, dst=dst)
    dist.gather(pad_val, val_list, dst=dst)
    if rank != dst:
        return None
    out = np.zeros((n_total,) + local.shape[1:], local.dtype)
    for r in range(world):
        k = counts[r]
        out[idx_list[r][:k].cpu().numpy()] = val_list[r][:k].cpu().numpy()
    return out


def cost_sum(costs: np.ndarray) -> int:
    """Sum of the alignment costs over all shards (the one reduction of the design, C1 in SURVEY.md 2c)."""
    import torch

    dist = _dist()
    total = int(np.asarray(costs, dtype=np.int64).sum())
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return total
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([total], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def run_sharded(worker: Callable[[np.ndarray], Dict[str, np.ndarray]], work: np.ndarray, dst: int = 0):
    """Runs `worker(pair_indices) -> {name: per-pair array}` on this rank's shard and gathers every array on
    rank `dst` in the original pair order.  Returns (gathered dict or None, this rank's indices)."""
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    idx = shard_indices(work, world, rank)
    local = worker(idx)
    out = {}
    for k in sorted(local):
        out[k] = gather_rows(np.asarray(local[k]), idx, len(work), dst)
    return (out if rank == dst else None), idx
