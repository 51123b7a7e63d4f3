#!/usr/bin/env python
"""bench.py -- GCUPS of the batched pairwise median DP (BASELINE.json `metric`) on N B200s.

A step = one pass of the hot path (algn_CAML_align_affine_3 for every pair: affine fill in the reference's band,
traceback, median / medianwg / aligned pair) over one batch of synthetic pairs -- configs[1] of BASELINE.json:
500 bp DNA, 10 % substitutions, 2 % indels, subst 1 / indel 2 / gap opening 3.

  value         cells / second with the batch already resident in HBM (kernels only, CUDA events, max over ranks)
  e2e           the same metric through the C ABI with HOST buffers: plan + H2D + kernels + D2H every step
  roofline      the dominant kernel against the bound that applies to an integer min-plus recurrence: the INT32 ALU
                (BASELINE.json north_star).  The peak is this repo's own add.s32 microbenchmark (no driver-measured
                INT32 peak exists), `executed_frac` is the executed-instruction view (warp instructions / issue slots);
                `roofline_hbm` is the same kernel against the driver-measured copy bandwidth
  workloads     every other BASELINE.json config, smaller batches, same legs: median-like operands, linear gaps,
                protein 300 aa (as the product computes deltaw, and with an explicit band), 300 bp triples (cube)
  cpu_baseline  the compiled reference algn.c (oracle/_ref) on cores + 1 processes (README:147-150), bounded sample

`--impl reference` times the reference's own CPU implementation instead (the driver computes the ratio).
Under torchrun (N > 1) every rank runs its own pairs (weak scaling, no data-path collective); rank 0 then also runs ONE
batch of N x pairs through the in-library sharded call (`sharded_call`: strong scaling over the same GPUs), and all ranks
run the cross-process shard / gather / cost-sum path of poyd_b200/sharding.py (`rank_sharded_gather`).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GCUPS (band cells/s, batched pairwise median DP: fill + traceback + median)"
UNIT = "GCUPS"

WORKLOADS = {
    # name: (description, mode, reference ops per cell, dominant kernel)
    "affine500": ("configs[1]: 1M DNA pairs 500 bp (10% subst, 2% indel), affine gaps (subst 1, indel 2, gap opening 3), "
                  "align_affine_3 = fill + traceback + median/medianwg/aligned pair", 3, 50, "aff_x2_kernel<5,8>"),
    "affine500_medianlike": ("configs[1], median-like operands: as affine500 plus 0.5% IUPAC ambiguities and 10% of positions "
                             "carrying the gap bit (what internal-node medians look like; exercises the block-diagonal "
                             "state)", 3, 50, "aff_ring_kernel<5,8,true,true>"),
    "linear500": ("cfg 2-lin: DNA pairs 500 bp, linear gaps (subst 1, indel 2), deltaw as Sequence.Align.cost_2 computes it, "
                  "align_2 + ancestor_2 + median_2_with_gaps", 1, 10, "lin_stripe_kernel<K,G,true>"),
    "protein300": ("configs[2] (3a): protein pairs 300 aa, 22x22 matrix 1/2, deltaw as the product computes it "
                   "(full matrix, SURVEY.md A15), align_2 + medians", 1, 10, "lin_rows_kernel<10,32,true>"),
    "protein300_band16": ("configs[2] (3b): protein pairs 300 aa, explicit deltaw 16, align_2 + medians", 1, 10,
                          "lin_stripe_kernel<K,G,true>"),
    "tree": ("configs[4]-style host-driver workload (one GPU per tree): Wagner build with batched candidate-edge sweeps, "
             "all-directions downpass, root selection, single assignment and adjusted cost of the built tree, then --spr SPR "
             "neighbours evaluated exactly in lockstep (poyd_b200/tree.py); "
             "synthetic --taxa x --bp DNA data set, affine gaps (1, 2, opening 3)", 3, 50, "aff_stripe_kernel<5,8,true>"),
}
SECONDARY = ("affine500_medianlike", "linear500", "protein300", "protein300_band16")


def workload(n_pairs: int, seed: int, name: str = "affine500"):
    """Returns (cm, pool, pairs, deltaw or None)."""
    from poyd_b200 import cost_matrix as CM, sequence as S, synth

    if name in ("affine500", "affine500_medianlike"):
        cm = CM.nucleotides(1, 2, 3)
        extra = dict(ambiguity=0.005, gap_ambiguity=0.10) if name.endswith("medianlike") else {}
        pool, pairs = synth.pair_batch(n_pairs, 500, seed=seed, min_len=450, stride=512, **extra)
        return cm, pool, pairs, None
    if name == "linear500":
        cm = CM.default_nucleotides()
        pool, pairs = synth.pair_batch(n_pairs, 500, seed=seed, min_len=450, stride=512)
    else:
        cm = CM.default_aminoacids()
        pool, pairs = synth.pair_batch(n_pairs, 300, seed=seed, alphabet="protein", subst=0.15, indel=0.02, stride=304)
    if name == "protein300_band16":
        return cm, pool, pairs, np.full(len(pairs), 16, np.int32)
    cnt = pool.count(cm.gap)
    la, lb = pool.len[pairs[:, 0]].astype(np.int64), pool.len[pairs[:, 1]].astype(np.int64)
    dw = np.maximum(cnt[pairs[:, 0]], cnt[pairs[:, 1]]) + S.deltaw_calc(np.maximum(la, lb), np.minimum(la, lb), None)
    return cm, pool, pairs, dw.astype(np.int32)


def total_cells(pool, pairs, deltaw=None) -> int:
    """Cells the reference visits for the batch: a pure function of the two lengths (and deltaw), SURVEY.md 8d.
    (poyd_b200.sequence.cells_* are host-side integer formulas of the library; no GPU work.)"""
    from poyd_b200 import sequence as S

    la, lb = pool.len[pairs[:, 0]].astype(np.int64), pool.len[pairs[:, 1]].astype(np.int64)
    if deltaw is None:
        key = la * 65536 + lb
        uniq, cnt = np.unique(key, return_counts=True)
        per = np.array([S.cells_affine(int(k >> 16), int(k & 65535)) for k in uniq], dtype=np.int64)
    else:
        l1, l2 = np.maximum(la, lb), np.minimum(la, lb)
        key = (l1 * 65536 + l2) * 65536 + deltaw.astype(np.int64)
        uniq, cnt = np.unique(key, return_counts=True)
        per = np.array([S.cells_linear(int(k >> 32), int((k >> 16) & 65535), int(k & 65535)) for k in uniq], dtype=np.int64)
    return int((per * cnt).sum())


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---- CPU arm: the reference's own implementation, cores + 1 single-threaded PROCESSES (README:147-150: N + 1 servants) ----
_CPU = {}


def _cpu_init(cm, pool, pairs, deltaw, mode):
    from oracle import oracle

    _CPU.update(chk=oracle.best_checker(cm), pool=pool, pairs=pairs, dw=deltaw, mode=mode)


def _cpu_shard(rng):
    lo, hi = rng
    c = _CPU
    t0 = time.perf_counter()
    c["chk"].batch(c["mode"], c["pool"].pool, c["pool"].off, c["pool"].len, c["pairs"][lo:hi],
                   deltaw=None if c["dw"] is None else c["dw"][lo:hi], nthreads=1)
    return time.perf_counter() - t0


def cpu_arm(cm, pool, pairs, sample: int, cores: int, steps: int = 1, warmup: int = 0, deltaw=None, mode: int = 3):
    """Times the reference's CPU implementation (compiled algn.c if oracle/_ref exists, else the port) on a bounded sample:
    P = cores + 1 processes, each a single-threaded loop over a contiguous shard of the pair list (exactly a servant's
    inner loop); wall time of the slowest shard, measured around the whole parallel region."""
    import multiprocessing as mp

    from oracle import oracle

    oracle.build(ref=True)
    kind = oracle.best_checker(cm).kind
    from poyd_b200 import sequence as S

    sample = min(sample, len(pairs))
    sub = pairs[:sample]
    dws = None if deltaw is None else deltaw[:sample]
    cells = total_cells(pool, sub, dws)
    # the sample's own sequences only: the worker processes inherit (fork) a pool of megabytes, not the whole workload
    used, inv = np.unique(sub.reshape(-1), return_inverse=True)
    pool = S.SeqPool([pool.seq(int(i)) for i in used])
    sub = inv.reshape(-1, 2).astype(np.int32)
    P = cores + 1
    cuts = [(sample * k // P, sample * (k + 1) // P) for k in range(P)]
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(P, initializer=_cpu_init, initargs=(cm, pool, sub, dws, mode)) as procs:
        procs.map(_cpu_shard, [(0, min(sample, 8))] * P)  # every process up and its checker built
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            procs.map(_cpu_shard, cuts, chunksize=1)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    return {"value": cells / sec * 1e-9, "unit": UNIT, "cores": cores, "processes": P, "kind": kind,
            "sample": f"first {sample} pairs of the workload, all outputs, {P} single-threaded processes on {cores} cores "
                      f"(N + 1 servants), {sec:.1f} s per pass"}, sec


def cpu_arm_cube(pool, triples, cm3, n: int):
    """The compiled reference's cube (algn_nw_3d + backtrack_3d + median) on the first n triples, one process."""
    from oracle import oracle

    oracle.build(ref=True)
    chk = oracle.best_checker_3(cm3)
    t0 = time.perf_counter()
    cells = 0
    for t in range(min(n, len(triples))):
        i1, i2, i3 = (int(x) for x in triples[t])
        chk.align_3(pool.seq(i1), pool.seq(i2), pool.seq(i3))
        cells += int(pool.len[i1]) * int(pool.len[i2]) * int(pool.len[i3])
    sec = time.perf_counter() - t0
    return {"value": cells / sec * 1e-9, "unit": UNIT, "cores": 1, "processes": 1, "kind": chk.kind,
            "sample": f"first {min(n, len(triples))} triples, one process, {sec:.1f} s"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            # median over the samples taken under load (above 60 % of the maximum seen)
            hi = [x for x in sm if x >= 0.6 * max(sm)]
            out["sm_mhz"] = float(np.median(hi))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def bench_tree(args, rank, local_rank, world, threads):
    """Secondary workload: the tree-level driver end to end (host buffers in every call).  One step = one Wagner build
    + one full evaluation of the built tree; every rank works on its own data set (independent replicas)."""
    from poyd_b200 import cost_matrix as CM, synth, tree as T

    cm = CM.nucleotides(1, 2, 3)
    leaves = synth.taxa_on_random_tree(args.taxa, args.bp, seed=5 + rank)

    def one_pass(engine, lv):
        ev = T.Evaluator(engine, cm)
        engine.log.clear()
        t0 = time.perf_counter()
        topo, steps = ev.wagner(lv)
        r = ev.evaluate(topo, lv, keep=True)
        t1 = time.perf_counter()
        n1, c1 = sum(len(x[1]) for x in engine.log), len(engine.log)
        if args.spr > 0:  # exact evaluation of a sample of the SPR neighbourhood, all trees in lockstep
            nb = T.spr_neighbours(topo, limit=args.spr, seed=7)
            rs = ev.evaluate_many(nb, lv, keep=True)
            r.stats["spr_trees"] = len(nb)
            r.stats["spr_best"] = min(x.adjusted for x in rs) if rs else None
        dt = time.perf_counter() - t0
        r.stats["phases"] = {"wagner_and_evaluation": {"seconds": round(t1 - t0, 3), "pairs": n1, "calls": c1},
                             "spr_lockstep": {"seconds": round(t0 + dt - t1, 3), "pairs": sum(len(x[1]) for x in engine.log) - n1,
                                              "calls": len(engine.log) - c1}}
        return dt, r, T.logged_cells(engine.log, True), sum(len(x[1]) for x in engine.log), len(engine.log)

    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import oracle
        from oracle.tree_engine import OracleEngine

        oracle.build(ref=True)
        sub = {k: v for k, v in leaves.items() if k <= min(args.taxa, 60)}
        dt, r, cells, npairs, ncalls = one_pass(OracleEngine(cm, nthreads=threads), sub)
        v = cells / dt * 1e-9
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
                          "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "int32", "data": "synthetic",
                          "config": {"workload": WORKLOADS["tree"][0], "taxa": len(sub), "bp": args.bp, "pairs": npairs,
                                     "batches": ncalls, "adjusted_cost": r.adjusted, "spr_trees": r.stats.get("spr_trees"),
                                     "spr_best": r.stats.get("spr_best")},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                                           "sample": f"the first {len(sub)} taxa, same driver, CPU checker engine"},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = T.GpuEngine(cm, device=local_rank)
    small = {k: v for k, v in leaves.items() if k <= min(args.taxa, 24)}
    for _ in range(max(1, min(args.warmup, 2))):
        one_pass(eng, small)
    l0 = eng.al.launch_count()
    tot_dt, tot_cells = 0.0, 0
    steps = max(1, min(args.steps, 2))
    for _ in range(steps):
        dt, r, cells, npairs, ncalls = one_pass(eng, leaves)
        tot_dt += dt
        tot_cells += cells
    launches = eng.al.launch_count() - l0
    t = torch.tensor([tot_dt / steps, float(tot_cells / steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        sec, cells_all = float(tm[0]), float(t[1])
    else:
        sec, cells_all = float(t[0]), float(t[1])
    if rank == 0:
        v = cells_all / sec * 1e-9
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": min(args.warmup, 2),
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic",
                "config": {"workload": WORKLOADS["tree"][0], "taxa": args.taxa, "bp": args.bp, "pairs": npairs, "batches": ncalls,
                           "adjusted_cost": r.adjusted, "spr_trees": r.stats.get("spr_trees"), "spr_best": r.stats.get("spr_best"), "phases": r.stats.get("phases"), "timing": "host wall clock around the driver (every call moves host "
                           "buffers in and out); small batches, so latency- not roofline-bound"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None},
                "gpu_launches": int(launches)}
        if not args.skip_cpu and world == 1:
            from oracle import oracle
            from oracle.tree_engine import OracleEngine

            oracle.build(ref=True)
            sub = {k: v_ for k, v_ in leaves.items() if k <= min(args.taxa, 60)}
            dt, rc, cells, _, _ = one_pass(OracleEngine(cm, nthreads=threads), sub)
            line["cpu_baseline"] = {"value": cells / dt * 1e-9, "unit": UNIT, "cores": threads, "kind": "reference",
                                    "sample": f"the first {len(sub)} taxa, same driver, CPU checker engine, {dt:.1f} s"}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def bench_cfg5(args, rank, local_rank, world, cores):
    """BASELINE.json configs[4]: synthetic --taxa x --bp unaligned DNA, Wagner build + SPR search with batched downpass medians,
    through the NATIVE driver (include/poyb200_tree.h) on device-resident stores.  N > 1: rank 0 builds the Wagner tree and
    broadcasts it; every SPR round each rank sweeps and evaluates its shard of the tree's neighbourhood (the breaks k with
    k mod N = rank), the ranks' finds are all-gathered and the candidate the unsharded search would have taken (smallest key)
    becomes the next tree on every rank.  The only collectives are those few integers and the winning topology."""
    import torch
    import torch.distributed as dist

    from poyd_b200 import cost_matrix as CM, synth, tree as T, tree_native as TN

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cm = CM.nucleotides(1, 2, 3)
    leaves = synth.taxa_on_random_tree(args.taxa, args.bp, seed=5)
    ev = TN.NativeEvaluator(cm, leaves, device=local_rank)
    # warm-up: a small build on the same evaluator type (kernels loaded, buffers grown)
    small = {k: v for k, v in leaves.items() if k <= 16}
    w = TN.NativeEvaluator(cm, small, device=local_rank)
    w.evaluate(w.wagner(sorted(small))[0])
    w.close()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n_nodes = 2 * args.taxa - 2
    if rank == 0:
        topo, steps = ev.wagner(sorted(leaves))
        ids, nbr = TN._pack(topo)
        pack = torch.from_numpy(np.concatenate([ids, nbr.reshape(-1), [topo.handle]]).astype(np.int32)).cuda()
    else:
        pack = torch.zeros(4 * n_nodes + 1, dtype=torch.int32, device="cuda")
    if world > 1:
        dist.broadcast(pack, src=0)
        p = pack.cpu().numpy()
        topo = TN._unpack(p[:n_nodes], p[n_nodes:4 * n_nodes].reshape(-1, 3), int(p[-1]), args.taxa)
    t_build = time.perf_counter() - t0
    start = ev.evaluate(topo, keep=True)
    best, rounds, totals = start.adjusted, 0, {"breaks": 0, "joins_swept": 0, "exact_evaluated": 0}
    t1 = time.perf_counter()
    while rounds < args.spr_rounds:
        found, key, cost, joined, st = ev.spr_round(topo, best, rank, world, window=args.spr_window)
        for k in totals:
            totals[k] += st[k]
        if world > 1:
            mine = torch.tensor([key if found else (1 << 62), cost if found else 0, rank], dtype=torch.int64, device="cuda")
            allv = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)
            win = min(allv, key=lambda v: int(v[0]))
            if int(win[0]) >= (1 << 62):
                break
            src = int(win[2])
            if rank == src:
                ids, nbr = TN._pack(joined)
                pack = torch.from_numpy(np.concatenate([ids, nbr.reshape(-1), [joined.handle]]).astype(np.int32)).cuda()
            else:
                pack = torch.zeros(4 * n_nodes + 1, dtype=torch.int32, device="cuda")
            dist.broadcast(pack, src=src)
            p = pack.cpu().numpy()
            topo = TN._unpack(p[:n_nodes], p[n_nodes:4 * n_nodes].reshape(-1, 3), int(p[-1]), args.taxa)
            best = int(win[1])
        else:
            if not found:
                break
            topo, best = joined, cost
        rounds += 1
    torch.cuda.synchronize()
    t_spr = time.perf_counter() - t1
    total = time.perf_counter() - t0
    stt = ev.stats()
    vals = torch.tensor([float(stt["cells"]), float(stt["pairs"]), float(stt["calls"]), float(totals["joins_swept"]),
                         float(totals["exact_evaluated"]), float(totals["breaks"])], dtype=torch.float64, device="cuda")
    tmax = torch.tensor([total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    launches = ev.al.launch_count()
    if rank == 0:
        sec = float(tmax[0])
        v = float(vals[0]) / sec * 1e-9
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": "configs[4]: synthetic %d-taxon x %d bp unaligned DNA, Wagner build + SPR search (first-best, whole-sweep "
                                       "evaluation, in-order replay) with batched downpass medians, affine gaps (1, 2, opening 3); native "
                                       "driver on device-resident sequence stores; one neighbourhood shard per GPU" % (args.taxa, args.bp),
                           "taxa": args.taxa, "bp": args.bp, "spr_rounds_limit": args.spr_rounds, "spr_window": args.spr_window},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": int(float(vals[1]) * 56), "d2h_bytes_per_step": int(float(vals[1]) * 12),
                        "note": "the driver IS the public call: leaves go up once, per batch only task descriptors go up and costs / lengths "
                                "/ gap counts come back"},
                "tree": {"wagner_seconds": t_build, "spr_seconds": t_spr, "start_cost": start.adjusted, "final_cost": best,
                         "accepted_rounds": rounds, "breaks": int(vals[5]), "joins_swept": int(vals[3]), "exact_evaluated": int(vals[4]),
                         "pairs": int(vals[1]), "batches": int(vals[2]), "cells": int(vals[0]),
                         "host_us_per_pair": sec * 1e6 / max(1.0, float(vals[1])) * world},
                "gpu_launches": int(launches)}
        if not args.skip_cpu and world == 1:
            from oracle import oracle
            from oracle.tree_engine import OracleEngine

            oracle.build(ref=True)
            sub = {k: v_ for k, v_ in leaves.items() if k <= min(args.taxa, 40)}
            eng = OracleEngine(cm, nthreads=cores)
            evc = T.Evaluator(eng, cm)
            tc0 = time.perf_counter()
            tp, _ = evc.wagner(sub)
            evc.evaluate(tp, sub, keep=True)
            dt = time.perf_counter() - tc0
            cells = T.logged_cells(eng.log, True)
            line["cpu_baseline"] = {"value": cells / dt * 1e-9, "unit": UNIT, "cores": cores, "kind": "reference",
                                    "sample": f"Wagner build + evaluation of the first {len(sub)} taxa, Python driver over the compiled "
                                              f"reference on {cores} threads, {dt:.1f} s"}
        print(json.dumps(line))
    ev.close()
    if world > 1:
        dist.destroy_process_group()


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Multi-rank runs: pin this process (and so its pinned host buffers, first touch) to the CPUs of the NUMA node its GPU
    hangs off, so the 8 ranks' host copies do not all cross one socket.  No-op where sysfs has no answer."""
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        node = int(open("/sys/bus/pci/devices/" + bdf[4:] + "/numa_node").read())
        if node < 0:
            return "numa_node unknown (-1): not bound"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"node {node}: none of its CPUs available, not bound"
        os.sched_setaffinity(0, cpus)
        return f"rank bound to NUMA node {node} ({len(cpus)} CPUs)"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"


def kernel_metrics():
    """Per-pair ncu figures of THIS round's build (profiles/r02_kernel_metrics.json, written from the committed
    `ncu --set full` captures by tools/ncu_summary.py): DRAM bytes and executed warp instructions per pair, per kernel."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_metrics.json")))
    except Exception:  # noqa: BLE001
        return {}


class Dist:
    """The collectives the bench needs: barrier, max and sum over ranks (NCCL under torchrun, identity at N = 1)."""

    def __init__(self, world):
        import torch

        self.torch, self.world = torch, world
        if world > 1:
            import torch.distributed as dist

            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX if self.world > 1 else None)

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM if self.world > 1 else None)


def pinned_batch(S, al, pool, pairs, dw, want, torch):
    """The batch descriptor over PINNED host buffers (inputs and outputs).  Returns (batch, tensors kept alive, info)."""
    def pinned(arr):
        return torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()

    ppool = S.SeqPool.__new__(S.SeqPool)
    keep = [pinned(pool.pool), pinned(pool.off), pinned(pool.len), pinned(pairs)]
    ppool.pool, ppool.off, ppool.len = keep[0].numpy(), keep[1].numpy(), keep[2].numpy()
    ppairs = keep[3].numpy()
    if dw is not None:
        keep.append(pinned(dw))
    batch, _ = al.make_batch(ppool, ppairs, deltaw=None if dw is None else keep[-1].numpy(), want=want, outputs=False)
    n = len(ppairs)
    stride = (int((ppool.len[ppairs[:, 0]].astype(np.int64) + ppool.len[ppairs[:, 1]]).max()) + 2 + 15) // 16 * 16
    bstride = ((stride + 7) // 8 + 3) // 4 * 4
    cost_t = torch.empty(n, dtype=torch.int32, pin_memory=True)
    lens_t = torch.empty((n, 4), dtype=torch.int32, pin_memory=True)
    batch.cost, batch.out_len, batch.out_stride = cost_t.data_ptr(), lens_t.data_ptr(), stride
    out = {"cost": cost_t, "lens": lens_t}
    keep += [cost_t, lens_t]
    return batch, keep, {"n": n, "stride": stride, "bstride": bstride, "out": out, "h2d": int(ppool.pool.nbytes + n * 64)}


def set_outputs(S, batch, info, keep, torch, payload):
    """payload 'four': median, medianwg and both aligned sequences; 'dos': what SeqCS.DOS.median keeps."""
    n, stride, bstride = info["n"], info["stride"], info["bstride"]
    batch.median = batch.medianwg = batch.aligned_a = batch.aligned_b = None
    batch.bits_a = batch.bits_b = batch.bits_wg = None
    if payload == "four":
        t = [torch.empty((n, stride), dtype=torch.uint8, pin_memory=True) for _ in range(4)]
        batch.want = S.WANT_MEDIAN | S.WANT_MEDIANWG | S.WANT_ALIGNED
        batch.median, batch.medianwg, batch.aligned_a, batch.aligned_b = (x.data_ptr() for x in t)
        d2h = int(4 * n * stride + n * 4 + n * 16)
    else:
        t = [torch.empty((n, stride), dtype=torch.uint8, pin_memory=True)] + \
            [torch.empty((n, bstride), dtype=torch.uint8, pin_memory=True) for _ in range(3)]
        batch.want = S.WANT_MEDIAN | S.WANT_BITSETS
        batch.median = t[0].data_ptr()
        batch.bits_a, batch.bits_b, batch.bits_wg = (x.data_ptr() for x in t[1:])
        batch.bits_stride = bstride
        d2h = int(n * stride + 3 * n * bstride + n * 4 + n * 16)
    keep.append(t)
    return d2h


def run_pairs_workload(name, n_pairs, args, rank, local_rank, D, payloads=("four", "dos"), sampler=False):
    """Device-resident and end-to-end legs of one pair workload on this rank's GPU; aggregates over ranks.
    Returns a dict (identical on every rank)."""
    import torch

    from poyd_b200 import sequence as S

    desc, mode, ops_per_cell, kernel = WORKLOADS[name]
    cm, pool, pairs, dw = workload(n_pairs, seed=2 + rank, name=name)  # each rank owns its own pairs
    cells = total_cells(pool, pairs, dw)
    al = S.Align(cm, device=local_rank)
    batch, keep, info = pinned_batch(S, al, pool, pairs, dw, 0, torch)
    d2h = {p: 0 for p in payloads}
    d2h[payloads[0]] = set_outputs(S, batch, info, keep, torch, payloads[0])
    stream = torch.cuda.ExternalStream(al.L.poyb200_stream(al.h), device=torch.device("cuda", local_rank))

    # ---- device-resident leg: stage once, time K passes of the kernels
    al.stage(mode, batch)
    al.sync()
    for _ in range(args.warmup):
        al.run()
    al.sync()
    D.barrier()
    smp = ClockSampler(local_rank) if (sampler and rank == 0) else None
    launches0 = al.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        al.run()
    e1.record(stream)
    al.sync()
    D.barrier()
    launches = al.launch_count() - launches0
    dev_ms = D.max(e0.elapsed_time(e1) / args.steps)
    f_ms, t_ms = al.last_run_ms()  # the last pass, per phase (CUDA events on the library's streams)
    cells_all = D.sum(float(cells))
    res = {"workload": desc, "kernel": kernel, "pairs_per_gpu_per_step": info["n"], "cells_per_gpu_per_step": cells,
           "value": cells_all / (dev_ms * 1e-3) * 1e-9, "ms_per_step": dev_ms, "phase_ms": {"fill": f_ms, "traceback": t_ms},
           "gpu_launches": int(launches), "launches_per_step": int(launches // max(1, args.steps))}

    # ---- end-to-end legs: the call a user of the C ABI makes -- host buffers in, host buffers out, every step
    fn = al.L.poyb200_batch_align_affine_3 if mode == 3 else al.L.poyb200_batch_align_2
    checksum = None
    for p in payloads:
        if d2h[p] == 0:
            d2h[p] = set_outputs(S, batch, info, keep, torch, p)
        for _ in range(max(1, args.warmup - 1)):
            al._check(fn(al.h, ctypes.byref(batch)))
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            al._check(fn(al.h, ctypes.byref(batch)))
        torch.cuda.synchronize()
        sec = D.max((time.perf_counter() - t0) / args.steps)
        D.barrier()
        cs = int(info["out"]["cost"].numpy().astype(np.int64).sum())
        assert checksum is None or cs == checksum
        checksum = cs
        key = "e2e" if p == "four" else "e2e_dos_median"
        res[key] = {"value": cells_all / sec * 1e-9, "unit": UNIT, "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": d2h[p],
                    "ms_per_step": sec * 1e3,
                    "outputs": "cost, median, medianwg, aligned a, aligned b (pinned host buffers)" if p == "four" else
                               "what SeqCS.DOS.median keeps (src/seqCS.ml:769-776): cost, median, gap bitsets of aligned a / "
                               "aligned b / medianwg (POYB200_WANT_BITSETS); same kernels, the aligned sequences stay on the device"}
    res["cost_checksum"] = checksum
    if smp:
        res["clocks"] = smp.stop()
    res["_objs"] = (al, cm, pool, pairs, dw, mode, ops_per_cell, cells)
    return res


def run_cube_workload(args, rank, local_rank, D):
    """configs[3]: 300 bp triples through poyb200_batch_align_3 (cost, aligned triple, median; host buffers in and out),
    a sample of --triples per GPU.  Only a one-shot entry point exists for the cube, so value = e2e (one untimed call of
    the same size first, as for every other workload: the device buffers are allocated by then)."""
    from poyd_b200 import cost_matrix as CM, sequence as S, synth

    cm = CM.default_nucleotides()
    cm3 = CM.of_two_dim(cm)
    pool, triples = synth.triple_batch(args.triples, 300, seed=40 + rank)
    cells = int(np.prod(pool.len[triples].astype(np.int64), axis=1).sum())
    al = S.Align3(cm, cm3, device=local_rank)
    al.align_3(pool, triples, want=3)  # warm-up at full size: the 16 GB of direction cubes stay with the context
    D.barrier()
    l0 = al.launch_count()
    t0 = time.perf_counter()
    g = al.align_3(pool, triples, want=3)
    sec = D.max(time.perf_counter() - t0)
    launches = al.launch_count() - l0
    cells_all = D.sum(float(cells))
    res = {"workload": "configs[3]: three-sequence medians (algn_fill_cube as the reference executes it + backtrack_3d + "
                       "algn_get_median_3d), 300 bp DNA triples, a sample of the 100k", "kernel": "cube_fill_kernel",
           "triples_per_gpu_per_step": len(triples), "cells_per_gpu_per_step": cells, "value": cells_all / sec * 1e-9,
           "ms_per_step": sec * 1e3, "gpu_launches": int(launches),
           "e2e": {"value": cells_all / sec * 1e-9, "unit": UNIT, "ms_per_step": sec * 1e3,
                   "outputs": "cost, three aligned sequences, median, status (host buffers)"},
           "walks_in_bounds": int((g.status == 0).sum()), "cost_checksum": int(g.cost.astype(np.int64).sum()),
           "extrapolated_seconds_for_100k_triples_per_gpu": sec * 100000 / max(1, len(triples))}
    if not args.skip_cpu and D.world == 1:
        res["cpu_baseline"] = cpu_arm_cube(pool, triples, cm3, args.cpu_triples)
    al.close()
    return res


def _powell_triples(count, n, p, seed):
    """`count` triples (a, two mutated copies of a), length n, substitution / indel rate p: what readjust_3d sees at an
    interior vertex (two children and the parent's sequence)."""
    rng = np.random.default_rng(seed)
    bases = np.array([1, 2, 4, 8], np.uint8)

    def mutate(a):
        r = rng.random(len(a) - 1)
        out = [16]
        for x, q in zip(a[1:], r):
            if q < p / 3:
                continue
            if q < 2 * p / 3:
                out.append(int(rng.choice(bases)))
            out.append(int(rng.choice(bases)) if q < p else int(x))
        return np.array(out, np.uint8)

    seqs = []
    for _ in range(count):
        a = np.concatenate([[16], rng.choice(bases, size=n)]).astype(np.uint8)
        seqs += [a, mutate(a), mutate(a)]
    return seqs


def _powell_ref_worker(job):
    """One process of the CPU arm: the compiled reference's powell_3D_align on its share of the sample."""
    import ctypes as C

    seqs, costs = job
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "libpoyref.so"))
    u8 = C.POINTER(C.c_uint8)
    out = []
    for k in range(0, len(seqs), 3):
        a, b, c = seqs[k:k + 3]
        cap = len(a) + len(b) + len(c)
        rows = [np.zeros(cap, np.uint8) for _ in range(3)]
        n = C.c_int(0)
        out.append(lib.camlrt_powell(a.ctypes.data_as(u8), len(a), b.ctypes.data_as(u8), len(b), c.ctypes.data_as(u8), len(c), *costs,
                                     *[r.ctypes.data_as(u8) for r in rows], C.byref(n)))
    return out


def run_powell_workload(args, rank, local_rank, D, cores):
    """Powell's three-sequence affine aligner (readjust_3d's aligner), 300 bp triples at 3 % divergence,
    through poyb200_batch_powell_3 with host buffers in and out (rows + median).  Only a one-shot entry point exists."""
    from poyd_b200 import cost_matrix as CM, sequence as S

    cm = CM.nucleotides(1, 2, 3)
    count = args.triples
    seqs = _powell_triples(count, 300, 0.03, seed=70 + rank)
    pool = S.SeqPool(seqs)
    triples = np.arange(3 * count, dtype=np.int32).reshape(-1, 3)
    al = S.Align3(cm, CM.of_two_dim(cm), device=local_rank)
    al.align_3_powell_inter(pool, triples)  # warm-up at full size: the workspaces stay with the context
    D.barrier()
    l0 = al.launch_count()
    t0 = time.perf_counter()
    g = al.align_3_powell_inter(pool, triples)
    sec = D.max(time.perf_counter() - t0)
    n_all = D.sum(float(count))
    res = {"workload": "8f #3: Powell 3-D affine Ukkonen aligner (powell_3D_align behind readjust_3d), 300 bp DNA triples, 3 % "
                       "substitutions + indels, costs (1, 3, 2), rows + median", "kernel": "powell_kernel<256,2>",
           "triples_per_gpu_per_step": count, "value": n_all / sec, "unit": "triples/s", "ms_per_step": sec * 1e3,
           "gpu_launches": int(al.launch_count() - l0), "mean_cost": float(g.cost.mean()), "errors": int((g.status != 0).sum()),
           "e2e": {"value": n_all / sec, "unit": "triples/s", "ms_per_step": sec * 1e3, "outputs": "cost, three rows, median (host buffers)"},
           "cost_checksum": int(g.cost.astype(np.int64).sum())}
    if not args.skip_cpu and D.world == 1 and os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "libpoyref.so")):
        import multiprocessing as mp

        procs = cores + 1
        per = 6
        sample = seqs[: 3 * per * procs]
        jobs = [(sample[3 * per * k: 3 * per * (k + 1)], (1, 3, 2)) for k in range(procs)]
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(procs) as pl:
            outs = pl.map(_powell_ref_worker, jobs)
        dt = time.perf_counter() - t0
        flat = [c for o in outs for c in o]
        res["cpu_baseline"] = {"value": len(flat) / dt, "unit": "triples/s", "cores": cores, "processes": procs, "kind": "reference",
                               "sample": f"first {len(flat)} triples over {procs} processes, {dt:.1f} s",
                               "same_costs_as_gpu": bool(np.array_equal(np.array(flat), g.cost[: len(flat)]))}
    al.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU per step (headline workload)")
    ap.add_argument("--secondary-pairs", type=int, default=262_144, help="pairs per GPU per step of the `workloads` block")
    ap.add_argument("--triples", type=int, default=592, help="triples per GPU of the cube workload")
    ap.add_argument("--cpu-triples", type=int, default=6)
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the `workloads`, `sharded_call` and gather blocks")
    ap.add_argument("--taxa", type=int, default=150, help="tree workload: taxa")
    ap.add_argument("--bp", type=int, default=1500, help="tree workload: bases per taxon")
    ap.add_argument("--spr", type=int, default=400, help="tree workload: SPR neighbours evaluated exactly, in lockstep")
    ap.add_argument("--spr-rounds", type=int, default=4, help="cfg5 workload: accepted SPR moves at most")
    ap.add_argument("--spr-window", type=int, default=16, help="cfg5 workload: candidates evaluated exactly per lockstep call")
    ap.add_argument("--workload", default="affine500", choices=sorted(WORKLOADS) + ["cfg5"],
                    help="the headline line's workload (affine500 = BASELINE.json configs[1])")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = host_threads()

    if args.workload == "tree":
        return bench_tree(args, rank, local_rank, world, cores)
    if args.workload == "cfg5":
        if args.impl == "reference":
            args.workload = "tree"
            return bench_tree(args, rank, local_rank, world, cores)
        return bench_cfg5(args, rank, local_rank, world, cores)

    wl_desc, wl_mode, ops_per_cell, wl_kernel = WORKLOADS[args.workload]

    def config_of(n, cells):
        return {"workload": wl_desc, "pairs_per_gpu_per_step": n, "cells_per_gpu_per_step": cells, "sharding": f"pairs x{args.gpus}",
                "l2": "inputs (pool + direction bands, > 1 GB) exceed the 126 MB L2 between iterations"}

    if args.impl == "reference":
        # rank 0 alone runs the CPU arm; the other ranks exit without work
        if rank != 0:
            return
        cm, pool, pairs, dw = workload(args.pairs, seed=2, name=args.workload)  # the GPU arm's rank-0 batch
        cells = total_cells(pool, pairs, dw)
        sample = args.cpu_sample or max(2000, 1500 * cores)
        base, sec = cpu_arm(cm, pool, pairs, sample, cores, steps=args.steps, warmup=args.warmup, deltaw=dw, mode=wl_mode)
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": config_of(len(pairs), cells),
                "sample_pairs_per_step": min(sample, len(pairs)),
                "note": "CPU arm: each step is a bounded sample of the same workload (the metric is a rate)",
                "cpu_baseline": dict(base),
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa_note = bind_to_gpu_numa_node(local_rank) if world > 1 else "single process: no binding"
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")  # barriers that must not occupy a GPU (see sharded_call below)
    D = Dist(world)

    from poyd_b200 import build as _build, sequence as S

    if rank == 0:
        _build.build()
    D.barrier()

    # ---- headline workload ----------------------------------------------------------------------------------------------
    head = run_pairs_workload(args.workload, args.pairs, args, rank, local_rank, D, sampler=True)
    al, cm, pool, pairs, dw, _, _, cells = head.pop("_objs")
    n = head["pairs_per_gpu_per_step"]

    # ---- roofline of the dominant kernel (rank 0's numbers) -------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    add_g, mm_g, mix_g = al.int32_peak()
    f_ms = head["phase_ms"]["fill"]
    fill_s = f_ms * 1e-3
    # launches per chunk: affine = aff_x2_kernel (pair2, the default) + aff_fast_kernel over the batches it declined + the full
    # ring instance over the batches that one declined + traceback kernel; linear = fill + traceback
    from poyd_b200 import _lib as _L
    per_chunk = (3 + (1 if _L.make_config(None).pair2 else 0)) if wl_mode == 3 else 2
    fill_launches = max(1, head["launches_per_step"] // per_chunk)
    km = kernel_metrics().get(wl_kernel, {})
    clocks = head.pop("clocks", None)
    sm_hz = ((clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))) * 1e6
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    issue_peak = sm_count * 4 * sm_hz  # warp instructions per second: 4 schedulers per SM, one issue per clock each
    gops = cells * ops_per_cell / fill_s * 1e-9
    executed = km.get("warp_instr_per_pair")
    roof = {"bound": "int32_alu", "achieved": gops, "peak": add_g, "unit": "Gop/s", "frac": gops / add_g,
            "peak_source": "builder microbenchmark, measured live in this run: dependent-free add.s32 chains "
                           "(poyb200_int32_peak; no driver-measured INT32 peak exists)",
            "peak_minmax_gops": mm_g, "peak_minplus_mix_gops": mix_g, "ops_per_cell": ops_per_cell,
            "kernel": wl_kernel, "kernel_ms_per_step": f_ms, "launches_per_step": fill_launches,
            "kernel_gcups": cells / fill_s * 1e-9,
            "executed_frac": None if executed is None else executed * n / fill_s / issue_peak,
            "executed_warp_instr_per_pair": executed, "issue_peak_warp_instr_per_s": issue_peak,
            "traffic": None if "dram_bytes_per_pair" not in km else km["dram_bytes_per_pair"] * n / fill_launches,
            "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum per pair of this round's "
                            "capture, profiles/r02_kernel_metrics.json, x pairs per launch)",
            "note": "ops_per_cell is the REFERENCE's operation count per cell (SURVEY.md 8d), i.e. algorithmic work; the kernel "
                    "executes fewer instructions than that, so frac can exceed 1; executed_frac = executed warp instructions "
                    "per second / issue slots per second is the utilisation"}
    # HBM view of the fill kernel: both operands in, one direction byte per band cell out (the band is re-read by the
    # traceback kernel, not by this one)
    alg_bytes = float(pool.len[pairs[:, 0]].astype(np.int64).sum() + pool.len[pairs[:, 1]].astype(np.int64).sum()) + float(cells)
    roof_hbm = {"bound": "hbm", "achieved": alg_bytes / fill_s * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / fill_s * 1e-9 / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes / fill_launches,
                "traffic": roof["traffic"], "peak_source": hbm_src, "note": "not the bound: integer min-plus work is ALU-bound"}
    al.close()
    del pool, pairs

    # ---- every other BASELINE config ----------------------------------------------------------------------------------------
    blocks = {}
    if not args.headline_only:
        saved = (args.steps, args.warmup)
        args.steps, args.warmup = max(2, min(args.steps, 3)), 3
        for name in SECONDARY:
            if name == args.workload:
                continue
            r = run_pairs_workload(name, args.secondary_pairs, args, rank, local_rank, D, payloads=("dos",))
            al2, cm2, pool2, pairs2, dw2, mode2, _, _ = r.pop("_objs")
            al2.close()
            if rank == 0 and not args.skip_cpu and world == 1:
                r["cpu_baseline"], _ = cpu_arm(cm2, pool2, pairs2, max(800, 400 * cores), cores, deltaw=dw2, mode=mode2)
            blocks[name] = r
            del pool2, pairs2
        blocks["triples300"] = run_cube_workload(args, rank, local_rank, D)
        blocks["powell300"] = run_powell_workload(args, rank, local_rank, D, cores)
        args.steps, args.warmup = saved

    # ---- N > 1: the two multi-GPU paths ----------------------------------------------------------------------------------------
    gather_block = sharded_block = None
    if world > 1 and not args.headline_only:
        from poyd_b200 import sharding

        # (1) across processes: shard one common batch by work, every rank runs its shard on its GPU, results gathered on rank 0,
        #     cost total by the one collective of the design (all_reduce over NCCL)
        cmg, poolg, pairsg, _ = workload(65_536 * world, seed=77, name="affine500")
        work = (poolg.len[pairsg[:, 0]].astype(np.int64) + poolg.len[pairsg[:, 1]])
        alg = S.Align(cmg, device=local_rank)

        def worker(idx):
            g = alg.align_affine_3(poolg, pairsg[idx], S.WANT_MEDIAN)
            return {"cost": g.cost, "median_len": g.lens[:, 0].copy()}

        D.barrier()
        t0 = time.perf_counter()
        got, idx = sharding.run_sharded(worker, work)
        total = sharding.cost_sum(worker(idx)["cost"])
        sec = D.max(time.perf_counter() - t0)
        if rank == 0:
            assert int(got["cost"].astype(np.int64).sum()) == total
            gather_block = {"pairs": len(pairsg), "ranks": world, "seconds": sec, "cost_total_allreduce": total,
                            "what": "poyd_b200/sharding.py run_sharded + cost_sum on the GPUs (NCCL): shard by work, gather on "
                                    "rank 0, all_reduce of the shard cost totals; checked against the gathered costs"}
        alg.close()
        D.barrier()
        # While rank 0 drives every GPU from one process, the other ranks must not hold a kernel on theirs: an NCCL barrier
        # is a spinning kernel, and two processes on one GPU time-slice.  They wait on the host (gloo) instead.
        # (2) inside one process: rank 0 alone runs ONE batch (world x pairs, at most 2 M pairs) through poyb200_multi_batch
        #     over all GPUs (host thread per device, results land in one set of pinned buffers); the other ranks wait
        if rank == 0:
            cmm, poolm, pairsm, _ = workload(min(args.pairs * world, 2_000_000), seed=91, name="affine500")
            cellsm = total_cells(poolm, pairsm, None)
            ma = S.MultiAlign(cmm, list(range(world)))
            batch, keep, info = pinned_batch(S, ma, poolm, pairsm, None, 0, torch)
            d2h = set_outputs(S, batch, info, keep, torch, "dos")
            l0 = ma.launch_count()
            ma._one_shot(3, batch)
            t0 = time.perf_counter()
            reps = max(2, min(args.steps, 3))
            for _ in range(reps):
                ma._one_shot(3, batch)
            sec = (time.perf_counter() - t0) / reps
            sharded_block = {"value": cellsm / sec * 1e-9, "unit": UNIT, "ms_per_step": sec * 1e3, "pairs": len(pairsm),
                             "devices": world, "scaling": "strong", "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": d2h,
                             "gpu_launches": int(ma.launch_count() - l0), "shard_begin": [int(x) for x in ma.shards()],
                             "cost_checksum": int(info["out"]["cost"].numpy().astype(np.int64).sum()),
                             "what": "one poyb200_multi_batch call (csrc/multi.cu) on one batch of N x pairs, DOS.median payload, "
                                     "pinned host buffers in and out: the north_star's shard-by-pair-index + host gather"}
            ma.close()
        torch.cuda.synchronize()
        dist.barrier(group=host_group)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    base = None
    if not args.skip_cpu and world == 1:  # the CPU baseline is an N=1 leg (rank 0, all host cores)
        cm_, pool_, pairs_, dw_ = workload(min(args.pairs, max(4000, 3000 * cores)), seed=2, name=args.workload)
        base, _ = cpu_arm(cm_, pool_, pairs_, args.cpu_sample or max(2000, 1500 * cores), cores, deltaw=dw_, mode=wl_mode)
    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": config_of(n, cells),
        "e2e": head["e2e"], "e2e_dos_median": head["e2e_dos_median"], "numa": numa_note,
        "gpu_launches": head["gpu_launches"], "phase_ms": head["phase_ms"],
        "roofline": roof, "roofline_hbm": roof_hbm, "clocks": clocks, "cost_checksum": head["cost_checksum"],
    }
    if blocks:
        line["workloads"] = blocks
    if gather_block:
        line["rank_sharded_gather"] = gather_block
    if sharded_block:
        line["sharded_call"] = sharded_block
    if base is not None:
        line["cpu_baseline"] = base
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
