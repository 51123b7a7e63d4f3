#!/usr/bin/env python
"""bench.py -- GCUPS of the batched pairwise median DP (BASELINE.json `metric`) on N B200s.

A step = one pass of the hot path (algn_CAML_align_affine_3 for every pair: affine fill in the reference's band,
traceback, median / medianwg / aligned pair) over one batch of synthetic pairs -- configs[1] of BASELINE.json:
500 bp DNA, 10 % substitutions, 2 % indels, subst 1 / indel 2 / gap opening 3.

  value      cells / second with the batch already resident in HBM (kernels only, CUDA events, max over ranks)
  e2e        the same metric through the C ABI with HOST buffers: plan + H2D + kernels + D2H every step
  roofline   the dominant kernel (the affine stripe fill) against the measured HBM peak, as the contract asks,
             plus `roofline_int32`: the same kernel against the measured INT32 ALU throughput, which is the
             bound that actually applies to this integer min-plus recurrence (BASELINE.json north_star)
  cpu_baseline  the compiled reference algn.c (oracle/_ref) on all host cores, on a bounded sample

`--impl reference` times the reference's own CPU implementation instead (the driver computes the ratio).
"""
from __future__ import annotations

import argparse
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
OPS_PER_CELL = 50  # integer ops per affine_3 cell with traceback, counted from src/algn.c


NCU_DRAM_BYTES_PER_PAIR = {"affine500": 6.617161e9 / 100000}

WORKLOADS = {
    # name: (description, mode, ops per cell)
    "affine500": ("configs[1]: 1M DNA pairs 500 bp (10% subst, 2% indel), affine gaps (subst 1, indel 2, gap opening 3), "
                  "align_affine_3 = fill + traceback + median/medianwg/aligned pair", 3, 50),
    "affine500_medianlike": ("configs[1], median-like operands: as affine500 plus 0.5% IUPAC ambiguities and 10% of positions "
                             "carrying the gap bit (what internal-node medians look like; exercises the block-diagonal "
                             "state)", 3, 50),
    "linear500": ("cfg 2-lin: DNA pairs 500 bp, linear gaps (subst 1, indel 2), deltaw as Sequence.Align.cost_2 computes it, "
                  "align_2 + ancestor_2 + median_2_with_gaps", 1, 10),
    "protein300": ("configs[2] (3a): protein pairs 300 aa, 22x22 matrix 1/2, deltaw as the product computes it "
                   "(full matrix, SURVEY.md A15), align_2 + medians", 1, 10),
    "protein300_band16": ("configs[2] (3b): protein pairs 300 aa, explicit deltaw 16, align_2 + medians", 1, 10),
    "tree": ("configs[4]-style host-driver workload (one GPU per tree): Wagner build with batched candidate-edge sweeps, "
             "all-directions downpass, root selection, single assignment and adjusted cost of the built tree, then --spr SPR "
             "neighbours evaluated exactly in lockstep (poyd_b200/tree.py); "
             "synthetic --taxa x --bp DNA data set, affine gaps (1, 2, opening 3)", 3, 50),
}


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
    """Cells the reference visits for the batch: a pure function of the two lengths (and deltaw), SURVEY.md 8d."""
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


def cpu_arm(cm, pool, pairs, sample: int, threads: int, steps: int = 1, warmup: int = 0, deltaw=None, mode: int = 3):
    """Times the reference's CPU implementation (compiled algn.c if oracle/_ref exists, else the port)."""
    from oracle import oracle

    oracle.build(ref=True)
    chk = oracle.best_checker(cm)
    sample = min(sample, len(pairs))
    sub = pairs[:sample]
    dws = None if deltaw is None else deltaw[:sample]
    cells = total_cells(pool, sub, dws)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        chk.batch(mode, pool.pool, pool.off, pool.len, sub, deltaw=dws, nthreads=threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"value": cells / sec * 1e-9, "unit": UNIT, "cores": threads, "kind": chk.kind,
            "sample": f"first {sample} pairs of the workload, all outputs, {threads} threads"}, sec


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


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Multi-rank runs: pin this process (and so its pinned host buffers, first touch) to the CPUs of the NUMA node its GPU
    hangs off, so the 8 ranks' host copies do not all cross one socket.  No-op where sysfs has no answer."""
    try:
        import subprocess

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--taxa", type=int, default=150, help="tree workload: taxa")
    ap.add_argument("--bp", type=int, default=1500, help="tree workload: bases per taxon")
    ap.add_argument("--spr", type=int, default=400, help="tree workload: SPR neighbours evaluated exactly, in lockstep")
    ap.add_argument("--workload", default="affine500", choices=sorted(WORKLOADS),
                    help="affine500 is the headline configuration; the others are reported in DESIGN.md")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = host_threads()

    if args.workload == "tree":
        return bench_tree(args, rank, local_rank, world, threads)

    if args.impl == "reference":
        # rank 0 alone runs the CPU arm; the other ranks exit without work
        if rank != 0:
            return
        sample = args.cpu_sample or max(2000, 1500 * threads)
        cm, pool, pairs, dw = workload(sample, seed=2, name=args.workload)
        base, sec = cpu_arm(cm, pool, pairs, sample, threads, steps=args.steps, warmup=args.warmup, deltaw=dw,
                            mode=WORKLOADS[args.workload][1])
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": WORKLOADS[args.workload][0], "pairs_per_step": sample,
                           "note": "CPU arm: bounded sample of the same workload per step"},
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from poyd_b200 import build as _build, sequence as S

    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()

    cm, pool, pairs, dw = workload(args.pairs, seed=2 + rank, name=args.workload)  # each rank owns its own pairs
    wl_desc, wl_mode, ops_per_cell = WORKLOADS[args.workload]
    cells = total_cells(pool, pairs, dw)
    al = S.Align(cm, device=local_rank)
    want = S.WANT_MEDIAN | S.WANT_MEDIANWG | S.WANT_ALIGNED

    # pinned host buffers for the end-to-end leg
    def pinned(arr):
        return torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()

    ppool = S.SeqPool.__new__(S.SeqPool)
    keep = [pinned(pool.pool), pinned(pool.off), pinned(pool.len), pinned(pairs)]
    ppool.pool, ppool.off, ppool.len = keep[0].numpy(), keep[1].numpy(), keep[2].numpy()
    ppairs = keep[3].numpy()
    if dw is not None:
        keep.append(pinned(dw))
    batch, res = al.make_batch(ppool, ppairs, deltaw=None if dw is None else keep[-1].numpy(), want=want, outputs=False)
    n = len(ppairs)
    stride = (int((ppool.len[ppairs[:, 0]].astype(np.int64) + ppool.len[ppairs[:, 1]]).max()) + 2 + 15) // 16 * 16
    out_t = {k: torch.empty((n, stride), dtype=torch.uint8, pin_memory=True) for k in ("median", "medianwg", "a", "b")}
    cost_t = torch.empty(n, dtype=torch.int32, pin_memory=True)
    lens_t = torch.empty((n, 4), dtype=torch.int32, pin_memory=True)
    batch.cost, batch.out_len, batch.out_stride = cost_t.data_ptr(), lens_t.data_ptr(), stride
    batch.median, batch.medianwg = out_t["median"].data_ptr(), out_t["medianwg"].data_ptr()
    batch.aligned_a, batch.aligned_b = out_t["a"].data_ptr(), out_t["b"].data_ptr()
    h2d = int(ppool.pool.nbytes + n * 64)
    d2h = int(4 * n * stride + n * 4 + n * 16)

    stream = torch.cuda.ExternalStream(al.L.poyb200_stream(al.h), device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident leg: stage once, time K passes of the kernels ------------------------------------
    al.stage(wl_mode, batch)
    al.sync()
    for _ in range(args.warmup):
        al.run()
    al.sync()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = al.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fill_ms, trace_ms = 0.0, 0.0
    e0.record(stream)
    for _ in range(args.steps):
        al.run()
    e1.record(stream)
    al.sync()
    barrier()
    launches = al.launch_count() - launches0
    dev_ms = e0.elapsed_time(e1) / args.steps
    f_ms, t_ms = al.last_run_ms()  # the last pass, per phase
    # launches per chunk: affine = aff_fast_kernel + aff_stripe_kernel over the declined list + traceback; linear = fill + traceback
    fill_launches = max(1, (launches // args.steps) // (3 if wl_mode == 3 else 2))
    dev_ms_max = max_over_ranks(dev_ms)
    total_cells_all = sum_over_ranks(float(cells))
    value = total_cells_all / (dev_ms_max * 1e-3) * 1e-9

    # ---- end-to-end leg: host buffers in, host buffers out, every step ---------------------------------------
    import ctypes

    def e2e_call():
        # the call a user of the C ABI makes: host buffers in, host buffers out
        fn = al.L.poyb200_batch_align_affine_3 if wl_mode == 3 else al.L.poyb200_batch_align_2
        al._check(fn(al.h, ctypes.byref(batch)))

    for _ in range(max(1, args.warmup - 1)):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    barrier()
    clocks = sampler.stop() if sampler else None
    e2e_max = max_over_ranks(e2e_s)
    e2e_value = total_cells_all / e2e_max * 1e-9
    checksum = int(cost_t.numpy().astype(np.int64).sum())

    # ---- end-to-end leg 2: the payload SeqCS.DOS.median keeps (cost, median, three gap bitsets; src/seqCS.ml:769-776)
    # instead of the four sequences -- same kernels, same traceback, smaller result
    bstride = ((stride + 7) // 8 + 3) // 4 * 4
    bits_t = [torch.empty((n, bstride), dtype=torch.uint8, pin_memory=True) for _ in range(3)]
    batch.want = S.WANT_MEDIAN | S.WANT_BITSETS
    batch.medianwg = batch.aligned_a = batch.aligned_b = None
    batch.bits_a, batch.bits_b, batch.bits_wg = (t.data_ptr() for t in bits_t)
    batch.bits_stride = bstride
    d2h_dos = int(n * stride + 3 * n * bstride + n * 4 + n * 16)
    for _ in range(max(1, args.warmup - 1)):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()
    torch.cuda.synchronize()
    dos_s = (time.perf_counter() - t0) / args.steps
    barrier()
    dos_max = max_over_ranks(dos_s)
    dos_value = total_cells_all / dos_max * 1e-9
    assert int(cost_t.numpy().astype(np.int64).sum()) == checksum

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    add_g, mm_g, mix_g = al.int32_peak()
    # algorithmic bytes of the fill kernel per pair: both operands in, one direction byte per band cell out
    # (the band is re-read by the traceback kernel, not by this one)
    alg_bytes = float(pool.len[pairs[:, 0]].astype(np.int64).sum() + pool.len[pairs[:, 1]].astype(np.int64).sum()) + float(cells)
    fill_s = f_ms * 1e-3
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture (profiles/r01_fill_fast.csv:
    # dram__bytes_read.sum + dram__bytes_write.sum = 6.617 GB for one launch over 100 000 pairs of this workload),
    # scaled to the pairs one launch of this run covers; null for workloads without a capture.
    traffic = NCU_DRAM_BYTES_PER_PAIR.get(args.workload)
    if traffic is not None:
        traffic = traffic * n / fill_launches
    roof = {"bound": "hbm", "achieved": alg_bytes / fill_s * 1e-9, "peak": hbm_peak, "unit": "GB/s",
            "frac": alg_bytes / fill_s * 1e-9 / hbm_peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu, profiles/r01_fill_fast.csv)",
            "algorithmic_bytes_per_launch": alg_bytes / fill_launches, "peak_source": hbm_src,
            "kernel": ("aff_fast_kernel<5,8,true>" if args.workload == "affine500" else "aff_stripe_kernel<K,G,true>") if wl_mode == 3 else "lin_stripe_kernel<K,G,true>", "kernel_ms_per_step": f_ms, "launches_per_step": fill_launches,
            "note": "integer min-plus recurrence: ALU-bound, see roofline_int32"}
    gops = cells * ops_per_cell / fill_s * 1e-9
    roof_int = {"bound": "int32_alu", "achieved": gops, "peak": add_g, "unit": "Gop/s", "frac": gops / add_g,
                "ops_per_cell": ops_per_cell, "peak_source": "measured live: dependent-free add.s32 chains (poyb200_int32_peak)",
                "peak_minmax_gops": mm_g, "peak_minplus_mix_gops": mix_g, "kernel_gcups": cells / fill_s * 1e-9,
                "note": "ops_per_cell is the REFERENCE's operation count per cell (SURVEY.md 8d), i.e. algorithmic work; the kernel "
                        "executes fewer (ncu: 30.1 warp instructions per 32 cells incl. loads/stores on affine500, "
                        "profiles/r01_fill_fast.csv), so frac can exceed 1; the executed-instruction view is ALU pipe 74 % / "
                        "issue 81 % active"}
    base = None
    if not args.skip_cpu and world == 1:  # the CPU baseline is an N=1 leg (rank 0, all host threads)
        sample = args.cpu_sample or max(2000, 1500 * threads)
        base, _ = cpu_arm(cm, pool, pairs, sample, threads, deltaw=dw, mode=wl_mode)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": wl_desc,
                   "pairs_per_gpu_per_step": n, "cells_per_gpu_per_step": cells, "sharding": f"pairs x{world}",
                   "l2": "inputs (pool + direction bands, > 1 GB) exceed the 126 MB L2 between iterations"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_max * 1e3, "outputs": "cost, median, medianwg, aligned a, aligned b (pinned host buffers)"},
        "e2e_dos_median": {"value": dos_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_dos,
                           "ms_per_step": dos_max * 1e3,
                           "outputs": "what SeqCS.DOS.median keeps (src/seqCS.ml:769-776): cost, median, gap bitsets of "
                                      "aligned a / aligned b / medianwg (POYB200_WANT_BITSETS); same kernels, the aligned "
                                      "sequences stay on the device"},
        "numa": numa_note,
        "gpu_launches": int(launches),
        "phase_ms": {"fill": f_ms, "traceback": t_ms},
        "roofline": roof, "roofline_int32": roof_int, "clocks": clocks, "cost_checksum": checksum,
    }
    if base is not None:
        line["cpu_baseline"] = base
    print(json.dumps(line))
    al.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
