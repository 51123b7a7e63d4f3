"""Host-link probe for the multi-GPU end-to-end numbers: aggregate H2D / D2H bandwidth with 1 and with all ranks active.
Run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
GB = 1 << 30
host = torch.empty(GB, dtype=torch.uint8, pin_memory=True)
host2 = torch.empty(GB, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(GB, dtype=torch.uint8, device="cuda")
dev2 = torch.empty(GB, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()


def bar():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(kind, active):
    bar()
    t0 = time.perf_counter()
    if active:
        for _ in range(4):
            if kind in ("d2h", "both"):
                host.copy_(dev, non_blocking=True)
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s2):
                    dev2.copy_(host2, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt if active else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {}
for kind in ("d2h", "h2d", "both"):
    mult = 2 if kind == "both" else 1
    run(kind, True)
    t_all = run(kind, True)
    t_one = run(kind, rank == 0)
    out[kind] = {"all_ranks_GBps": round(world * 4 * mult / t_all, 1), "rank0_alone_GBps": round(4 * mult / t_one, 1)}
if rank == 0:
    numa = {}
    try:
        import subprocess
        ids = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.split()
        for i, b in enumerate(ids):
            p = "/sys/bus/pci/devices/" + b.lower()[4:] + "/numa_node"
            numa[i] = open(p).read().strip() if os.path.exists(p) else "?"
    except Exception as e:
        numa = {"error": str(e)}
    out["gpu_numa_node"] = numa
    out["cpus"] = len(os.sched_getaffinity(0))
    print(json.dumps({"pcie_probe": out, "world": world}))
if world > 1:
    dist.destroy_process_group()
