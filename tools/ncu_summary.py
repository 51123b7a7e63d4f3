#!/usr/bin/env python
"""Reads `ncu --set full` captures (.ncu-rep) and writes the per-pair figures bench.py quotes
(profiles/r02_kernel_metrics.json): DRAM bytes and executed warp instructions per pair, registers, duration, pipe
utilisation.  usage: tools/ncu_summary.py OUT.json NAME=REPORT.ncu-rep:PAIRS_IN_THE_CAPTURED_LAUNCH ..."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
    "smsp__inst_executed.sum": "warp_instr", "gpu__time_duration.sum": "duration_ns",
    "launch__registers_per_thread": "registers", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__occupancy_limit_registers": "occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instr",
}


def read(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, vals = rows[0], rows[1], rows[2]
    res = {"kernel_name": vals[head.index("Kernel Name")]}
    for k, name in WANT.items():
        if k in head:
            i = head.index(k)
            v = float(vals[i].replace(",", ""))
            u = units[i]
            if name.startswith("dram_"):
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            if name == "duration_ns":
                v *= {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1}.get(u, 1)
            res[name] = v
    return res


def main():
    out_path, specs = sys.argv[1], sys.argv[2:]
    try:
        data = json.load(open(out_path))
    except Exception:  # noqa: BLE001
        data = {}
    for spec in specs:
        name, rest = spec.split("=", 1)
        report, pairs = rest.rsplit(":", 1)
        r = read(report)
        pairs = int(pairs)
        r["pairs_in_launch"] = pairs
        if "dram_read_bytes" in r:
            r["dram_bytes_per_pair"] = (r["dram_read_bytes"] + r["dram_write_bytes"]) / pairs
        if "warp_instr" in r:
            r["warp_instr_per_pair"] = r["warp_instr"] / pairs
        r["source"] = report
        data[name] = r
    json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
