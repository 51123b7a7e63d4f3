#!/bin/bash
# Round 2, closing call: default bench.py on the final build (both edits since the last full call: x2_usable guard, cube warm-up).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 170 python bench.py > gpurun_out/r02x2g_bench.json 2> gpurun_out/r02x2g_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02x2g_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02x2g_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['kernel'], d['gpu_launches'], d['clocks']); print({k:(round(v['value'],1), round(v['ms_per_step'],1)) for k,v in d['workloads'].items()})"
