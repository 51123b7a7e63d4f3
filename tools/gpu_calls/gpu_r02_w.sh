#!/bin/bash
# Round 2, GPU call W: linear kernels without a boundary phase when prepend / tail costs are the defaults -- parity and timing.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02w_pytest.log
for wl in linear500 protein300 protein300_band16; do
  timeout 300 python bench.py --workload $wl --pairs 262144 --skip-cpu --headline-only > gpurun_out/r02w_$wl.json 2> gpurun_out/r02w_tmp.err; echo -n "$wl rc=$? "
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['phase_ms'])" gpurun_out/r02w_$wl.json
done
