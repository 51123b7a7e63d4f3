#!/bin/bash
# Round 2, GPU call Y: 256 traceback walkers per SM as the default -- the other workloads, priority and 384 as variants.
mkdir -p gpurun_out
for cfg in "" "traceback_priority=0" "traceback_threads_per_sm=384" "traceback_threads_per_sm=512"; do
  POYB200_CONFIG="$cfg" timeout 300 python bench.py --skip-cpu --headline-only > gpurun_out/r02y_tmp.json 2> gpurun_out/r02y_tmp.err; echo -n "affine500 [$cfg] rc=$? "
  python -c "
import json; d=json.load(open('gpurun_out/r02y_tmp.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['phase_ms'])"
done 2>&1 | tee gpurun_out/r02y_tb.log
for wl in affine500_medianlike protein300 protein300_band16; do
  for cfg in "" "traceback_threads_per_sm=512"; do
    POYB200_CONFIG="$cfg" timeout 300 python bench.py --workload $wl --pairs 262144 --skip-cpu --headline-only > gpurun_out/r02y_tmp.json 2> gpurun_out/r02y_tmp.err; echo -n "$wl [$cfg] rc=$? "
    python -c "
import json; d=json.load(open('gpurun_out/r02y_tmp.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['phase_ms'])"
  done
done 2>&1 | tee -a gpurun_out/r02y_tb.log
