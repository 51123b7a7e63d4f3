#!/bin/bash
# Round 2, GPU call X: how the concurrent traceback costs the fill -- walkers per SM, serial traceback, chunk size (config only).
mkdir -p gpurun_out
for cfg in "" "traceback_threads_per_sm=256" "traceback_threads_per_sm=128" "traceback_threads_per_sm=1024" "overlap_traceback=0" "chunk_pairs=131072" "traceback_block=64,traceback_threads_per_sm=256"; do
  POYB200_CONFIG="$cfg" timeout 300 python bench.py --skip-cpu --headline-only > gpurun_out/r02x_tmp.json 2> gpurun_out/r02x_tmp.err; echo -n "affine500 [$cfg] rc=$? "
  python -c "
import json; d=json.load(open('gpurun_out/r02x_tmp.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['phase_ms'])"
done 2>&1 | tee gpurun_out/r02x_tb.log
for cfg in "" "traceback_threads_per_sm=256" "traceback_threads_per_sm=128" "overlap_traceback=0" "chunk_pairs=131072"; do
  POYB200_CONFIG="$cfg" timeout 300 python bench.py --workload linear500 --pairs 262144 --skip-cpu --headline-only > gpurun_out/r02x_tmp.json 2> gpurun_out/r02x_tmp.err; echo -n "linear500 [$cfg] rc=$? "
  python -c "
import json; d=json.load(open('gpurun_out/r02x_tmp.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['phase_ms'])"
done 2>&1 | tee -a gpurun_out/r02x_tb.log
