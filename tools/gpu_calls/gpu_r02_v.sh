#!/bin/bash
# Round 2, GPU call V: chunk size against the 262 144-pair linear workloads (tails of the tapered chunks).
mkdir -p gpurun_out
for cp in 65536 131072 262144; do
  for wl in linear500 protein300; do
    POYB200_CONFIG="chunk_pairs=$cp" timeout 300 python bench.py --workload $wl --pairs 262144 --skip-cpu --headline-only > gpurun_out/r02v_tmp.json 2> gpurun_out/r02v_tmp.err; echo -n "chunk_pairs=$cp $wl rc=$? "
    python -c "
import json; d=json.load(open('gpurun_out/r02v_tmp.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['phase_ms'], d['gpu_launches'])"
  done
done 2>&1 | tee gpurun_out/r02v_chunks.log
POYB200_CONFIG="chunk_pairs=262144" timeout 300 python bench.py --pairs 1000000 --skip-cpu --headline-only > gpurun_out/r02v_tmp.json 2> gpurun_out/r02v_tmp.err; python -c "
import json; d=json.load(open('gpurun_out/r02v_tmp.json')); print('affine500 chunk 262144:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['phase_ms'], d['gpu_launches'])"
