#!/bin/bash
# Round 2, GPU call T: lin_stripe_kernel occupancy variants (launch bounds 2 / 3 / 4 CTAs per SM) and an ncu capture of it.
mkdir -p gpurun_out
for v in "" build/lib_lin3.so build/lib_lin4.so; do
  for wl in linear500 protein300_band16; do
    POYB200_SO=$v timeout 300 python bench.py --workload $wl --pairs 262144 --skip-cpu --headline-only > gpurun_out/r02t_tmp.json 2> gpurun_out/r02t_tmp.err; echo -n "[${v:-default}] $wl rc=$? "
    python -c "
import json; d=json.load(open('gpurun_out/r02t_tmp.json')); print(round(d['value'],1), round(d['ms_per_step'],2), d['phase_ms'])"
  done
done 2>&1 | tee gpurun_out/r02t_variants.log
POYB200_CONFIG="chunk_pairs=1048576" timeout 600 ncu --set full --import-source on --clock-control none -k regex:lin_stripe_kernel -c 1 -o gpurun_out/r02t_linstripe python bench.py --workload linear500 --pairs 131072 --steps 1 --warmup 1 --skip-cpu --headline-only > gpurun_out/r02t_ncu.log 2>&1; echo "ncu rc=$?"
