#!/bin/bash
# Round 2, GPU call R: 6-bit direction band for the default affine path -- parity, A/B timing, ncu capture.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02r_pytest.log
for cfg in "" "dir6=0" ""; do
  POYB200_CONFIG="$cfg" timeout 600 python bench.py --skip-cpu --headline-only > "gpurun_out/r02r_bench_${cfg:-default}.json" 2> gpurun_out/r02r_bench.err; echo "bench[$cfg] rc=$?"
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['phase_ms'], d['roofline']['kernel_ms_per_step'])" "gpurun_out/r02r_bench_${cfg:-default}.json"
done
POYB200_CONFIG="chunk_pairs=1048576" timeout 600 ncu --set full --import-source on --clock-control none -k regex:aff_fast_kernel -c 1 -o gpurun_out/r02r_fast6 python bench.py --pairs 100000 --steps 1 --warmup 1 --skip-cpu --headline-only > gpurun_out/r02r_ncu.log 2>&1; echo "ncu rc=$?"
POYB200_CONFIG="chunk_pairs=1048576" timeout 600 ncu --set full --clock-control none -k regex:aff_traceback_kernel -c 1 -o gpurun_out/r02r_tb6 python bench.py --pairs 100000 --steps 1 --warmup 1 --skip-cpu --headline-only > gpurun_out/r02r_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out | grep r02r
