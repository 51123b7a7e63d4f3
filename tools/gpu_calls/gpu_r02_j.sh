#!/bin/bash
# Round 2, GPU call J: full parity suite with the column-striped kernel in place, default bench, ncu capture of lin_rows_kernel.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02j_pytest.log
timeout 900 python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02j_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['traffic']); print({k:(round(v['value'],1), round(v.get('e2e_dos_median',v.get('e2e'))['value'],1), round(v.get('cpu_baseline',{}).get('value',0),2)) for k,v in d['workloads'].items()})"
POYB200_CONFIG="chunk_pairs=1048576" timeout 600 ncu --set full --import-source on --clock-control none -k regex:lin_rows_kernel -c 1 -o gpurun_out/r02j_rows python bench.py --workload protein300 --pairs 131072 --steps 1 --warmup 1 --skip-cpu --headline-only > gpurun_out/r02j_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | grep r02j
