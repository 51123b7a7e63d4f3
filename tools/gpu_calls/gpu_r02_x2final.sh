#!/bin/bash
# Round 2, last 1-GPU call: whole parity suite and smoke on the build with aff_x2_kernel, default bench, ncu evidence
# for the new dominant kernel (one --set full capture, launch list of the headline command).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02x2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02x2f_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02x2f_bench.json 2> gpurun_out/r02x2f_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02x2f_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02x2f_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['kernel'], d['gpu_launches']); print({k:(round(v['value'],1), v.get('unit','GCUPS')) for k,v in d['workloads'].items()}); print(d['cpu_baseline'])"
export POYB200_CONFIG=chunk_pairs=1048576
timeout 200 ncu --set full --clock-control none --import-source on -k regex:aff_x2 -s 3 -c 1 -o gpurun_out/r02x2f_prof python bench.py --pairs 100000 --steps 1 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02x2f_prof.log 2>&1; echo "ncu rc=$?"
unset POYB200_CONFIG
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02x2f_launches.csv python bench.py --pairs 200000 --steps 2 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02x2f_launches.log 2>&1; echo "launch list rc=$?"
