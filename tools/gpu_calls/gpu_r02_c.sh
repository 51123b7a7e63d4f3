#!/bin/bash
# Round 2, GPU call C: ring kernels with the staged (cp.async) walk window.
mkdir -p gpurun_out
timeout 300 python tools/ring_check.py > gpurun_out/r02c_ring_check.log 2>&1; echo "ring_check rc=$?"; grep -E "mismatch|differ|RING" gpurun_out/r02c_ring_check.log | sort | uniq -c | sort -rn | head -8
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02c_pytest.log
timeout 600 python bench.py --skip-cpu --headline-only > gpurun_out/r02c_bench_head.json 2> gpurun_out/r02c_bench_head.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_head.json')); print('ring:', d['value'], d['ms_per_step'], d['phase_ms'], 'e2e', d['e2e']['value'], d['e2e_dos_median']['value'], d['gpu_launches'])"
timeout 600 python bench.py --skip-cpu --headline-only --workload affine500_medianlike > gpurun_out/r02c_bench_ml.json 2> gpurun_out/r02c_bench_ml.err; python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_ml.json')); print('ring medianlike:', d['value'], d['ms_per_step'], d['phase_ms'], 'e2e', d['e2e']['value'], d['e2e_dos_median']['value'])"
export POYB200_CONFIG=chunk_pairs=1048576
timeout 400 ncu --set full --clock-control none --import-source on -k regex:aff_ring -s 6 -c 1 -o gpurun_out/r02c_prof_ring python bench.py --pairs 100000 --steps 1 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02c_prof.log 2>&1
ls -la gpurun_out | grep r02c
