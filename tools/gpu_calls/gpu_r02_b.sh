#!/bin/bash
# Round 2, GPU call B: first run of the ring kernels (fill + walk in one kernel): parity, then speed.
mkdir -p gpurun_out
timeout 300 python tools/ring_check.py > gpurun_out/r02b_ring_check.log 2>&1; echo "ring_check rc=$?"; tail -25 gpurun_out/r02b_ring_check.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02b_pytest.log
timeout 600 python bench.py --skip-cpu --headline-only > gpurun_out/r02b_bench_head.json 2> gpurun_out/r02b_bench_head.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_head.json')); print('ring:', d['value'], d['ms_per_step'], d['phase_ms'], 'e2e', d['e2e']['value'], d['e2e_dos_median']['value'], d['gpu_launches'])"
timeout 600 python bench.py --skip-cpu --headline-only --workload affine500_medianlike > gpurun_out/r02b_bench_ml.json 2> gpurun_out/r02b_bench_ml.err; python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_ml.json')); print('ring medianlike:', d['value'], d['ms_per_step'], d['phase_ms'], 'e2e', d['e2e']['value'], d['e2e_dos_median']['value'])"
for v in eb3; do
POYB200_SO=build/lib_$v.so timeout 600 python bench.py --skip-cpu --headline-only --workload affine500_medianlike > gpurun_out/r02b_bench_ml_$v.json 2> gpurun_out/r02b_bench_ml_$v.err; python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_ml_$v.json')); print('$v medianlike:', d['value'], d['ms_per_step'], d['phase_ms'])"
done
POYB200_CONFIG=use_ring=0 timeout 600 python bench.py --skip-cpu --headline-only > gpurun_out/r02b_bench_legacy.json 2> gpurun_out/r02b_bench_legacy.err; python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_legacy.json')); print('legacy:', d['value'], d['ms_per_step'], d['phase_ms'])"
export POYB200_CONFIG=chunk_pairs=1048576
timeout 400 ncu --set full --clock-control none --import-source on -k regex:aff_ring -s 6 -c 1 -o gpurun_out/r02b_prof_ring python bench.py --pairs 100000 --steps 1 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02b_prof.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:aff_ring -s 7 -c 1 -o gpurun_out/r02b_prof_ring_ml python bench.py --workload affine500_medianlike --pairs 100000 --steps 1 --warmup 3 --skip-cpu --headline-only >> gpurun_out/r02b_prof.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --pairs 200000 --steps 2 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02b_launches.log 2>&1
ls -la gpurun_out | grep r02b
