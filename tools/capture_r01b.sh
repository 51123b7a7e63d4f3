#!/bin/bash
# One GPU call: GPU tests, both bench arms, the ncu launch list and full captures of the fast fill + traceback kernels.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo "bench rc=$?"; cat gpurun_out/bench_b200.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --pairs 200000 --steps 2 --warmup 3 --skip-cpu > gpurun_out/launches_r01b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:aff_fast -s 4 -c 1 -o gpurun_out/prof_fast_r01b python bench.py --pairs 100000 --steps 1 --warmup 3 --skip-cpu > gpurun_out/prof_r01b.log 2>&1
ls -la gpurun_out | tail -12
