#!/bin/bash
# One GPU call: the pending racecheck diagnostic (plain-load staging) + the ncu evidence of the current kernels.
mkdir -p gpurun_out
POYB200_SO=build/lib_notma.so timeout 450 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/racecheck_notma2.log 2>&1
echo "notma2: $(grep -E 'RACECHECK SUMMARY' gpurun_out/racecheck_notma2.log) races=$(grep -cE 'Race reported' gpurun_out/racecheck_notma2.log)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --pairs 200000 --steps 2 --warmup 3 --skip-cpu > gpurun_out/launches_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aff_stripe -s 4 -c 1 -o gpurun_out/prof_fill_final python bench.py --pairs 100000 --steps 1 --warmup 3 --skip-cpu > gpurun_out/prof_final.log 2>&1
ncu --set full --clock-control none -k regex:aff_traceback -s 4 -c 1 -o gpurun_out/prof_trace_final python bench.py --pairs 200000 --steps 1 --warmup 3 --skip-cpu >> gpurun_out/prof_final.log 2>&1
ls gpurun_out | tail -8
