#!/bin/bash
# Round 2, GPU call A: parity suite, racecheck (control program + product kernels), both bench arms, ncu evidence.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02a_pytest.log
for m in 0 1 2; do
  timeout 300 compute-sanitizer --tool racecheck build/racecheck_control $m > gpurun_out/r02a_racecheck_control_$m.log 2>&1
  echo "control mode $m: $(grep -E 'racecheck_control mode' gpurun_out/r02a_racecheck_control_$m.log) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/r02a_racecheck_control_$m.log)"
done
timeout 300 compute-sanitizer --tool racecheck build/racecheck_control 0 1 > gpurun_out/r02a_racecheck_control_0_oneslot.log 2>&1
echo "control mode 0 one slot: $(grep -E 'RACECHECK SUMMARY' gpurun_out/r02a_racecheck_control_0_oneslot.log)"
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r02a_racecheck_product.log 2>&1
echo "product racecheck: $(grep -E 'RACECHECK SUMMARY' gpurun_out/r02a_racecheck_product.log)"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r02a_memcheck_product.log 2>&1
echo "product memcheck: $(grep -E 'ERROR SUMMARY' gpurun_out/r02a_memcheck_product.log)"
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02a_bench.json; tail -3 gpurun_out/r02a_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02a_bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --pairs 200000 --steps 2 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02a_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:aff_fast -s 4 -c 1 -o gpurun_out/r02a_prof_fast python bench.py --pairs 100000 --steps 1 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02a_prof.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:aff_stripe -s 8 -c 1 -o gpurun_out/r02a_prof_stripe_ml python bench.py --workload affine500_medianlike --pairs 100000 --steps 1 --warmup 3 --skip-cpu --headline-only >> gpurun_out/r02a_prof.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:aff_traceback -s 4 -c 1 -o gpurun_out/r02a_prof_trace python bench.py --pairs 200000 --steps 1 --warmup 3 --skip-cpu --headline-only >> gpurun_out/r02a_prof.log 2>&1
ls -la gpurun_out | grep r02a
