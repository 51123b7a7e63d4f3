#!/bin/bash
# Round 2, GPU call F: parity of the final build, the full default bench, cfg5, ncu captures for profiles/r02_kernel_metrics.json.
mkdir -p gpurun_out
timeout 300 python tools/ring_check.py > gpurun_out/r02f_ring_check.log 2>&1; echo "ring_check rc=$?"; grep -E "RING|mismatches [1-9]|[1-9][0-9]* rows differ" gpurun_out/r02f_ring_check.log | head
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02f_pytest.log
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02f_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['roofline']['frac'], d['roofline']['executed_frac']); print({k:(round(v['value'],1), round(v.get('e2e_dos_median',v.get('e2e'))['value'],1), round(v.get('cpu_baseline',{}).get('value',0),2)) for k,v in d['workloads'].items()})"
for v in ebun6; do
POYB200_SO=build/lib_$v.so timeout 600 python bench.py --skip-cpu --headline-only --workload affine500_medianlike > gpurun_out/r02f_bench_ml_$v.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_ml_$v.json')); print('$v medianlike:', d['value'], d['ms_per_step'])"
done
timeout 900 python bench.py --workload cfg5 --taxa 500 --bp 1500 --spr-rounds 3 --skip-cpu > gpurun_out/r02f_cfg5.json 2> gpurun_out/r02f_cfg5.err; echo "cfg5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02f_cfg5.json')); print(d['value'], d['ms_per_step'], d['tree'])"
export POYB200_CONFIG=chunk_pairs=1048576
timeout 400 ncu --set full --clock-control none --import-source on -k regex:aff_fast -s 4 -c 1 -o gpurun_out/r02f_prof_fast python bench.py --pairs 100000 --steps 1 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02f_prof.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:aff_ring -s 4 -c 1 -o gpurun_out/r02f_prof_ring_ml python bench.py --workload affine500_medianlike --pairs 100000 --steps 1 --warmup 3 --skip-cpu --headline-only >> gpurun_out/r02f_prof.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:aff_traceback -s 4 -c 1 -o gpurun_out/r02f_prof_trace python bench.py --pairs 100000 --steps 1 --warmup 3 --skip-cpu --headline-only >> gpurun_out/r02f_prof.log 2>&1
unset POYB200_CONFIG
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --pairs 200000 --steps 2 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02f_launches.log 2>&1
ls -la gpurun_out | grep r02f
