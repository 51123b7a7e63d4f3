#!/bin/bash
# Round 2, final 1-GPU call: full parity suite, smoke, default bench (both arms), configs[4], Powell probe.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02final_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02final_bench.json 2> gpurun_out/r02final_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02final_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02final_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['traffic'], d['roofline_hbm']['algorithmic_bytes_per_launch']); print({k:(round(v['value'],1), v.get('unit','GCUPS'), round(v.get('cpu_baseline',{}).get('value',0),2)) for k,v in d['workloads'].items()}); print(d['cpu_baseline'])"
timeout 600 python bench.py --impl reference > gpurun_out/r02final_bench_ref.json 2> gpurun_out/r02final_bench_ref.err; echo "ref rc=$?"; cut -c1-220 gpurun_out/r02final_bench_ref.json
timeout 900 python bench.py --workload cfg5 --taxa 500 --bp 1500 --spr-rounds 3 > gpurun_out/r02final_cfg5.json 2> gpurun_out/r02final_cfg5.err; echo "cfg5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02final_cfg5.json')); print(d['value'], d['ms_per_step'], d['tree'], d.get('cpu_baseline'))"
PROBE_NO_REF=1 timeout 300 python tools/powell_probe.py 100,0.05,2368 300,0.03,592 300,0.10,148 500,0.05,148 > gpurun_out/r02final_probe.log 2>&1; cat gpurun_out/r02final_probe.log
