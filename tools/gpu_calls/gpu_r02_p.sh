#!/bin/bash
# Round 2, GPU call P: full parity suite, smoke, default bench (with the powell300 block), Powell probe.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python tools/powell_probe.py 100,0.05,2368 300,0.03,592 300,0.10,148 500,0.05,148 300,0.10,16 > gpurun_out/r02p_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02p_probe.log
timeout 900 python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02p_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02p_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['traffic']); print({k:(round(v['value'],1), v.get('unit','GCUPS'), v.get('cpu_baseline',{}).get('value',0)) for k,v in d['workloads'].items()})"
