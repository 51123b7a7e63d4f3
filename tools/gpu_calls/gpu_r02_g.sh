#!/bin/bash
# Round 2, GPU call G: final build -- parity, the default bench (both arms) with profiles/r02_kernel_metrics.json in place, cfg5, smoke.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02g_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['roofline']['frac'], d['roofline']['executed_frac'], d['roofline']['traffic']); print({k:(round(v['value'],1), round(v.get('e2e_dos_median',v.get('e2e'))['value'],1), round(v.get('cpu_baseline',{}).get('value',0),2)) for k,v in d['workloads'].items()})"
timeout 600 python bench.py --impl reference > gpurun_out/r02g_bench_ref.json 2> gpurun_out/r02g_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r02g_bench_ref.json
timeout 900 python bench.py --workload cfg5 --taxa 500 --bp 1500 --spr-rounds 3 > gpurun_out/r02g_cfg5.json 2> gpurun_out/r02g_cfg5.err; echo "cfg5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02g_cfg5.json')); print(d['value'], d['ms_per_step'], d['tree'], d.get('cpu_baseline'))"
timeout 300 python tools/long_pairs_probe.py > gpurun_out/r02g_long.log 2>&1; tail -8 gpurun_out/r02g_long.log
