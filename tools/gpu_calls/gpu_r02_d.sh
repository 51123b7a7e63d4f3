#!/bin/bash
# Round 2, GPU call D: native tree driver + store, mixed ring mode, cfg5.
mkdir -p gpurun_out
timeout 300 python tools/ring_check.py > gpurun_out/r02d_ring_check.log 2>&1; echo "ring_check rc=$?"; grep -E "mismatch|differ|RING" gpurun_out/r02d_ring_check.log | sort | uniq -c | sort -rn | head -6
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02d_pytest.log
for mode in 2 1 0; do
for wl in affine500 affine500_medianlike; do
POYB200_CONFIG=use_ring=$mode timeout 600 python bench.py --skip-cpu --headline-only --workload $wl > gpurun_out/r02d_bench_${wl}_$mode.json 2> gpurun_out/r02d_bench_${wl}_$mode.err; python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_${wl}_$mode.json')); print('use_ring=$mode $wl:', round(d['value'],1), round(d['ms_per_step'],2), d['phase_ms'], 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1))"
done; done
timeout 900 python bench.py --workload cfg5 --taxa 120 --bp 1500 --spr-rounds 3 > gpurun_out/r02d_cfg5_small.json 2> gpurun_out/r02d_cfg5_small.err; echo "cfg5 small rc=$?"; cut -c1-1800 gpurun_out/r02d_cfg5_small.json; tail -3 gpurun_out/r02d_cfg5_small.err
timeout 1500 python bench.py --workload cfg5 --taxa 500 --bp 1500 --spr-rounds 3 --skip-cpu > gpurun_out/r02d_cfg5.json 2> gpurun_out/r02d_cfg5.err; echo "cfg5 rc=$?"; cut -c1-1800 gpurun_out/r02d_cfg5.json; tail -3 gpurun_out/r02d_cfg5.err
ls -la gpurun_out | grep r02d
