#!/bin/bash
# Round 2, GPU call Q: Powell probe after the single-step fallback; latency of small calls per kernel path; cfg5 with small calls on the ring kernels.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_stubs.py -m gpu -x -q -k "powell" ) > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02q_pytest.log
PROBE_NO_REF=1 timeout 600 python tools/powell_probe.py 100,0.05,2368 300,0.03,592 300,0.10,148 500,0.05,148 > gpurun_out/r02q_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02q_probe.log
timeout 600 python tools/small_batch_probe.py 1500 > gpurun_out/r02q_small.log 2>&1; echo "small rc=$?"; cat gpurun_out/r02q_small.log
for cfg in "" "small_ring_pairs=4096"; do
  POYB200_CONFIG="$cfg" timeout 900 python bench.py --workload cfg5 --taxa 500 --bp 1500 --spr-rounds 3 --skip-cpu > "gpurun_out/r02q_cfg5_${cfg:-default}.json" 2> gpurun_out/r02q_cfg5.err; echo "cfg5[$cfg] rc=$?"
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(d['value'], d['ms_per_step'], d['tree'])" "gpurun_out/r02q_cfg5_${cfg:-default}.json"
done
