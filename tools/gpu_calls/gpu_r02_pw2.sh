#!/bin/bash
# Round 2: Powell kernel with the workspaces kept by the context -- parity and probe.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_stubs.py -m gpu -x -q -k "powell" ) > gpurun_out/r02pw2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02pw2_pytest.log
timeout 300 python tools/powell_probe.py 100,0.05,2368 300,0.03,592 300,0.10,148 500,0.05,148 > gpurun_out/r02pw2_probe.log 2>&1; cat gpurun_out/r02pw2_probe.log
timeout 300 python bench.py --skip-cpu --steps 2 --warmup 3 2> gpurun_out/r02pw2_bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), {k:(round(v['value'],1), v.get('unit','GCUPS')) for k,v in d['workloads'].items()})"
