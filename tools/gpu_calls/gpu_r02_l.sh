#!/bin/bash
# Round 2, GPU call L: Powell kernel after the latency work (batched reads, sequences in shared memory, 4 CTAs per SM).
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_stubs.py -m gpu -x -q -k "powell" ) > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02l_pytest.log
timeout 600 python tools/powell_probe.py > gpurun_out/r02l_probe.log 2>&1; echo "probe rc=$?"; tail -12 gpurun_out/r02l_probe.log
