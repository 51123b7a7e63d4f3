#!/bin/bash
# Round 2, GPU call O: Powell kernel with the strided first phase and the two launch shapes; parity + probe (with the reference sample).
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_stubs.py -m gpu -x -q -k "powell" ) > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02o_pytest.log
timeout 600 python tools/powell_probe.py 100,0.05,2368 100,0.05,592 300,0.03,592 300,0.10,148 500,0.05,148 300,0.10,16 > gpurun_out/r02o_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02o_probe.log
