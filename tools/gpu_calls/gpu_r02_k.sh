#!/bin/bash
# Round 2, GPU call K: the Powell kernel -- parity against the compiled reference, then a timing of a batch.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "powell" ) > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02k_pytest.log
timeout 600 python tools/powell_probe.py > gpurun_out/r02k_probe.log 2>&1; echo "probe rc=$?"; tail -12 gpurun_out/r02k_probe.log
