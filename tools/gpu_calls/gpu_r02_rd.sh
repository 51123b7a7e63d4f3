#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_stubs.py -m gpu -x -q -k "powell" ) > gpurun_out/r02rd_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02rd_pytest.log
