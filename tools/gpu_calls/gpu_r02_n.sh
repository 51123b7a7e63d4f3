#!/bin/bash
# Round 2, GPU call N: Powell kernel with single-shot sub-passes / strided second phase; launch-shape variants; parity.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_stubs.py -m gpu -x -q -k "powell" ) > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02n_pytest.log
export PROBE_NO_REF=1
for v in "" build/lib_pw512x1.so build/lib_pw1024x1.so; do
  echo "== variant ${v:-default (256 x 2)}"
  POYB200_SO=$v timeout 300 python tools/powell_probe.py 100,0.05,592 300,0.03,592 300,0.10,148 500,0.05,148 2>&1 | tail -4
done > gpurun_out/r02n_variants.log 2>&1
cat gpurun_out/r02n_variants.log
