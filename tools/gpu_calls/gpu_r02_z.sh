#!/bin/bash
# Round 2, GPU call Z: traceback CTAs of 64 threads x 64 registers (fit next to three fill CTAs) against the default.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "affine" ) > gpurun_out/r02z_pytest.log 2>&1; echo "pytest(default) rc=$?"; tail -2 gpurun_out/r02z_pytest.log
( POYB200_CONFIG="traceback_block=64" timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "affine_align_ragged or headline" ) > gpurun_out/r02z_pytest64.log 2>&1; echo "pytest(block 64) rc=$?"; tail -2 gpurun_out/r02z_pytest64.log
for cfg in "" "traceback_block=64" "traceback_block=64,traceback_threads_per_sm=512" "traceback_block=64,traceback_threads_per_sm=128" "traceback_block=64,traceback_threads_per_sm=1024"; do
  POYB200_CONFIG="$cfg" timeout 300 python bench.py --skip-cpu --headline-only > gpurun_out/r02z_tmp.json 2> gpurun_out/r02z_tmp.err; echo -n "affine500 [$cfg] rc=$? "
  python -c "
import json; d=json.load(open('gpurun_out/r02z_tmp.json')); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e_dos_median']['value'],1), d['phase_ms'])"
done 2>&1 | tee gpurun_out/r02z_tb.log
