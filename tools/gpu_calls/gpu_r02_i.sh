#!/bin/bash
# Round 2, GPU call I: the column-striped kernel for full linear matrices -- parity, then protein300 before / after.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "linear or protein or closest or custom_tail" ) > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02i_pytest.log
for cfg in "" "allow_rows=0"; do
  POYB200_CONFIG="$cfg" timeout 300 python bench.py --workload protein300 --pairs 262144 --skip-cpu --headline-only > "gpurun_out/r02i_protein_${cfg:-default}.json" 2> gpurun_out/r02i_protein.err; echo "bench[$cfg] rc=$?"
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d.get('phase_ms'), d['roofline'].get('kernel'), d['roofline'].get('kernel_ms_per_step'))" "gpurun_out/r02i_protein_${cfg:-default}.json"
done
POYB200_CONFIG="chunk_pairs=1048576" timeout 600 ncu --set full --import-source on --clock-control none -k regex:lin_rows_kernel -c 1 -o gpurun_out/r02i_rows python bench.py --workload protein300 --pairs 131072 --steps 1 --warmup 1 --skip-cpu --headline-only > gpurun_out/r02i_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | grep r02i
