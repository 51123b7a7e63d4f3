#!/bin/bash
# chunk sizes that are whole multiples of the x2 grid's 14 208 pairs per pass (444 CTAs x 4 warps x 8 pairs)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 65536 71040 56832; do
  POYB200_CONFIG=chunk_pairs=$c timeout 60 python bench.py --headline-only --skip-cpu --steps 3 --warmup 3 --pairs 568320 > gpurun_out/r02x2c_$c.json 2> gpurun_out/r02x2c_$c.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02x2c_$c.json') if l.startswith('{')][-1]); print($c, round(d['value'],1), round(d['ms_per_step'],2), d['phase_ms'], round(d['e2e_dos_median']['value'],1))"
done
