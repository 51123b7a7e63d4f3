#!/bin/bash
# Runs the default bench (traceback not overlapped, no CPU arm) with each experimental library in build/.
for so in "" $(ls build/lib_*.so 2>/dev/null); do
  r=$(POYB200_SO=$so POYB200_OVERLAP_TB=0 timeout 200 python bench.py --skip-cpu "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f GCUPS %.2f ms fill %.2f trace %.2f e2e %.2f chk %d' % (d['value'], d['ms_per_step'], d['phase_ms']['fill'], d['phase_ms']['traceback'], d['e2e']['ms_per_step'], d['cost_checksum']))")
  echo "${so:-default}: $r"
done
