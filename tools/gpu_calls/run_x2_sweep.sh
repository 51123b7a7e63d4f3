#!/bin/bash
# aff_x2_kernel: unroll / flag-pipe variants, then traceback configurations, headline workload, device-resident
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 120 python bench.py --headline-only --skip-cpu --steps 3 --warmup 3 --pairs 524288 > gpurun_out/r02x2s_$name.json 2> gpurun_out/r02x2s_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02x2s_{n}.json") if l.startswith("{")][-1])
    print(f"{n:28s} {d['value']:8.1f} GCUPS  {d['ms_per_step']:7.2f} ms  fill {d['phase_ms']['fill']:.2f}  e2e {d['e2e']['value']:.0f}  dos {d['e2e_dos_median']['value']:.0f}  chk {d['cost_checksum']}")
except Exception as e:
    print(n, "ERR", e)
PY
}
run base X=1
for v in un3 un3f un2f un6f; do run $v POYB200_SO=build/lib_$v.so; done
run tb128 POYB200_CONFIG=traceback_threads_per_sm=128
run tb512 POYB200_CONFIG=traceback_threads_per_sm=512
run tb384 POYB200_CONFIG=traceback_threads_per_sm=384
run prio0 POYB200_CONFIG=traceback_priority=0
run noov_768 POYB200_CONFIG=overlap_traceback=0,traceback_threads_per_sm=768
run noov_1280 POYB200_CONFIG=overlap_traceback=0,traceback_threads_per_sm=1280
run chunk128k POYB200_CONFIG=chunk_pairs=131072
