#!/bin/bash
mkdir -p gpurun_out
for k in aff lin_stripe lin_rows powell; do
  timeout 200 compute-sanitizer --tool synccheck --error-exitcode 3 python tools/sync_probe.py $k > gpurun_out/r02_synccheck_$k.txt 2>&1; echo "synccheck $k rc=$?"; grep -E "ERROR SUMMARY|Barrier error|done" gpurun_out/r02_synccheck_$k.txt | sort | uniq -c | head -5
done
