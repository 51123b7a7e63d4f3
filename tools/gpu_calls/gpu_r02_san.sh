#!/bin/bash
# Round 2: compute-sanitizer memcheck + synccheck over every kernel family of the final build (small batches).
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_memcheck.txt
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_synccheck.txt 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_synccheck.txt
