#!/bin/bash
# Round 2, GPU call M: Powell kernel launch-shape variants and an ncu capture (stall reasons, source hot spots).
mkdir -p gpurun_out
export PROBE_NO_REF=1
for v in "" build/lib_pw128x2.so build/lib_pw256x2.so build/lib_pw512x1.so; do
  echo "== variant ${v:-default (128 x 4)}"
  POYB200_SO=$v timeout 300 python tools/powell_probe.py 100,0.05,592 300,0.03,592 300,0.10,148 2>&1 | tail -3
done > gpurun_out/r02m_variants.log 2>&1
cat gpurun_out/r02m_variants.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:powell_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r02m_powell python tools/powell_probe.py 300,0.03,148 > gpurun_out/r02m_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02m_ncu.log
ls -la gpurun_out | grep r02m
