#!/bin/bash
# aff_x2_kernel: one ncu --set full capture of one launch over 100 k pairs (configs[1] shape)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export POYB200_CONFIG=chunk_pairs=1048576
timeout 400 ncu --set full --clock-control none --import-source on -k regex:aff_x2 -s 3 -c 1 -o gpurun_out/r02x2_prof python bench.py --pairs 100000 --steps 1 --warmup 3 --skip-cpu --headline-only > gpurun_out/r02x2_prof.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02x2_prof.log; ls -la gpurun_out/r02x2_prof.ncu-rep
