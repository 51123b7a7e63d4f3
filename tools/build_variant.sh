#!/bin/bash
# usage: tools/build_variant.sh NAME [-DMACRO ...]   -> build/lib_NAME.so (experimental builds; POYB200_SO selects one)
name=$1; shift
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" -o build/lib_$name.so poyd_b200/csrc/api.cu
