#!/bin/bash
# build_variant.sh NAME file.cu [-DFLAG ...]: rebuild ONE translation unit with extra flags and link it with the objects of the
# regular build into variants/libicw_b200_NAME.so (git-ignored, travels with gpurun) for A/B runs on the GPU box:
#   cp variants/libicw_b200_NAME.so in_cwave_b200/libicw_b200.so && python bench.py ...
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p variants
stem=$(basename "$src" .cu)
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v "$@" \
     -c -o variants/${stem}_${name}.o in_cwave_b200/csrc/$src 2> variants/${stem}_${name}.log
objs=$(ls in_cwave_b200/build/*.o | grep -v "/${stem}.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/libicw_b200_${name}.so $objs variants/${stem}_${name}.o -ldl
grep -h "registers" variants/${stem}_${name}.log | sort | uniq -c | head -8
