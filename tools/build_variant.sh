#!/bin/bash
# tools/build_variant.sh <name> [-DMACRO=..]...  -> sketchy_b200/build/variants/lib_<name>.so  (kernels_predict.cu rebuilt with the macros)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p sketchy_b200/build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-Wall --expt-relaxed-constexpr "$@" \
  -c sketchy_b200/csrc/kernels_predict.cu -o sketchy_b200/build/variants/kp_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o sketchy_b200/build/variants/lib_$name.so \
  sketchy_b200/build/api.o sketchy_b200/build/pack_avx2.o sketchy_b200/build/kernels_sketch.o sketchy_b200/build/variants/kp_$name.o -lcudart
echo built lib_$name.so
