#!/bin/bash
# tools/build_variant_sketch.sh <name> [-DMACRO=..]...  -> sketchy_b200/build/variants/lib_<name>.so
# kernels_sketch.cu rebuilt with the macros (SKB_X_HASH_UNROLL k-mers per unrolled body of hash_kernel); the other
# objects come from the in-tree build. Load a variant with SKB_LIB=<path> (sketchy_b200/_lib.py).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p sketchy_b200/build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-Wall --expt-relaxed-constexpr"
nvcc $FLAGS "$@" -c sketchy_b200/csrc/kernels_sketch.cu -o sketchy_b200/build/variants/ks_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o sketchy_b200/build/variants/lib_$name.so \
  sketchy_b200/build/api.o sketchy_b200/build/pack_avx2.o sketchy_b200/build/variants/ks_$name.o sketchy_b200/build/kernels_predict.o -lcudart
rm -f sketchy_b200/build/variants/ks_$name.o
echo built lib_$name.so
