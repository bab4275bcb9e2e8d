#!/bin/bash
# tools/build_variant.sh <name> [-DMACRO=..]...  -> sketchy_b200/build/variants/lib_<name>.so
# api.cu and kernels_predict.cu are rebuilt with the macros (SKB_X_CW warps per CTA, SKB_X_SUB hashes per sub-tile,
# SKB_X_STAGES staging buffers per warp, SKB_X_ROWBUF rows in flight, SKB_BLOOM_K filter bits, SKB_X_IDBITS read-id
# width = log2 of the largest pass, SKB_X_ABLATE experiment switches); the other objects come from the in-tree build.
# Load a variant with SKB_LIB=<path> (sketchy_b200/_lib.py).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p sketchy_b200/build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-Wall --expt-relaxed-constexpr"
nvcc $FLAGS "$@" -c sketchy_b200/csrc/api.cu -o sketchy_b200/build/variants/api_$name.o &
nvcc $FLAGS "$@" -Xptxas -v -c sketchy_b200/csrc/kernels_predict.cu -o sketchy_b200/build/variants/kp_$name.o 2>&1 | grep -A2 "fused_kernelILi4" | grep -E "registers|spill" | tr '\n' ' '
echo
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o sketchy_b200/build/variants/lib_$name.so \
  sketchy_b200/build/variants/api_$name.o sketchy_b200/build/pack_avx2.o sketchy_b200/build/kernels_sketch.o sketchy_b200/build/variants/kp_$name.o -lcudart
rm -f sketchy_b200/build/variants/api_$name.o sketchy_b200/build/variants/kp_$name.o
echo built lib_$name.so
