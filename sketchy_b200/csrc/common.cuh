// common.cuh — shared host/device helpers for libsketchy_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sketchy_b200.h"

#define SKB_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: a launcher remembers, per device, the largest size it
// has opted in to, so a process that drives several GPUs (one context each) configures every one of them.
struct SkbSmemOptIn {
  size_t bytes[64] = {};
  // true when `smem` bytes still have to be opted in to on the current device (and records it)
  bool needs(size_t smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& have = bytes[dev & 63];
    if (smem <= have) return false;
    have = smem;
    return true;
  }
};

// ---- packed sequence layout -----------------------------------------------------------------------------
// codes: 2 bit/base, 16 bases per u32, base p lives in word p/16 at bits 2*(p%16)   (A=0 C=1 G=2 T=3)
// nmask: 1 bit/base, 32 bases per u32, bit p%32 set = base p is NOT one of ACGT (or is padding / a separator)
// Every record starts on a 32-base boundary and is followed by >= 1 invalid base, so k-mer windows never span
// records and no two records share a word. A "chunk" is 32 consecutive base positions; a "segment" is up to
// 32 consecutive chunks of one group and is the unit of work of one warp.
#define SKB_CHUNK 32

struct SkbPackedView {
  const uint32_t* codes;
  const uint32_t* nmask;
  const uint32_t* seg_group;   // [nseg]
  const uint32_t* seg_chunk0;  // [nseg]
  const uint8_t* seg_n;        // [nseg] chunks in the segment (1..32)
  uint32_t nseg;
};

// One candidate produced by the rank warps: reference row `idx` has cumulative sum `sum` after the read whose
// bucket the record sits in.
struct __align__(16) SkbCand {
  unsigned long long sum;
  uint32_t idx;
  uint32_t pad;
};

// Run-length form the rank warps produce: row `idx` has cumulative sum `sum` and is a top-N candidate for every read
// in [b0, b1) of the pass (span = b0 | b1 << 16).
struct __align__(16) SkbInterval {
  unsigned long long sum;
  uint32_t idx;
  uint32_t span;
};

// (sum desc, idx asc) — the order of the reference's stable descending sort (src/sketchy.rs:310, 348).
__host__ __device__ __forceinline__ bool skb_key_better(unsigned long long sa, uint32_t ia, unsigned long long sb,
                                                        uint32_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

static inline uint64_t skb_next_pow2(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t skb_lane() { return threadIdx.x & 31u; }
#endif
