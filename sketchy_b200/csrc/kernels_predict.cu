// kernels_predict.cu — streaming predict on sm_100a:
//   table_*      : per-pass hash set of the batch's query hashes (global table + 128 KB shared-memory filter)
//   stream_kernel: the HBM-bound kernel. Streams the flat reference hash matrix once per pass through a
//                  cp.async.bulk (TMA) + mbarrier shared-memory ring, probes every streamed hash, and adds the
//                  hits into the per-pass (row, read) count matrix.
//   rank_*       : cumulative sums over the reads of the pass + exact per-read top-N by (sum desc, index asc).
// Replaces `_common_hashes` x N + `sum[i] += shared` + stable sort + `[..top]`
// (reference src/sketchy.rs:337-348, 391, 419-459). Rows are strictly increasing (checked at upload) and each
// read's query list is distinct, so the two-pointer merge count equals the set-intersection size computed here.
#include "kernels.h"

namespace {

__device__ __forceinline__ uint32_t table_home(uint64_t h, uint32_t log2cap) {
  return (uint32_t)((h * 0x9E3779B97F4A7C15ull) >> (64 - log2cap));
}
__device__ __forceinline__ uint32_t bloom_word(uint32_t lo) { return (lo >> 5) & (SKB_BLOOM_WORDS - 1u); }
__device__ __forceinline__ uint32_t bloom_mask(uint32_t lo) { return (1u << (lo & 31u)) | (1u << ((lo >> 20) & 31u)); }

// ---------------------------------------------------------------------------------------------------------
// query table
// ---------------------------------------------------------------------------------------------------------
__global__ void table_insert_kernel(SkbTable t, const uint64_t* __restrict__ qh, uint32_t n_keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys) return;
  const uint64_t h = qh[i];
  uint32_t slot;
  if (h == SKB_EMPTY_KEY) {
    slot = t.cap;
  } else {
    slot = table_home(h, t.log2cap);
    for (;;) {
      const unsigned long long prev = atomicCAS((unsigned long long*)&t.keys[slot], SKB_EMPTY_KEY, h);
      if (prev == SKB_EMPTY_KEY || prev == h) break;
      slot = (slot + 1) & (t.cap - 1);
    }
  }
  atomicAdd(&t.cnt[slot], 1u);
  t.slot_of[i] = slot;
  const uint32_t lo = (uint32_t)h;
  atomicOr(&t.bloom[bloom_word(lo)], bloom_mask(lo));
}

__global__ void table_alloc_kernel(SkbTable t) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > t.cap) return;
  const uint32_t c = t.cnt[s];
  if (c) t.start[s] = atomicAdd(t.cursor, c);
}

__global__ void table_fill_kernel(SkbTable t, const uint32_t* __restrict__ qread, uint32_t n_keys,
                                  uint32_t read_base) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys) return;
  const uint32_t s = t.slot_of[i];
  const uint32_t p = t.start[s] + atomicAdd(&t.fill[s], 1u);
  t.reads[p] = qread[i] - read_base;
}

// ---------------------------------------------------------------------------------------------------------
// stream kernel: mbarrier / bulk-copy primitives
// ---------------------------------------------------------------------------------------------------------
constexpr int ST_TILE = 2048;  // hashes per stage (16 KB)
constexpr int ST_STAGES = 4;
constexpr int ST_CONSUMER_WARPS = 16;
constexpr int ST_THREADS = (ST_CONSUMER_WARPS + 1) * 32;  // warp 0 is the bulk-copy producer
constexpr int ST_QCAP = 64;                                // per-warp queue of filter passers
constexpr size_t ST_SMEM_BLOOM = (size_t)SKB_BLOOM_WORDS * 4;
constexpr size_t ST_SMEM_RING = (size_t)ST_STAGES * ST_TILE * 8;
constexpr size_t ST_SMEM_QUEUE = (size_t)ST_CONSUMER_WARPS * ST_QCAP * 16;
constexpr size_t ST_SMEM_TOTAL = ST_SMEM_BLOOM + ST_SMEM_RING + ST_SMEM_QUEUE;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier; streamed data is marked
// evict-first so the query table and the count matrix keep their place in L2.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                          uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

struct StreamShared {
  uint32_t* bloom;
  uint64_t* ring;
  uint64_t* qh;    // [warps][ST_QCAP]
  uint64_t* qpos;  // [warps][ST_QCAP]
};

__device__ __forceinline__ uint32_t row_of(const SkbStreamArgs& a, uint64_t pos) {
  if (a.uniform_len) return (uint32_t)(pos / a.uniform_len);
  uint32_t lo = 0, hi = a.n_rows;  // largest r with row_off[r] <= pos
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a.row_off[mid] <= pos) lo = mid; else hi = mid;
  }
  return lo;
}

// exact verification of the filter passers queued by one warp, and the count updates for true hits
__device__ __forceinline__ void drain_queue(const SkbStreamArgs& a, const uint64_t* qh, const uint64_t* qpos,
                                            uint32_t qn) {
  __syncwarp();
  const SkbTable& t = a.table;
  for (uint32_t i = skb_lane(); i < qn; i += 32) {
    const uint64_t h = qh[i];
    uint32_t slot;
    bool found = false;
    if (h == SKB_EMPTY_KEY) {
      slot = t.cap;
      found = t.cnt[slot] != 0;
    } else {
      slot = table_home(h, t.log2cap);
      for (;;) {
        const uint64_t key = t.keys[slot];
        if (key == h) { found = true; break; }
        if (key == SKB_EMPTY_KEY) break;
        slot = (slot + 1) & (t.cap - 1);
      }
    }
    if (found) {
      const uint32_t row = row_of(a, qpos[i]);
      const uint32_t st = t.start[slot], c = t.cnt[slot];
      uint32_t* crow = reinterpret_cast<uint32_t*>(a.counts + (size_t)row * a.row_stride);
      for (uint32_t j = 0; j < c; ++j) {
        const uint32_t rd = t.reads[st + j];
        atomicAdd(crow + (rd >> 1), 1u << (16 * (rd & 1u)));  // two u16 counters per word; no carry: count <= 65535
      }
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(ST_THREADS, 1) stream_kernel(const SkbStreamArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[ST_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[ST_STAGES];

  uint32_t* bloom = reinterpret_cast<uint32_t*>(smem_raw);
  uint64_t* ring = reinterpret_cast<uint64_t*>(smem_raw + ST_SMEM_BLOOM);
  uint64_t* queue = reinterpret_cast<uint64_t*>(smem_raw + ST_SMEM_BLOOM + ST_SMEM_RING);

  // tiles of this CTA: contiguous range
  const uint64_t n_tiles = (a.ref_len + ST_TILE - 1) / ST_TILE;
  const uint64_t t_begin = n_tiles * blockIdx.x / gridDim.x;
  const uint64_t t_end = n_tiles * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], ST_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {  // stage the filter
    const uint4* src = reinterpret_cast<const uint4*>(a.table.bloom);
    uint4* dst = reinterpret_cast<uint4*>(bloom);
    for (uint32_t i = threadIdx.x; i < SKB_BLOOM_WORDS / 4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();

  const uint32_t warp = threadIdx.x >> 5, lane = skb_lane();
  if (warp == 0) {
    // ===== producer: one elected lane issues the bulk copies =====
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      uint32_t it = 0;
      for (uint64_t tile = t_begin; tile < t_end; ++tile, ++it) {
        const uint32_t stage = it % ST_STAGES, phase = (it / ST_STAGES) & 1u;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        const uint64_t first = tile * ST_TILE;
        uint64_t n = a.ref_len - first;
        if (n > ST_TILE) n = ST_TILE;
        const uint32_t bytes = (uint32_t)(((n + 1) & ~1ull) * 8);  // multiple of 16; the array is padded to even
        mbar_arrive_expect_tx(&full_bar[stage], bytes);
        bulk_load(ring + (size_t)stage * ST_TILE, a.ref + first, bytes, &full_bar[stage], policy);
      }
    }
    return;
  }

  // ===== consumers =====
  const uint32_t cw = warp - 1;
  uint64_t* qh = queue + (size_t)cw * ST_QCAP * 2;
  uint64_t* qpos = qh + ST_QCAP;
  uint32_t qn = 0;
  const uint32_t ct = threadIdx.x - 32;  // 0 .. 511
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t it = 0;
  for (uint64_t tile = t_begin; tile < t_end; ++tile, ++it) {
    const uint32_t stage = it % ST_STAGES, phase = (it / ST_STAGES) & 1u;
    mbar_wait(&full_bar[stage], phase);
    const uint4* tp = reinterpret_cast<const uint4*>(ring + (size_t)stage * ST_TILE);
    uint4 v[ST_TILE / (2 * ST_CONSUMER_WARPS * 32)];
#pragma unroll
    for (int r = 0; r < ST_TILE / (2 * ST_CONSUMER_WARPS * 32); ++r) v[r] = tp[ct + r * ST_CONSUMER_WARPS * 32];
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);  // data is in registers: hand the slot back

    const uint64_t first = tile * ST_TILE;
    uint64_t n = a.ref_len - first;
    if (n > ST_TILE) n = ST_TILE;
#pragma unroll
    for (int r = 0; r < ST_TILE / (2 * ST_CONSUMER_WARPS * 32); ++r) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const uint32_t lo = e ? v[r].z : v[r].x, hi = e ? v[r].w : v[r].y;
        const uint32_t idx = 2u * (ct + r * ST_CONSUMER_WARPS * 32) + e;
        const uint32_t w = bloom[bloom_word(lo)];
        const uint32_t m = bloom_mask(lo);
        const bool pass = ((w & m) == m) && (idx < n);
        const uint32_t bal = __ballot_sync(0xffffffffu, pass);
        if (bal) {
          const uint32_t np = __popc(bal);
          if (qn + np > ST_QCAP) { drain_queue(a, qh, qpos, qn); qn = 0; }
          if (pass) {
            const uint32_t q = qn + __popc(bal & lt_mask);
            qh[q] = ((uint64_t)hi << 32) | lo;
            qpos[q] = first + idx;
          }
          qn += np;
        }
      }
    }
  }
  drain_queue(a, qh, qpos, qn);
}

// ---------------------------------------------------------------------------------------------------------
// rank kernels
// ---------------------------------------------------------------------------------------------------------
// inclusive scan of one u32 per thread across the CTA; returns inclusive prefix, *total = sum
__device__ __forceinline__ uint32_t cta_scan_u32(uint32_t v, uint32_t* warp_sums, uint32_t* total) {
  const uint32_t lane = skb_lane(), wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, v, o);
    if ((int)lane >= o) v += y;
  }
  if (lane == 31) warp_sums[wid] = v;
  __syncthreads();
  uint32_t off = 0, tot = 0;
  for (uint32_t w = 0; w < nw; ++w) {
    const uint32_t s = warp_sums[w];
    if (w < wid) off += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return v + off;
}

// Lower bound of every read's top-th key: the worst key among `n_tracked` fixed rows (the previous top rows).
// Any `top` distinct rows give a valid bound because sums never decrease.
__global__ void __launch_bounds__(1024) rank_bounds_kernel(const SkbRankArgs a) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t per = (a.n_reads + blockDim.x - 1) / blockDim.x;
  const uint32_t b0 = threadIdx.x * per;
  const uint32_t b1 = min(b0 + per, a.n_reads);
  for (uint32_t t = 0; t < a.n_tracked; ++t) {
    const uint32_t row = a.tracked[t];
    const uint16_t* c = a.counts + (size_t)row * a.row_stride;
    uint32_t local = 0;
    for (uint32_t b = b0; b < b1; ++b) local += c[b];
    uint32_t tot;
    const uint32_t incl = cta_scan_u32(local, warp_sums, &tot);
    unsigned long long s = a.sums_in[row] + (incl - local);
    const uint32_t gi = a.row_base + row;
    for (uint32_t b = b0; b < b1; ++b) {
      s += c[b];
      if (t == 0 || skb_key_better(a.lb_sum[b], a.lb_idx[b], s, gi)) {
        a.lb_sum[b] = s;
        a.lb_idx[b] = gi;
      }
    }
  }
}

// One warp per reference row: prefix sums of the row's per-read counts, candidate test against the bounds,
// new running sum. 8 reads per lane per round (one 16-byte load).
__global__ void __launch_bounds__(256) rank_scan_kernel(const SkbRankArgs a) {
  const uint32_t lane = skb_lane();
  const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (uint32_t row = warp0; row < a.n_rows; row += nwarps) {
    const unsigned long long carry = a.sums_in[row];
    const uint32_t gi = a.row_base + row;
    const uint16_t* c = a.counts + (size_t)row * a.row_stride;
    uint32_t run = 0;
    for (uint32_t r0 = 0; r0 < a.n_reads; r0 += 256) {
      const uint32_t bl = r0 + lane * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (bl < a.row_stride) v = *reinterpret_cast<const uint4*>(c + bl);
      uint32_t p[8];
      p[0] = v.x & 0xFFFFu; p[1] = p[0] + (v.x >> 16);
      p[2] = p[1] + (v.y & 0xFFFFu); p[3] = p[2] + (v.y >> 16);
      p[4] = p[3] + (v.z & 0xFFFFu); p[5] = p[4] + (v.z >> 16);
      p[6] = p[5] + (v.w & 0xFFFFu); p[7] = p[6] + (v.w >> 16);
      uint32_t incl = p[7];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += y;
      }
      const uint32_t round_total = __shfl_sync(0xffffffffu, incl, 31);
      const uint32_t base = run + incl - p[7];
      // sums and bounds are non-decreasing along the reads: if the row's sum at the END of the round is below
      // the bound at the START of the round, no read of the round can take it
      const unsigned long long lb_first = a.lb_sum[r0];
      if (carry + run + round_total >= lb_first) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t b = bl + e;
          bool is_cand = false;
          unsigned long long s = 0;
          if (b < a.n_reads) {
            s = carry + base + p[e];
            const unsigned long long ls = a.lb_sum[b];
            is_cand = s > ls || (s == ls && gi <= a.lb_idx[b]);
          }
          const uint32_t bal = __ballot_sync(0xffffffffu, is_cand);
          if (bal) {
            uint32_t slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(a.cand_total, (uint32_t)__popc(bal));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            if (is_cand) {
              const uint32_t slot = slot0 + __popc(bal & lt_mask);
              if (slot < a.cand_cap) {
                SkbCand cd;
                cd.sum = s; cd.idx = gi; cd.read = b;
                a.cand[slot] = cd;
              }
              atomicAdd(&a.cand_cnt[b], 1u);
            }
          }
        }
      }
      run += round_total;
    }
    if (lane == 0) a.sums_out[row] = carry + run;
  }
}

__global__ void __launch_bounds__(1024) rank_offsets_kernel(const SkbRankArgs a) {
  __shared__ uint32_t warp_sums[32];
  const bool overflow = *a.cand_total > a.cand_cap;
  const uint32_t per = (a.n_reads + blockDim.x - 1) / blockDim.x;
  const uint32_t b0 = threadIdx.x * per;
  const uint32_t b1 = min(b0 + per, a.n_reads);
  uint32_t local = 0;
  if (!overflow)
    for (uint32_t b = b0; b < b1; ++b) local += a.cand_cnt[b];
  uint32_t tot;
  const uint32_t incl = cta_scan_u32(local, warp_sums, &tot);
  uint32_t off = incl - local;
  for (uint32_t b = b0; b < b1; ++b) {
    a.cand_off[b] = off;
    if (!overflow) off += a.cand_cnt[b];
  }
  if (threadIdx.x == 0) a.cand_off[a.n_reads] = tot;
}

__global__ void rank_scatter_kernel(const SkbRankArgs a) {
  const uint32_t total = *a.cand_total;
  if (total > a.cand_cap) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const SkbCand c = a.cand[i];
    const uint32_t p = a.cand_off[c.read] + atomicAdd(&a.cand_fill[c.read], 1u);
    a.cand_sorted[p] = c;
  }
}

__device__ __forceinline__ void warp_best(unsigned long long& s, uint32_t& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const uint32_t i2 = __shfl_xor_sync(0xffffffffu, i, o);
    if (skb_key_better(s2, i2, s, i)) { s = s2; i = i2; }
  }
}

// One warp per read: the `top` best candidates in order. Candidate rows are distinct, so "best key strictly
// worse than the previous pick" enumerates them without marking. Missing entries are (sum 0, idx UINT32_MAX).
__global__ void __launch_bounds__(256) rank_select_kernel(const SkbRankArgs a) {
  const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= a.n_reads) return;
  const uint32_t lane = skb_lane();
  const uint32_t lo = a.cand_off[b], hi = a.cand_off[b + 1];
  unsigned long long last_s = 0;
  uint32_t last_i = 0;
  for (uint32_t t = 0; t < a.top; ++t) {
    unsigned long long bs = 0;
    uint32_t bi = 0xFFFFFFFFu;
    for (uint32_t i = lo + lane; i < hi; i += 32) {
      const SkbCand c = a.cand_sorted[i];
      if ((t == 0 || skb_key_better(last_s, last_i, c.sum, c.idx)) && skb_key_better(c.sum, c.idx, bs, bi)) {
        bs = c.sum; bi = c.idx;
      }
    }
    warp_best(bs, bi);
    if (lane == 0) {
      a.out_idx[(size_t)b * a.top + t] = bi;
      a.out_sum[(size_t)b * a.top + t] = bs;
      if (b == a.n_reads - 1 && a.tracked_next && bi != 0xFFFFFFFFu) a.tracked_next[t] = bi - a.row_base;
    }
    last_s = bs; last_i = bi;
    if (bi == 0xFFFFFFFFu) {  // exhausted: pad the rest
      for (uint32_t u = t + 1 + lane; u < a.top; u += 32) {
        a.out_idx[(size_t)b * a.top + u] = 0xFFFFFFFFu;
        a.out_sum[(size_t)b * a.top + u] = 0;
      }
      break;
    }
  }
}

// top-N of an array by (value desc, index asc); one CTA.
__global__ void __launch_bounds__(1024) rank_full_kernel(const unsigned long long* __restrict__ vals, uint32_t n,
                                                         uint32_t top, uint32_t idx_base, uint32_t* out_idx,
                                                         unsigned long long* out_val, uint32_t* out_local) {
  __shared__ unsigned long long sh_s[32];
  __shared__ uint32_t sh_i[32];
  __shared__ unsigned long long pick_s;
  __shared__ uint32_t pick_i;
  const uint32_t lane = skb_lane(), wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned long long last_s = 0;
  uint32_t last_i = 0;
  for (uint32_t t = 0; t < top; ++t) {
    unsigned long long bs = 0;
    uint32_t bi = 0xFFFFFFFFu;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long v = vals[i];
      if ((t == 0 || skb_key_better(last_s, last_i, v, i)) && skb_key_better(v, i, bs, bi)) { bs = v; bi = i; }
    }
    warp_best(bs, bi);
    if (lane == 0) { sh_s[wid] = bs; sh_i[wid] = bi; }
    __syncthreads();
    if (wid == 0) {
      bs = lane < nw ? sh_s[lane] : 0ull;
      bi = lane < nw ? sh_i[lane] : 0xFFFFFFFFu;
      warp_best(bs, bi);
      if (lane == 0) { pick_s = bs; pick_i = bi; }
    }
    __syncthreads();
    last_s = pick_s; last_i = pick_i;
    if (threadIdx.x == 0) {
      if (out_idx) out_idx[t] = last_i == 0xFFFFFFFFu ? last_i : last_i + idx_base;
      if (out_val) out_val[t] = last_s;
      if (out_local && last_i != 0xFFFFFFFFu) out_local[t] = last_i;
    }
    __syncthreads();
  }
}

// multi-GPU: per read, merge n_parts lists of `top` (sum, idx) into the best `top`. One warp per read.
__global__ void __launch_bounds__(256) merge_topn_kernel(const uint32_t* __restrict__ idx_parts,
                                                         const unsigned long long* __restrict__ sum_parts,
                                                         uint32_t n_parts, uint64_t n_reads, uint32_t top,
                                                         uint32_t* out_idx, unsigned long long* out_sum) {
  const uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= n_reads) return;
  const uint32_t lane = skb_lane();
  const uint32_t total = n_parts * top;
  unsigned long long last_s = 0;
  uint32_t last_i = 0;
  for (uint32_t t = 0; t < top; ++t) {
    unsigned long long bs = 0;
    uint32_t bi = 0xFFFFFFFFu;
    for (uint32_t e = lane; e < total; e += 32) {
      const uint32_t part = e / top, j = e - part * top;
      const size_t at = ((size_t)part * n_reads + b) * top + j;
      const unsigned long long s = sum_parts[at];
      const uint32_t i = idx_parts[at];
      if ((t == 0 || skb_key_better(last_s, last_i, s, i)) && skb_key_better(s, i, bs, bi)) { bs = s; bi = i; }
    }
    warp_best(bs, bi);
    if (lane == 0) {
      out_idx[b * top + t] = bi;
      out_sum[b * top + t] = bs;
    }
    last_s = bs; last_i = bi;
  }
}

// rows strictly increasing? + maximum hash. One warp per row.
__global__ void __launch_bounds__(256) ref_check_kernel(const uint64_t* __restrict__ ref,
                                                        const uint64_t* __restrict__ row_off, uint32_t n_rows,
                                                        uint32_t* bad, unsigned long long* hmax) {
  const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = warp0; r < n_rows; r += nwarps) {
    const uint64_t lo = row_off[r], hi = row_off[r + 1];
    bool ok = true;
    for (uint64_t i = lo + skb_lane(); i + 1 < hi; i += 32) ok = ok && (ref[i] < ref[i + 1]);
    if (!ok) atomicOr(bad, 1u);
    if (skb_lane() == 0 && hi > lo) atomicMax(hmax, (unsigned long long)ref[hi - 1]);
  }
}

// dense shared counts (reference `shared`, src/sketchy.rs:238-279): one warp per (reference row, query) pair;
// each lane binary-searches its share of the query in the row. Inputs strictly increasing => == merge count.
__global__ void __launch_bounds__(256) shared_kernel(const uint64_t* __restrict__ ref,
                                                     const uint64_t* __restrict__ row_off, uint32_t n_rows,
                                                     const uint64_t* __restrict__ q,
                                                     const uint64_t* __restrict__ q_off, uint32_t Q,
                                                     unsigned long long* out) {
  const uint64_t pair = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pair >= (uint64_t)n_rows * Q) return;
  const uint32_t i = (uint32_t)(pair / Q), j = (uint32_t)(pair % Q);
  const uint64_t r0 = row_off[i], r1 = row_off[i + 1];
  const uint64_t q0 = q_off[j], q1 = q_off[j + 1];
  uint32_t c = 0;
  for (uint64_t x = q0 + skb_lane(); x < q1; x += 32) {
    const uint64_t h = q[x];
    uint64_t lo = r0, hi = r1;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (ref[mid] < h) lo = mid + 1; else hi = mid;
    }
    c += (lo < r1 && ref[lo] == h) ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (skb_lane() == 0) out[pair] = c;
}

}  // namespace

// =========================================================================================================
// launch wrappers
// =========================================================================================================
void skb_launch_table_build(const SkbTable& t, const uint64_t* qh, const uint32_t* qread, uint32_t n_keys,
                            uint32_t read_base, cudaStream_t st) {
  cudaMemsetAsync(t.keys, 0xFF, ((size_t)t.cap + 1) * sizeof(uint64_t), st);
  cudaMemsetAsync(t.cnt, 0, ((size_t)t.cap + 1) * sizeof(uint32_t), st);
  cudaMemsetAsync(t.fill, 0, ((size_t)t.cap + 1) * sizeof(uint32_t), st);
  cudaMemsetAsync(t.bloom, 0, (size_t)SKB_BLOOM_WORDS * sizeof(uint32_t), st);
  cudaMemsetAsync(t.cursor, 0, sizeof(uint32_t), st);
  if (n_keys == 0) return;
  const int th = 256;
  table_insert_kernel<<<(n_keys + th - 1) / th, th, 0, st>>>(t, qh, n_keys);
  table_alloc_kernel<<<(t.cap + 1 + th - 1) / th, th, 0, st>>>(t);
  table_fill_kernel<<<(n_keys + th - 1) / th, th, 0, st>>>(t, qread, n_keys, read_base);
}

size_t skb_stream_smem_bytes() { return ST_SMEM_TOTAL; }

void skb_launch_stream(const SkbStreamArgs& a, cudaStream_t st) {
  if (a.ref_len == 0) return;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM_TOTAL);
    configured = true;
  }
  stream_kernel<<<a.num_ctas, ST_THREADS, ST_SMEM_TOTAL, st>>>(a);
}

void skb_launch_rank_bounds(const SkbRankArgs& a, cudaStream_t st) { rank_bounds_kernel<<<1, 1024, 0, st>>>(a); }

void skb_launch_rank_scan(const SkbRankArgs& a, cudaStream_t st) {
  const int th = 256;
  unsigned blocks = (a.n_rows + (th / 32) - 1) / (th / 32);
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  if (blocks == 0) return;
  rank_scan_kernel<<<blocks, th, 0, st>>>(a);
}

void skb_launch_rank_group(const SkbRankArgs& a, cudaStream_t st) {
  rank_offsets_kernel<<<1, 1024, 0, st>>>(a);
  rank_scatter_kernel<<<148 * 4, 256, 0, st>>>(a);
}

void skb_launch_rank_select(const SkbRankArgs& a, cudaStream_t st) {
  const int th = 256;
  const unsigned blocks = (unsigned)(((uint64_t)a.n_reads * 32 + th - 1) / th);
  rank_select_kernel<<<blocks, th, 0, st>>>(a);
}

void skb_launch_rank_full(const unsigned long long* vals, uint32_t n, uint32_t top, uint32_t idx_base,
                          uint32_t* out_idx, unsigned long long* out_val, uint32_t* out_local, cudaStream_t st) {
  rank_full_kernel<<<1, 1024, 0, st>>>(vals, n, top, idx_base, out_idx, out_val, out_local);
}

void skb_launch_merge_topn(const uint32_t* idx_parts, const unsigned long long* sum_parts, uint32_t n_parts,
                           uint64_t n_reads, uint32_t top, uint32_t* out_idx, unsigned long long* out_sum,
                           cudaStream_t st) {
  if (n_reads == 0) return;
  const int th = 256;
  const unsigned blocks = (unsigned)((n_reads * 32 + th - 1) / th);
  merge_topn_kernel<<<blocks, th, 0, st>>>(idx_parts, sum_parts, n_parts, n_reads, top, out_idx, out_sum);
}

void skb_launch_ref_check(const uint64_t* ref, const uint64_t* row_off, uint32_t n_rows, uint32_t* bad,
                          unsigned long long* hmax, cudaStream_t st) {
  if (n_rows == 0) return;
  unsigned blocks = (n_rows + 7) / 8;
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  ref_check_kernel<<<blocks, 256, 0, st>>>(ref, row_off, n_rows, bad, hmax);
}

void skb_launch_shared(const uint64_t* ref, const uint64_t* row_off, uint32_t n_rows, const uint64_t* q,
                       const uint64_t* q_off, uint32_t Q, unsigned long long* out, cudaStream_t st) {
  const uint64_t pairs = (uint64_t)n_rows * Q;
  if (pairs == 0) return;
  const int th = 256;
  const unsigned blocks = (unsigned)((pairs * 32 + th - 1) / th);
  shared_kernel<<<blocks, th, 0, st>>>(ref, row_off, n_rows, q, q_off, Q, out);
}
