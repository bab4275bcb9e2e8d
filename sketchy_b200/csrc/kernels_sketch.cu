// kernels_sketch.cu — K1 (canonical k-mer + MurmurHash3_x64_128.h1 + threshold filter) and K2 (bottom-s:
// sort, dedup, occurrence counts, truncate). sm_100a.
//
// K1 replaces needletail `canonical_kmers` + finch `hash_f` reached through `sketcher.process(record)`
// (reference src/sketchy.rs:296, 333, 477); K2 replaces finch `MashSketcher::push` / `to_vec()`
// (src/sketchy.rs:302, 335, 480) using the closed form "s smallest distinct hashes with exact counts"
// (SURVEY.md Appendix A.4).
#include "kernels.h"

namespace {

// ---------------------------------------------------------------------------------------------------------
// MurmurHash3_x64_128, first 64 bits
// ---------------------------------------------------------------------------------------------------------
constexpr uint64_t MM_C1 = 0x87c37b91114253d5ull;
constexpr uint64_t MM_C2 = 0x4cf5ad432745937full;

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__device__ __forceinline__ uint64_t fmix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

__device__ __forceinline__ void mm3_block(uint64_t& h1, uint64_t& h2, uint64_t k1, uint64_t k2) {
  k1 *= MM_C1; k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
  h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ull;
  k2 *= MM_C2; k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
  h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ull;
}

__device__ __forceinline__ uint64_t mm3_finish_h1(uint64_t h1, uint64_t h2, uint64_t len) {
  h1 ^= len; h2 ^= len;
  h1 += h2; h2 += h1;
  h1 = fmix64(h1); h2 = fmix64(h2);
  return h1 + h2;
}

// generic length <= 32: w[i] holds bytes 8i..8i+7 little-endian, bytes >= len are zero.
__device__ __forceinline__ uint64_t mm3_h1_upto32(const uint64_t w[4], uint32_t len, uint64_t seed) {
  uint64_t h1 = seed, h2 = seed;
  uint32_t tb = 0;
  if (len >= 16) { mm3_block(h1, h2, w[0], w[1]); tb = 2; }
  if (len == 32) { mm3_block(h1, h2, w[2], w[3]); tb = 4; }
  const uint32_t t = len & 15u;
  if (t > 8) {
    uint64_t k2 = (tb == 0) ? w[1] : w[3];
    k2 *= MM_C2; k2 = rotl64(k2, 33); k2 *= MM_C1; h2 ^= k2;
  }
  if (t > 0) {
    uint64_t k1 = (tb == 0) ? w[0] : w[2];
    k1 *= MM_C1; k1 = rotl64(k1, 31); k1 *= MM_C2; h1 ^= k1;
  }
  return mm3_finish_h1(h1, h2, len);
}

// 4 packed bases (one byte, base i at bits 2i) -> 4 ASCII bytes, byte i = "ACGT"[code_i]
__device__ __forceinline__ uint32_t ascii4(uint32_t v) {
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) r |= ((0x54474341u >> (8 * ((v >> (2 * i)) & 3u))) & 0xFFu) << (8 * i);
  return r;
}

// OR of k consecutive bits: bit p of the result = any(m[p .. p+k))
__device__ __forceinline__ uint64_t window_or(uint64_t m, uint32_t k) {
  uint32_t r = 1;
  while (2 * r <= k) { m |= m >> r; r *= 2; }
  if (k > r) m |= m >> (k - r);
  return m;
}

template <bool DUMP>
__device__ __forceinline__ void emit_hash(const SkbHashArgs& a, uint32_t g, uint64_t tau, uint64_t base,
                                          uint32_t cap, uint64_t pos, uint64_t h, bool valid) {
  if (DUMP) {
    a.dump_hash[pos] = valid ? h : 0ull;
    a.dump_valid[pos] = valid ? 1 : 0;
  } else if (valid && h <= tau) {
    const uint32_t slot = atomicAdd(&a.cand_cnt[g], 1u);
    if (slot < cap) a.cand[base + slot] = h;
  }
}

// ---- k = 16 fast path: MurmurHash3_x64_128.h1 of one 16-byte block on explicit 32-bit halves -------------------
// The integer pipes bound this kernel, so the arithmetic is spelled the way the SASS should come out: a 64-bit
// multiply by a constant is IMAD.WIDE + 2 IMAD, a rotate is two funnel shifts, x ^= x >> 33 touches the low word only.
struct U2 { uint32_t lo, hi; };
__device__ __forceinline__ U2 mulc(U2 a, uint64_t c) {
  const uint32_t clo = (uint32_t)c, chi = (uint32_t)(c >> 32);
  U2 r;
  // spelled in PTX so that the three multiply-adds stay three (the compiler's own form is four: two products, the wide
  // product and an add)
  asm("{\n\t"
      ".reg .u64 t;\n\t"
      ".reg .u32 th;\n\t"
      "mul.wide.u32 t, %2, %4;\n\t"
      "mov.b64 {%0, th}, t;\n\t"
      "mad.lo.u32 th, %2, %5, th;\n\t"
      "mad.lo.u32 %1, %3, %4, th;\n\t"
      "}"
      : "=r"(r.lo), "=r"(r.hi)
      : "r"(a.lo), "r"(a.hi), "r"(clo), "r"(chi));
  return r;
}
template <int R>
__device__ __forceinline__ U2 rotl2(U2 a) {
  U2 o;
  if constexpr (R < 32) { o.hi = __funnelshift_l(a.lo, a.hi, R); o.lo = __funnelshift_l(a.hi, a.lo, R); }
  else { o.hi = __funnelshift_l(a.hi, a.lo, R - 32); o.lo = __funnelshift_l(a.lo, a.hi, R - 32); }
  return o;
}
__device__ __forceinline__ U2 add2(U2 a, U2 b) {
  const uint64_t x = (((uint64_t)a.hi << 32) | a.lo) + (((uint64_t)b.hi << 32) | b.lo);
  U2 r; r.lo = (uint32_t)x; r.hi = (uint32_t)(x >> 32);
  return r;
}
// a * 5 + c; c64 = c in a 64-bit register pair the caller keeps live (IMAD.WIDE takes it as its addend)
__device__ __forceinline__ U2 mul5add(U2 a, uint64_t c64) {
  const uint64_t t = (uint64_t)a.lo * 5u + c64;
  U2 r; r.lo = (uint32_t)t; r.hi = a.hi * 5u + (uint32_t)(t >> 32);
  return r;
}
// fmix64 without its last step (x ^= x >> 33 changes the low word only: the caller applies it when it needs the
// low word at all)
__device__ __forceinline__ U2 fmix_head(U2 x) {
  x.lo ^= x.hi >> 1;
  x = mulc(x, 0xff51afd7ed558ccdull);
  x.lo ^= x.hi >> 1;
  return mulc(x, 0xc4ceb9fe1a85ec53ull);
}
// A, B with h = (A ^ (A >> 33)) + (B ^ (B >> 33)); k1 = bytes 0..7 of the k-mer (little endian), k2 = bytes 8..15
template <bool SEED0>
__device__ __forceinline__ void mm3_k16_heads(U2 k1, U2 k2, uint32_t seed_lo, uint32_t seed_hi, uint64_t c52, uint64_t c38,
                                              U2& A, U2& B) {
  k1 = mulc(k1, MM_C1); k1 = rotl2<31>(k1); k1 = mulc(k1, MM_C2);
  k2 = mulc(k2, MM_C2); k2 = rotl2<33>(k2); k2 = mulc(k2, MM_C1);
  U2 h1 = k1, h2 = k2;
  if (!SEED0) { h1.lo ^= seed_lo; h1.hi ^= seed_hi; h2.lo ^= seed_lo; h2.hi ^= seed_hi; }
  h1 = rotl2<27>(h1);
  if (!SEED0) { U2 sd; sd.lo = seed_lo; sd.hi = seed_hi; h1 = add2(h1, sd); }
  h1 = mul5add(h1, c52);
  h2 = rotl2<31>(h2);
  h2 = add2(h2, h1);
  h2 = mul5add(h2, c38);
  h1.lo ^= 16u; h2.lo ^= 16u;  // len
  h1 = add2(h1, h2);
  h2 = add2(h2, h1);
  A = fmix_head(h1);
  B = fmix_head(h2);
}
__device__ __forceinline__ uint64_t mm3_k16_finish(U2 A, U2 B) {
  A.lo ^= A.hi >> 1; B.lo ^= B.hi >> 1;
  return (((uint64_t)A.hi << 32) | A.lo) + (((uint64_t)B.hi << 32) | B.lo);
}

#ifndef SKB_X_HASH_UNROLL
#define SKB_X_HASH_UNROLL 8  // 4, 8 or 16: k-mers per unrolled body (the fully unrolled chunk was 64 KB of SASS: instruction-fetch stalls)
#endif

// One warp per segment (<= 32 chunks of one group); one lane per chunk of 32 k-mer start positions.
template <bool K16, bool DUMP, bool SEED0>
__global__ void __launch_bounds__(256) hash_kernel(const SkbHashArgs a) {
  __shared__ __align__(1024) uint32_t lut[256];
  if (K16) {
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = ascii4(i);
    __syncthreads();
  }
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= a.pv.nseg) return;
  const uint32_t lane = skb_lane();
  const uint32_t g = a.pv.seg_group[warp];
  if (a.active && !a.active[g]) return;
  const uint32_t nchunk = a.pv.seg_n[warp];
  uint32_t nvalid = 0;
  if (lane < nchunk) {
    const uint64_t chunk = (uint64_t)a.pv.seg_chunk0[warp] + lane;
    const uint32_t m0 = __ldg(a.pv.nmask + chunk), m1 = __ldg(a.pv.nmask + chunk + 1);
    const uint64_t inv = window_or((uint64_t)m0 | ((uint64_t)m1 << 32), a.k);
    const uint32_t valid = ~(uint32_t)inv;
    nvalid = __popc(valid);
    if (valid != 0u || DUMP) {
      const uint64_t tau = DUMP ? 0 : a.tau[g];
      const uint64_t pos0 = chunk * SKB_CHUNK;
      const uint32_t* cw = a.pv.codes + chunk * 2;
      if (K16) {
        const uint32_t w0 = __ldg(cw), w1 = __ldg(cw + 1), w2 = __ldg(cw + 2);
        // Values the loop keeps in registers; they pass through an empty asm so that the compiler does not rebuild them
        // from their definitions for every k-mer. lut_base: byte address of the table in the shared window, 1 KB
        // aligned, so (index bits | base) is one LOP3. tau_hi1: a hash can be <= tau only if its high word (without the
        // carry of the low words) + 1 <= tau's high word + 1 (saturating).
        uint32_t lut_base = (uint32_t)__cvta_generic_to_shared(lut);
        uint32_t tau_hi1 = (uint32_t)(tau >> 32) == 0xFFFFFFFFu ? 0xFFFFFFFFu : (uint32_t)(tau >> 32) + 1u;
        uint64_t c52 = 0x52dce729ull, c38 = 0x38495ab5ull;
        asm volatile("" : "+r"(lut_base), "+r"(tau_hi1), "+l"(c52), "+l"(c38));
        const uint32_t seed_lo = (uint32_t)a.seed, seed_hi = (uint32_t)(a.seed >> 32);
        uint32_t x = __brev(w0);
        uint32_t fwd_be = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);  // the window, first base most significant
        fwd_be >>= 2;  // (the loop shifts the incoming base in before use)
#pragma unroll 1
        for (uint32_t j0 = 0; j0 < 32; j0 += SKB_X_HASH_UNROLL) {
          // bases j0 .. j0 + 31 of the lane's 48-base window (w3 would be the next lane's: not needed, k - 1 + 32 <= 48)
          const bool up = j0 >= 16;
          const uint32_t sh0 = 2u * (j0 & 15u);
          const uint32_t wa = up ? w1 : w0, wb = up ? w2 : w1, wc = up ? 0u : w2;
          const uint32_t lo = __funnelshift_r(wa, wb, sh0), hi = __funnelshift_r(wb, wc, sh0);
#pragma unroll
          for (uint32_t jj = 0; jj < SKB_X_HASH_UNROLL; ++jj) {
            const uint32_t j = j0 + jj;
            const uint32_t fwd_le = jj ? __funnelshift_r(lo, hi, 2 * jj) : lo;  // base j + i at bits 2i
            fwd_be = (fwd_be << 2) | (fwd_le >> 30);
            // reverse complement, first base most significant, is ~fwd_le; in the little-endian layout it is ~fwd_be
            const uint32_t canon = (~fwd_le) < fwd_be ? ~fwd_be : fwd_le;
            uint32_t a0, a1, a2, a3;
            asm("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(((canon << 2) & 0x3FCu) | lut_base));
            asm("ld.shared.u32 %0, [%1];" : "=r"(a1) : "r"(((canon >> 6) & 0x3FCu) | lut_base));
            asm("ld.shared.u32 %0, [%1];" : "=r"(a2) : "r"(((canon >> 14) & 0x3FCu) | lut_base));
            asm("ld.shared.u32 %0, [%1];" : "=r"(a3) : "r"(((canon >> 22) & 0x3FCu) | lut_base));
            U2 k1, k2, A, B;
            k1.lo = a0; k1.hi = a1; k2.lo = a2; k2.hi = a3;
            mm3_k16_heads<SEED0>(k1, k2, seed_lo, seed_hi, c52, c38, A, B);
            if (DUMP) {
              const uint64_t h = mm3_k16_finish(A, B);
              const bool ok = (valid >> j) & 1u;
              a.dump_hash[pos0 + j] = ok ? h : 0ull;
              a.dump_valid[pos0 + j] = ok ? 1 : 0;
            } else if (A.hi + B.hi + 1u <= tau_hi1) {  // rare
              const uint64_t h = mm3_k16_finish(A, B);
              if (((valid >> j) & 1u) && h <= tau) {
                const uint32_t slot = atomicAdd(&a.cand_cnt[g], 1u);
                if (slot < __ldg(a.cand_cap + g)) a.cand[__ldg(a.cand_base + g) + slot] = h;
              }
            }
          }
        }
      } else {
        const uint64_t base = DUMP ? 0 : a.cand_base[g];
        const uint32_t cap = DUMP ? 0 : a.cand_cap[g];
        const uint32_t k = a.k;
        const uint64_t lo = (uint64_t)__ldg(cw) | ((uint64_t)__ldg(cw + 1) << 32);
        const uint64_t hi = (uint64_t)__ldg(cw + 2) | ((uint64_t)__ldg(cw + 3) << 32);
        const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
        uint64_t fwd = 0, rc = 0;
        // prime with the first k-1 bases, then one base per k-mer
        for (uint32_t i = 0; i < 31 + k; ++i) {
          const uint64_t c = (i < 32) ? ((lo >> (2 * i)) & 3ull) : ((hi >> (2 * (i - 32))) & 3ull);
          fwd = ((fwd << 2) | c) & kmask;
          rc = (rc >> 2) | ((3ull - c) << (2 * (k - 1)));
          if (i + 1 >= k) {
            const uint32_t j = i + 1 - k;
            const uint64_t canon = fwd < rc ? fwd : rc;  // first base most significant == lexicographic
            uint64_t w[4] = {0, 0, 0, 0};
#pragma unroll
            for (uint32_t b = 0; b < 32; ++b) {
              if (b < k) {
                const uint32_t code = (uint32_t)(canon >> (2 * (k - 1 - b))) & 3u;
                w[b >> 3] |= (uint64_t)((0x54474341u >> (8 * code)) & 0xFFu) << (8 * (b & 7));
              }
            }
            const uint64_t h = mm3_h1_upto32(w, k, a.seed);
            emit_hash<DUMP>(a, g, tau, base, cap, pos0 + j, h, (valid >> j) & 1u);
          }
        }
      }
    }
  }
  if (!DUMP) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    if (lane == 0 && nvalid) atomicAdd(&a.kmers[g], (unsigned long long)nvalid);
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2: one CTA per group. Sort the group's candidates (bitonic; shared memory when they fit, else in place in
// global memory), drop duplicates keeping occurrence counts, keep the s smallest.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cta_bitonic_sort(uint64_t* buf, uint32_t n2) {
  for (uint32_t size = 2; size <= n2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const uint32_t i = 2 * t - (t & (stride - 1));
        const uint32_t j = i + stride;
        const bool up = (i & size) == 0;
        const uint64_t x = buf[i], y = buf[j];
        if ((x > y) == up) { buf[i] = y; buf[j] = x; }
      }
      __syncthreads();
    }
  }
}

// exclusive scan of one flag per thread across the CTA; returns this thread's offset, *total = CTA sum
__device__ __forceinline__ uint32_t cta_scan_flag(bool flag, uint32_t* warp_sums, uint32_t* total) {
  const uint32_t lane = skb_lane(), wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const uint32_t bal = __ballot_sync(0xffffffffu, flag);
  const uint32_t in_warp = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) warp_sums[wid] = __popc(bal);
  __syncthreads();
  uint32_t off = 0, tot = 0;
  for (uint32_t w = 0; w < nw; ++w) {
    const uint32_t v = warp_sums[w];
    if (w < wid) off += v;
    tot += v;
  }
  __syncthreads();
  *total = tot;
  return off + in_warp;
}

__global__ void select_kernel(const SkbSelectArgs a) {
  extern __shared__ __align__(16) uint64_t sbuf[];
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t start_s_sh;
  const uint32_t g = blockIdx.x;
  if (a.active && !a.active[g]) return;
  const uint32_t cnt = a.cand_cnt[g], cap = a.cand_cap[g];
  if (cnt > cap) {
    if (threadIdx.x == 0) {
      a.status[g] = SKB_ST_OVERFLOW; a.out_n[g] = 0;
      if (a.any_bad) atomicOr(a.any_bad, 1u);
    }
    return;
  }
  uint64_t* gbuf = a.cand + a.cand_base[g];
  uint32_t n2 = 1;
  while (n2 < cnt) n2 <<= 1;
  const bool in_smem = n2 <= a.smem_elems;
  uint64_t* buf = in_smem ? sbuf : gbuf;
  if (in_smem) {
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) sbuf[i] = i < cnt ? gbuf[i] : SKB_EMPTY_KEY;
  } else {
    for (uint32_t i = cnt + threadIdx.x; i < n2; i += blockDim.x) gbuf[i] = SKB_EMPTY_KEY;  // n2 <= cap (pow2)
  }
  if (threadIdx.x == 0) start_s_sh = cnt;
  __syncthreads();
  cta_bitonic_sort(buf, n2);

  uint64_t* out = a.out_hashes + (a.out_off ? a.out_off[g] : (uint64_t)g * a.s);
  uint32_t* cnt_out = a.out_counts ? a.out_counts + (uint64_t)g * a.s : nullptr;
  uint32_t distinct = 0;
  for (uint32_t base = 0; base < cnt; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    uint64_t v = 0;
    bool head = false;
    if (i < cnt) {
      v = buf[i];
      head = (i == 0) || (buf[i - 1] != v);
    }
    __syncthreads();  // all reads of this tile done before any in-place write
    uint32_t tot;
    const uint32_t rank = distinct + cta_scan_flag(head, warp_sums, &tot);
    if (head) {
      if (rank < a.s) {
        out[rank] = v;
        if (cnt_out) cnt_out[rank] = i;  // start index of the run; turned into a count below
      } else if (rank == a.s) {
        start_s_sh = i;
      }
    }
    distinct += tot;
    __syncthreads();
  }
  const uint32_t n_out = distinct < a.s ? distinct : a.s;
  if (cnt_out) {
    const uint32_t end_all = start_s_sh;  // start of run #s if it exists, else cnt
    for (uint32_t base = 0; base < n_out; base += blockDim.x) {
      const uint32_t r = base + threadIdx.x;
      uint32_t st = 0, nx = 0;
      if (r < n_out) {
        st = cnt_out[r];
        nx = (r + 1 < n_out) ? cnt_out[r + 1] : end_all;
      }
      __syncthreads();
      if (r < n_out) cnt_out[r] = nx - st;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    a.out_n[g] = n_out;
    const uint32_t st = (a.check_underfill && distinct < a.s && a.tau[g] != SKB_EMPTY_KEY) ? SKB_ST_UNDERFILL : SKB_ST_OK;
    a.status[g] = st;
    if (a.any_bad && st != SKB_ST_OK) atomicOr(a.any_bad, 1u);
  }
}

// block-wide sum of one value per thread (1024 threads)
__device__ __forceinline__ unsigned long long cta_sum_u64(unsigned long long v, unsigned long long* warp_tot) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (skb_lane() == 0) warp_tot[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long t = 0;
  for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) t += warp_tot[w];
  __syncthreads();
  return t;
}

// plan, first launch: one CTA per tile of 1024 groups. Capacities, the tile's total, the per-group scratch reset.
__global__ void __launch_bounds__(1024) plan_caps_kernel(const SkbPlanArgs a) {
  __shared__ unsigned long long warp_tot[32];
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t cap = 0;
  if (g < a.n_groups) {
    const float n = fmaxf((float)a.g_len[g], 1.0f);
    const float est = fminf(4.0f * (n * (float)a.frac) + 16.0f, n);  // (single precision: a capacity plan, not a result)
    cap = 32;
    while ((float)cap < est && cap < 0x80000000u) cap <<= 1;
    a.cap[g] = cap;
    a.tau_out[g] = a.tau;
    a.cnt[g] = 0; a.kmers[g] = 0; a.active[g] = 1; a.status[g] = 0;
  }
  const unsigned long long tot = cta_sum_u64(cap, warp_tot);
  if (threadIdx.x == 0) a.tile_tot[blockIdx.x] = tot;
  if (g == 0) *a.any_bad = 0;
}

// plan, second launch: every group's place in the pool = the totals of the tiles before its own + an exclusive scan
// within the tile.
__global__ void __launch_bounds__(1024) plan_bases_kernel(const SkbPlanArgs a) {
  __shared__ unsigned long long warp_tot[32];
  unsigned long long before = 0;
  for (uint32_t t = threadIdx.x; t < blockIdx.x; t += blockDim.x) before += a.tile_tot[t];
  before = cta_sum_u64(before, warp_tot);
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x, lane = skb_lane(), wid = threadIdx.x >> 5;
  const uint32_t cap = g < a.n_groups ? a.cap[g] : 0u;
  unsigned long long incl = cap;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane >= o) incl += y;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  for (uint32_t w = 0; w < wid; ++w) before += warp_tot[w];
  if (g < a.n_groups) a.base[g] = before + incl - cap;
}

__global__ void compact_queries_kernel(const uint64_t* __restrict__ cand, const uint64_t* __restrict__ cand_base,
                                       const uint32_t* __restrict__ out_n, const uint64_t* __restrict__ q_off,
                                       uint32_t n_reads, uint64_t* __restrict__ qh) {
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_reads) return;
  const uint32_t n = out_n[r];
  const uint64_t src = cand_base[r], dst = q_off[r];
  for (uint32_t j = skb_lane(); j < n; j += 32) qh[dst + j] = cand[src + j];
}

// qread[j] = the pass read whose query list holds flat position j
__global__ void fill_qread_kernel(const uint64_t* __restrict__ q_off, uint32_t n_reads, uint32_t* __restrict__ qread) {
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_reads) return;
  for (uint64_t j = q_off[r] + skb_lane(); j < q_off[r + 1]; j += 32) qread[j] = r;
}

// rows of the scratch ranking that belong to the last piece of a read go to the caller's arrays
__global__ void report_pieces_kernel(const uint32_t* __restrict__ idx, const unsigned long long* __restrict__ sum,
                                     const uint32_t* __restrict__ out_row, uint32_t n_pieces, uint32_t top,
                                     uint32_t* __restrict__ out_idx, unsigned long long* __restrict__ out_sum) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (uint64_t)n_pieces * top) return;
  const uint32_t p = (uint32_t)(e / top), t = (uint32_t)(e % top);
  const uint32_t row = out_row[p];
  if (row == 0xFFFFFFFFu) return;
  out_idx[(size_t)row * top + t] = idx[e];
  out_sum[(size_t)row * top + t] = sum[e];
}

}  // namespace

void skb_launch_hash(const SkbHashArgs& a, cudaStream_t st) {
  if (a.pv.nseg == 0) return;
  const int threads = 256;
  const unsigned blocks = (unsigned)(((uint64_t)a.pv.nseg * 32 + threads - 1) / threads);
  const bool dump = a.dump_hash != nullptr;
  if (a.k == 16) {
    if (dump) hash_kernel<true, true, false><<<blocks, threads, 0, st>>>(a);
    else if (a.seed == 0) hash_kernel<true, false, true><<<blocks, threads, 0, st>>>(a);
    else hash_kernel<true, false, false><<<blocks, threads, 0, st>>>(a);
  } else {
    if (dump) hash_kernel<false, true, false><<<blocks, threads, 0, st>>>(a);
    else hash_kernel<false, false, false><<<blocks, threads, 0, st>>>(a);
  }
}

void skb_launch_select(const SkbSelectArgs& a, cudaStream_t st) {
  if (a.n_groups == 0) return;
  const size_t smem = (size_t)a.smem_elems * sizeof(uint64_t);
  static SkbSmemOptIn opt_in;
  if (smem > 48 * 1024 && opt_in.needs(smem))
    cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  select_kernel<<<a.n_groups, a.threads, smem, st>>>(a);
}

void skb_launch_plan_query(const SkbPlanArgs& a, cudaStream_t st) {
  if (a.n_groups == 0) return;
  const unsigned tiles = (a.n_groups + 1023) / 1024;
  plan_caps_kernel<<<tiles, 1024, 0, st>>>(a);
  plan_bases_kernel<<<tiles, 1024, 0, st>>>(a);
}

void skb_launch_compact_queries(const uint64_t* cand, const uint64_t* cand_base, const uint32_t* out_n,
                                const uint64_t* q_off, uint32_t n_reads, uint64_t* qh, cudaStream_t st) {
  if (n_reads == 0) return;
  const int threads = 256;
  const unsigned blocks = (unsigned)(((uint64_t)n_reads * 32 + threads - 1) / threads);
  compact_queries_kernel<<<blocks, threads, 0, st>>>(cand, cand_base, out_n, q_off, n_reads, qh);
}

void skb_launch_fill_qread(const uint64_t* q_off, uint32_t n_reads, uint32_t* qread, cudaStream_t st) {
  if (n_reads == 0) return;
  const int threads = 256;
  const unsigned blocks = (unsigned)(((uint64_t)n_reads * 32 + threads - 1) / threads);
  fill_qread_kernel<<<blocks, threads, 0, st>>>(q_off, n_reads, qread);
}

void skb_launch_report_pieces(const uint32_t* idx, const unsigned long long* sum, const uint32_t* out_row,
                              uint32_t n_pieces, uint32_t top, uint32_t* out_idx, unsigned long long* out_sum,
                              cudaStream_t st) {
  const uint64_t n = (uint64_t)n_pieces * top;
  if (n == 0) return;
  report_pieces_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(idx, sum, out_row, n_pieces, top, out_idx, out_sum);
}
