// kernels_predict.cu — streaming predict on sm_100a.
//
//   table_*       : per-pass hash set of the pass's query hashes: a global open-addressing table of 16-byte slots
//                   (L2 resident; a slot holds the key, its count and up to three read ids inline) plus a 64 KB bit
//                   filter (three bits per key in one 32-bit word) that every CTA keeps in shared memory. A table is
//                   cleared through the keys of its previous build.
//   fused_kernel  : the HBM-bound kernel. Persistent, one CTA per SM, each owning a contiguous range of reference
//                   rows. A row is cut into claims of two 256-hash sub-tiles; the CTA's 32 warps take claims in order
//                   from a shared counter and stream them through a PRIVATE two-stage cp.async.bulk (TMA) + mbarrier
//                   ring that the warp refills itself. A warp probes every staged hash against the filter and
//                   resolves the few passers on the spot: the lanes that hold one look the hash up in the table (one
//                   16-byte load) and add the key's reads to the row's counters in shared memory. The warp whose
//                   claim completes a row does the row's rank step (new running sum; lane segments of reads that
//                   reach their bound are copied out as records / intervals for the post-pass kernels) and hands
//                   the counter buffer back; 4 or 8 rows are in flight. No FIFO, no rank warps, no polling.
//                   HBM traffic per pass = the reference matrix, once.
//   rank_*        : bounds from the tracked rows, candidate lists from the records and intervals, exact per-read
//                   top-N by (sum desc, index asc); brute-force ranking from prefix sums for the first pass after
//                   a reset, small shards and redone passes.
//
// Replaces `_common_hashes` x N + `sum[i] += shared` + stable sort + `[..top]`
// (reference src/sketchy.rs:337-348, 391, 419-459). Rows are strictly increasing (checked at upload) and each
// read's query list is distinct, so the two-pointer merge count equals the set-intersection size computed here.
#include "kernels.h"

namespace {

// Home slot of a key: the low word is already MurmurHash3 output, one 32-bit multiply spreads it further (the choice
// affects probe lengths only, never results).
__device__ __forceinline__ uint32_t table_home(uint64_t h, uint32_t log2cap) {
  return ((uint32_t)h * 0x9E3779B1u) >> (32 - log2cap);
}
// Filter geometry, chosen for the probe's instruction count: the word's BYTE offset is lo & 0xFFFC (one LOP); a key
// sets SKB_BLOOM_K bits of its word at positions 31 - f, f = 5-bit fields of the hash (low bits of hi, bits 16-20 and
// 21-25 of lo). Shifting the word LEFT by f (wrap-mode shifts take f mod 32 without masking) brings the bit to position
// 31, so the conjunction of the shifted words has the verdict in its top bit and one funnel shift appends it to the
// chunk's passer mask. All of these hash bits are uniform for any reference maximum >= 2^37.
#ifndef SKB_BLOOM_K
#define SKB_BLOOM_K 3
#endif
__device__ __forceinline__ uint32_t bloom_word(uint32_t lo) { return (lo >> 2) & (SKB_BLOOM_WORDS - 1u); }
__device__ __forceinline__ uint32_t bloom_mask(uint32_t lo, uint32_t hi) {
  uint32_t m = 0x80000000u >> (hi & 31u);
  if (SKB_BLOOM_K >= 2) m |= 0x80000000u >> ((lo >> 16) & 31u);
  if (SKB_BLOOM_K >= 3) m |= 0x80000000u >> ((lo >> 21) & 31u);
  if (SKB_BLOOM_K >= 4) m |= 0x80000000u >> ((lo >> 26) & 31u);
  return m;
}
// top bit set when every bit of the key (lo, hi) is set in its filter word (the lower bits are garbage)
__device__ __forceinline__ uint32_t bloom_probe(const uint32_t* bloom, uint32_t lo, uint32_t hi) {
  const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(bloom) + (lo & (4u * SKB_BLOOM_WORDS - 4u)));
  uint32_t r = w << (hi & 31u);
  if (SKB_BLOOM_K >= 2) r &= w << ((lo >> 16) & 31u);
  if (SKB_BLOOM_K >= 3) r &= w << ((lo >> 21) & 31u);
  if (SKB_BLOOM_K >= 4) r &= w << ((lo >> 26) & 31u);
  return r;
}

__device__ __forceinline__ SkbSlot load_slot(const SkbSlot* p) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  SkbSlot s;
  s.key = ((unsigned long long)v.y << 32) | v.x;
  s.meta = ((unsigned long long)v.w << 32) | v.z;
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// reference membership filter (query-side prefilter)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void memb_pos(uint64_t h, uint32_t log2w, uint32_t& word, uint32_t& mask) {
  const uint64_t x = h * 0xD6E8FEB86659FD93ull;
  word = (uint32_t)(x >> (64 - log2w));
  mask = (1u << (x & 31u)) | (1u << ((x >> 5) & 31u)) | (1u << ((x >> 10) & 31u));
}
__device__ __forceinline__ bool memb_test(const SkbTable& t, uint64_t h) {
  uint32_t word, mask;
  memb_pos(h, t.memb_log2, word, mask);
  return (__ldg(t.memb + word) & mask) == mask;
}
__global__ void __launch_bounds__(256) memb_build_kernel(const SkbRefView rv, uint32_t* memb, uint32_t log2w) {
  const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = warp0; r < rv.n_rows; r += nwarps) {
    const uint64_t* row = rv.ref + rv.row_start[r];
    const uint32_t n = rv.row_len[r];
    for (uint32_t i = skb_lane(); i < n; i += 32) {
      uint32_t word, mask;
      memb_pos(row[i], log2w, word, mask);
      if ((memb[word] & mask) != mask) atomicOr(memb + word, mask);  // near-identical rows: most bits are set already
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// query table
// ---------------------------------------------------------------------------------------------------------
// A table is cleared through the keys of its previous build (a pass fills a few per cent of the slots): every slot a
// key was put in goes back to empty. n_prev == UINT32_MAX: the table is new, every slot is cleared.
__global__ void table_clear_kernel(SkbTable t, uint32_t n_prev) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_prev == 0xFFFFFFFFu) {
    if (i <= t.cap) {
      reinterpret_cast<uint4*>(t.slots)[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
      t.fill[i] = 0;
    }
  } else if (i < n_prev) {
    const uint32_t s = t.slot_of[i];
    if (s != 0xFFFFFFFFu) {
      reinterpret_cast<uint4*>(t.slots)[s & 0x7FFFFFFFu] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
      t.fill[s & 0x7FFFFFFFu] = 0;
    }
  }
  if (i < SKB_BLOOM_WORDS) t.bloom[i] = 0;
  if (i == 0) *t.cursor = 0;
}

// slot_of[i] = the key's slot, bit 31 set for the key that opened the slot (it sizes the slot's read list afterwards)
__global__ void table_insert_kernel(SkbTable t, const uint64_t* __restrict__ qh, uint32_t n_keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys) return;
  const uint64_t h = qh[i];
  if (t.memb && !memb_test(t, h)) {  // in no reference row: it can never be hit
    t.slot_of[i] = 0xFFFFFFFFu;
    return;
  }
  if (t.memb_kept) atomicAdd(t.memb_kept, 1ull);
  uint32_t slot;
  if (h == SKB_EMPTY_KEY) {
    slot = t.cap;
  } else {
    slot = table_home(h, t.log2cap);
    for (;;) {
      const unsigned long long prev = atomicCAS(&t.slots[slot].key, SKB_EMPTY_KEY, (unsigned long long)h);
      if (prev == SKB_EMPTY_KEY || prev == h) break;
      slot = (slot + 1) & (t.cap - 1);
    }
  }
  // cnt lives in the low bits; a read holds a hash at most once. The first to count opened the slot.
  const bool opened = SKB_SLOT_CNT(atomicAdd(&t.slots[slot].meta, 1ull)) == 0u;
  t.slot_of[i] = slot | (opened ? 0x80000000u : 0u);
  const uint32_t lo = (uint32_t)h;
  atomicOr(&t.bloom[bloom_word(lo)], bloom_mask(lo, (uint32_t)(h >> 32)));
}

// read lists longer than the inline ids get their place in `reads`: done by the key that opened the slot
__global__ void table_alloc_kernel(SkbTable t, uint32_t n_keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys) return;
  const uint32_t so = t.slot_of[i];
  if (so == 0xFFFFFFFFu || !(so & 0x80000000u)) return;
  const uint32_t s = so & 0x7FFFFFFFu;
  const unsigned long long m = t.slots[s].meta;
  const uint32_t c = SKB_SLOT_CNT(m);
  if (c > SKB_SLOT_INLINE) t.slots[s].meta = m | ((unsigned long long)atomicAdd(t.cursor, c) << SKB_SLOT_CNT_BITS);
}

__global__ void table_fill_kernel(SkbTable t, const uint32_t* __restrict__ qread, uint32_t n_keys,
                                  uint32_t read_base) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_keys) return;
  const uint32_t so = t.slot_of[i];
  if (so == 0xFFFFFFFFu) return;  // dropped by the membership prefilter
  const uint32_t s = so & 0x7FFFFFFFu;
  const uint32_t rd = qread[i] - read_base;
  // cnt (and, for long lists, the start) are final here; only the inline id bits are still being OR-ed in
  const unsigned long long m = *reinterpret_cast<volatile unsigned long long*>(&t.slots[s].meta);
  const uint32_t pos = atomicAdd(&t.fill[s], 1u);
  if (SKB_SLOT_CNT(m) <= SKB_SLOT_INLINE) {
    atomicOr(&t.slots[s].meta, (unsigned long long)rd << SKB_SLOT_ID_SHIFT(pos));
  } else {
    t.reads[SKB_SLOT_START(m) + pos] = rd;
  }
}

// exact lookup; returns true and the slot contents when h is a query hash of this pass
__device__ __forceinline__ bool table_lookup(const SkbTable& t, uint64_t h, SkbSlot& out) {
  if (h == SKB_EMPTY_KEY) {
    out = load_slot(&t.slots[t.cap]);
    return SKB_SLOT_CNT(out.meta) != 0;
  }
  uint32_t slot = table_home(h, t.log2cap);
  for (;;) {
    out = load_slot(&t.slots[slot]);
    if (out.key == h) return true;
    if (out.key == SKB_EMPTY_KEY) return false;
    slot = (slot + 1) & (t.cap - 1);
  }
}

// ---------------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives
// ---------------------------------------------------------------------------------------------------------
#ifndef SKB_X_SUB
#define SKB_X_SUB 256
#endif
#ifndef SKB_X_STAGES
#define SKB_X_STAGES 2
#endif
#ifndef SKB_X_CW
#define SKB_X_CW 32
#endif
#ifndef SKB_X_ABLATE
#define SKB_X_ABLATE 0  // experiments only (never in the shipped build): 1 = no filter probe, 2 = probe but drop the passers, 4 = no candidate walk
#endif
constexpr int FS_SUB = SKB_X_SUB;        // hashes per sub-tile (2 KB): chunks of 8 per lane
constexpr int FS_STAGES = SKB_X_STAGES;  // staging buffers per warp
constexpr int FS_WARPS = SKB_X_CW;
constexpr int FS_THREADS = FS_WARPS * 32;
constexpr int FS_ROWBUF_MAX = 8;         // rows in flight per CTA: a.rowbuf (4 or 8) counter buffers, indexed by row % a.rowbuf
constexpr int FS_NHASH = 8;              // hashes per lane per chunk (one chunk = 256 hashes)
constexpr int FS_CHUNKS = FS_SUB / (32 * FS_NHASH);
static_assert(FS_SUB % (32 * FS_NHASH) == 0, "a sub-tile is a whole number of chunks");
constexpr size_t FS_SMEM_BLOOM = (size_t)SKB_BLOOM_WORDS * 4;
constexpr size_t FS_SMEM_RING = (size_t)FS_WARPS * FS_STAGES * FS_SUB * 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// the try_wait suspends the warp in hardware (up to the hint) instead of spinning on the issue port
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(1000000u)
      : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier; streamed data is marked
// evict-first so the query table keeps its place in L2.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                          uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint32_t lds_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

// two u16 counters per 32-bit word; a per-(row, read) count never exceeds 65535 (checked on the host)
template <int CPW>
__device__ __forceinline__ void count_hit(uint32_t* cbuf, uint32_t rd) {
  if (CPW == 2) atomicAdd(cbuf + (rd >> 1), 1u << (16 * (rd & 1u)));
  else atomicAdd(cbuf + (rd >> 2), 1u << (8 * (rd & 3u)));  // u8 counters: the host guarantees counts <= 255
}

// add the slot's reads to the row's counters (cnt == 0: the reserved slot of the all-ones key when no read holds it)
template <int CPW>
__device__ __forceinline__ void apply_hit(const SkbTable& t, unsigned long long meta, uint32_t* cbuf) {
  const uint32_t c = SKB_SLOT_CNT(meta);
  if (c <= SKB_SLOT_INLINE) {
    unsigned long long ids = meta >> SKB_SLOT_CNT_BITS;
#pragma unroll 1
    for (uint32_t j = 0; j < c; ++j) {
      count_hit<CPW>(cbuf, (uint32_t)ids & ((1u << SKB_SLOT_ID_BITS) - 1u));
      ids >>= SKB_SLOT_ID_BITS;
    }
  } else {
    const uint32_t st = SKB_SLOT_START(meta);
#pragma unroll 1
    for (uint32_t j = 0; j < c; ++j) count_hit<CPW>(cbuf, t.reads[st + j]);
  }
}

// Row bookkeeping of one CTA (static shared memory).
struct FsCtl {
  unsigned long long lb_seg[32];  // bound of the first read of each lane segment (bounds never decrease along the reads)
  uint32_t li_seg[32];            // largest bound index within each lane segment of the reads
  uint32_t done[FS_ROWBUF_MAX];   // sub-tiles of the buffer's current row that are fully counted
  uint32_t freed[FS_ROWBUF_MAX];  // rows of this buffer that have been ranked (the buffer is zero again)
  uint32_t next_tile;             // next unclaimed claim unit of the CTA's rows (claimed in order)
};
constexpr uint32_t FS_NONE = 0xFFFFFFFFu;

constexpr uint32_t FS_CLAIM = (uint32_t)FS_SUB * FS_STAGES;       // hashes per claim: one ring's worth of a row
__device__ __forceinline__ uint32_t fs_claims_of(uint32_t len) {  // an empty row still has one (empty) claim:
  const uint32_t n = (len + FS_CLAIM - 1) / FS_CLAIM;             // somebody has to close and rank it
  return n ? n : 1u;
}

// Rank work for one finished row, done by the warp whose sub-tile completed it. It is deliberately short, because
// the row's counter buffer is held meanwhile: the counters' total gives the row's new running sum; a lane segment
// of the reads (cnt_stride / 32 consecutive reads) whose final sum reaches the bound of the segment's first read
// may hold top-N candidates and is handed to the post-pass kernels, which test every read against its exact bound:
//   - a segment with hits is copied out as a record {sum at segment start, row, first read, the segment's counters}
//     (walk_kernel);
//   - a segment without hits holds one sum for all its reads: an interval (expand_kernel).
// Sums and bounds never decrease along the reads, so a segment (or a row) whose FINAL sum is under its FIRST bound
// cannot hold a candidate. On a tie the row index decides: the segment passes when the row's index does not exceed
// the largest bound index within it (a superset; the exact test comes later).
template <int CPW>
__device__ __forceinline__ void rank_row(const SkbFusedArgs& a, const FsCtl& ctl, uint32_t* cpar, uint32_t cwords,
                                         uint32_t row_in_shard) {
  const uint32_t lane = skb_lane();
  const uint32_t per = a.cnt_stride >> 5;  // reads per lane segment (multiple of 16)
  const uint32_t seg0 = lane * per;
  const uint32_t segw = per / CPW;  // words in this lane's segment (multiple of 4)
  const unsigned long long carry = a.sums_in[row_in_shard];
  const uint32_t* cseg = cpar ? cpar + seg0 / CPW : nullptr;
  uint32_t tot = 0;
  if (cpar) {
    for (uint32_t i = 0; i < segw; i += 4) {
      const uint4 x = *reinterpret_cast<const uint4*>(cseg + i);
      if (CPW == 2) {
        tot += (x.x & 0xFFFFu) + (x.x >> 16) + (x.y & 0xFFFFu) + (x.y >> 16) + (x.z & 0xFFFFu) + (x.z >> 16) +
               (x.w & 0xFFFFu) + (x.w >> 16);
      } else {  // four u8 counters per word: sum the byte lanes with a masked add
        const uint32_t w4[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t pair = (w4[q] & 0x00FF00FFu) + ((w4[q] >> 8) & 0x00FF00FFu);
          tot += (pair & 0xFFFFu) + (pair >> 16);
        }
      }
    }
  }
  const uint32_t row_total = __reduce_add_sync(0xffffffffu, tot);
  if (lane == 0) a.sums_out[row_in_shard] = carry + row_total;
  const uint32_t gi = a.row_base + row_in_shard;
  const unsigned long long fin = carry + row_total;
  const unsigned long long lb_min = ctl.lb_seg[0];
  if (a.dense) {
    // Dense pass (no bounds yet, or the bounds let too many rows through): the row's sum after every read of the pass,
    // relative to `carry`, goes out as one prefix-sum vector (u16 per read with u8 counters, u32 with u16 counters);
    // dense_topk_kernel ranks every read over all rows from these.
    if (row_total) {  // (a row without a hit keeps sums_out == sums_in: dense_topk_kernel never reads its vector)
      uint32_t incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += y;
      }
      uint32_t run = incl - tot;
      if (CPW == 4) {
        if (row_total > 0xFFFFu) *a.dense_overflow = 1u;  // does not fit 16 bits: the host redoes the pass with <= 256 reads
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.dense) + (size_t)row_in_shard * a.cnt_stride + seg0);
        for (uint32_t i = 0; i < segw; i += 2) {
          uint32_t o4[4];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t x = cseg[i + q];
            const uint32_t p0 = run + (x & 0xFFu), p1 = p0 + ((x >> 8) & 0xFFu), p2 = p1 + ((x >> 16) & 0xFFu);
            run = p2 + (x >> 24);
            o4[2 * q] = (p0 & 0xFFFFu) | (p1 << 16);
            o4[2 * q + 1] = (p2 & 0xFFFFu) | (run << 16);
          }
          dst[i >> 1] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      } else {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(a.dense) + (size_t)row_in_shard * a.cnt_stride + seg0);
        for (uint32_t i = 0; i < segw; i += 2) {
          const uint32_t x0 = cseg[i], x1 = cseg[i + 1];
          const uint32_t p0 = run + (x0 & 0xFFFFu), p1 = p0 + (x0 >> 16), p2 = p1 + (x1 & 0xFFFFu);
          run = p2 + (x1 >> 16);
          dst[i >> 1] = make_uint4(p0, p1, p2, run);
        }
      }
    }
  } else if (fin > lb_min || (fin == lb_min && gi <= __reduce_max_sync(0xffffffffu, ctl.li_seg[lane]))) {
    uint32_t incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += y;
    }
    const unsigned long long seg_end_sum = carry + incl, lbs = ctl.lb_seg[lane];
    const bool pass = seg0 < a.n_reads && (seg_end_sum > lbs || (seg_end_sum == lbs && gi <= ctl.li_seg[lane]));
    const uint32_t seg_hi = min(seg0 + per, a.n_reads);
    if (pass && tot == 0u) {  // one sum for the whole segment
      const uint32_t slot = atomicAdd(a.ivl_total, 1u);
      if (slot < a.ivl_cap) {
        SkbInterval iv;
        iv.sum = seg_end_sum; iv.idx = gi; iv.span = seg0 | (seg_hi << 16);
        a.ivl[slot] = iv;
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, pass && tot != 0u);
    if (bal) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(a.seg_total, (uint32_t)__popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      const uint32_t slot = base + __popc(bal & ((1u << lane) - 1u));
      if (((bal >> lane) & 1u) && slot < a.seg_cap) {
        a.seg_hdr[slot] = make_uint4((uint32_t)(seg_end_sum - tot), (uint32_t)((seg_end_sum - tot) >> 32), gi, seg0);
        uint4* dst = reinterpret_cast<uint4*>(a.seg_words + (size_t)slot * segw);
        for (uint32_t i = 0; i < segw; i += 4) dst[i >> 2] = *reinterpret_cast<const uint4*>(cseg + i);
      }
    }
  }
  if (row_total) {  // clear the buffer for the next row that uses it (lane-interleaved 16-byte stores)
    __syncwarp();
    uint4* z = reinterpret_cast<uint4*>(cpar);
    for (uint32_t i = lane; i < (cwords >> 2); i += 32) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
}

// CPW = counters per 32-bit word of a row buffer: 2 (u16, any pass) or 4 (u8, when no read of the pass keeps more than
// 255 query hashes; more reads fit a pass).
//
// Persistent, one CTA per SM, each owning a contiguous range of reference rows. A row is cut into FS_SUB-hash
// sub-tiles; the CTA's warps claim sub-tiles in order from a shared counter, one claim ahead of every free slot of
// their PRIVATE ring of cp.async.bulk (TMA) staging buffers, so the claim's round trip hides behind the copy. A warp
// probes each staged hash against the filter in shared memory (one LDS + a handful of integer instructions) and
// resolves the few passers on the spot: the lanes that hold one look their hash up in the L2-resident table (a
// 16-byte load; two passers per trip, both loads in flight together) and add the key's reads to the row's counters
// in shared memory. The warp whose sub-tile completes a row does the row's short rank step (rank_row) and hands the
// counter buffer back; a.rowbuf rows are in flight, and since sub-tiles are claimed in order the warps cannot run
// far apart, so nobody waits for a buffer. Nothing in the loop is warp-collective except that hand-over.
// HBM traffic per pass = the reference matrix, once.
template <int CPW>
__global__ void __launch_bounds__(FS_THREADS, 1) fused_kernel(const SkbFusedArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[FS_WARPS][FS_STAGES];
  __shared__ __align__(16) FsCtl ctl;

  if (*a.abort) return;  // an earlier pass of this batch has to be redone: leave sums and candidates alone

  uint32_t* bloom = reinterpret_cast<uint32_t*>(smem_raw);
  uint64_t* ring = reinterpret_cast<uint64_t*>(smem_raw + FS_SMEM_BLOOM);
  uint32_t* cnt32 = reinterpret_cast<uint32_t*>(smem_raw + FS_SMEM_BLOOM + FS_SMEM_RING);
  const uint32_t cwords = a.cnt_stride / CPW;  // 32-bit words per row buffer

  // every `row` below is a CTA-local number; c0 + row is the row of the shard
  const uint32_t c0 = a.cta_row[blockIdx.x];
  const uint32_t r1 = a.cta_row[blockIdx.x + 1] - c0;
  const uint32_t tile0 = a.rv.uniform_len ? 0u : a.tile_cum[c0];
  const uint32_t n_tiles = a.rv.uniform_len ? r1 * fs_claims_of(a.rv.uniform_len) : a.tile_cum[c0 + r1] - tile0;

  if (threadIdx.x == 0) {
    for (int w = 0; w < FS_WARPS; ++w)
      for (int s = 0; s < FS_STAGES; ++s) mbar_init(&full_bar[w][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ctl.next_tile = 0;
  }
  if (threadIdx.x < FS_ROWBUF_MAX) { ctl.done[threadIdx.x] = 0; ctl.freed[threadIdx.x] = 0; }
  if (threadIdx.x < 32) {
    const uint32_t seg0 = threadIdx.x * (a.cnt_stride >> 5);
    ctl.li_seg[threadIdx.x] = 0;
    ctl.lb_seg[threadIdx.x] = seg0 < a.n_reads ? a.lb_sum[seg0] : ~0ull;
  }
  if (!a.skip_stream) {  // stage the filter
    const uint4* src = reinterpret_cast<const uint4*>(a.table.bloom);
    uint4* dst = reinterpret_cast<uint4*>(bloom);
    for (uint32_t i = threadIdx.x; i < SKB_BLOOM_WORDS / 4; i += blockDim.x) dst[i] = src[i];
  }
  for (uint32_t i = threadIdx.x; i < a.rowbuf * cwords; i += blockDim.x) cnt32[i] = 0;
  __syncthreads();
  {
    const uint32_t per = a.cnt_stride >> 5;
    for (uint32_t b = threadIdx.x; b < a.n_reads; b += blockDim.x) atomicMax(&ctl.li_seg[b / per], a.lb_idx[b]);
  }
  __syncthreads();

  const uint32_t warp = threadIdx.x >> 5, lane = skb_lane();

  if (a.skip_stream) {  // the pass has no query hashes: rows are ranked from their running sums alone
    for (uint32_t row = warp; row < r1; row += FS_WARPS) rank_row<CPW>(a, ctl, nullptr, cwords, c0 + row);
    return;
  }

  const SkbTable& t = a.table;
  const uint4* tslots = reinterpret_cast<const uint4*>(t.slots);
  const uint32_t tcap = t.cap, tlog2 = t.log2cap;
  uint8_t* my_ring = reinterpret_cast<uint8_t*>(ring + (size_t)warp * FS_STAGES * FS_SUB);
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

  // A claim = FS_STAGES consecutive sub-tiles of one row (its "claim unit", one ring's worth); the warp's ring slot j
  // holds sub-tile j of the current claim. All lanes keep the claim in registers: {row, first sub-tile, row length}.
  auto claim = [&](uint32_t& row, uint32_t& t0, uint32_t& len, const uint64_t*& p) {
    uint32_t n = 0;
    if (lane == 0) n = atomicAdd(&ctl.next_tile, 1u);
    n = __shfl_sync(0xffffffffu, n, 0);
    row = FS_NONE; t0 = 0; len = 0; p = a.rv.ref;
    if (n >= n_tiles) return;
    if (a.rv.uniform_len) {
      row = a.tpr == 1u ? n : __umulhi(n, a.tpr_magic);  // n / claims-per-row by a precomputed reciprocal (exact while n * tpr < 2^32)
      t0 = (n - row * a.tpr) * FS_STAGES;
      len = a.rv.uniform_len;
      p = a.rv.ref + (size_t)(c0 + row) * a.rv.uniform_pitch;
    } else {  // ragged rows: the row whose cumulative claim range holds n
      uint32_t lo = 0, hi = r1;
      const uint32_t key = tile0 + n;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.tile_cum[c0 + mid] <= key) lo = mid; else hi = mid;
      }
      row = lo;
      t0 = (key - a.tile_cum[c0 + row]) * FS_STAGES;
      len = a.rv.row_len[c0 + row];
      p = a.rv.ref + a.rv.row_start[c0 + row];
    }
  };
  // lane 0: bulk copy of sub-tile (t0 + j) of a claimed row into ring slot j (nothing to copy past the row's end)
  auto issue = [&](uint32_t j, uint32_t t0, uint32_t len, const uint64_t* p) {
    const uint32_t first = (t0 + j) * FS_SUB;
    uint32_t cnt = first < len ? len - first : 0u;
    if (cnt > (uint32_t)FS_SUB) cnt = FS_SUB;
    const uint32_t bytes = ((cnt + 1u) & ~1u) * 8u;  // multiple of 16; rows start on even offsets
    mbar_arrive_expect_tx(&full_bar[warp][j], bytes);
    if (bytes) bulk_load(my_ring + (size_t)j * FS_SUB * 8, p + first, bytes, &full_bar[warp][j], policy);
  };

  uint32_t row, t0, len;
  const uint64_t* rp;
  claim(row, t0, len, rp);
  if (row != FS_NONE && lane == 0)
    for (uint32_t j = 0; j < (uint32_t)FS_STAGES; ++j) issue(j, t0, len, rp);  // prime the private ring

  for (uint32_t k = 0; row != FS_NONE; ++k) {  // claims consumed by this warp: every ring slot's barrier flips once per claim
    const uint32_t phase = k & 1u;
    // the next claim is made now, so that its round trip is over when the first ring slot is free again
    uint32_t nrow, nt0, nlen;
    const uint64_t* np;
    claim(nrow, nt0, nlen, np);
    const uint32_t par = row & (a.rowbuf - 1u);
    uint32_t* cb = cnt32 + par * cwords;
    if (row >= a.rowbuf) {  // the buffer's previous row must have been ranked and cleared (it almost always has)
      const uint32_t need = row >> a.rowbuf_log2;
      if (lds_acquire_u32(&ctl.freed[par]) < need) {
        if (lane == 0)
          while (lds_acquire_u32(&ctl.freed[par]) < need) __nanosleep(100);
        __syncwarp();
      }
    }
#pragma unroll 1
    for (uint32_t j = 0; j < (uint32_t)FS_STAGES; ++j) {
      mbar_wait(&full_bar[warp][j], phase);
      const uint8_t* tile = my_ring + (size_t)j * FS_SUB * 8;
      const uint32_t first = (t0 + j) * FS_SUB;
      const uint32_t n_sub = first < len ? len - first : 0u;  // >= FS_SUB for every sub-tile but a row's last
#pragma unroll 1
      for (uint32_t ch = 0; ch < (uint32_t)FS_CHUNKS; ++ch) {
        const uint32_t base = ch * FS_NHASH * 32;  // first hash index of the chunk within the sub-tile
        if (base >= n_sub) break;
        const uint8_t* cbase = tile + (size_t)base * 8 + lane * 16;
        uint4 v[FS_NHASH / 2];
#pragma unroll
        for (int r = 0; r < FS_NHASH / 2; ++r) v[r] = *reinterpret_cast<const uint4*>(cbase + 512 * r);
        uint32_t pm = 0;  // bit 7 - j: this lane's j-th hash of the chunk passed the filter
        if (!(SKB_X_ABLATE & 1)) {
          if (n_sub >= base + FS_NHASH * 32) {
#pragma unroll
            for (int q = 0; q < FS_NHASH; ++q) {
              const uint32_t lo = (q & 1) ? v[q >> 1].z : v[q >> 1].x;
              const uint32_t hi = (q & 1) ? v[q >> 1].w : v[q >> 1].y;
              pm = __funnelshift_l(bloom_probe(bloom, lo, hi), pm, 1);
            }
          } else {
#pragma unroll
            for (int q = 0; q < FS_NHASH; ++q) {
              const uint32_t lo = (q & 1) ? v[q >> 1].z : v[q >> 1].x;
              const uint32_t idx = base + 2u * (lane + 32 * (q >> 1)) + (q & 1);
              const uint32_t hi = (q & 1) ? v[q >> 1].w : v[q >> 1].y;
              pm = __funnelshift_l(idx < n_sub ? bloom_probe(bloom, lo, hi) : 0u, pm, 1);
            }
          }
        }
        if (SKB_X_ABLATE & 2) pm = 0;
        // Divergent: only the lanes that hold a passer run this, two passers per trip. A passer's hash is read back
        // from the staging buffer by its position (hash q of the chunk sits at byte (q >> 1) * 512 + (q & 1) * 8 of
        // the lane's column), both table loads are issued before either is resolved, so a lane pays the L2 latency
        // once per trip; a slot owned by another key is chased in place (load factor 0.125: rare).
        while (pm) {
          const uint32_t b0 = 31u - (uint32_t)__clz((int)pm);
          pm &= ~(1u << b0);
          const bool two = pm != 0u;
          const uint32_t b1 = two ? 31u - (uint32_t)__clz((int)pm) : b0;
          pm &= ~(1u << b1);
          // bit b holds hash q = 7 - b: offset 0x608 - ((b * 0x108) & 0x608)
          const uint2 e0 = *reinterpret_cast<const uint2*>(cbase + 0x608u - ((b0 * 0x108u) & 0x608u));
          const uint2 e1 = *reinterpret_cast<const uint2*>(cbase + 0x608u - ((b1 * 0x108u) & 0x608u));
          const uint64_t h0 = ((uint64_t)e0.y << 32) | e0.x, h1 = ((uint64_t)e1.y << 32) | e1.x;
          uint32_t s0 = h0 == SKB_EMPTY_KEY ? tcap : table_home(h0, tlog2);
          uint32_t s1 = h1 == SKB_EMPTY_KEY ? tcap : table_home(h1, tlog2);
          uint4 r0 = __ldg(tslots + s0);
          uint4 r1v = __ldg(tslots + s1);
          for (;;) {  // walk the probe sequence until the key or an empty slot
            const uint64_t key = ((uint64_t)r0.y << 32) | r0.x;
            if (key == h0) { apply_hit<CPW>(t, ((unsigned long long)r0.w << 32) | r0.z, cb); break; }
            if (key == SKB_EMPTY_KEY) break;
            s0 = (s0 + 1) & (tcap - 1);
            r0 = __ldg(tslots + s0);
          }
          if (two) {
            for (;;) {
              const uint64_t key = ((uint64_t)r1v.y << 32) | r1v.x;
              if (key == h1) { apply_hit<CPW>(t, ((unsigned long long)r1v.w << 32) | r1v.z, cb); break; }
              if (key == SKB_EMPTY_KEY) break;
              s1 = (s1 + 1) & (tcap - 1);
              r1v = __ldg(tslots + s1);
            }
          }
        }
      }
      // ring slot j is fully read: refill it with sub-tile j of the next claim
      __syncwarp();
      if (lane == 0 && nrow != FS_NONE) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(j, nt0, nlen, np);
      }
    }
    // this claim's hits are in shared memory: count it done; the warp whose claim completes the row ranks it
    uint32_t last = 0;
    if (lane == 0) {
      __threadfence_block();
      last = atomicAdd(&ctl.done[par], 1u) + 1u == fs_claims_of(len) ? 1u : 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {  // hand the buffer back after the row's short rank step
      __threadfence_block();
      rank_row<CPW>(a, ctl, cb, cwords, c0 + row);
      if (lane == 0) {
        ctl.done[par] = 0;
        __threadfence_block();
        atomicAdd(&ctl.freed[par], 1u);
      }
    }
    row = nrow; t0 = nt0; len = nlen; rp = np;
  }
}

// Post-pass: segment records -> per-read candidate buckets. One thread per 16-byte unit of a record's counters: it
// sums the units before it (a record is at most 160 bytes), then tests each of its reads against the read's exact
// bound; a row that is at least as good as the bound is a candidate for that read.
__device__ __forceinline__ void walk_records(const SkbRankArgs& a) {
  const uint32_t total = min(*a.seg_total, a.seg_cap);
  const uint32_t upr = a.seg_words_per / 4;  // 16-byte units per record
  const uint32_t cpw = a.seg_cpw;            // counters per word: 2 (u16) or 4 (u8)
  const uint64_t n_units = (uint64_t)total * upr;
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t rec = (uint32_t)(u / upr), unit = (uint32_t)(u - (uint64_t)rec * upr);
    const uint4 hd = a.seg_hdr[rec];
    const uint4* w = reinterpret_cast<const uint4*>(a.seg_words + (size_t)rec * a.seg_words_per);
    unsigned long long run = ((unsigned long long)hd.y << 32) | hd.x;
    for (uint32_t q = 0; q < unit; ++q) {
      const uint4 x = w[q];
      const uint32_t w4[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (cpw == 2) run += (w4[i] & 0xFFFFu) + (w4[i] >> 16);
        else { const uint32_t pair = (w4[i] & 0x00FF00FFu) + ((w4[i] >> 8) & 0x00FF00FFu); run += (pair & 0xFFFFu) + (pair >> 16); }
      }
    }
    const uint4 x = w[unit];
    const uint32_t w4[4] = {x.x, x.y, x.z, x.w};
    const uint32_t gi = hd.z;
    uint32_t b = hd.w + unit * 4 * cpw;
    for (int i = 0; i < 4; ++i) {
      for (uint32_t h = 0; h < cpw; ++h, ++b) {
        run += cpw == 2 ? ((w4[i] >> (16 * h)) & 0xFFFFu) : ((w4[i] >> (8 * h)) & 0xFFu);
        if (b < a.n_reads && !skb_key_better(a.lb_sum[b], a.lb_idx[b], run, gi)) {
          const uint32_t slot = atomicAdd(&a.cand_cnt[b], 1u);
          if (slot < a.cand_cap) {
            SkbCand cd;
            cd.sum = run; cd.idx = gi; cd.pad = 0;
            a.cand[(size_t)b * a.cand_cap + slot] = cd;
          } else {
            *a.cand_total = 1u;  // bucket overflow: the host redoes the pass densely
          }
        }
      }
    }
  }
}

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

// Bounds, step 1: the per-read counts of one tracked row against the pass's table, accumulated in shared memory, then
// their inclusive prefix sums over the reads of the pass (what the row adds to its running sum after each read).
// One CTA per tracked row (they define the bounds; at most SKB_MAX_TRACKED of them).
__global__ void __launch_bounds__(1024) tracked_prefix_kernel(const SkbRefView rv, const SkbTable t, const SkbRankArgs a, int has_keys) {
  extern __shared__ __align__(16) uint32_t tk_cnt[];  // [row_stride / 2] u16 counters, two per word
  __shared__ uint32_t warp_sums[32];
  const uint32_t tr = blockIdx.x;
  if (tr >= *a.n_tracked) return;
  const uint32_t row = a.tracked[tr];
  for (uint32_t i = threadIdx.x; i < a.row_stride / 2; i += blockDim.x) tk_cnt[i] = 0;
  __syncthreads();
  if (has_keys) {
    const uint64_t* src = rv.ref + rv.row_start[row];
    const uint32_t len = rv.row_len[row];
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
      SkbSlot s;
      if (table_lookup(t, src[i], s)) apply_hit<2>(t, s.meta, tk_cnt);
    }
  }
  __syncthreads();
  const uint16_t* c = reinterpret_cast<const uint16_t*>(tk_cnt);
  uint32_t* p = a.tracked_prefix + (size_t)tr * a.row_stride;
  const uint32_t per = (a.n_reads + blockDim.x - 1) / blockDim.x;
  const uint32_t b0 = threadIdx.x * per;
  const uint32_t b1 = min(b0 + per, a.n_reads);
  uint32_t local = 0;
  for (uint32_t b = b0; b < b1; ++b) local += c[b];
  uint32_t tot;
  const uint32_t incl = cta_scan_u32(local, warp_sums, &tot);
  uint32_t run = incl - local;
  for (uint32_t b = b0; b < b1; ++b) {
    run += c[b];
    p[b] = run;
  }
}

// total hits of every tracked row against a pass's table (summed over the reads of that pass); 8 CTAs share a row
__global__ void __launch_bounds__(256) tracked_totals_kernel(const SkbRefView rv, const uint32_t* __restrict__ tracked,
                                                             const uint32_t* __restrict__ n_tracked, const SkbTable t,
                                                             unsigned long long* extra) {
  const uint32_t tr = blockIdx.x / 8, part = blockIdx.x % 8;
  if (tr >= *n_tracked) return;
  const uint32_t row = tracked[tr];
  const uint64_t* src = rv.ref + rv.row_start[row];
  const uint32_t len = rv.row_len[row];
  uint32_t n = 0;
  for (uint32_t i = part * blockDim.x + threadIdx.x; i < len; i += 8 * blockDim.x) {
    SkbSlot s;
    if (table_lookup(t, src[i], s)) n += SKB_SLOT_CNT(s.meta);
  }
  n = __reduce_add_sync(0xffffffffu, n);
  if (skb_lane() == 0 && n) atomicAdd(&extra[tr], (unsigned long long)n);
}

// ---------------------------------------------------------------------------------------------------------
// rank kernels
// ---------------------------------------------------------------------------------------------------------
// Bounds, step 2: for every read the `top`-th best key among the tracked rows (exact sums, so a valid lower bound of
// the read's true `top`-th key: sums never decrease and the tracked rows are distinct). One warp per read: every lane
// holds up to six tracked rows' keys at this read, the warp takes out its best key `top` times. With fewer than `top`
// tracked rows the bound is their worst.
__global__ void __launch_bounds__(256) rank_bounds_kernel(const SkbRankArgs a) {
  __shared__ unsigned long long base_s[SKB_MAX_TRACKED];
  __shared__ uint32_t base_i[SKB_MAX_TRACKED];
  const uint32_t nt = *a.n_tracked;
  for (uint32_t t = threadIdx.x; t < nt; t += blockDim.x) {
    const uint32_t row = a.tracked[t];
    base_s[t] = a.sums_in[row] + a.tracked_extra[t];
    base_i[t] = a.row_base + row;
  }
  __syncthreads();
  const uint32_t lane = skb_lane();
  const uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= a.n_reads) return;
  constexpr int PER = (SKB_MAX_TRACKED + 31) / 32;
  unsigned long long ks[PER];
  uint32_t ki[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const uint32_t t = lane + 32 * j;
    const bool have = t < nt;
    ks[j] = have ? base_s[t] + a.tracked_prefix[(size_t)t * a.row_stride + b] : 0ull;
    ki[j] = have ? base_i[t] : 0xFFFFFFFFu;   // (0, UINT32_MAX): worse than any real key
  }
  const uint32_t keep = a.top < nt ? a.top : nt;
  unsigned long long bs = 0;
  uint32_t bi = 0xFFFFFFFFu;
  for (uint32_t r = 0; r < keep; ++r) {
    // this lane's best remaining key, then the warp's
    unsigned long long ls = ks[0];
    uint32_t li = ki[0];
    int lj = 0;
#pragma unroll
    for (int j = 1; j < PER; ++j)
      if (skb_key_better(ks[j], ki[j], ls, li)) { ls = ks[j]; li = ki[j]; lj = j; }
    bs = ls; bi = li;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long s2 = __shfl_xor_sync(0xffffffffu, bs, o);
      const uint32_t i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      if (skb_key_better(s2, i2, bs, bi)) { bs = s2; bi = i2; }
    }
    if (li == bi && ls == bs) {  // the owner retires it (keys are distinct: the row index is part of the key)
#pragma unroll
      for (int j = 0; j < PER; ++j)
        if (j == lj) { ks[j] = 0; ki[j] = 0xFFFFFFFFu; }
    }
  }
  if (lane == 0) {
    a.lb_sum[b] = keep ? bs : 0ull;
    a.lb_idx[b] = keep ? bi : 0xFFFFFFFFu;
  }
}

// intervals (one sum over a run of reads) -> per-read candidate buckets. One warp per interval; every read is tested
// against its exact bound; the per-read counter is the slot allocator.
__device__ __forceinline__ void expand_intervals(const SkbRankArgs& a) {
  const uint32_t total = min(*a.ivl_total, a.ivl_cap);
  const uint32_t lane = skb_lane();
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += nwarps) {
    const SkbInterval iv = a.ivl[w];
    const uint32_t b0 = iv.span & 0xFFFFu, b1 = iv.span >> 16;
    for (uint32_t b = b0 + lane; b < b1; b += 32) {
      if (skb_key_better(a.lb_sum[b], a.lb_idx[b], iv.sum, iv.idx)) continue;  // under this read's bound
      const uint32_t slot = atomicAdd(&a.cand_cnt[b], 1u);
      if (slot < a.cand_cap) {
        SkbCand cd;
        cd.sum = iv.sum; cd.idx = iv.idx; cd.pad = 0;
        a.cand[(size_t)b * a.cand_cap + slot] = cd;
      } else {
        *a.cand_total = 1u;  // bucket overflow: the host redoes the pass densely
      }
    }
  }
}

// post-pass, first launch: segment records and intervals -> per-read candidate buckets
__global__ void __launch_bounds__(256) candidates_kernel(const SkbRankArgs a) {
  walk_records(a);
  expand_intervals(a);
}

// Post-pass, last launch (one CTA): the overflow verdict of the pass, then the tracked rows of the next passes.
//   verdict: the fullest candidate bucket and the record counts go to `abort` for the host; a pass whose buckets or
//            record lists overflowed marks itself there (the passes behind it then do nothing); the counters and the
//            bucket fill counts are cleared for the slot's next pass.
//   update:  the union of the top lists of 16 evenly spaced reads of this pass, the last read's list first (it alone
//            guarantees `top` distinct rows). Rows that led at any point of the pass stay tracked, so a lineage that
//            overtakes and falls back does not loosen the bounds. Skipped while a pass is being redone.
__global__ void __launch_bounds__(1024) verdict_update_kernel(const SkbRankArgs a, int with_verdict) {
  __shared__ uint32_t wmax[32];
  __shared__ uint32_t list[16 * SKB_MAX_TOP];
  __shared__ uint32_t keep[16 * SKB_MAX_TOP];
  __shared__ uint32_t stop;
  if (with_verdict) {
    uint32_t m = 0;
    for (uint32_t b = threadIdx.x; b < a.n_reads; b += blockDim.x) { m = max(m, a.cand_cnt[b]); a.cand_cnt[b] = 0; }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31u) == 0) wmax[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (uint32_t w = 1; w < (blockDim.x >> 5); ++w) m = max(m, wmax[w]);
      if (a.abort[0] == 0u) {
        a.abort[2] = m;
        a.abort[3] = *a.ivl_total;
        a.abort[4] = *a.seg_total;
        if (*a.cand_total != 0u || *a.ivl_total > a.ivl_cap || *a.seg_total > a.seg_cap) {
          a.abort[1] = a.seq;
          a.abort[0] = 1u;
        }
      }
      *a.cand_total = 0u;
      *const_cast<uint32_t*>(a.ivl_total) = 0u;
      *const_cast<uint32_t*>(a.seg_total) = 0u;
    }
  }
  if (threadIdx.x == 0) stop = a.abort[0];
  __syncthreads();
  if (stop) return;  // this pass (or one before it) is redone: keep the tracked rows it started from
  const uint32_t n_s = 16, total = n_s * a.top;
  for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
    const uint32_t j = e / a.top, tpos = e % a.top;
    // sample j = 0 is the last read; the others are spread over the pass
    const uint32_t b = j == 0 ? a.n_reads - 1 : (uint32_t)(((unsigned long long)j * a.n_reads) / n_s);
    list[e] = a.out_idx[(size_t)b * a.top + tpos];
  }
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
    const uint32_t v = list[e];
    bool first = v != 0xFFFFFFFFu;
    for (uint32_t f = 0; f < e && first; ++f) first = list[f] != v;
    keep[e] = first ? 1u : 0u;
  }
  __syncthreads();
  if (threadIdx.x < 32) {  // order-preserving compaction by one warp
    uint32_t n = 0;
    for (uint32_t e0 = 0; e0 < total && n < SKB_MAX_TRACKED; e0 += 32) {
      const uint32_t e = e0 + threadIdx.x;
      const bool k = e < total && keep[e];
      const uint32_t bal = __ballot_sync(0xffffffffu, k);
      const uint32_t at = n + __popc(bal & ((1u << threadIdx.x) - 1u));
      if (k && at < SKB_MAX_TRACKED) a.tracked_next[at] = list[e] - a.row_base;
      n += __popc(bal);
    }
    if (threadIdx.x == 0) *a.n_tracked_next = n < SKB_MAX_TRACKED ? n : SKB_MAX_TRACKED;
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

// One warp per read: the `top` best candidates of the read's bucket, in order. Candidate rows are distinct, so
// "best key strictly worse than the previous pick" enumerates them without marking. The first RS_CACHE records of
// the bucket are staged in shared memory once; missing entries are (sum 0, idx UINT32_MAX).
constexpr int RS_WARPS = 8;
constexpr int RS_CACHE = 512;
__global__ void __launch_bounds__(RS_WARPS * 32) rank_select_kernel(const SkbRankArgs a) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  const uint32_t wid = threadIdx.x >> 5, lane = skb_lane();
  const uint32_t b = blockIdx.x * RS_WARPS + wid;
  if (b >= a.n_reads) return;
  uint4* cache = reinterpret_cast<uint4*>(rs_smem) + (size_t)wid * RS_CACHE;
  const uint32_t produced = a.cand_cnt[b];
  const uint32_t n = produced < a.cand_cap ? produced : a.cand_cap;
  const uint4* list = reinterpret_cast<const uint4*>(a.cand + (size_t)b * a.cand_cap);
  const uint32_t nc = n < (uint32_t)RS_CACHE ? n : (uint32_t)RS_CACHE;
  for (uint32_t i = lane; i < nc; i += 32) cache[i] = list[i];
  if (lane == 0 && a.cand_stat) atomicAdd(a.cand_stat, (unsigned long long)produced);
  __syncwarp();
  // A bucket larger than the cache (loose bounds: a few reads per pass) is first cut down to the records that can still
  // be in the top: the top-th best of the 32 per-lane maxima is a lower bound of the top-th best record, and the
  // records at least that good are compacted into the cache. Two coalesced passes instead of `top` of them.
  uint32_t m = n;       // records the selection rounds look at
  bool cached = n <= (uint32_t)RS_CACHE;
  if (!cached && a.top <= 32u) {  // (the 32 lane maxima bound the top-th best only for top <= 32)
    unsigned long long ls = 0;
    uint32_t li = 0xFFFFFFFFu;
    for (uint32_t i = lane; i < n; i += 32) {
      const uint4 c = list[i];
      const unsigned long long cs = ((unsigned long long)c.y << 32) | c.x;
      if (skb_key_better(cs, c.z, ls, li)) { ls = cs; li = c.z; }
    }
    unsigned long long ts = 0;  // threshold key = the top-th best lane maximum (n > RS_CACHE >= 32: every lane has one)
    uint32_t ti = 0xFFFFFFFFu;
    for (uint32_t t = 0; t < a.top; ++t) {
      unsigned long long bs = ls;
      uint32_t bi = li;
      warp_best(bs, bi);
      ts = bs; ti = bi;
      if (ls == bs && li == bi) { ls = 0; li = 0xFFFFFFFFu; }  // the winner's lane steps aside
    }
    uint32_t kept = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
      const uint32_t i = i0 + lane;
      uint4 c = make_uint4(0, 0, 0, 0);
      bool keep = false;
      if (i < n) {
        c = list[i];
        const unsigned long long cs = ((unsigned long long)c.y << 32) | c.x;
        keep = !skb_key_better(ts, ti, cs, c.z);  // at least as good as the threshold
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, keep);
      const uint32_t at = kept + __popc(bal & ((1u << lane) - 1u));
      if (keep && at < (uint32_t)RS_CACHE) cache[at] = c;
      kept += __popc(bal);
    }
    __syncwarp();
    if (kept <= (uint32_t)RS_CACHE) { cached = true; m = kept; }  // else (never seen): rounds over the whole bucket
  }
  unsigned long long last_s = 0;
  uint32_t last_i = 0;
  for (uint32_t t = 0; t < a.top; ++t) {
    unsigned long long bs = 0;
    uint32_t bi = 0xFFFFFFFFu;
    for (uint32_t i = lane; i < m; i += 32) {
      const uint4 c = cached ? cache[i] : list[i];
      const unsigned long long cs = ((unsigned long long)c.y << 32) | c.x;
      if ((t == 0 || skb_key_better(last_s, last_i, cs, c.z)) && skb_key_better(cs, c.z, bs, bi)) {
        bs = cs; bi = c.z;
      }
    }
    warp_best(bs, bi);
    if (lane == 0) {
      a.out_idx[(size_t)b * a.top + t] = bi;
      a.out_sum[(size_t)b * a.top + t] = bs;
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

// Dense ranking: per read, the `top` best rows by (sum desc, index asc) among one group of rows, from the per-row
// prefix-sum vectors the streaming kernel wrote (dense[row][read], relative to sums_in[row]). One thread per
// (read, row group): consecutive threads hold consecutive reads, so every row step is one coalesced load. Rows are
// visited in increasing index, so a later row must be strictly better to displace an earlier one (the tie rule of
// the reference's stable sort). Output: parts[g][read][top] for merge_topn_kernel. A vector of rows without a hit in
// the pass is not written: such a row has sums_out == sums_in and counts as all-zero.
template <int KREG, typename P>
__global__ void __launch_bounds__(128) dense_topk_kernel(const SkbDenseArgs a) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t g = blockIdx.y;
  const uint32_t r_lo = (uint32_t)(((uint64_t)a.n_rows * g) / gridDim.y), r_hi = (uint32_t)(((uint64_t)a.n_rows * (g + 1)) / gridDim.y);
  if (b >= a.n_reads) return;
  const P* col = reinterpret_cast<const P*>(a.dense) + b;
  const uint32_t top = a.top;
  unsigned long long ks[KREG];
  uint32_t ki[KREG];
#pragma unroll
  for (int t = 0; t < KREG; ++t) { ks[t] = 0; ki[t] = 0xFFFFFFFFu; }
  for (uint32_t row = r_lo; row < r_hi; ++row) {
    const unsigned long long s_in = a.sums_in[row];
    const unsigned long long v = s_in + (a.sums_out[row] != s_in ? (unsigned long long)col[(size_t)row * a.cnt_stride] : 0ull);
    const uint32_t gi = a.row_base + row;
    if (KREG <= 16) {  // the list lives in registers: compare-and-shift through the whole list
      if (skb_key_better(v, gi, ks[KREG - 1], ki[KREG - 1])) {
        unsigned long long cs = v;
        uint32_t ci = gi;
#pragma unroll
        for (int t = 0; t < KREG; ++t) {
          const bool better = skb_key_better(cs, ci, ks[t], ki[t]);
          const unsigned long long ts = ks[t];
          const uint32_t ti = ki[t];
          if (better) { ks[t] = cs; ki[t] = ci; cs = ts; ci = ti; }
        }
      }
    } else {  // long lists (local memory): insertion from the tail
      if (skb_key_better(v, gi, ks[top - 1], ki[top - 1])) {
        uint32_t pos = top - 1;
        while (pos > 0 && skb_key_better(v, gi, ks[pos - 1], ki[pos - 1])) { ks[pos] = ks[pos - 1]; ki[pos] = ki[pos - 1]; --pos; }
        ks[pos] = v; ki[pos] = gi;
      }
    }
  }
  const size_t o = ((size_t)g * a.n_reads + b) * top;
  if (KREG <= 16) {
#pragma unroll
    for (int t = 0; t < KREG; ++t)
      if ((uint32_t)t < top) { a.part_sum[o + t] = ks[t]; a.part_idx[o + t] = ki[t]; }
  } else {
    for (uint32_t t = 0; t < top; ++t) { a.part_sum[o + t] = ks[t]; a.part_idx[o + t] = ki[t]; }
  }
}

// Dense ranking for top <= 32, the fast form: a CTA takes 64 consecutive columns and one group of rows; row tiles of
// 32 rows x 64 columns are staged in shared memory (coalesced row segments), then every warp ranks its 8 columns with
// one lane per row of the tile. The `top` best rows of a column so far live across the warp's lanes (lane t holds the
// t-th best); a row enters only when it beats the column's threshold (a ballot: almost always empty once the list has
// warmed up), by a shuffle-insert. Whole keys are compared, so the order of the rows does not matter.
// Two launches per pass (skb_launch_dense_topk):
//   anchors  column c = read 64 c (col_stride 64): the exact top lists of every 64th read, over many small row groups.
//   reads    column c = read c: the top-th key of the anchor read of a CTA's 64 reads is a lower bound of the top-th
//            key of each of them (sums never decrease along the reads), so the lists start with that threshold and
//            only the handful of rows that reach it are ever inserted. Without it every (read, row group) list warms
//            up on its own, about top * ln(rows / top) insertions each: 80 % of the kernel's instructions.
template <typename P>
__global__ void __launch_bounds__(256) dense_topk_warp_kernel(const SkbDenseArgs a) {
  constexpr int PADW = sizeof(P) == 2 ? 33 : 65;  // 32-bit words per tile row: 64 columns + one word of padding (no bank conflicts)
  __shared__ uint32_t tile[32 * PADW];
  const uint32_t b0 = blockIdx.x * 64, g = blockIdx.y;
  const uint32_t r_lo = (uint32_t)(((uint64_t)a.n_rows * g) / gridDim.y), r_hi = (uint32_t)(((uint64_t)a.n_rows * (g + 1)) / gridDim.y);
  const uint32_t warp = threadIdx.x >> 5, lane = skb_lane(), top = a.top;
  const uint32_t n_cols = a.col_stride == 1 ? a.n_reads : (a.n_reads + a.col_stride - 1) / a.col_stride;
  unsigned long long ks[8], thr_s[8];
  uint32_t ki[8], thr_i[8];
  unsigned long long t0s = 0;
  uint32_t t0i = 0xFFFFFFFFu;
  if (a.thr_sum) {  // "at least as good as the anchor's top-th key (s, i)" == "better than (s, i + 1)"
    const uint32_t i = a.thr_idx[(size_t)blockIdx.x * top + top - 1];
    if (i != 0xFFFFFFFFu) { t0s = a.thr_sum[(size_t)blockIdx.x * top + top - 1]; t0i = i + 1u; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { ks[j] = 0; ki[j] = 0xFFFFFFFFu; thr_s[j] = t0s; thr_i[j] = t0i; }
  const uint32_t lrow = threadIdx.x >> 3, lcol = (threadIdx.x & 7u) * 8;  // tile loader: 8 threads per row, 8 columns each
  for (uint32_t r0 = r_lo; r0 < r_hi; r0 += 32) {
    __syncthreads();  // the previous tile has been consumed
    {
      const uint32_t row = r0 + lrow;
      uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // 8 values: 4 words (u16) or 8 words (u32)
      if (row < r_hi && a.sums_out[row] != a.sums_in[row]) {  // (a row without a hit in the pass: all zero, never written)
        if (a.col_stride == 1) {
          const P* src = reinterpret_cast<const P*>(a.dense) + (size_t)row * a.cnt_stride + b0 + lcol;
          const uint4 x = *reinterpret_cast<const uint4*>(src);
          w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
          if (sizeof(P) == 4) {
            const uint4 y = *reinterpret_cast<const uint4*>(src + 4);
            w[4] = y.x; w[5] = y.y; w[6] = y.z; w[7] = y.w;
          }
        } else {  // gather: one value per column
          const P* src = reinterpret_cast<const P*>(a.dense) + (size_t)row * a.cnt_stride;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t c = b0 + lcol + q;
            const uint32_t v = c < n_cols ? (uint32_t)src[(size_t)c * a.col_stride] : 0u;
            if (sizeof(P) == 2) w[q >> 1] |= v << (16 * (q & 1));
            else w[q] = v;
          }
        }
      }
      uint32_t* dst = tile + lrow * PADW + lcol * sizeof(P) / 4;
#pragma unroll
      for (int q = 0; q < (int)(2 * sizeof(P)); ++q) dst[q] = w[q];
    }
    __syncthreads();
    const uint32_t row = r0 + lane;
    const bool valid = row < r_hi;
    const unsigned long long s_in = valid ? a.sums_in[row] : 0ull;
    const uint32_t gi = a.row_base + row;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t col = warp * 8 + j;
      if (b0 + col >= n_cols) break;
      uint32_t pv;
      if (sizeof(P) == 2) pv = (tile[lane * PADW + (col >> 1)] >> (16 * (col & 1u))) & 0xFFFFu;
      else pv = tile[lane * PADW + col];
      const unsigned long long v = s_in + pv;
      uint32_t bal = __ballot_sync(0xffffffffu, valid && skb_key_better(v, gi, thr_s[j], thr_i[j]));
      while (bal) {
        const int src = __ffs((int)bal) - 1;
        bal &= bal - 1;
        const unsigned long long cs = __shfl_sync(0xffffffffu, v, src);
        const uint32_t ci = __shfl_sync(0xffffffffu, gi, src);
        if (!skb_key_better(cs, ci, thr_s[j], thr_i[j])) continue;  // the threshold rose since the ballot
        const uint32_t pos = __popc(__ballot_sync(0xffffffffu, lane < top && skb_key_better(ks[j], ki[j], cs, ci)));
        const unsigned long long us = __shfl_up_sync(0xffffffffu, ks[j], 1);
        const uint32_t ui = __shfl_up_sync(0xffffffffu, ki[j], 1);
        if (lane > pos) { ks[j] = us; ki[j] = ui; }
        else if (lane == pos) { ks[j] = cs; ki[j] = ci; }
        // the list's top-th entry once the list is full; never below the threshold the list started with
        const unsigned long long ns = __shfl_sync(0xffffffffu, ks[j], top - 1);
        const uint32_t ni = __shfl_sync(0xffffffffu, ki[j], top - 1);
        if (skb_key_better(ns, ni, thr_s[j], thr_i[j])) { thr_s[j] = ns; thr_i[j] = ni; }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t b = b0 + warp * 8 + j;
    if (b < n_cols && lane < top) {
      const size_t o = ((size_t)g * n_cols + b) * top + lane;
      a.part_sum[o] = ks[j];
      a.part_idx[o] = ki[j];
    }
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

__global__ void __launch_bounds__(256) relayout_kernel(const uint64_t* __restrict__ src,
                                                       const uint64_t* __restrict__ src_off, uint64_t* dst,
                                                       const uint64_t* __restrict__ dst_start, uint32_t n_rows) {
  const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = warp0; r < n_rows; r += nwarps) {
    const uint64_t s0 = src_off[r], n = src_off[r + 1] - s0, d0 = dst_start[r];
    for (uint64_t i = skb_lane(); i < n; i += 32) dst[d0 + i] = src[s0 + i];
  }
}

// rows strictly increasing? + maximum hash. One warp per row.
__global__ void __launch_bounds__(256) ref_check_kernel(const SkbRefView rv, uint32_t* bad, unsigned long long* hmax) {
  const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = warp0; r < rv.n_rows; r += nwarps) {
    const uint64_t* row = rv.ref + rv.row_start[r];
    const uint32_t n = rv.row_len[r];
    bool ok = true;
    for (uint32_t i = skb_lane(); i + 1 < n; i += 32) ok = ok && (row[i] < row[i + 1]);
    if (!ok) atomicOr(bad, 1u);
    if (skb_lane() == 0 && n) atomicMax(hmax, (unsigned long long)row[n - 1]);
  }
}

// dense shared counts (reference `shared`, src/sketchy.rs:238-279): one warp per (reference row, query) pair;
// each lane binary-searches its share of the query in the row. Inputs strictly increasing => == merge count.
__global__ void __launch_bounds__(256) shared_kernel(const SkbRefView rv, const uint64_t* __restrict__ q,
                                                     const uint64_t* __restrict__ q_off, uint32_t Q,
                                                     unsigned long long* out) {
  const uint64_t pair = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pair >= (uint64_t)rv.n_rows * Q) return;
  const uint32_t i = (uint32_t)(pair / Q), j = (uint32_t)(pair % Q);
  const uint64_t* row = rv.ref + rv.row_start[i];
  const uint32_t rn = rv.row_len[i];
  const uint64_t q0 = q_off[j], q1 = q_off[j + 1];
  uint32_t c = 0;
  for (uint64_t x = q0 + skb_lane(); x < q1; x += 32) {
    const uint64_t h = q[x];
    uint32_t lo = 0, hi = rn;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (row[mid] < h) lo = mid + 1; else hi = mid;
    }
    c += (lo < rn && row[lo] == h) ? 1u : 0u;
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
                            uint32_t read_base, uint32_t n_prev, cudaStream_t st) {
  const int th = 256;
  uint32_t n_clear = n_prev == 0xFFFFFFFFu ? t.cap + 1 : n_prev;
  if (n_clear < SKB_BLOOM_WORDS) n_clear = SKB_BLOOM_WORDS;
  table_clear_kernel<<<(n_clear + th - 1) / th, th, 0, st>>>(t, n_prev);
  if (n_keys == 0) return;
  table_insert_kernel<<<(n_keys + th - 1) / th, th, 0, st>>>(t, qh, n_keys);
  table_alloc_kernel<<<(n_keys + th - 1) / th, th, 0, st>>>(t, n_keys);
  table_fill_kernel<<<(n_keys + th - 1) / th, th, 0, st>>>(t, qread, n_keys, read_base);
}

void skb_launch_memb_build(const SkbRefView& rv, uint32_t* memb, uint32_t memb_log2, cudaStream_t st) {
  if (rv.n_rows == 0) return;
  unsigned blocks = (rv.n_rows + 7) / 8;
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  memb_build_kernel<<<blocks, 256, 0, st>>>(rv, memb, memb_log2);
}

size_t skb_fused_smem_bytes(uint32_t cnt_stride, int narrow, uint32_t rowbuf) {
  return FS_SMEM_BLOOM + FS_SMEM_RING + (size_t)cnt_stride * (narrow ? 1 : 2) * rowbuf;
}
uint32_t skb_fused_tile() { return FS_CLAIM; }  // hashes per claim unit (what tile_cum / tpr count)
// Counter buffers (rows in flight) a pass of `reads` reads gets: 8 when they fit next to the filter and the staging
// rings in the 227 KB of a CTA, else 4; 0 = the pass does not fit at all.
uint32_t skb_fused_rowbuf(uint32_t cnt_stride, int narrow) {
  const size_t budget = 232448 - 2048;  // opt-in maximum per CTA on sm_100 minus the static barriers, row bookkeeping, slack
  for (uint32_t rb : {8u, 4u})
    if (skb_fused_smem_bytes(cnt_stride, narrow, rb) <= budget) return rb;
  return 0;
}
// Largest pass the kernel's shared memory holds (with 4 counter buffers); pass-local read ids are SKB_SLOT_ID_BITS
// wide in a table slot: that caps it either way.
uint32_t skb_fused_max_reads(int narrow) {
  const uint32_t gran = narrow ? 512u : 256u;  // cnt_stride granularity (api.cu)
  uint32_t b = 1u << SKB_SLOT_ID_BITS;
  while (b >= gran && skb_fused_rowbuf(b, narrow) == 0) b -= gran;
  return b / gran * gran;
}

void skb_launch_fused(const SkbFusedArgs& a, cudaStream_t st) {
  if (a.rv.n_rows == 0) return;
  const size_t smem = skb_fused_smem_bytes(a.cnt_stride, a.narrow, a.rowbuf);
  static SkbSmemOptIn opt_in;
  if (opt_in.needs(smem)) {
    cudaFuncSetAttribute(fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (a.narrow) fused_kernel<4><<<a.num_ctas, FS_THREADS, smem, st>>>(a);
  else fused_kernel<2><<<a.num_ctas, FS_THREADS, smem, st>>>(a);
}

void skb_launch_tracked_totals(const SkbRefView& rv, const uint32_t* tracked, const uint32_t* n_tracked,
                               const SkbTable& t, unsigned long long* extra, cudaStream_t st) {
  tracked_totals_kernel<<<SKB_MAX_TRACKED * 8, 256, 0, st>>>(rv, tracked, n_tracked, t, extra);
}

void skb_launch_rank_bounds(const SkbRefView& rv, const SkbTable& t, const SkbRankArgs& a, bool has_keys, cudaStream_t st) {
  tracked_prefix_kernel<<<SKB_MAX_TRACKED, 1024, (size_t)a.row_stride * 2, st>>>(rv, t, a, has_keys ? 1 : 0);
  rank_bounds_kernel<<<(a.n_reads + 7) / 8, 256, 0, st>>>(a);
}

void skb_launch_rank_expand(const SkbRankArgs& a, cudaStream_t st) { candidates_kernel<<<148 * 8, 256, 0, st>>>(a); }

void skb_launch_verdict_update(const SkbRankArgs& a, bool with_verdict, cudaStream_t st) {
  verdict_update_kernel<<<1, 1024, 0, st>>>(a, with_verdict ? 1 : 0);
}

void skb_launch_rank_select(const SkbRankArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)RS_WARPS * RS_CACHE * 16;
  static SkbSmemOptIn opt_in;
  if (opt_in.needs(smem)) cudaFuncSetAttribute(rank_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const unsigned blocks = (a.n_reads + RS_WARPS - 1) / RS_WARPS;
  rank_select_kernel<<<blocks, RS_WARPS * 32, smem, st>>>(a);
}

void skb_launch_rank_full(const unsigned long long* vals, uint32_t n, uint32_t top, uint32_t idx_base,
                          uint32_t* out_idx, unsigned long long* out_val, uint32_t* out_local, cudaStream_t st) {
  rank_full_kernel<<<1, 1024, 0, st>>>(vals, n, top, idx_base, out_idx, out_val, out_local);
}

void skb_launch_dense_topk(const SkbDenseArgs& a, cudaStream_t st) {
  if (a.n_reads == 0 || a.groups == 0) return;
  if (a.top <= 32) {
    SkbDenseArgs r = a;
    r.col_stride = 1; r.thr_sum = nullptr; r.thr_idx = nullptr;
    if (a.anchor_idx && a.n_reads > 64) {
      // every 64th read first, over small row groups (enough CTAs for the few columns), merged into exact top lists
      SkbDenseArgs an = a;
      const uint32_t n_anchor = (a.n_reads + 63) / 64;
      an.col_stride = 64; an.thr_sum = nullptr; an.thr_idx = nullptr;
      an.groups = std::max<uint32_t>(1, std::min<uint32_t>(a.anchor_groups, a.n_rows / 64));
      an.part_idx = a.anchor_part_idx; an.part_sum = a.anchor_part_sum;
      const dim3 agrid((n_anchor + 63) / 64, an.groups);
      if (a.wide) dense_topk_warp_kernel<uint32_t><<<agrid, 256, 0, st>>>(an);
      else dense_topk_warp_kernel<uint16_t><<<agrid, 256, 0, st>>>(an);
      skb_launch_merge_topn(an.part_idx, an.part_sum, an.groups, n_anchor, a.top, a.anchor_idx, a.anchor_sum, st);
      r.thr_sum = a.anchor_sum; r.thr_idx = a.anchor_idx;
    }
    const dim3 wgrid((a.n_reads + 63) / 64, a.groups);
    if (a.wide) dense_topk_warp_kernel<uint32_t><<<wgrid, 256, 0, st>>>(r);
    else dense_topk_warp_kernel<uint16_t><<<wgrid, 256, 0, st>>>(r);
    return;
  }
  const dim3 grid((a.n_reads + 127) / 128, a.groups);
  if (a.top <= 16) {
    if (a.wide) dense_topk_kernel<16, uint32_t><<<grid, 128, 0, st>>>(a);
    else dense_topk_kernel<16, uint16_t><<<grid, 128, 0, st>>>(a);
  } else {
    if (a.wide) dense_topk_kernel<SKB_MAX_TOP, uint32_t><<<grid, 128, 0, st>>>(a);
    else dense_topk_kernel<SKB_MAX_TOP, uint16_t><<<grid, 128, 0, st>>>(a);
  }
}

void skb_launch_merge_topn(const uint32_t* idx_parts, const unsigned long long* sum_parts, uint32_t n_parts,
                           uint64_t n_reads, uint32_t top, uint32_t* out_idx, unsigned long long* out_sum,
                           cudaStream_t st) {
  if (n_reads == 0) return;
  const int th = 256;
  const unsigned blocks = (unsigned)((n_reads * 32 + th - 1) / th);
  merge_topn_kernel<<<blocks, th, 0, st>>>(idx_parts, sum_parts, n_parts, n_reads, top, out_idx, out_sum);
}

void skb_launch_relayout(const uint64_t* src, const uint64_t* src_off, uint64_t* dst, const uint64_t* dst_start,
                         uint32_t n_rows, cudaStream_t st) {
  if (n_rows == 0) return;
  unsigned blocks = (n_rows + 7) / 8;
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  relayout_kernel<<<blocks, 256, 0, st>>>(src, src_off, dst, dst_start, n_rows);
}

void skb_launch_ref_check(const SkbRefView& rv, uint32_t* bad, unsigned long long* hmax, cudaStream_t st) {
  if (rv.n_rows == 0) return;
  unsigned blocks = (rv.n_rows + 7) / 8;
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  ref_check_kernel<<<blocks, 256, 0, st>>>(rv, bad, hmax);
}

void skb_launch_shared(const SkbRefView& rv, const uint64_t* q, const uint64_t* q_off, uint32_t Q,
                       unsigned long long* out, cudaStream_t st) {
  const uint64_t pairs = (uint64_t)rv.n_rows * Q;
  if (pairs == 0) return;
  const int th = 256;
  const unsigned blocks = (unsigned)((pairs * 32 + th - 1) / th);
  shared_kernel<<<blocks, th, 0, st>>>(rv, q, q_off, Q, out);
}
