// kernels.h — host-callable launch wrappers (one per kernel); implemented in kernels_sketch.cu / kernels_predict.cu.
#pragma once
#include "common.cuh"

// status codes written by the select kernel
enum : uint32_t { SKB_ST_OK = 0, SKB_ST_OVERFLOW = 1, SKB_ST_UNDERFILL = 2 };

struct SkbHashArgs {
  SkbPackedView pv;
  uint32_t k;
  uint64_t seed;
  const uint64_t* tau;     // [G] keep hashes <= tau[g]
  const uint8_t* active;   // [G] or null (all groups)
  uint64_t* cand;          // candidate pool
  const uint64_t* cand_base;  // [G]
  const uint32_t* cand_cap;   // [G]
  uint32_t* cand_cnt;         // [G] (may exceed cap: overflow)
  unsigned long long* kmers;  // [G] number of valid k-mer windows
  // dump mode (debug): positional output
  uint64_t* dump_hash;
  uint8_t* dump_valid;
};
void skb_launch_hash(const SkbHashArgs& a, cudaStream_t st);

struct SkbSelectArgs {
  uint32_t n_groups;
  uint64_t* cand;
  const uint64_t* cand_base;
  const uint32_t* cand_cap;
  const uint32_t* cand_cnt;
  const uint8_t* active;
  const uint64_t* tau;
  uint32_t s;
  int check_underfill;
  uint64_t* out_hashes;     // sketch mode: [G*s]; in-place mode: == cand
  const uint64_t* out_off;  // null => g*s ; else per-group offset into out_hashes
  uint32_t* out_counts;     // [G*s] or null
  uint32_t* out_n;          // [G]
  uint32_t* status;         // [G]
  uint32_t threads;         // block size (32..1024, power of two)
  uint32_t smem_elems;      // power of two; sort happens in shared memory when the group fits
  uint32_t* any_bad;        // [1] or null: set when a group ends with a status other than SKB_ST_OK
};
void skb_launch_select(const SkbSelectArgs& a, cudaStream_t st);

// Query-mode plan on the device (one threshold for every group): candidate capacity of a group from its length
// (4 x the expected number of hashes <= tau, + 16, a power of two >= 32, at most the length), the groups' places in the
// pool (exclusive scan), and the per-group scratch of hash_kernel / select_kernel reset. Two launches.
struct SkbPlanArgs {
  uint32_t n_groups;
  const uint64_t* g_len;  // [G] bases per group
  uint64_t tau;
  double frac;            // (tau + 1) / 2^64
  uint64_t* tau_out;      // [G]
  uint64_t* base;         // [G]
  uint32_t* cap;          // [G]
  uint32_t* cnt;          // [G] = 0
  unsigned long long* kmers;  // [G] = 0
  uint8_t* active;        // [G] = 1
  uint32_t* status;       // [G] = 0
  uint32_t* any_bad;      // [1] = 0
  unsigned long long* tile_tot;  // [ceil(G / 1024)] scratch: capacity totals of the tiles of 1024 groups
};
void skb_launch_plan_query(const SkbPlanArgs& a, cudaStream_t st);

// predict: gather every read's selected hashes into one flat list (read after read, at q_off[read])
void skb_launch_compact_queries(const uint64_t* cand, const uint64_t* cand_base, const uint32_t* out_n,
                                const uint64_t* q_off, uint32_t n_reads, uint64_t* qh, cudaStream_t st);
// qread[j] = pass read of flat query position j (q_off: [n_reads + 1])
void skb_launch_fill_qread(const uint64_t* q_off, uint32_t n_reads, uint32_t* qread, cudaStream_t st);
// copy the scratch ranking of every piece that ends a read (out_row != UINT32_MAX) to the caller's row
void skb_launch_report_pieces(const uint32_t* idx, const unsigned long long* sum, const uint32_t* out_row,
                              uint32_t n_pieces, uint32_t top, uint32_t* out_idx, unsigned long long* out_sum,
                              cudaStream_t st);

// ---- predict ---------------------------------------------------------------------------------------------
#define SKB_BLOOM_WORDS 16384u  // 64 KB shared-memory filter (2^19 bits)

// One slot of the per-pass query table (open addressing on `key`). A pass holds up to 2^SKB_SLOT_ID_BITS reads.
// meta: the low SKB_SLOT_CNT_BITS bits = cnt, the number of reads of the pass that hold `key` (0 .. 2^ID_BITS);
// cnt <= SKB_SLOT_INLINE: their pass-local read ids sit inline above cnt, so a lookup is one 16-byte load;
// more: the 32 bits above cnt are the start of the slot's read list in `reads`.
// Default: 13-bit ids (up to 8192 reads per pass), 14-bit cnt, 3 inline ids. SKB_X_IDBITS=12: 4096 reads, 13-bit cnt, 4 inline.
#ifndef SKB_X_IDBITS
#define SKB_X_IDBITS 13
#endif
#define SKB_SLOT_ID_BITS SKB_X_IDBITS
#define SKB_SLOT_CNT_BITS (SKB_X_IDBITS + 1)
struct __align__(16) SkbSlot {
  unsigned long long key;
  unsigned long long meta;
};
#define SKB_SLOT_CNT(m) ((uint32_t)((m) & ((1ull << SKB_SLOT_CNT_BITS) - 1ull)))
#define SKB_SLOT_INLINE ((uint32_t)((64 - SKB_SLOT_CNT_BITS) / SKB_SLOT_ID_BITS))
#define SKB_SLOT_ID_SHIFT(i) (SKB_SLOT_CNT_BITS + SKB_SLOT_ID_BITS * (i))
#define SKB_SLOT_ID(m, i) ((uint32_t)(((m) >> SKB_SLOT_ID_SHIFT(i)) & ((1ull << SKB_SLOT_ID_BITS) - 1ull)))
#define SKB_SLOT_START(m) ((uint32_t)(((m) >> SKB_SLOT_CNT_BITS) & 0xFFFFFFFFull))
// Reads per pass are bounded by the fused kernel's shared memory: skb_fused_max_reads(narrow) (kernels_predict.cu):
// 8192 with u8 counters (every read of the pass keeps <= 255 query hashes) and 4 counter buffers, 4096 with u16
// counters; passes of half that size get 8 counter buffers. The library's default pass is 4096 reads.

struct SkbTable {
  SkbSlot* slots;     // [cap + 1]; slot `cap` is reserved for the key that equals SKB_EMPTY_KEY
  uint32_t* fill;     // [cap + 1]
  uint32_t* reads;    // [max_keys] read ids grouped by slot
  uint32_t* slot_of;  // [max_keys]
  uint32_t* bloom;    // [SKB_BLOOM_WORDS]
  uint32_t* cursor;   // [1]
  uint32_t cap;       // power of two
  uint32_t log2cap;
  // membership filter of the whole reference shard (built once per upload): a query hash that is in no reference row
  // (sequencing errors make most read k-mers novel) is left out of the table and of the shared-memory filter
  const uint32_t* memb;   // [1 << memb_log2] words, null = no prefilter
  uint32_t memb_log2;
  unsigned long long* memb_kept;  // [1] statistics: keys that passed, or null
};
// three bits of one word per reference hash
void skb_launch_memb_build(const struct SkbRefView& rv, uint32_t* memb, uint32_t memb_log2, cudaStream_t st);
// n_prev = keys of the table's previous build (its slots are cleared through them), UINT32_MAX = clear every slot
void skb_launch_table_build(const SkbTable& t, const uint64_t* qh, const uint32_t* qread, uint32_t n_keys,
                            uint32_t read_base, uint32_t n_prev, cudaStream_t st);

// Reference shard as laid out in HBM: row r = ref[row_start[r] .. row_start[r] + row_len[r]), row_start even
// (16-byte aligned rows, the granularity of the bulk copies).
struct SkbRefView {
  const uint64_t* ref;
  const uint64_t* row_start;  // [n_rows]
  const uint32_t* row_len;    // [n_rows]
  uint32_t n_rows;
  uint32_t uniform_len;    // != 0: every row holds exactly this many hashes ...
  uint32_t uniform_pitch;  // ... and row r starts at r * uniform_pitch (no per-row loads needed)
};

struct SkbFusedArgs {
  SkbRefView rv;
  const uint32_t* cta_row;  // [num_ctas + 1] contiguous row range of every CTA (balanced by tiles)
  int num_ctas;
  SkbTable table;
  uint32_t n_reads;     // reads in this pass
  uint32_t cnt_stride;  // counters per row buffer (multiple of 512, >= n_reads)
  int narrow;           // counters are u8 (every read of the pass keeps <= 255 query hashes) instead of u16
  uint32_t rowbuf, rowbuf_log2;  // counter buffers (rows in flight) per CTA: 4 or 8 (skb_fused_rowbuf)
  int skip_stream;      // the pass has no query hashes: rows are ranked without being streamed
  uint32_t row_base;    // global index of local row 0
  const unsigned long long* sums_in;  // [n_rows]
  unsigned long long* sums_out;       // [n_rows]
  const unsigned long long* lb_sum;   // [n_reads] lower bound of every read's top-th key
  const uint32_t* lb_idx;             // [n_reads]
  SkbInterval* ivl;                   // [ivl_cap] candidate intervals (a lane segment without hits that meets its bound)
  uint32_t ivl_cap;
  uint32_t* ivl_total;                // [1] intervals produced; > ivl_cap = overflow
  uint4* seg_hdr;                     // [seg_cap] segment records: {sum at segment start (lo, hi), global row, first read}
  uint32_t* seg_words;                // [seg_cap][cnt_stride / 32 / counters-per-word] the segment's counters
  uint32_t seg_cap;
  uint32_t* seg_total;                // [1] records produced; > seg_cap = overflow
  void* dense;                        // non-null: dense pass, [n_rows][cnt_stride] prefix sums (u16 / u32) instead of candidates
  uint32_t* dense_overflow;           // [1] set when a u16 prefix sum overflows
  const uint32_t* tile_cum;           // [n_rows + 1] sub-tiles before each row of the shard (ragged shards only)
  uint32_t tpr, tpr_magic;            // uniform shards: sub-tiles per row and ceil(2^32 / tpr) (tile -> row by a multiply)
  const uint32_t* abort;              // [2] {set once an earlier pass of the batch overflowed, its sequence number}: skip
};
void skb_launch_fused(const SkbFusedArgs& a, cudaStream_t st);
size_t skb_fused_smem_bytes(uint32_t cnt_stride, int narrow, uint32_t rowbuf);
uint32_t skb_fused_rowbuf(uint32_t cnt_stride, int narrow);
#define SKB_IVL_CAP ((4u << 20) << (SKB_X_IDBITS - 12))  // candidate intervals per pass (4 M per 4096 reads); more than that shrinks the pass
#define SKB_SEG_CAP (2u << 20)     // segment records per pass (16 + up to 160 bytes each)
// counters of one lane segment, in words: reads / 32 / 4 with u8 counters, at most 2560 / 32 / 2 = 40 with u16 counters
#define SKB_SEG_WORDS_MAX ((1u << SKB_X_IDBITS) / 128u > 40u ? (1u << SKB_X_IDBITS) / 128u : 40u)
uint32_t skb_fused_tile();
uint32_t skb_fused_max_reads(int narrow);

// totals of the tracked rows against a pass's table (all reads of that pass): extra[t] += hits of tracked row t
void skb_launch_tracked_totals(const SkbRefView& rv, const uint32_t* tracked, const uint32_t* n_tracked,
                               const SkbTable& t, unsigned long long* extra, cudaStream_t st);
#define SKB_CAND_BUDGET ((24u << 20) << (SKB_X_IDBITS - 12))  // candidate records per pass over all reads (384 MB per 4096 reads); bucket = budget / reads

#define SKB_MAX_TRACKED 192u  // rows whose exact per-read sums define the bounds

struct SkbRankArgs {
  uint16_t* tracked_counts;        // [SKB_MAX_TRACKED][row_stride] per-read counts of the tracked rows
  uint32_t* tracked_prefix;        // [SKB_MAX_TRACKED][row_stride] inclusive prefix sums of the above
  unsigned long long* tracked_extra;  // [SKB_MAX_TRACKED] added to sums_in of each tracked row (its total of the pass that is still streaming; zero otherwise)
  uint32_t row_stride;
  uint32_t n_reads;  // reads in this pass
  uint32_t row_base; // global index of local row 0
  const unsigned long long* sums_in;   // [n_rows]
  const uint32_t* tracked;             // [*n_tracked] local rows
  const uint32_t* n_tracked;           // device scalar, <= SKB_MAX_TRACKED
  unsigned long long* lb_sum;          // [n_reads]
  uint32_t* lb_idx;                    // [n_reads] (global index)
  const SkbInterval* ivl;              // candidate intervals of the pass
  uint32_t ivl_cap;
  const uint32_t* ivl_total;
  const uint4* seg_hdr;                // segment records of the pass (see SkbFusedArgs)
  const uint32_t* seg_words;
  uint32_t seg_cap, seg_words_per, seg_cpw;
  const uint32_t* seg_total;
  SkbCand* cand;                       // [n_reads][cand_cap] per-read candidate buckets
  uint32_t cand_cap;
  uint32_t* cand_total;                // [1] overflow flag
  uint32_t* cand_cnt;                  // [n_reads]
  unsigned long long* cand_stat;       // [1] running total of candidates (statistics only)
  uint32_t top;
  uint32_t* out_idx;                   // [n_reads * top] device
  unsigned long long* out_sum;         // [n_reads * top]
  uint32_t* tracked_next;              // [SKB_MAX_TRACKED] tracked rows for the next pass
  uint32_t* n_tracked_next;            // device scalar
  uint32_t* abort;                     // [2] see SkbFusedArgs::abort
  uint32_t seq;                        // sequence number of this pass within the call
};
// bounds of a sparse pass: per-read counts of the tracked rows against the pass's table and their prefix sums (one CTA
// per tracked row), then every read's top-th best tracked key (one warp per read)
void skb_launch_rank_bounds(const SkbRefView& rv, const SkbTable& t, const SkbRankArgs& a, bool has_keys, cudaStream_t st);
void skb_launch_rank_expand(const SkbRankArgs& a, cudaStream_t st);  // segment records + intervals -> per-read candidate buckets
void skb_launch_rank_select(const SkbRankArgs& a, cudaStream_t st);  // per-read top-N
// last launch of a pass (one CTA). with_verdict: record the first pass whose candidates overflowed (abort[0] = 1,
// abort[1] = seq) and clear the pass's counters and bucket fill counts; then the tracked rows for the next passes = union
// of the top lists of 16 sampled reads of this pass (last read first), unless a pass is being redone
void skb_launch_verdict_update(const SkbRankArgs& a, bool with_verdict, cudaStream_t st);

// top-N of a plain value array by (value desc, index asc); one CTA. idx_base is added to reported indices.
void skb_launch_rank_full(const unsigned long long* vals, uint32_t n, uint32_t top, uint32_t idx_base,
                          uint32_t* out_idx, unsigned long long* out_val, uint32_t* out_local, cudaStream_t st);

// dense ranking of one pass: per (read, row group) top lists from the prefix-sum vectors (see dense_topk_kernel)
struct SkbDenseArgs {
  const void* dense;                   // [n_rows][cnt_stride] u16 (wide == 0) or u32 (wide == 1)
  int wide;
  uint32_t cnt_stride, n_rows, n_reads, row_base, top, groups;
  const unsigned long long* sums_in;   // [n_rows] before the pass
  const unsigned long long* sums_out;  // [n_rows] after the pass (== sums_in: the row had no hit, its vector is all zero)
  uint32_t* part_idx;                  // [groups][n_reads][top]
  unsigned long long* part_sum;
  // top <= 32: the exact lists of every 64th read are made first and give the other reads their starting thresholds
  uint32_t anchor_groups;              // row groups of the anchor launch
  uint32_t* anchor_part_idx;           // [anchor_groups][ceil(n_reads / 64)][top]; null = no anchor launch
  unsigned long long* anchor_part_sum;
  uint32_t* anchor_idx;                // [ceil(n_reads / 64)][top] merged
  unsigned long long* anchor_sum;
  // set by the launcher:
  uint32_t col_stride;                 // column c of the kernel = read c * col_stride
  const unsigned long long* thr_sum;   // starting threshold of CTA x = entry [x * top + top - 1] (or null)
  const uint32_t* thr_idx;
};
void skb_launch_dense_topk(const SkbDenseArgs& a, cudaStream_t st);

void skb_launch_merge_topn(const uint32_t* idx_parts, const unsigned long long* sum_parts, uint32_t n_parts,
                           uint64_t n_reads, uint32_t top, uint32_t* out_idx, unsigned long long* out_sum,
                           cudaStream_t st);

// copy ragged rows into the aligned device layout
void skb_launch_relayout(const uint64_t* src, const uint64_t* src_off, uint64_t* dst, const uint64_t* dst_start,
                         uint32_t n_rows, cudaStream_t st);

// reference validation: rows strictly increasing; writes flag != 0 on violation, and the max hash
void skb_launch_ref_check(const SkbRefView& rv, uint32_t* bad, unsigned long long* hmax, cudaStream_t st);

// dense shared counts: out[i*Q + j] = |ref_i ∩ q_j| (warp per pair, binary-search merge)
void skb_launch_shared(const SkbRefView& rv, const uint64_t* q, const uint64_t* q_off, uint32_t Q,
                       unsigned long long* out, cudaStream_t st);
