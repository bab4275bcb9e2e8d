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
};
void skb_launch_select(const SkbSelectArgs& a, cudaStream_t st);

// predict: gather every read's selected hashes into one flat (hash, read) list
void skb_launch_compact_queries(const uint64_t* cand, const uint64_t* cand_base, const uint32_t* out_n,
                                const uint64_t* q_off, uint32_t n_reads, uint64_t* qh, uint32_t* qread,
                                cudaStream_t st);

// ---- predict ---------------------------------------------------------------------------------------------
#define SKB_BLOOM_WORDS 32768u  // 128 KB shared-memory filter (2^20 bits)
struct SkbTable {
  uint64_t* keys;     // [cap + 1] open addressing, SKB_EMPTY_KEY = free; slot `cap` is reserved for key == EMPTY
  uint32_t* cnt;      // [cap + 1]
  uint32_t* start;    // [cap + 1]
  uint32_t* fill;     // [cap + 1]
  uint32_t* reads;    // [max_keys] read ids grouped by slot
  uint32_t* slot_of;  // [max_keys]
  uint32_t* bloom;    // [SKB_BLOOM_WORDS]
  uint32_t* cursor;   // [1]
  uint32_t cap;       // power of two
  uint32_t log2cap;
};
void skb_launch_table_build(const SkbTable& t, const uint64_t* qh, const uint32_t* qread, uint32_t n_keys,
                            uint32_t read_base, cudaStream_t st);

struct SkbStreamArgs {
  const uint64_t* ref;   // flat reference hashes, 16-byte aligned, padded to a multiple of 2
  uint64_t ref_len;      // number of hashes
  const uint64_t* row_off;  // [n_rows + 1]
  uint32_t n_rows;
  uint32_t uniform_len;  // != 0: every row has exactly this many hashes
  SkbTable table;
  uint16_t* counts;      // [n_rows][row_stride] per-pass (row, read) shared-hash counts
  uint32_t row_stride;   // in u16 elements, multiple of 8
  int num_ctas;
};
void skb_launch_stream(const SkbStreamArgs& a, cudaStream_t st);
size_t skb_stream_smem_bytes();

struct SkbRankArgs {
  const uint16_t* counts;
  uint32_t row_stride;
  uint32_t n_rows;
  uint32_t n_reads;  // reads in this pass
  uint32_t row_base; // global index of local row 0
  const unsigned long long* sums_in;   // [n_rows]
  unsigned long long* sums_out;        // [n_rows]
  const uint32_t* tracked;             // [n_tracked] local rows
  uint32_t n_tracked;
  unsigned long long* lb_sum;          // [n_reads]
  uint32_t* lb_idx;                    // [n_reads] (global index)
  SkbCand* cand;                       // [cand_cap]
  uint32_t cand_cap;
  uint32_t* cand_total;                // [1] (keeps counting past cap: overflow)
  uint32_t* cand_cnt;                  // [n_reads]
  uint32_t* cand_off;                  // [n_reads + 1]
  uint32_t* cand_fill;                 // [n_reads]
  SkbCand* cand_sorted;                // [cand_cap]
  uint32_t top;
  uint32_t* out_idx;                   // [n_reads * top] device
  unsigned long long* out_sum;         // [n_reads * top]
  uint32_t* tracked_next;              // [top] local rows of the last read's top
};
void skb_launch_rank_bounds(const SkbRankArgs& a, cudaStream_t st);
void skb_launch_rank_scan(const SkbRankArgs& a, cudaStream_t st);
void skb_launch_rank_group(const SkbRankArgs& a, cudaStream_t st);   // offsets + scatter by read
void skb_launch_rank_select(const SkbRankArgs& a, cudaStream_t st);  // per-read top-N

// top-N of a plain value array by (value desc, index asc); one CTA. idx_base is added to reported indices.
void skb_launch_rank_full(const unsigned long long* vals, uint32_t n, uint32_t top, uint32_t idx_base,
                          uint32_t* out_idx, unsigned long long* out_val, uint32_t* out_local, cudaStream_t st);

void skb_launch_merge_topn(const uint32_t* idx_parts, const unsigned long long* sum_parts, uint32_t n_parts,
                           uint64_t n_reads, uint32_t top, uint32_t* out_idx, unsigned long long* out_sum,
                           cudaStream_t st);

// reference validation: rows strictly increasing; writes flag != 0 on violation, and the max hash
void skb_launch_ref_check(const uint64_t* ref, const uint64_t* row_off, uint32_t n_rows, uint32_t* bad,
                          unsigned long long* hmax, cudaStream_t st);

// dense shared counts: out[i*Q + j] = |ref_i ∩ q_j| (warp per pair, binary-search merge)
void skb_launch_shared(const uint64_t* ref, const uint64_t* row_off, uint32_t n_rows, const uint64_t* q,
                       const uint64_t* q_off, uint32_t Q, unsigned long long* out, cudaStream_t st);
