/*
 * sketchy_b200.h — C ABI of libsketchy_b200.so: the B200 (sm_100a) implementation of sketchy's MinHash hot path.
 *
 * This is the drop-in boundary. The reference (esteinig/sketchy 0.6.0, Rust) has no FFI of its own; the seam this
 * library sits behind is (i) finch's `SketchScheme` trait object — create_sketcher() -> process(record)* ->
 * to_vec()/total_bases_and_kmers() — used at reference src/sketchy.rs:291-302, 331-335, 473-481, and (ii) the
 * in-tree compute `_common_hashes` + accumulate + stable sort + slice at src/sketchy.rs:305-310, 337-348, 419-459.
 * Each entry point below cites the reference lines it replaces. INTEGRATION.md shows the Rust `extern "C"` block a
 * maintainer would add to bind them.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns SKB_OK (0) or a negative SKB_ERR_* code and never
 *     throws, aborts or prints. skb_last_error(ctx) gives the message of the last failing call on that context.
 *   - a context is NOT thread-safe: one context per host thread (the reference's callers are single-threaded per
 *     sketcher too: one sketcher per rayon task, src/sketchy.rs:470-473).
 *   - there is NO CPU fallback: without a CUDA device skb_create fails with SKB_ERR_NO_DEVICE.
 *   - hashes are u64 (finch ItemHash), ascending and distinct inside one sketch (finch `to_vec()`).
 */
#ifndef SKETCHY_B200_H
#define SKETCHY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKB_OK 0
#define SKB_ERR_INVALID_ARG (-1)
#define SKB_ERR_CUDA (-2)
#define SKB_ERR_NO_DEVICE (-3)
#define SKB_ERR_REF_NOT_SORTED (-4) /* a reference row is not strictly increasing */
#define SKB_ERR_TOP_GT_N (-5)       /* reference panics on result_vec[..top], src/sketchy.rs:371,391 */
#define SKB_ERR_NO_REFERENCE (-6)
#define SKB_ERR_UNSUPPORTED_K (-7) /* k must be 1..32 */
#define SKB_ERR_OOM (-8)
#define SKB_ERR_INTERNAL (-9)
#define SKB_ERR_STATE (-10) /* call order violated (e.g. adding to a staged batch) */
#define SKB_ERR_COMM (-11)  /* NCCL could not be loaded or a collective failed */

#define SKB_MAX_K 32
#define SKB_MAX_TOP 128

typedef struct skb_ctx skb_ctx;
typedef struct skb_batch skb_batch;

/* ---- context ------------------------------------------------------------------------------------------------ */

/* Bind a context to CUDA device `device` (one process per GPU: pass LOCAL_RANK). Owns the stream, all device
 * memory, the resident reference shard and the running sums. */
int skb_create(int device, skb_ctx** out);
void skb_destroy(skb_ctx* ctx);
const char* skb_last_error(const skb_ctx* ctx);
const char* skb_version(void);
/* The CUDA stream (cudaStream_t) every kernel of this context is launched on, for callers that time with events. */
void* skb_stream(skb_ctx* ctx);
int skb_synchronize(skb_ctx* ctx);

/* ---- ingest: pinned, batched 2-bit packing (replaces needletail record iteration + `normalize(false)`,
 *      reached through `sketcher.process(record)` at src/sketchy.rs:296, 333, 477) ------------------------------ */

/* A batch holds records normalised and packed 2 bit/base (+1 validity bit/base) in pinned host memory, grouped:
 * group = input file for `sketch` (one sketcher per file, src/sketchy.rs:473-478), = read for streaming predict
 * (one sketcher per read, :331), = 0 for read-set predict (one sketcher for all reads, :291). */
int skb_batch_create(skb_ctx* ctx, skb_batch** out);
void skb_batch_destroy(skb_batch* b);
int skb_batch_clear(skb_batch* b);
/* Append n records given as one blob + offsets[n+1]. groups[n] must be non-decreasing and continue from the last
 * group already in the batch (equal = same sketcher, +1.. = next); NULL means "every record is its own new group".
 * Bytes are mapped exactly as needletail's normalize(false): ACGT kept, acgt upper-cased, u/U -> T, whitespace
 * removed, everything else breaks k-mer windows. nthreads host threads do the packing (0 = all cores). */
int skb_batch_add(skb_batch* b, const uint8_t* blob, const uint64_t* offsets, const uint32_t* groups, uint64_t n,
                  uint32_t nthreads);
/* The same with the records given one by one: record r = lens[r] bytes at recs[r], anywhere in host memory (slices of
 * file buffers read in place: a FASTA file goes from the page cache to the packer without an intermediate copy, which
 * is how the `sketch` host reads its files where the reference iterates needletail records, src/sketchy.rs:474-478). */
int skb_batch_add_records(skb_batch* b, const uint8_t* const* recs, const uint64_t* lens, const uint32_t* groups,
                          uint64_t n, uint32_t nthreads);
/* What a group's total_bases counts (finch `total_bases_and_kmers`, src/sketchy.rs:481 -> `.msh` length / `info`):
 * SKB_BASES_RAW (default) = the bytes of the record's sequence as the reader hands them over, line breaks of a
 * multi-line FASTA record included (needletail 0.4.1 passes its raw slice on);
 * SKB_BASES_STRIPPED = the bases left after blanks, tabs, CR and LF are removed. Hashes never depend on it.
 * Set on an empty batch. */
#define SKB_BASES_RAW 0
#define SKB_BASES_STRIPPED 1
int skb_batch_set_base_count(skb_batch* b, int mode);
uint32_t skb_batch_num_groups(const skb_batch* b);
uint64_t skb_batch_num_records(const skb_batch* b);
uint64_t skb_batch_num_bases(const skb_batch* b); /* sum of raw record lengths (finch total_bases) */
/* Start copying the packed batch to the device (H2D on the context's copy stream; the call does not wait, the kernels
 * that read the batch do, on the device). Calls that consume a batch stage it on demand; staging ahead lets a caller
 * keep inputs resident in HBM. skb_batch_add / skb_batch_stage of one batch may run on another host thread while the
 * context works on a different batch (a ring of batches: packing, copy and kernels overlap). skb_batch_clear waits for
 * the batch's copies before its page-locked buffers are reused. */
int skb_batch_stage(skb_batch* b);

/* ---- sketch (replaces `_sketch_files`, src/sketchy.rs:465-494: create_sketcher :473, process :477, to_vec :480,
 *      total_bases_and_kmers :481) ----------------------------------------------------------------------------- */

/* One bottom-s sketch per group. out_hashes/out_counts are [G*s] (row g at g*s, first out_n[g] valid, ascending,
 * distinct; counts = occurrences, finch KmerCount.count; may be NULL). out_bases/out_kmers = finch totals. */
int skb_sketch(skb_ctx* ctx, skb_batch* b, uint32_t k, uint32_t s, uint64_t seed, uint64_t* out_hashes,
               uint32_t* out_counts, uint32_t* out_n, uint64_t* out_bases, uint64_t* out_kmers);

/* ---- reference residency (replaces the in-memory Vec<Sketch> built by `_read_sketch`, src/sketchy.rs:497-536) - */

/* Upload this GPU's shard of the reference: n_rows ragged rows, row i = hashes[off[i] .. off[i+1]).
 * Rows must be strictly increasing (what makes the merge of :419-459 a set intersection) or the call fails with
 * SKB_ERR_REF_NOT_SORTED. global_row_base is added to local row numbers in every reported index (multi-GPU:
 * contiguous row ranges per rank). Resets the running sums. */
int skb_ref_upload(skb_ctx* ctx, const uint64_t* hashes, const uint64_t* off, uint32_t n_rows,
                   uint32_t global_row_base);
/* Same, from device memory (hashes: device pointer to off[n_rows] u64 values; off: HOST pointer). */
int skb_ref_upload_device(skb_ctx* ctx, const uint64_t* d_hashes, const uint64_t* off, uint32_t n_rows,
                          uint32_t global_row_base);
uint32_t skb_ref_rows(const skb_ctx* ctx);

/* ---- streaming predict (replaces `_sum_of_shared_hashes`, src/sketchy.rs:317-356: per-read sketcher :331-335,
 *      `_common_hashes` vs every reference :337-339, `sum[i] += shared` :341, stable sort :348, `[..top]` :391) -- */

/* Every group of `b` is one read, in stream order. For read r (0-based in this call) writes the `top` best
 * references by (cumulative sum desc, reference index asc) — the order of the reference's stable sort — to
 * out_idx[r*top ..] / out_sum[r*top ..] (host memory). Running sums stay resident in the context across calls.
 * s_query = len(hashes of GLOBAL reference #0) as the reference derives it (src/sketchy.rs:82, 522).
 * top > total rows of this shard is SKB_ERR_TOP_GT_N unless the shard is part of a larger reference, in which
 * case the caller passes pad=1 and missing entries are (idx=UINT32_MAX, sum=0) and sort last. */
int skb_predict_stream(skb_ctx* ctx, skb_batch* b, uint32_t k, uint32_t s_query, uint64_t seed, uint32_t top,
                       int pad, uint32_t* out_idx, uint64_t* out_sum);
/* Same with DEVICE output pointers (no D2H; results are complete when the call returns). */
int skb_predict_stream_device(skb_ctx* ctx, skb_batch* b, uint32_t k, uint32_t s_query, uint64_t seed,
                              uint32_t top, int pad, uint32_t* d_out_idx, uint64_t* d_out_sum);
int skb_sums_reset(skb_ctx* ctx);
int skb_sums_download(skb_ctx* ctx, uint64_t* out /* [n_rows] */);
int skb_sums_upload(skb_ctx* ctx, const uint64_t* in /* [n_rows] */);
/* Reads per streaming pass: 0 = automatic (4096 for shards of 2 GB and more: the streaming kernel's best fraction of
 * the HBM roofline; 8192 for smaller shards, where the per-pass work that does not depend on the shard dominates), up
 * to 8192 (fewer passes over the matrix: more reads per second at a lower fraction). Any value gives identical results.
 * skb_pass_reads returns the size in effect. */
int skb_set_pass_reads(skb_ctx* ctx, uint32_t max_reads_per_pass);
uint32_t skb_pass_reads(const skb_ctx* ctx);
/* How a pass turns its counts into every read's top-N; every mode gives identical results (tests run all of them).
 * 0 = automatic (default): brute-force ranking over all rows right after a reset, for small shards and to redo a
 * pass whose candidate lists overflowed, candidate lists from per-read lower bounds otherwise;
 * 1 = candidate lists wherever they are possible (even on small shards); 2 = brute force always. */
int skb_set_rank_mode(skb_ctx* ctx, int mode);

/* ---- read-set predict (replaces `_shared_hashes`, src/sketchy.rs:281-315) and `shared` (:238-279) ------------- */

/* Shared-hash counts of Q query sketches (ragged, strictly increasing rows, HOST memory) against every resident
 * reference row: out[i*Q + j] = |ref_i ∩ query_j|, the print order of src/sketchy.rs:251-252. */
int skb_shared_counts(skb_ctx* ctx, const uint64_t* q_hashes, const uint64_t* q_off, uint32_t Q, uint64_t* out);
/* Rank one count vector as the reference does (stable, descending): first `top` of (count desc, index asc). */
int skb_rank_counts(skb_ctx* ctx, const uint64_t* counts, uint32_t n, uint32_t top, uint32_t* out_idx,
                    uint64_t* out_sum);

/* ---- multi-GPU merge (no reference equivalent; exact because every global top-N member is in its shard's
 *      local top-N and (sum desc, index asc) is a total order) ------------------------------------------------- */

/* d_idx_parts/d_sum_parts: [n_parts][n_reads][top] device arrays (e.g. the all-gathered per-rank outputs);
 * writes the merged [n_reads][top] to d_out_*. */
int skb_merge_topn_device(skb_ctx* ctx, const uint32_t* d_idx_parts, const uint64_t* d_sum_parts,
                          uint32_t n_parts, uint64_t n_reads, uint32_t top, uint32_t* d_out_idx,
                          uint64_t* d_out_sum);

/* ---- multi-GPU predict: one process per GPU, the reference rows sharded by contiguous range over the ranks (each
 *      rank uploads its range with its global_row_base), one NCCL communicator over NVLink / NVSwitch. The reference
 *      is single-process; this is the scale-out of `_sum_of_shared_hashes` (src/sketchy.rs:317-356): every rank ranks
 *      every read against its rows, the per-rank top-N lists are gathered and merged. ---------------------------- */

#define SKB_COMM_ID_BYTES 128
/* On ONE rank: make the communicator's id; hand the bytes to every rank by any means (a file, MPI, a socket). */
int skb_comm_unique_id(uint8_t id[SKB_COMM_ID_BYTES]);
/* On every rank (collective): join the communicator as `rank` of `world`. libnccl.so.2 is loaded at this point. */
int skb_comm_init(skb_ctx* ctx, const uint8_t id[SKB_COMM_ID_BYTES], int rank, int world);
int skb_comm_destroy(skb_ctx* ctx);
int skb_comm_rank(const skb_ctx* ctx);
int skb_comm_world(const skb_ctx* ctx);
/* The contiguous split every collective call assumes: items [*begin, *begin + *count) of n belong to `rank`
 * (ceil(n / world) items per rank, the last ranks may get fewer or none). Use it for reference rows and for reads. */
void skb_dist_range(uint64_t n, int rank, int world, uint64_t* begin, uint64_t* count);
/* Collective: every rank contributes `bytes` bytes (host memory); recv ([world * bytes], host) holds rank r's
 * contribution at r * bytes on every rank. For the small host-side exchanges of a sharded run (local top-N of the
 * read-set mode, the sketches of a rank's share of the input files). */
int skb_comm_allgather_host(skb_ctx* ctx, const void* send, void* recv, uint64_t bytes);
/* Collective streaming predict of `reads_total` reads. `local` holds ONLY this rank's reads,
 * skb_dist_range(reads_total, rank, world), one group each: a rank packs, copies and hashes 1/world of the reads and
 * the per-read query-hash lists are exchanged. out_idx / out_sum ([reads_total * top], host; may be NULL on ranks that
 * do not report) receive the merged ranking of EVERY read, identical on all ranks and identical to a single GPU
 * holding all rows. Without a communicator it is skb_predict_stream. */
int skb_predict_stream_dist(skb_ctx* ctx, skb_batch* local, uint64_t reads_total, uint32_t k, uint32_t s_query,
                            uint64_t seed, uint32_t top, uint32_t* out_idx, uint64_t* out_sum);
/* Same with DEVICE output pointers. */
int skb_predict_stream_dist_device(skb_ctx* ctx, skb_batch* local, uint64_t reads_total, uint32_t k,
                                   uint32_t s_query, uint64_t seed, uint32_t top, uint32_t* d_out_idx,
                                   uint64_t* d_out_sum);

/* ---- measurement hooks (bench.py) ---------------------------------------------------------------------------- */

enum skb_kernel_id {
  SKB_K_HASH = 0,      /* canonical k-mer + MurmurHash3 + threshold filter */
  SKB_K_SELECT = 1,    /* bottom-s sort/dedup/count */
  SKB_K_TABLE = 2,     /* query hash-set build */
  SKB_K_STREAM = 3,    /* reference matrix stream + probe (the HBM-bound kernel) */
  SKB_K_RANK = 4,      /* cumulative sums + candidate filter + top-N */
  SKB_K_MERGE = 5,     /* multi-GPU top-N merge */
  SKB_K_SHARED = 6,    /* dense shared-count kernel */
  SKB_K_MISC = 7,
  SKB_K_COUNT = 8
};
int skb_prof_enable(skb_ctx* ctx, int on); /* CUDA-event timing of every launch, grouped by skb_kernel_id */
int skb_prof_reset(skb_ctx* ctx);
int skb_prof_get(skb_ctx* ctx, int kernel_id, double* total_ms, uint64_t* launches);
uint64_t skb_launch_count(const skb_ctx* ctx); /* kernels launched by this context since creation */
/* Bytes of reference hashes the last predict call streamed per pass, and the number of passes it made. */
int skb_last_predict_stats(const skb_ctx* ctx, uint64_t* ref_bytes_per_pass, uint64_t* passes,
                           uint64_t* query_hashes, uint64_t* candidates);
/* Query hashes of the last predict call that occur in at least one reference row of the shard according to the
 * membership prefilter (the others never entered a pass's table). Equals query_hashes when the prefilter is off. */
uint64_t skb_last_predict_member_hashes(const skb_ctx* ctx);

/* ---- debug / parity hooks (used only by tests) --------------------------------------------------------------- */

/* Internal knobs for tests and experiments (every setting gives identical results): "cand_budget" = candidate
 * records per pass over all reads (forces overflow / redo paths), "trace_passes" = one stderr line per checkpoint,
 * "dense_after_reset" = brute-force ranked passes after the first one of a reset, "pipeline" = bounds taken two passes
 * back, "stream_ctas" = CTAs of the streaming kernel (takes effect at the next upload). The same names, upper-cased
 * with an SKB_ prefix, are read from the environment once when the context is created. */
int skb_debug_set(skb_ctx* ctx, const char* key, uint64_t value);

/* Hash of the canonical k-mer starting at every packed position of the batch; valid[p] = 0 where the window holds
 * a non-ACGT base or crosses a record end. Arrays are [skb_batch_packed_len(b)]. */
uint64_t skb_batch_packed_len(const skb_batch* b);
int skb_batch_record_start(const skb_batch* b, uint64_t record, uint64_t* packed_pos, uint64_t* packed_len);
int skb_debug_kmer_hashes(skb_ctx* ctx, skb_batch* b, uint32_t k, uint64_t seed, uint64_t* out_hash,
                          uint8_t* out_valid);

#ifdef __cplusplus
}
#endif
#endif /* SKETCHY_B200_H */
