// mock_abi.cpp — TEST INFRASTRUCTURE: a recording stand-in for the part of the C ABI the `sketchy` CLI host calls.
// It computes nothing. Every call is appended to the file named by $MOCK_ABI_LOG so that tests/test_host_cli.py can
// check, without a GPU, WHAT the host hands to the library (records, groups, parameters, call order); outputs are
// filled with fixed patterns. The `-m gpu` tests run the same CLI against the real library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/sketchy_b200.h"

struct skb_ctx { std::string err; uint32_t n_rows = 0; };
struct skb_batch { skb_ctx* ctx; uint32_t groups = 0; uint64_t records = 0, bases = 0; long last_group = -1; };

static void logf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
#include <cstdarg>
static void logf(const char* fmt, ...) {
  const char* p = getenv("MOCK_ABI_LOG");
  if (!p) return;
  FILE* f = fopen(p, "a");
  if (!f) return;
  va_list ap;
  va_start(ap, fmt);
  vfprintf(f, fmt, ap);
  va_end(ap);
  fclose(f);
}
static unsigned long long fnv(const uint8_t* p, size_t n) {
  unsigned long long h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

extern "C" {
int skb_create(int device, skb_ctx** out) { *out = new skb_ctx(); logf("create %d\n", device); return SKB_OK; }
void skb_destroy(skb_ctx* c) { delete c; }
const char* skb_last_error(const skb_ctx* c) { return c ? c->err.c_str() : "null"; }
int skb_batch_create(skb_ctx* c, skb_batch** out) { *out = new skb_batch(); (*out)->ctx = c; return SKB_OK; }
void skb_batch_destroy(skb_batch* b) { delete b; }
int skb_batch_clear(skb_batch* b) { b->groups = 0; b->records = 0; b->bases = 0; b->last_group = -1; logf("batch_clear\n"); return SKB_OK; }
static int add_impl(skb_batch* b, const uint8_t* const* recs, const uint64_t* lens, const uint32_t* groups, uint64_t n) {
  // one line per record: group, length, checksum of the raw slice
  logf("batch_add n=%llu groups=%s\n", (unsigned long long)n, groups ? "given" : "null");
  for (uint64_t r = 0; r < n; ++r) {
    long g = groups ? (long)groups[r] : b->last_group + 1;
    if (g < b->last_group) { b->ctx->err = "groups must be non-decreasing"; return SKB_ERR_INVALID_ARG; }
    if (g != b->last_group) b->groups = (uint32_t)g + 1;
    b->last_group = g;
    if (getenv("MOCK_ABI_LOG")) logf("  rec group=%ld len=%llu fnv=%llu\n", g, (unsigned long long)lens[r], fnv(recs[r], lens[r]));
    b->bases += lens[r];
  }
  b->records += n;
  return SKB_OK;
}
int skb_batch_add(skb_batch* b, const uint8_t* blob, const uint64_t* off, const uint32_t* groups, uint64_t n, uint32_t) {
  std::vector<const uint8_t*> recs(n);
  std::vector<uint64_t> lens(n);
  for (uint64_t r = 0; r < n; ++r) { recs[r] = blob + off[r]; lens[r] = off[r + 1] - off[r]; }
  return add_impl(b, recs.data(), lens.data(), groups, n);
}
// the same records one by one: logged like skb_batch_add (the tests check WHAT reaches the library, not through which door)
int skb_batch_add_records(skb_batch* b, const uint8_t* const* recs, const uint64_t* lens, const uint32_t* groups, uint64_t n, uint32_t) {
  return add_impl(b, recs, lens, groups, n);
}
int skb_batch_stage(skb_batch*) { return SKB_OK; }  // (not logged: the copy is an overlap detail, not part of what is handed over)
uint32_t skb_batch_num_groups(const skb_batch* b) { return b->groups; }
int skb_sketch(skb_ctx*, skb_batch* b, uint32_t k, uint32_t s, uint64_t seed, uint64_t* oh, uint32_t* oc, uint32_t* on,
               uint64_t* ob, uint64_t* ok) {
  logf("sketch k=%u s=%u seed=%llu groups=%u records=%llu\n", k, s, (unsigned long long)seed, b->groups, (unsigned long long)b->records);
  for (uint32_t g = 0; g < b->groups; ++g) {  // pattern: group g gets the single hash g+1 with count 1
    oh[(size_t)g * s] = g + 1;
    if (oc) oc[(size_t)g * s] = 1;
    on[g] = 1; ob[g] = 100 + g; ok[g] = 10 + g;
  }
  return SKB_OK;
}
int skb_ref_upload(skb_ctx* c, const uint64_t*, const uint64_t* off, uint32_t n_rows, uint32_t base) {
  c->n_rows = n_rows;
  logf("ref_upload rows=%u hashes=%llu base=%u\n", n_rows, (unsigned long long)(n_rows ? off[n_rows] : 0), base);
  return SKB_OK;
}
int skb_predict_stream(skb_ctx*, skb_batch* b, uint32_t k, uint32_t s_query, uint64_t seed, uint32_t top, int pad,
                       uint32_t* oi, uint64_t* os) {
  logf("predict_stream k=%u s_query=%u seed=%llu top=%u pad=%d reads=%u\n", k, s_query, (unsigned long long)seed, top, pad, b->groups);
  for (uint32_t r = 0; r < b->groups; ++r)
    for (uint32_t t = 0; t < top; ++t) { oi[(size_t)r * top + t] = t; os[(size_t)r * top + t] = 7; }
  return SKB_OK;
}
// multi-GPU entry points: one rank, no exchange (the real ones are exercised by tests/test_multi_gpu.py on GPUs)
int skb_comm_unique_id(uint8_t* id) { memset(id, 0, SKB_COMM_ID_BYTES); return SKB_OK; }
int skb_comm_init(skb_ctx*, const uint8_t*, int rank, int world) { logf("comm_init %d %d\n", rank, world); return SKB_OK; }
int skb_comm_destroy(skb_ctx*) { return SKB_OK; }
int skb_comm_rank(const skb_ctx*) { return 0; }
int skb_comm_world(const skb_ctx*) { return 1; }
void skb_dist_range(uint64_t n, int rank, int world, uint64_t* begin, uint64_t* count) {
  const uint64_t per = (n + world - 1) / world;
  const uint64_t b = per * rank < n ? per * rank : n, e = per * (rank + 1) < n ? per * (rank + 1) : n;
  if (begin) *begin = b;
  if (count) *count = e - b;
}
int skb_comm_allgather_host(skb_ctx*, const void* send, void* recv, uint64_t bytes) { memcpy(recv, send, bytes); return SKB_OK; }
int skb_predict_stream_dist(skb_ctx*, skb_batch* b, uint64_t reads_total, uint32_t k, uint32_t s_query, uint64_t seed, uint32_t top,
                            uint32_t* oi, uint64_t* os) {
  logf("predict_stream k=%u s_query=%u seed=%llu top=%u reads_total=%llu reads=%u\n", k, s_query, (unsigned long long)seed, top,
       (unsigned long long)reads_total, b->groups);
  for (uint64_t r = 0; r < reads_total; ++r)
    for (uint32_t t = 0; t < top; ++t) { oi[(size_t)r * top + t] = t; os[(size_t)r * top + t] = 7; }
  return SKB_OK;
}
int skb_shared_counts(skb_ctx* c, const uint64_t*, const uint64_t* qoff, uint32_t Q, uint64_t* out) {
  logf("shared_counts Q=%u qhashes=%llu\n", Q, (unsigned long long)qoff[Q]);
  for (size_t i = 0; i < (size_t)c->n_rows * Q; ++i) out[i] = i;
  return SKB_OK;
}
int skb_rank_counts(skb_ctx*, const uint64_t*, uint32_t n, uint32_t top, uint32_t* oi, uint64_t* os) {
  logf("rank_counts n=%u top=%u\n", n, top);
  for (uint32_t t = 0; t < top; ++t) { oi[t] = t; os[t] = 3; }
  return SKB_OK;
}
}
