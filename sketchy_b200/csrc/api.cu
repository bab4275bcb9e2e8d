// api.cu — context, pinned 2-bit batch packer, pass orchestration and the exported C ABI (include/sketchy_b200.h).
// No CPU fallback lives here: every compute entry point launches the sm_100a kernels or fails.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <thread>

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: libnccl is loaded with dlopen when a communicator is asked for

#include "kernels.h"

#define SKB_VERSION_STR "sketchy_b200 0.1.0 (sm_100a)"

// ---------------------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------------------
// AVX2 block packer (pack_avx2.cpp); internal, exported only so that the CPU test-suite can call it directly
extern "C" uint64_t skb_pack_blocks_avx2(const uint8_t* s, uint64_t nbytes, uint32_t* codes, uint32_t* nmask);
extern "C" uint64_t skb_pack_blocks_avx512(const uint8_t* s, uint64_t nbytes, uint32_t* codes, uint32_t* nmask);

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  // grows keeping the first `keep` bytes
  cudaError_t ensure(size_t bytes, size_t keep) {
    if (bytes <= cap) return cudaSuccess;
    size_t ncap = std::max(bytes, cap * 2);
    void* np = nullptr;
    cudaError_t e = cudaHostAlloc(&np, ncap, cudaHostAllocDefault);
    if (e != cudaSuccess) return e;
    if (p && keep) std::memcpy(np, p, keep);
    if (p) cudaFreeHost(p);
    p = np; cap = ncap;
    return cudaSuccess;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

inline uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

// needletail normalize(false) folded into the 2-bit packer: 0..3 = ACGT code, 4 = breaks k-mers, 5 = removed
struct PackLut {
  uint8_t t[256];
  PackLut() {
    std::memset(t, 4, sizeof t);
    t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2;
    t['T'] = t['t'] = t['U'] = t['u'] = 3;
    t[' '] = t['\t'] = t['\r'] = t['\n'] = 5;
  }
};
const PackLut kPackLut;

// Two input bytes at a time: lut2[b0 | b1 << 8] = code0 | code1 << 2 when both bytes are ACGT/acgt/Uu, 0xFF otherwise
// (the rare chunk with anything else takes the byte-wise path).
struct PackLut2 {
  uint8_t t[65536];
  PackLut2() {
    for (int b0 = 0; b0 < 256; ++b0)
      for (int b1 = 0; b1 < 256; ++b1) {
        const uint8_t c0 = kPackLut.t[b0], c1 = kPackLut.t[b1];
        t[b0 | (b1 << 8)] = (c0 < 4 && c1 < 4) ? (uint8_t)(c0 | (c1 << 2)) : 0xFF;
      }
  }
};
const PackLut2 kPackLut2;

// Whole 32-byte blocks without a removed byte: two code words and one mask word each. Returns the bytes consumed (a
// multiple of 32); stops at the first block it cannot take. AVX2 version in pack_avx2.cpp, chosen at load time.

uint64_t pack_blocks_scalar(const uint8_t* s, uint64_t nbytes, uint32_t* codes, uint32_t* nmask) {
  uint64_t i = 0;
  for (; i + 32 <= nbytes; i += 32) {  // only blocks of 32 clean bases (anything else goes byte by byte)
    uint32_t w[2];
    for (int h = 0; h < 2; ++h) {
      uint32_t cw = 0;
      const uint8_t* p = s + i + 16 * h;
      for (int j = 0; j < 8; ++j) {
        const uint8_t v = kPackLut2.t[p[2 * j] | ((uint32_t)p[2 * j + 1] << 8)];
        if (v == 0xFF) return i;
        cw |= (uint32_t)v << (4 * j);
      }
      w[h] = cw;
    }
    codes[i >> 4] = w[0];
    codes[(i >> 4) + 1] = w[1];
    nmask[i >> 5] = 0;
  }
  return i;
}

typedef uint64_t (*pack_blocks_fn)(const uint8_t*, uint64_t, uint32_t*, uint32_t*);
pack_blocks_fn choose_pack_blocks() {
  if (const char* e = getenv("SKB_NO_AVX2")) { if (e[0] == '1') return pack_blocks_scalar; }
  if (!__builtin_cpu_supports("avx2")) return pack_blocks_scalar;
  const char* no512 = getenv("SKB_NO_AVX512");
  if (!(no512 && no512[0] == '1') && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
      __builtin_cpu_supports("avx512vbmi"))
    return skb_pack_blocks_avx512;  // 64 bases per step (1.2x the AVX2 path from memory, 1.6x from cache)
  return skb_pack_blocks_avx2;
}
const pack_blocks_fn kPackBlocks = choose_pack_blocks();

// Pack one record at base position P (multiple of 32); everything up to Pnext (multiple of 32) not covered by a
// valid base is marked invalid. The words touched belong to this record alone.
// Returns the number of bases the record keeps (its length without the removed bytes).
uint64_t pack_record(const uint8_t* s, uint64_t len, uint64_t P, uint64_t Pnext, uint32_t* codes, uint32_t* nmask) {
  uint64_t pos = P;
  uint32_t cw = 0, mw = 0;
  for (uint64_t i = 0; i < len; ++i) {
    // block path whenever the output is block-aligned (cw/mw are empty then) and 32 bytes are left
    if ((pos & 31) == 0 && i + 32 <= len) {
      const uint64_t n = kPackBlocks(s + i, len - i, codes + (pos >> 4), nmask + (pos >> 5));
      pos += n;
      i += n;
      if (i >= len) break;
    }
    const uint32_t v = kPackLut.t[s[i]];
    if (v == 5) continue;
    if (v < 4) cw |= v << (2 * (pos & 15));
    else mw |= 1u << (pos & 31);
    ++pos;
    if ((pos & 15) == 0) { codes[(pos >> 4) - 1] = cw; cw = 0; }
    if ((pos & 31) == 0) { nmask[(pos >> 5) - 1] = mw; mw = 0; }
  }
  const uint64_t kept = pos - P;
  // finish the current 32-base block as invalid, then whole invalid blocks
  while (pos < Pnext && (pos & 31)) {
    mw |= 1u << (pos & 31);
    ++pos;
    if ((pos & 15) == 0) { codes[(pos >> 4) - 1] = cw; cw = 0; }
    if ((pos & 31) == 0) { nmask[(pos >> 5) - 1] = mw; mw = 0; }
  }
  for (; pos < Pnext; pos += 32) {
    codes[pos >> 4] = 0; codes[(pos >> 4) + 1] = 0;
    nmask[pos >> 5] = 0xFFFFFFFFu;
  }
  return kept;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// context / batch
// ---------------------------------------------------------------------------------------------------------
// NCCL, bound at run time (skb_comm_init): a single-GPU caller never needs the library to be present, and inside a
// process that already carries NCCL (PyTorch) the same copy is used.
struct SkbNccl {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
      if (!lib) lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) { ok = false; err = std::string("libnccl lacks ") + n; } return p; };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
    Broadcast = reinterpret_cast<decltype(Broadcast)>(sym("ncclBroadcast"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) { lib = nullptr; }
    return ok;
  }
};
SkbNccl g_nccl;

struct ProfEvent { int id; cudaEvent_t a, b; };
// Default reads per pass. A pass costs the stream of the shard plus work that does not depend on the shard (table
// build, bounds, candidate lists: ~0.15 ms); more reads per pass always give more reads per second, at a lower
// fraction of the HBM roofline for the streaming kernel (more hits per streamed byte). The default keeps the kernel
// above 0.6 of the roofline where the stream dominates (shards of 2 GB and more: 4096 reads, measured 0.67 on 3.2 GB
// and 0.74 on 10 GB) and takes the largest pass on smaller shards, where the fixed work dominates.
#define SKB_DEFAULT_PASS_READS 4096u
#define SKB_SMALL_SHARD_BYTES (2ull << 30)
#define SKB_PASS_KEY_BUDGET (1ull << 18)  // query hashes per pass before the membership prefilter (which typically drops more than half of them; the 2^19-bit filter is designed for ~2^17 keys at three bits each)
constexpr uint32_t kAnchorGroups = 64;   // row groups of the dense ranking's anchor launch (every 64th read)
#define SKB_NSUMS 4
#define SKB_NTRACK 3
#define SKB_NTAB 3

struct skb_ctx {
  int device = 0;
  int num_sms = 148;
  int stream_ctas = 148;  // CTAs of the streaming kernel (one per SM it may take; the rest of the GPU is left to the pre-/post-pass kernels)
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // batch staging (H2D): lets skb_batch_stage overlap a predict running on `stream`
  std::string err;
  // reference shard
  DevBuf ref, row_start, row_len, cta_row, tile_cum, memb;
  uint32_t memb_log2 = 0;  // 0 = no membership prefilter
  uint64_t ref_len = 0;  // hashes in the shard (without alignment padding)
  uint32_t n_rows = 0, row_base = 0, uniform_len = 0, uniform_pitch = 0;
  uint64_t hmax = 0;
  bool has_ref = false;
  // Rotating state of the pass pipeline (see predict_device): the running sums after the last SKB_NSUMS passes, the
  // tracked-row sets the last SKB_NTRACK passes proposed, the query tables of the last SKB_NTAB passes.
  DevBuf sums[SKB_NSUMS];
  int sums_cur = 0;
  DevBuf tracked[SKB_NTRACK];  // [SKB_MAX_TRACKED] local rows + (at index SKB_MAX_TRACKED) their count
  int tracked_cur = 0;
  DevBuf tprefix, textra;
  cudaStream_t side = nullptr;  // pre-pass / post-pass kernels, overlapping the streaming kernel of the neighbouring passes
  cudaStream_t tabs = nullptr;  // query-table builds: table(i) is built next to the post-pass kernels of pass i-1
  cudaEvent_t ev_tab[SKB_NTAB] = {nullptr, nullptr, nullptr}, ev_post = nullptr;
  cudaEvent_t ev_pre[SKB_NTAB] = {nullptr, nullptr, nullptr}, ev_fused[2] = {nullptr, nullptr}, ev_join = nullptr;
  int tab_cur = 0;              // slot of the newest query table
  bool pass_proven = false;     // a full-size sparse pass has been checked and did not overflow: batch them
  int rank_mode = 0;            // skb_set_rank_mode
  uint32_t dense_left = 0;      // upcoming passes that are ranked densely (bounds too loose to be worth candidates)
  uint32_t dense_after_reset = 0;  // dense passes after the first one of a reset (measured: the bounds from the first pass's tracked rows hold, no second dense pass is needed; SKB_DENSE_AFTER_RESET: experiments)
  DevBuf dense, part_idx, part_sum, anch_part_idx, anch_part_sum, anch_idx, anch_sum, piece_idx, piece_sum, piece_row;
  bool pipeline = false;        // steady-state passes take their bounds from two passes back (pre-pass work enqueued next to the stream)
  // multi-GPU: one process per GPU, reference rows sharded by contiguous range, NCCL over NVLink for the exchange steps
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  DevBuf qn_all, gath_idx, gath_sum, hmax_all;
  uint64_t tau_all = 0;         // largest reference hash of any rank's shard
  bool tau_all_valid = false;   // reset by an upload and by skb_comm_init
  uint32_t tracked_top = 0;  // 0 = invalid
  // per-group scratch (hash/select)
  DevBuf g_tau, g_cap, g_base, g_cnt, g_kmers, g_active, g_outn, g_status, g_tiles, cand_pool;
  DevBuf sk_hashes, sk_counts;
  // predict scratch
  DevBuf q_off, qh, qread, counts, lb_sum[SKB_NTAB], lb_idx[SKB_NTAB], cand[2], cand_cnt[2], ivl[2], seg_hdr[2], seg_words[2], scal;
  DevBuf t_slots[SKB_NTAB], t_fill[SKB_NTAB], t_reads[SKB_NTAB], t_slot[SKB_NTAB], t_bloom[SKB_NTAB];
  uint32_t t_cap = 0, t_maxkeys = 0;
  uint32_t t_built[SKB_NTAB] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};  // keys of each table's last build (UINT32_MAX: never built, clear all)
  DevBuf out_idx, out_sum, misc;
  uint32_t pass_user = 0;                      // skb_set_pass_reads (0 = automatic)
  uint32_t pass_max = SKB_DEFAULT_PASS_READS;  // reads per pass in effect (choose_pass_max)
  uint32_t cand_cap = 0;
  uint64_t cand_budget = SKB_CAND_BUDGET;  // candidate records per pass over all reads (SKB_CAND_BUDGET: tests)
  bool trace_passes = false;               // SKB_TRACE_PASSES: one line per checkpoint on stderr
  // stats / profiling
  bool prof_on = false;
  std::vector<ProfEvent> pending;
  std::vector<cudaEvent_t> ev_pool;  // timing events ready for reuse
  double prof_ms[SKB_K_COUNT] = {0};
  uint64_t prof_n[SKB_K_COUNT] = {0};
  uint64_t launches = 0;
  uint64_t st_ref_bytes = 0, st_passes = 0, st_qhashes = 0, st_cands = 0, st_members = 0;
  uint32_t* h_scal = nullptr;  // pinned, 64 bytes
  PinBuf h_outn, h_off;        // page-locked staging of per-read arrays (query-hash counts down, list offsets up)
};

struct skb_batch {
  skb_ctx* ctx = nullptr;
  PinBuf codes, nmask;
  uint64_t cur = 0;  // packed length in bases (multiple of 32)
  std::vector<uint64_t> rec_pos, rec_len;
  std::vector<uint64_t> g_first, g_end;  // chunk ranges
  std::vector<uint64_t> g_raw, g_packed; // raw bytes, packed bases
  uint64_t total_raw = 0;
  int base_count = SKB_BASES_RAW;  // what total_bases counts (skb_batch_set_base_count)
  // device mirror
  bool staged = false;
  DevBuf d_codes, d_nmask, d_seg_group, d_seg_chunk0, d_seg_n, d_g_len;
  PinBuf h_seg_group, h_seg_chunk0, h_seg_n, h_g_len;  // page-locked sources of the small arrays (true async copies)
  cudaEvent_t ev_staged = nullptr;  // recorded on the copy stream behind the batch's H2D copies
  uint32_t nseg = 0;
  uint64_t max_group_len = 0;  // longest group (bases), known once the batch is staged
};

namespace {

void choose_pass_max(skb_ctx* c) {
  if (c->pass_user) c->pass_max = std::min<uint32_t>(c->pass_user, skb_fused_max_reads(1));
  else c->pass_max = (c->has_ref && c->ref_len * 8 < SKB_SMALL_SHARD_BYTES) ? skb_fused_max_reads(1) : SKB_DEFAULT_PASS_READS;
}

int fail(skb_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    c->err = buf;
  }
  return code;
}

#define CU(c, call)                                                                                  \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail((c), e__ == cudaErrorMemoryAllocation ? SKB_ERR_OOM : SKB_ERR_CUDA, "%s: %s", #call, \
                  cudaGetErrorString(e__));                                                          \
  } while (0)

struct ProfScope {
  skb_ctx* c; int id; cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  ProfScope(skb_ctx* ctx, int kid, int n_launches, cudaStream_t on = nullptr) : c(ctx), id(kid), st(on ? on : ctx->stream) {
    c->launches += n_launches;
    c->prof_n[id] += n_launches;
    if (c->prof_on) {  // timing events come from a pool: creating a pair per scope costs more host time than the launches it brackets
      for (cudaEvent_t* e : {&a, &b}) {
        if (!c->ev_pool.empty()) { *e = c->ev_pool.back(); c->ev_pool.pop_back(); }
        else cudaEventCreate(e);
      }
      cudaEventRecord(a, st);
    }
  }
  ~ProfScope() {
    if (c->prof_on) {
      cudaEventRecord(b, st);
      c->pending.push_back({id, a, b});
    }
  }
};

void prof_resolve(skb_ctx* c) {
  if (c->pending.empty()) return;
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->side);
  if (c->tabs) cudaStreamSynchronize(c->tabs);
  for (auto& e : c->pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) c->prof_ms[e.id] += ms;
    c->ev_pool.push_back(e.a); c->ev_pool.push_back(e.b);
  }
  c->pending.clear();
}

int check_launch(skb_ctx* c, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(c, SKB_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return SKB_OK;
}

// Enqueue the batch's H2D copies on the context's copy stream and record the batch's event behind them. Nothing here
// waits: the kernels that read the batch wait for the event on their own stream (use_batch), so a caller's packing
// thread goes on to its next batch while the DMA runs. The page-locked sources stay untouched until the batch is
// cleared (skb_batch_clear waits for the event).
int stage_batch(skb_batch* b) {
  skb_ctx* c = b->ctx;
  if (b->staged) return SKB_OK;
  const uint64_t ncode = b->cur / 16 + 4, nmask = b->cur / 32 + 2;
  CU(c, b->codes.ensure(ncode * 4, (b->cur / 16) * 4));
  CU(c, b->nmask.ensure(nmask * 4, (b->cur / 32) * 4));
  for (uint64_t i = b->cur / 16; i < ncode; ++i) b->codes.as<uint32_t>()[i] = 0;
  for (uint64_t i = b->cur / 32; i < nmask; ++i) b->nmask.as<uint32_t>()[i] = 0xFFFFFFFFu;
  // segments: <= 32 chunks of one group each
  const size_t G = b->g_first.size();
  uint64_t nseg = 0;
  b->max_group_len = 0;
  for (size_t g = 0; g < G; ++g) {
    nseg += (b->g_end[g] - b->g_first[g] + 31) / 32;
    b->max_group_len = std::max(b->max_group_len, b->g_packed[g]);
  }
  if (nseg >= 0xFFFFFFF0ull) return fail(c, SKB_ERR_INVALID_ARG, "batch too large (segments)");
  CU(c, b->h_seg_group.ensure(std::max<size_t>(4, nseg * 4), 0));
  CU(c, b->h_seg_chunk0.ensure(std::max<size_t>(4, nseg * 4), 0));
  CU(c, b->h_seg_n.ensure(std::max<size_t>(4, nseg), 0));
  CU(c, b->h_g_len.ensure(std::max<size_t>(8, G * 8), 0));
  uint32_t* sg = b->h_seg_group.as<uint32_t>();
  uint32_t* sc = b->h_seg_chunk0.as<uint32_t>();
  uint8_t* sn = b->h_seg_n.as<uint8_t>();
  uint64_t at = 0;
  for (size_t g = 0; g < G; ++g) {
    for (uint64_t ch = b->g_first[g]; ch < b->g_end[g]; ch += 32, ++at) {
      sg[at] = (uint32_t)g;
      sc[at] = (uint32_t)ch;
      sn[at] = (uint8_t)std::min<uint64_t>(32, b->g_end[g] - ch);
    }
  }
  if (G) std::memcpy(b->h_g_len.p, b->g_packed.data(), G * 8);
  b->nseg = (uint32_t)nseg;
  CU(c, b->d_codes.ensure(ncode * 4));
  CU(c, b->d_nmask.ensure(nmask * 4));
  CU(c, b->d_seg_group.ensure(std::max<size_t>(4, nseg * 4)));
  CU(c, b->d_seg_chunk0.ensure(std::max<size_t>(4, nseg * 4)));
  CU(c, b->d_seg_n.ensure(std::max<size_t>(4, nseg)));
  CU(c, b->d_g_len.ensure(std::max<size_t>(8, G * 8)));
  if (!b->ev_staged) CU(c, cudaEventCreateWithFlags(&b->ev_staged, cudaEventDisableTiming));
  CU(c, cudaMemcpyAsync(b->d_codes.p, b->codes.p, ncode * 4, cudaMemcpyHostToDevice, c->copy_stream));
  CU(c, cudaMemcpyAsync(b->d_nmask.p, b->nmask.p, nmask * 4, cudaMemcpyHostToDevice, c->copy_stream));
  if (G) CU(c, cudaMemcpyAsync(b->d_g_len.p, b->h_g_len.p, G * 8, cudaMemcpyHostToDevice, c->copy_stream));
  if (nseg) {
    CU(c, cudaMemcpyAsync(b->d_seg_group.p, sg, nseg * 4, cudaMemcpyHostToDevice, c->copy_stream));
    CU(c, cudaMemcpyAsync(b->d_seg_chunk0.p, sc, nseg * 4, cudaMemcpyHostToDevice, c->copy_stream));
    CU(c, cudaMemcpyAsync(b->d_seg_n.p, sn, nseg, cudaMemcpyHostToDevice, c->copy_stream));
  }
  CU(c, cudaEventRecord(b->ev_staged, c->copy_stream));
  b->staged = true;
  return SKB_OK;
}

// stage the batch if it is not staged yet and make the context's main stream wait for its copies
int use_batch(skb_batch* b) {
  if (int rc = stage_batch(b)) return rc;
  skb_ctx* c = b->ctx;
  CU(c, cudaStreamWaitEvent(c->stream, b->ev_staged, 0));
  return SKB_OK;
}

SkbPackedView view_of(const skb_batch* b) {
  SkbPackedView v;
  v.codes = b->d_codes.as<uint32_t>();
  v.nmask = b->d_nmask.as<uint32_t>();
  v.seg_group = b->d_seg_group.as<uint32_t>();
  v.seg_chunk0 = b->d_seg_chunk0.as<uint32_t>();
  v.seg_n = b->d_seg_n.as<uint8_t>();
  v.nseg = b->nseg;
  return v;
}

// ---- hash + select with the exactness loop ---------------------------------------------------------------
// mode sketch : tau per group from the group's size; groups that end with < s distinct hashes under a finite tau
//               are redone with a larger tau; out = [G*s] (+counts)
// mode query  : tau = hmax for every group, no underfill check; output stays in the candidate pool
struct SelectPlan {
  std::vector<uint64_t> tau, base;
  std::vector<uint32_t> cap;
};

int run_hash_select(skb_ctx* c, skb_batch* b, uint32_t k, uint32_t s, uint64_t seed, bool query_mode,
                    uint64_t query_tau, uint64_t* d_out_hashes, uint32_t* d_out_counts,
                    std::vector<uint32_t>& h_out_n, std::vector<uint64_t>& h_kmers, SelectPlan& plan) {
  const uint32_t G = (uint32_t)b->g_first.size();
  plan.tau.assign(G, 0); plan.base.assign(G, 0); plan.cap.assign(G, 0);
  h_out_n.assign(G, 0); h_kmers.assign(G, 0);
  if (G == 0) return SKB_OK;
  const double two64 = 18446744073709551616.0;
  if (query_mode) {
    // Fast path: one threshold for every group, so the plan (capacities, places in the pool, scratch reset) is made on
    // the device and the host reads back only the per-read counts and one "some group needs another round" flag,
    // through page-locked memory. A call of 100,000 reads otherwise spends a millisecond in the loop below and another in
    // pageable copies of its arrays. A group whose candidates overflow its capacity (rare: more than four times the
    // expected number of hashes under the threshold) sends the whole call through the general loop below.
    const double frac = ((double)query_tau + 1.0) / two64;
    const uint64_t pool_bound = (uint64_t)G * 96 + (uint64_t)(8.0 * frac * (double)b->total_raw) + 64;
    CU(c, c->g_tau.ensure(G * 8)); CU(c, c->g_base.ensure(G * 8)); CU(c, c->g_cap.ensure(G * 4));
    CU(c, c->g_cnt.ensure(G * 4)); CU(c, c->g_kmers.ensure(G * 8)); CU(c, c->g_active.ensure(G));
    CU(c, c->g_outn.ensure(G * 4)); CU(c, c->g_status.ensure(G * 4));
    CU(c, c->cand_pool.ensure(pool_bound * 8));
    CU(c, c->h_outn.ensure((size_t)G * 4 + 64, 0));
    uint32_t* d_flag = c->scal.as<uint32_t>() + 40;
    SkbPlanArgs pa{};
    pa.n_groups = G; pa.g_len = b->d_g_len.as<uint64_t>(); pa.tau = query_tau; pa.frac = frac;
    pa.tau_out = c->g_tau.as<uint64_t>(); pa.base = c->g_base.as<uint64_t>(); pa.cap = c->g_cap.as<uint32_t>();
    pa.cnt = c->g_cnt.as<uint32_t>(); pa.kmers = c->g_kmers.as<unsigned long long>(); pa.active = c->g_active.as<uint8_t>();
    pa.status = c->g_status.as<uint32_t>(); pa.any_bad = d_flag;
    CU(c, c->g_tiles.ensure(((size_t)G / 1024 + 1) * 8));
    pa.tile_tot = c->g_tiles.as<unsigned long long>();
    { ProfScope ps(c, SKB_K_SELECT, 2); skb_launch_plan_query(pa, c->stream); }
    SkbHashArgs ha{};
    ha.pv = view_of(b); ha.k = k; ha.seed = seed;
    ha.tau = pa.tau_out; ha.active = pa.active;
    ha.cand = c->cand_pool.as<uint64_t>(); ha.cand_base = pa.base; ha.cand_cap = pa.cap;
    ha.cand_cnt = pa.cnt; ha.kmers = pa.kmers;
    { ProfScope ps(c, SKB_K_HASH, 1); skb_launch_hash(ha, c->stream); }
    if (int rc = check_launch(c, "hash")) return rc;
    SkbSelectArgs sa{};
    sa.n_groups = G; sa.cand = ha.cand; sa.cand_base = ha.cand_base; sa.cand_cap = ha.cand_cap;
    sa.cand_cnt = ha.cand_cnt; sa.active = ha.active; sa.tau = ha.tau; sa.s = s;
    sa.check_underfill = 0; sa.out_hashes = ha.cand; sa.out_off = ha.cand_base; sa.out_counts = nullptr;
    sa.out_n = c->g_outn.as<uint32_t>(); sa.status = pa.status; sa.any_bad = d_flag;
    {
      const double n = (double)std::max<uint64_t>(b->max_group_len, 1);
      const double est = std::min(4.0 * (n * frac) + 16.0, n) * 1.001;  // (the device plans in single precision)
      const uint32_t want = (uint32_t)std::min<uint64_t>(16384, skb_next_pow2((uint64_t)std::max(32.0, est)));
      if (want <= 256) { sa.threads = 32; sa.smem_elems = 256; }
      else if (want <= 2048) { sa.threads = 256; sa.smem_elems = want; }
      else { sa.threads = 512; sa.smem_elems = want; }
    }
    { ProfScope ps(c, SKB_K_SELECT, 1); skb_launch_select(sa, c->stream); }
    if (int rc = check_launch(c, "select")) return rc;
    uint32_t* h = c->h_outn.as<uint32_t>();
    CU(c, cudaMemcpyAsync(h, c->g_outn.p, (size_t)G * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(h + G, d_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (h[G] == 0) {
      std::memcpy(h_out_n.data(), h, (size_t)G * 4);
      return SKB_OK;
    }
  }
  uint64_t pool = 0;
  for (uint32_t g = 0; g < G; ++g) {
    const double n = (double)std::max<uint64_t>(b->g_packed[g], 1);
    double est;
    if (query_mode) {
      plan.tau[g] = query_tau;
      est = n * (((double)query_tau + 1.0) / two64);
      est = 4.0 * est + 16.0;
    } else {
      const double m = 1.25 * ((double)s + 8.0 * std::sqrt((double)s) + 32.0);
      if (n <= 2.0 * m) { plan.tau[g] = SKB_EMPTY_KEY; est = n; }
      else { plan.tau[g] = (uint64_t)std::min(two64 * (m / n), 18446744073709549568.0); est = 2.0 * m; }
    }
    est = std::min(est, n);
    plan.cap[g] = (uint32_t)skb_next_pow2((uint64_t)std::max(32.0, est));
    plan.base[g] = pool;
    pool += plan.cap[g];
  }
  CU(c, c->g_tau.ensure(G * 8)); CU(c, c->g_base.ensure(G * 8)); CU(c, c->g_cap.ensure(G * 4));
  CU(c, c->g_cnt.ensure(G * 4)); CU(c, c->g_kmers.ensure(G * 8)); CU(c, c->g_active.ensure(G));
  CU(c, c->g_outn.ensure(G * 4)); CU(c, c->g_status.ensure(G * 4));
  CU(c, c->cand_pool.ensure(std::max<uint64_t>(pool, 1) * 8));
  CU(c, cudaMemsetAsync(c->g_cnt.p, 0, G * 4, c->stream));
  CU(c, cudaMemsetAsync(c->g_kmers.p, 0, G * 8, c->stream));
  CU(c, cudaMemsetAsync(c->g_active.p, 1, G, c->stream));
  CU(c, cudaMemsetAsync(c->g_status.p, 0, G * 4, c->stream));

  std::vector<uint32_t> status(G), cnt(G);
  std::vector<uint8_t> active(G, 1);
  uint64_t pool_used = pool;
  uint32_t max_cap = 32;
  for (uint32_t g = 0; g < G; ++g) max_cap = std::max(max_cap, plan.cap[g]);
  for (int iter = 0; iter < 12; ++iter) {
    CU(c, cudaMemcpyAsync(c->g_tau.p, plan.tau.data(), G * 8, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->g_base.p, plan.base.data(), G * 8, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->g_cap.p, plan.cap.data(), G * 4, cudaMemcpyHostToDevice, c->stream));
    uint64_t* cur_pool = c->cand_pool.as<uint64_t>();
    SkbHashArgs ha{};
    ha.pv = view_of(b); ha.k = k; ha.seed = seed;
    ha.tau = c->g_tau.as<uint64_t>(); ha.active = c->g_active.as<uint8_t>();
    ha.cand = cur_pool; ha.cand_base = c->g_base.as<uint64_t>(); ha.cand_cap = c->g_cap.as<uint32_t>();
    ha.cand_cnt = c->g_cnt.as<uint32_t>(); ha.kmers = c->g_kmers.as<unsigned long long>();
    { ProfScope ps(c, SKB_K_HASH, 1); skb_launch_hash(ha, c->stream); }
    if (int rc = check_launch(c, "hash")) return rc;
    SkbSelectArgs sa{};
    sa.n_groups = G; sa.cand = cur_pool; sa.cand_base = ha.cand_base; sa.cand_cap = ha.cand_cap;
    sa.cand_cnt = ha.cand_cnt; sa.active = ha.active; sa.tau = ha.tau; sa.s = s;
    sa.check_underfill = query_mode ? 0 : 1;
    sa.out_hashes = query_mode ? cur_pool : d_out_hashes;
    sa.out_off = query_mode ? ha.cand_base : nullptr;
    sa.out_counts = query_mode ? nullptr : d_out_counts;
    sa.out_n = c->g_outn.as<uint32_t>(); sa.status = c->g_status.as<uint32_t>();
    // the sort runs in shared memory when a group's candidates fit; bigger groups sort in place in global memory
    const uint32_t want = (uint32_t)std::min<uint64_t>(16384, skb_next_pow2(max_cap));
    if (want <= 256) { sa.threads = 32; sa.smem_elems = 256; }
    else if (want <= 2048) { sa.threads = 256; sa.smem_elems = want; }
    else { sa.threads = 512; sa.smem_elems = want; }
    { ProfScope ps(c, SKB_K_SELECT, 1); skb_launch_select(sa, c->stream); }
    if (int rc = check_launch(c, "select")) return rc;
    CU(c, cudaMemcpyAsync(status.data(), c->g_status.p, G * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(cnt.data(), c->g_cnt.p, G * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    std::vector<uint32_t> failed;
    for (uint32_t g = 0; g < G; ++g)
      if (active[g] && status[g] != SKB_ST_OK) failed.push_back(g);
    if (failed.empty()) break;
    if (iter == 11)
      return fail(c, SKB_ERR_INTERNAL, "bottom-s selection did not converge for %zu groups", failed.size());
    // redo only the failed groups with a corrected threshold / capacity, in fresh space at the end of the pool
    std::vector<uint32_t> outn(G);
    CU(c, cudaMemcpy(outn.data(), c->g_outn.p, G * 4, cudaMemcpyDeviceToHost));
    std::fill(active.begin(), active.end(), 0);
    const uint64_t pool_before = pool_used;
    max_cap = 32;
    for (uint32_t g : failed) {
      active[g] = 1;
      double expect = (double)cnt[g];
      if (status[g] == SKB_ST_UNDERFILL) {
        const double d = (double)outn[g];
        const double ratio = d < 1.0 ? 64.0 : std::max(2.0, 1.5 * (double)s / d);
        const double nt = (double)plan.tau[g] * ratio;
        plan.tau[g] = nt >= 18446744073709549568.0 ? SKB_EMPTY_KEY : (uint64_t)nt;
        expect = plan.tau[g] == SKB_EMPTY_KEY ? (double)b->g_packed[g] : expect * ratio * 1.5 + 64.0;
      }
      expect = std::min(expect, (double)std::max<uint64_t>(b->g_packed[g], 1));
      plan.cap[g] = (uint32_t)skb_next_pow2((uint64_t)std::max(32.0, expect));
      max_cap = std::max(max_cap, plan.cap[g]);
      plan.base[g] = pool_used;
      pool_used += plan.cap[g];
    }
    if (pool_used * 8 > c->cand_pool.cap) {  // grow, keeping the finished groups' data
      DevBuf nb;
      CU(c, nb.ensure(pool_used * 8));
      CU(c, cudaMemcpy(nb.p, c->cand_pool.p, pool_before * 8, cudaMemcpyDeviceToDevice));
      c->cand_pool.release();
      c->cand_pool = nb;
    }
    CU(c, cudaMemcpyAsync(c->g_active.p, active.data(), G, cudaMemcpyHostToDevice, c->stream));
    std::vector<uint32_t> z(G);
    std::vector<uint64_t> hk(G);
    CU(c, cudaMemcpy(z.data(), c->g_cnt.p, G * 4, cudaMemcpyDeviceToHost));
    CU(c, cudaMemcpy(hk.data(), c->g_kmers.p, G * 8, cudaMemcpyDeviceToHost));
    for (uint32_t g : failed) { z[g] = 0; hk[g] = 0; }
    CU(c, cudaMemcpy(c->g_cnt.p, z.data(), G * 4, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->g_kmers.p, hk.data(), G * 8, cudaMemcpyHostToDevice));
  }
  CU(c, cudaMemcpyAsync(h_out_n.data(), c->g_outn.p, G * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(h_kmers.data(), c->g_kmers.p, G * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SKB_OK;
}

int ensure_table(skb_ctx* c, uint32_t max_keys) {
  if (c->t_maxkeys && max_keys <= c->t_maxkeys) return SKB_OK;
  const uint32_t mk = (uint32_t)skb_next_pow2(std::max<uint32_t>(max_keys, 1u << 16));
  const uint32_t cap = mk * 8;  // load factor <= 0.125: a lookup is almost always one 16-byte load
  for (int i = 0; i < SKB_NTAB; ++i) {
    CU(c, c->t_slots[i].ensure(((size_t)cap + 1) * sizeof(SkbSlot)));
    CU(c, c->t_fill[i].ensure(((size_t)cap + 1) * 4));
    CU(c, c->t_reads[i].ensure((size_t)mk * 4));
    CU(c, c->t_slot[i].ensure((size_t)mk * 4));
    CU(c, c->t_bloom[i].ensure((size_t)SKB_BLOOM_WORDS * 4));
  }
  c->t_cap = cap;
  c->t_maxkeys = mk;
  for (int i = 0; i < SKB_NTAB; ++i) c->t_built[i] = 0xFFFFFFFFu;  // new tables: the first build clears every slot
  return SKB_OK;
}

SkbTable table_of(skb_ctx* c, int slot) {
  SkbTable t;
  t.slots = c->t_slots[slot].as<SkbSlot>(); t.fill = c->t_fill[slot].as<uint32_t>(); t.reads = c->t_reads[slot].as<uint32_t>();
  t.slot_of = c->t_slot[slot].as<uint32_t>(); t.bloom = c->t_bloom[slot].as<uint32_t>();
  t.cursor = c->scal.as<uint32_t>() + 4 + slot;
  t.cap = c->t_cap;
  t.memb = c->memb_log2 ? c->memb.as<uint32_t>() : nullptr;
  t.memb_log2 = c->memb_log2;
  t.memb_kept = reinterpret_cast<unsigned long long*>(c->scal.as<uint8_t>() + 72);
  uint32_t l = 0;
  while ((1u << l) < t.cap) ++l;
  t.log2cap = l;
  return t;
}

SkbRefView ref_view(const skb_ctx* c) {
  SkbRefView v;
  v.ref = c->ref.as<uint64_t>(); v.row_start = c->row_start.as<uint64_t>(); v.row_len = c->row_len.as<uint32_t>();
  v.n_rows = c->n_rows;
  v.uniform_len = c->uniform_len; v.uniform_pitch = c->uniform_pitch;
  return v;
}

int upload_ref_common(skb_ctx* c, const uint64_t* hashes, bool on_device, const uint64_t* off, uint32_t n_rows,
                      uint32_t global_row_base) {
  if (!off && n_rows) return fail(c, SKB_ERR_INVALID_ARG, "off is null");
  const uint64_t len = n_rows ? off[n_rows] : 0;
  if (n_rows && off[0] != 0) return fail(c, SKB_ERR_INVALID_ARG, "off[0] must be 0");
  for (uint32_t i = 0; i < n_rows; ++i) {
    if (off[i + 1] < off[i]) return fail(c, SKB_ERR_INVALID_ARG, "row offsets must be non-decreasing");
    if (off[i + 1] - off[i] > 0xFFFFFFF0ull) return fail(c, SKB_ERR_INVALID_ARG, "row %u too long", i);
  }
  if (len && !hashes) return fail(c, SKB_ERR_INVALID_ARG, "hashes is null");
  c->has_ref = false;
  // device layout: every row starts on an even element (16-byte aligned) so it can be bulk-copied
  std::vector<uint64_t> start(std::max<uint32_t>(n_rows, 1));
  std::vector<uint32_t> rlen(std::max<uint32_t>(n_rows, 1));
  bool same = true;
  uint64_t pos = 0;
  for (uint32_t i = 0; i < n_rows; ++i) {
    start[i] = pos;
    rlen[i] = (uint32_t)(off[i + 1] - off[i]);
    same = same && (pos == off[i]);
    pos += round_up(rlen[i], 2);
  }
  const uint64_t padded = round_up(pos + 2, 2048);
  CU(c, c->ref.ensure(padded * 8));
  CU(c, c->row_start.ensure(std::max<size_t>(8, (size_t)n_rows * 8)));
  CU(c, c->row_len.ensure(std::max<size_t>(4, (size_t)n_rows * 4)));
  CU(c, cudaMemsetAsync(c->ref.p, 0xFF, padded * 8, c->stream));
  if (n_rows) {
    CU(c, cudaMemcpyAsync(c->row_start.p, start.data(), (size_t)n_rows * 8, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->row_len.p, rlen.data(), (size_t)n_rows * 4, cudaMemcpyHostToDevice, c->stream));
  }
  if (len) {
    if (same) {
      CU(c, cudaMemcpyAsync(c->ref.p, hashes, len * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                            c->stream));
    } else {
      DevBuf tmp, doff;
      const uint64_t* src = hashes;
      if (!on_device) {
        CU(c, tmp.ensure(len * 8));
        CU(c, cudaMemcpyAsync(tmp.p, hashes, len * 8, cudaMemcpyHostToDevice, c->stream));
        src = tmp.as<uint64_t>();
      }
      cudaError_t e = doff.ensure(((size_t)n_rows + 1) * 8);
      if (e != cudaSuccess) { tmp.release(); return fail(c, SKB_ERR_OOM, "relayout: %s", cudaGetErrorString(e)); }
      e = cudaMemcpyAsync(doff.p, off, ((size_t)n_rows + 1) * 8, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) {
        ProfScope ps(c, SKB_K_MISC, 1);
        skb_launch_relayout(src, doff.as<uint64_t>(), c->ref.as<uint64_t>(), c->row_start.as<uint64_t>(), n_rows, c->stream);
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      tmp.release(); doff.release();
      if (e != cudaSuccess) return fail(c, SKB_ERR_CUDA, "relayout: %s", cudaGetErrorString(e));
    }
  }
  c->n_rows = n_rows;
  {
    uint32_t ul = n_rows ? rlen[0] : 0;
    for (uint32_t i = 0; i < n_rows && ul; ++i)
      if (rlen[i] != ul) ul = 0;
    c->uniform_len = ul;
    c->uniform_pitch = (uint32_t)round_up(ul, 2);
  }
  CU(c, c->scal.ensure(256));
  CU(c, cudaMemsetAsync(c->scal.p, 0, 256, c->stream));
  uint32_t* d_bad = c->scal.as<uint32_t>();
  unsigned long long* d_hmax = reinterpret_cast<unsigned long long*>(c->scal.as<uint8_t>() + 8);
  { ProfScope ps(c, SKB_K_MISC, 1); skb_launch_ref_check(ref_view(c), d_bad, d_hmax, c->stream); }
  if (int rc = check_launch(c, "ref_check")) return rc;
  uint32_t bad = 0;
  unsigned long long hmax = 0;
  CU(c, cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(&hmax, d_hmax, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (bad) return fail(c, SKB_ERR_REF_NOT_SORTED, "a reference row is not strictly increasing");
  c->ref_len = len; c->row_base = global_row_base; c->hmax = hmax;
  // membership filter of the shard: 8 bits per stored hash, 128 KB .. 2 GB (SKB_NO_PREFILTER=1 turns it off: experiments)
  {
    const char* off = getenv("SKB_NO_PREFILTER");
    c->memb_log2 = 0;
    if (!(off && off[0] == '1') && len > 0) {
      uint32_t l2 = 15;
      while (l2 < 29 && (32ull << l2) < len * 8ull) ++l2;
      CU(c, c->memb.ensure((size_t)4 << l2));
      CU(c, cudaMemsetAsync(c->memb.p, 0, (size_t)4 << l2, c->stream));
      { ProfScope ps(c, SKB_K_MISC, 1); skb_launch_memb_build(ref_view(c), c->memb.as<uint32_t>(), l2, c->stream); }
      if (int rc = check_launch(c, "memb_build")) return rc;
      c->memb_log2 = l2;
    }
  }
  // contiguous row ranges per CTA, balanced by ring tiles (a row costs at least one unit: its rank work)
  {
    const uint32_t G = (uint32_t)c->stream_ctas, tile = skb_fused_tile();
    std::vector<uint64_t> cum(n_rows + 1, 0);
    for (uint32_t i = 0; i < n_rows; ++i) cum[i + 1] = cum[i] + std::max<uint64_t>(1, (rlen[i] + tile - 1) / tile);
    std::vector<uint32_t> cta(G + 1, n_rows);
    cta[0] = 0;
    for (uint32_t g = 1; g < G; ++g) {
      const uint64_t target = cum[n_rows] * g / G;
      cta[g] = (uint32_t)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
      if (cta[g] < cta[g - 1]) cta[g] = cta[g - 1];
      if (cta[g] > n_rows) cta[g] = n_rows;
    }
    CU(c, c->cta_row.ensure((G + 1) * 4));
    CU(c, cudaMemcpy(c->cta_row.p, cta.data(), (G + 1) * 4, cudaMemcpyHostToDevice));
    // sub-tiles before each row: how a CTA of the streaming kernel maps a claimed sub-tile number to its row
    if (cum[n_rows] > 0xFFFFFFF0ull) return fail(c, SKB_ERR_INVALID_ARG, "reference shard too large (sub-tile count)");
    std::vector<uint32_t> cum32(cum.begin(), cum.end());
    CU(c, c->tile_cum.ensure(((size_t)n_rows + 1) * 4));
    CU(c, cudaMemcpy(c->tile_cum.p, cum32.data(), ((size_t)n_rows + 1) * 4, cudaMemcpyHostToDevice));
  }
  for (int i = 0; i < SKB_NSUMS; ++i) CU(c, c->sums[i].ensure(std::max<size_t>(8, (size_t)n_rows * 8)));
  CU(c, cudaMemsetAsync(c->sums[0].p, 0, std::max<size_t>(8, (size_t)n_rows * 8), c->stream));
  c->sums_cur = 0;
  c->tracked_top = 0;
  c->pass_proven = false;
  c->dense_left = c->dense_after_reset;  // passes after the first (always dense) one that are ranked densely too
  for (int i = 0; i < SKB_NTRACK; ++i) CU(c, c->tracked[i].ensure((SKB_MAX_TRACKED + 1) * 4));
  c->tracked_cur = 0;
  CU(c, cudaStreamSynchronize(c->stream));
  c->has_ref = true;
  c->tau_all_valid = false;
  choose_pass_max(c);
  return SKB_OK;
}

// ---- per-read query sets -----------------------------------------------------------------------------------
// The reads of a call as the pass loop sees them: R "pass reads" with their query hashes (<= tau, distinct, ascending,
// at most s_query) flat in c->qh / c->qread. A read that keeps more than 65535 query hashes (the range of the u16
// per-(row, read) counters) is cut into consecutive pass reads: the running sums are cumulative, so the ranking after
// its last piece is the read's ranking and the pieces before it are simply not reported (`last_piece`).
struct QuerySet {
  uint32_t R = 0;                    // pass reads
  std::vector<uint32_t> qn;          // [R] query hashes per pass read
  std::vector<uint64_t> q_off;       // [R + 1]
  std::vector<uint32_t> out_row;     // [R] caller's read (row of the output arrays) of a last piece, UINT32_MAX otherwise; empty = identity
};
constexpr uint32_t kMaxPieceKeys = 65535;

// qn_reads[i] = query hashes of caller read i (already on the device, compacted read after read in c->qh)
// offsets_on_device: c->q_off already holds the offsets of the unsplit reads (the caller uploaded them)
int finish_query_set(skb_ctx* c, const std::vector<uint32_t>& qn_reads, QuerySet& qs, bool offsets_on_device = false) {
  const uint32_t n = (uint32_t)qn_reads.size();
  bool split = false;
  for (uint32_t v : qn_reads) split = split || v > kMaxPieceKeys;
  if (!split) {
    qs.R = n; qs.qn = qn_reads;
  } else {
    for (uint32_t i = 0; i < n; ++i) {
      uint32_t left = qn_reads[i];
      do {
        const uint32_t take = std::min(left, kMaxPieceKeys);
        qs.qn.push_back(take);
        left -= take;
        qs.out_row.push_back(left == 0 ? i : 0xFFFFFFFFu);
      } while (left > 0);
    }
    qs.R = (uint32_t)qs.qn.size();
  }
  qs.q_off.assign((size_t)qs.R + 1, 0);
  for (uint32_t r = 0; r < qs.R; ++r) qs.q_off[r + 1] = qs.q_off[r] + qs.qn[r];
  const uint64_t QN = qs.q_off[qs.R];
  c->st_qhashes = QN;
  CU(c, c->qread.ensure(std::max<uint64_t>(QN, 1) * 4));
  if (split || !offsets_on_device) {
    CU(c, c->q_off.ensure(((size_t)qs.R + 1) * 8));
    CU(c, cudaStreamSynchronize(c->stream));  // (an earlier copy out of the staging buffer may still be in flight)
    CU(c, c->h_off.ensure(((size_t)qs.R + 1) * 8, 0));
    std::memcpy(c->h_off.p, qs.q_off.data(), ((size_t)qs.R + 1) * 8);
    CU(c, cudaMemcpyAsync(c->q_off.p, c->h_off.p, ((size_t)qs.R + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  }
  { ProfScope ps(c, SKB_K_SELECT, 1);
    skb_launch_fill_qread(c->q_off.as<uint64_t>(), qs.R, c->qread.as<uint32_t>(), c->stream); }
  return check_launch(c, "fill_qread");
}

int run_passes(skb_ctx* c, const QuerySet& qs, uint32_t top, uint32_t* d_out_idx, uint64_t* d_out_sum);

int predict_checks(skb_ctx* c, uint32_t k, uint32_t s_query, uint32_t top, int pad) {
  if (!c->has_ref) return fail(c, SKB_ERR_NO_REFERENCE, "no reference uploaded");
  if (k < 1 || k > SKB_MAX_K) return fail(c, SKB_ERR_UNSUPPORTED_K, "k=%u unsupported (1..%d)", k, SKB_MAX_K);
  if (top < 1 || top > SKB_MAX_TOP) return fail(c, SKB_ERR_INVALID_ARG, "top=%u unsupported (1..%d)", top, SKB_MAX_TOP);
  if (top > c->n_rows && !pad)
    return fail(c, SKB_ERR_TOP_GT_N, "top (%u) exceeds the number of reference sketches (%u)", top, c->n_rows);
  if (s_query < 1) return fail(c, SKB_ERR_INVALID_ARG, "s_query must be >= 1");
  return SKB_OK;
}

// passes over a query set whose pass reads may be pieces of the caller's reads: rank into scratch, report last pieces
int run_passes_reported(skb_ctx* c, const QuerySet& qs, uint32_t n_reads, uint32_t top, uint32_t* d_out_idx, uint64_t* d_out_sum) {
  if (qs.out_row.empty()) return run_passes(c, qs, top, d_out_idx, d_out_sum);
  CU(c, c->piece_idx.ensure((size_t)qs.R * top * 4));
  CU(c, c->piece_sum.ensure((size_t)qs.R * top * 8));
  CU(c, c->piece_row.ensure((size_t)qs.R * 4));
  CU(c, cudaMemcpyAsync(c->piece_row.p, qs.out_row.data(), (size_t)qs.R * 4, cudaMemcpyHostToDevice, c->stream));
  if (int rc = run_passes(c, qs, top, c->piece_idx.as<uint32_t>(), c->piece_sum.as<uint64_t>())) return rc;
  { ProfScope ps(c, SKB_K_RANK, 1);
    skb_launch_report_pieces(c->piece_idx.as<uint32_t>(), c->piece_sum.as<unsigned long long>(), c->piece_row.as<uint32_t>(), qs.R, top,
                             d_out_idx, reinterpret_cast<unsigned long long*>(d_out_sum), c->stream); }
  if (int rc = check_launch(c, "report_pieces")) return rc;
  CU(c, cudaStreamSynchronize(c->stream));
  (void)n_reads;
  return SKB_OK;
}

int predict_device(skb_ctx* c, skb_batch* b, uint32_t k, uint32_t s_query, uint64_t seed, uint32_t top, int pad,
                   uint32_t* d_out_idx, uint64_t* d_out_sum) {
  if (int rc = predict_checks(c, k, s_query, top, pad)) return rc;
  const auto t_0 = std::chrono::steady_clock::now();
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
  if (int rc = use_batch(b)) return rc;
  const uint32_t R = (uint32_t)b->g_first.size();
  c->st_passes = 0; c->st_qhashes = 0; c->st_cands = 0; c->st_ref_bytes = c->ref_len * 8;
  if (R == 0) return SKB_OK;
  if (c->n_rows == 0) {
    CU(c, cudaMemsetAsync(d_out_idx, 0xFF, (size_t)R * top * 4, c->stream));
    CU(c, cudaMemsetAsync(d_out_sum, 0, (size_t)R * top * 8, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return SKB_OK;
  }
  // ---- per-read query sets: hashes <= hmax, distinct, ascending, at most s_query
  std::vector<uint32_t> qn;
  std::vector<uint64_t> kmers;
  SelectPlan plan;
  if (int rc = run_hash_select(c, b, k, s_query, seed, true, c->hmax, nullptr, nullptr, qn, kmers, plan)) return rc;
  const double t_hash = ms_since(t_0);
  uint64_t QN = 0;
  for (uint32_t v : qn) QN += v;
  CU(c, c->qh.ensure(std::max<uint64_t>(QN, 1) * 8));
  {
    CU(c, c->h_off.ensure(((size_t)R + 1) * 8, 0));  // page-locked; rewritten by the next call only, which starts after this one has synchronised
    uint64_t* off = c->h_off.as<uint64_t>();
    off[0] = 0;
    for (uint32_t r = 0; r < R; ++r) off[r + 1] = off[r] + qn[r];
    CU(c, c->q_off.ensure(((size_t)R + 1) * 8));
    CU(c, cudaMemcpyAsync(c->q_off.p, off, ((size_t)R + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    ProfScope ps(c, SKB_K_SELECT, 1);
    skb_launch_compact_queries(c->cand_pool.as<uint64_t>(), c->g_base.as<uint64_t>(), c->g_outn.as<uint32_t>(), c->q_off.as<uint64_t>(),
                               R, c->qh.as<uint64_t>(), c->stream);
  }
  if (int rc = check_launch(c, "compact")) return rc;
  QuerySet qs;
  if (int rc = finish_query_set(c, qn, qs, true)) return rc;
  const double t_query = ms_since(t_0);
  const int rc = run_passes_reported(c, qs, R, top, d_out_idx, d_out_sum);
  if (c->trace_passes)
    fprintf(stderr, "[skb] predict call: %u reads, host clock: query lists ready at %.3f ms (hash + select %.3f), passes done at %.3f ms (%llu passes)\n",
            R, t_query, t_hash, ms_since(t_0), (unsigned long long)c->st_passes);
  return rc;
}

int run_passes(skb_ctx* c, const QuerySet& qs, uint32_t top, uint32_t* d_out_idx, uint64_t* d_out_sum) {
  const uint32_t R = qs.R;
  const std::vector<uint32_t>& qn = qs.qn;
  const std::vector<uint64_t>& q_off = qs.q_off;
  uint32_t* h_total = c->h_scal;
  uint32_t* d_scal = c->scal.as<uint32_t>();
  unsigned long long* d_cand_stat = reinterpret_cast<unsigned long long*>(c->scal.as<uint8_t>() + 64);  // statistics
  uint32_t* d_abort = d_scal + 24;
  uint32_t* d_dense_ovf = d_scal + 20;
  CU(c, cudaMemsetAsync(d_cand_stat, 0, 16, c->stream));  // candidates, member keys
  CU(c, cudaMemsetAsync(d_abort, 0, 32, c->stream));
  CU(c, cudaMemsetAsync(d_scal + 8, 0, 32, c->stream));  // per post-slot: bucket-overflow flag, interval count, segment-record count (the verdict kernel clears them after every pass)
  CU(c, cudaMemsetAsync(d_dense_ovf, 0, 4, c->stream));

  // ---- every buffer a pass may need, sized once for the largest pass: nothing is (re)allocated while passes are in flight
  const uint64_t budget = c->cand_budget;
  const uint32_t Bmax = std::max<uint32_t>(1, std::min<uint32_t>(c->pass_max, R));
  const uint32_t stride_max = (uint32_t)round_up(Bmax, 512);
  // dense ranking: row groups so that (reads x groups) threads fill the GPU, bounded by the part lists' size
  uint32_t groups_max = std::max<uint32_t>(1, std::min<uint32_t>(256, (uint32_t)((256ull << 20) / ((uint64_t)Bmax * top * 12))));
  groups_max = std::min<uint32_t>(groups_max, std::max<uint32_t>(1, c->n_rows / 64));
  {
    cudaError_t e = cudaSuccess;
    auto need = [&](DevBuf& b, size_t bytes) { if (e == cudaSuccess) e = b.ensure(bytes); };
    need(c->counts, (size_t)SKB_MAX_TRACKED * stride_max * 2);
    need(c->tprefix, (size_t)SKB_MAX_TRACKED * stride_max * 4);
    need(c->textra, (size_t)SKB_MAX_TRACKED * 8);
    for (int i = 0; i < SKB_NTAB; ++i) { need(c->lb_sum[i], (size_t)Bmax * 8); need(c->lb_idx[i], (size_t)Bmax * 4); }
    for (int i = 0; i < 2; ++i) {
      need(c->cand_cnt[i], (size_t)Bmax * 4);
      need(c->cand[i], (size_t)std::max<uint64_t>(budget, std::max<uint32_t>(c->n_rows, 64)) * sizeof(SkbCand));
      need(c->ivl[i], (size_t)SKB_IVL_CAP * sizeof(SkbInterval));
      need(c->seg_hdr[i], (size_t)SKB_SEG_CAP * 16);
      need(c->seg_words[i], (size_t)SKB_SEG_CAP * SKB_SEG_WORDS_MAX * 4);
    }
    // u16 per read with u8 counters, u32 per read with u16 counters (whose passes hold at most skb_fused_max_reads(0) reads)
    need(c->dense, (size_t)c->n_rows * std::max<size_t>((size_t)stride_max * 2, (size_t)std::min<uint32_t>(stride_max, skb_fused_max_reads(0)) * 4));
    need(c->part_idx, (size_t)groups_max * Bmax * top * 4);
    need(c->part_sum, (size_t)groups_max * Bmax * top * 8);
    {
      const size_t n_anchor = (Bmax + 63) / 64;
      need(c->anch_part_idx, (size_t)kAnchorGroups * n_anchor * top * 4);
      need(c->anch_part_sum, (size_t)kAnchorGroups * n_anchor * top * 8);
      need(c->anch_idx, n_anchor * top * 4);
      need(c->anch_sum, n_anchor * top * 8);
    }
    if (e != cudaSuccess) return fail(c, SKB_ERR_OOM, "pass buffers: %s", cudaGetErrorString(e));
    for (int i = 0; i < 2; ++i) CU(c, cudaMemsetAsync(c->cand_cnt[i].p, 0, (size_t)Bmax * 4, c->stream));  // (a pass's last kernel clears them again)
    const uint64_t key_budget = SKB_PASS_KEY_BUDGET;
    uint64_t max_keys = 0;
    for (uint32_t r0 = 0; r0 < R; r0 += Bmax) max_keys = std::max(max_keys, q_off[std::min(R, r0 + Bmax)] - q_off[r0]);
    if (int rc = ensure_table(c, (uint32_t)std::min<uint64_t>(std::max<uint64_t>(max_keys, 1), key_budget))) return rc;
  }

  // The pass loop. A pass = a block of consecutive reads against the whole shard:
  //   pre(i)    query table T_i (+ for a sparse pass: exact per-read sums of the tracked rows X_i -> per-read bounds L_i)
  //   stream(i) the HBM-bound kernel: sums S_{i-1} -> S_i, plus what the ranking needs
  //   post(i)   the top-N of every read of the pass, and the tracked rows U_i the next passes take their bounds from
  // Two ranking modes, same answer:
  //   sparse  rows are tested against per-read lower bounds of the top-th key (from the tracked rows' exact sums);
  //           the few that pass become per-read candidate lists (walk / expand / select). Cheap when the bounds are good.
  //   dense   the stream writes every hit row's prefix sums over the reads of the pass and every read is ranked over all
  //           rows by brute force (dense_topk + merge). Needs no bounds and cannot overflow: used right after a reset,
  //           for small shards, and to redo a sparse pass whose candidates overflowed.
  // stream() kernels run on the main stream, pre() and post() on the side stream. In steady state (sparse passes that
  // have been seen to fit) passes are enqueued in batches and checked on the host only every few passes; with
  // `pipeline` the bounds of pass i then come from X_i = U_{i-2} with sums S_{i-2} + (row totals against T_{i-1}), all
  // exact, so pre(i) does not wait for stream(i-1). A sparse pass whose candidates overflow records itself in `abort` on
  // the device, the passes behind it do nothing, and the host redoes it as a dense pass.
  struct PassRec { uint32_t r, B; int sums_in, tracked_in; bool full; };
  std::vector<PassRec> recs;
  struct Pending { bool on = false, dense = false; SkbRankArgs ra; SkbDenseArgs da; int fused_ev = 0; } post;
  const uint32_t kBatch = 8;
  bool prev_pipelined = false;  // the previous pass of this call was enqueued in steady state (S_{i-2}, T_{i-1} are its inputs)
  bool have_fused = false;      // a streaming kernel of this call has been enqueued (its event orders the next table build)
  bool have_post = false;       // a post-pass sequence of this call has been enqueued (ev_post is behind the newest one)
  uint32_t seq = 0, r = 0, dense_until = 0, dense_max_reads = Bmax;
  int rc_final = SKB_OK, tracked_prev = c->tracked_cur;
  const SkbRefView rv = ref_view(c);
  // (a small shard's prefix-sum vectors are a few MB: dense ranking costs less than the sparse launch train)
  const bool small_shard = c->rank_mode == 2 || (c->rank_mode == 0 && (uint64_t)c->n_rows * stride_max * 2 <= (32ull << 20));

  // event records / waits between the two streams: the first error is kept and ends the loop at its next iteration
  cudaError_t ev_err = cudaSuccess;
#define EV(call) do { const cudaError_t e__ = (call); if (ev_err == cudaSuccess) ev_err = e__; } while (0)
  auto enqueue_post = [&]() {  // post(i) on the side stream, behind stream(i)
    if (!post.on) return;
    EV(cudaStreamWaitEvent(c->side, c->ev_fused[post.fused_ev], 0));
    if (post.dense) {
      ProfScope ps(c, SKB_K_RANK, 5, c->side);
      skb_launch_dense_topk(post.da, c->side);
      skb_launch_merge_topn(post.da.part_idx, post.da.part_sum, post.da.groups, post.da.n_reads, top, post.ra.out_idx, post.ra.out_sum, c->side);
      skb_launch_verdict_update(post.ra, false, c->side);
    } else {
      ProfScope ps(c, SKB_K_RANK, 3, c->side);
      skb_launch_rank_expand(post.ra, c->side); skb_launch_rank_select(post.ra, c->side);
      skb_launch_verdict_update(post.ra, true, c->side);  // (the update is skipped on the device after an overflow)
    }
    EV(cudaEventRecord(c->ev_post, c->side));
    have_post = true;
    post.on = false;
  };
  auto join_streams = [&]() -> cudaError_t {  // main stream waits for everything enqueued on the side stream
    cudaError_t e = cudaEventRecord(c->ev_join, c->side);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->ev_join, 0);
    return e;
  };
  // the side stream starts behind whatever the main stream has done so far (hashing, selection, resets)
  EV(cudaEventRecord(c->ev_join, c->stream));
  EV(cudaStreamWaitEvent(c->side, c->ev_join, 0));
  EV(cudaStreamWaitEvent(c->tabs, c->ev_join, 0));

  while (r < R) {
    if (ev_err != cudaSuccess) { rc_final = fail(c, SKB_ERR_CUDA, "predict pass (stream ordering): %s", cudaGetErrorString(ev_err)); break; }
    const bool need_init = c->tracked_top != top;  // no tracked rows yet (first pass after an upload / a reset, or `top` changed)
    const bool dense = small_shard || need_init || c->dense_left > 0 || r < dense_until;
    uint32_t B = std::min<uint32_t>(c->pass_max, R - r);
    if (dense) B = std::min(B, dense_max_reads);
    const bool full = R - r >= c->pass_max;  // not the short last pass of a call
    // u8 counters need every read of the pass to keep <= 255 query hashes; otherwise u16 counters and fewer reads
    bool narrow = true;
    for (uint32_t i = r; i < r + B; ++i)
      if (qn[i] > 255u) { narrow = false; break; }
    if (!narrow) B = std::min<uint32_t>(B, skb_fused_max_reads(0));
    // keep the pass's key count inside the filter's design load (and the table)
    const uint64_t key_budget = SKB_PASS_KEY_BUDGET;
    while (B > 1 && q_off[r + B] - q_off[r] > key_budget) B = std::max(1u, B / 2);
    const uint32_t nkeys = (uint32_t)(q_off[r + B] - q_off[r]);
    if (nkeys > c->t_maxkeys) { rc_final = fail(c, SKB_ERR_INVALID_ARG, "a single read keeps %u query hashes; the pass table holds %u", nkeys, c->t_maxkeys); break; }
    // per-read candidate buckets share one fixed budget
    c->cand_cap = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(std::max<uint32_t>(c->n_rows, 64), budget / B));
    const uint32_t stride = (uint32_t)round_up(B, narrow ? 512 : 256);
    const bool steady = !dense && c->pass_proven;
    const bool lag2 = c->pipeline && steady && prev_pipelined;  // bounds from U_{i-2}: pre(i) does not wait for stream(i-1)
    const int tab = (c->tab_cur + 1) % SKB_NTAB, tab_prev = c->tab_cur;
    const int qs = seq & 1;                       // post slot (records, buckets, counters)
    const int s_in = c->sums_cur, s_out = (s_in + 1) % SKB_NSUMS, s_in2 = (s_in + SKB_NSUMS - 1) % SKB_NSUMS;
    const int x_in = lag2 ? tracked_prev : c->tracked_cur, x_out = (c->tracked_cur + 1) % SKB_NTRACK;
    if (!lag2) enqueue_post();  // X_i = U_{i-1}: post(i-1) first

    // ---- table(i) on its own stream: it needs nothing of pass i-1 but the end of its streaming kernel (every earlier
    // user of the table's slot is over by then), so it runs next to post(i-1) instead of behind it
    const SkbTable t = table_of(c, tab);
    EV(cudaStreamWaitEvent(c->tabs, have_fused ? c->ev_fused[(seq + 1) & 1] : c->ev_join, 0));
    { ProfScope ps(c, SKB_K_TABLE, nkeys ? 4 : 1, c->tabs);
      skb_launch_table_build(t, c->qh.as<uint64_t>() + q_off[r], c->qread.as<uint32_t>() + q_off[r], nkeys, r, c->t_built[tab], c->tabs);
      c->t_built[tab] = nkeys; }
    EV(cudaEventRecord(c->ev_tab[tab], c->tabs));
    // ---- pre(i): on the side stream behind post(i-1), whose tracked rows it uses; with bounds taken two passes back
    // (lag2) it needs nothing of post(i-1) and follows the table build on the table stream, next to post(i-1)
    cudaStream_t pre_st = lag2 ? c->tabs : c->side;
    if (lag2) {
      if (have_post) EV(cudaStreamWaitEvent(c->tabs, c->ev_post, 0));  // U_{i-2} (and every post before it) is complete
    } else {
      EV(cudaStreamWaitEvent(c->side, c->ev_tab[tab], 0));
    }
    SkbRankArgs ra{};
    ra.tracked_counts = c->counts.as<uint16_t>(); ra.tracked_prefix = c->tprefix.as<uint32_t>();
    ra.tracked_extra = c->textra.as<unsigned long long>();
    ra.row_stride = stride; ra.n_reads = B; ra.row_base = c->row_base;
    ra.sums_in = c->sums[lag2 ? s_in2 : s_in].as<unsigned long long>();
    ra.tracked = c->tracked[x_in].as<uint32_t>(); ra.n_tracked = ra.tracked + SKB_MAX_TRACKED;
    ra.lb_sum = c->lb_sum[tab].as<unsigned long long>(); ra.lb_idx = c->lb_idx[tab].as<uint32_t>();
    ra.ivl = c->ivl[qs].as<SkbInterval>(); ra.ivl_cap = SKB_IVL_CAP; ra.ivl_total = d_scal + 8 + 4 * qs + 1;
    ra.seg_hdr = c->seg_hdr[qs].as<uint4>(); ra.seg_words = c->seg_words[qs].as<uint32_t>(); ra.seg_cap = SKB_SEG_CAP;
    ra.seg_cpw = narrow ? 4 : 2; ra.seg_words_per = stride / 32 / ra.seg_cpw; ra.seg_total = d_scal + 8 + 4 * qs + 2;
    ra.cand = c->cand[qs].as<SkbCand>(); ra.cand_cap = c->cand_cap; ra.cand_total = d_scal + 8 + 4 * qs;
    ra.cand_cnt = c->cand_cnt[qs].as<uint32_t>(); ra.cand_stat = d_cand_stat;
    ra.top = top; ra.out_idx = d_out_idx + (size_t)r * top;
    ra.out_sum = reinterpret_cast<unsigned long long*>(d_out_sum) + (size_t)r * top;
    ra.tracked_next = c->tracked[x_out].as<uint32_t>(); ra.n_tracked_next = ra.tracked_next + SKB_MAX_TRACKED;
    ra.abort = d_abort; ra.seq = seq;
    cudaError_t e = cudaSuccess;
    if (!dense) {
      e = cudaMemsetAsync(c->textra.p, 0, (size_t)SKB_MAX_TRACKED * 8, pre_st);
      if (e != cudaSuccess) { rc_final = fail(c, SKB_ERR_CUDA, "predict pass: %s", cudaGetErrorString(e)); break; }
      ProfScope ps(c, SKB_K_RANK, 2 + (lag2 ? 1 : 0), pre_st);
      if (lag2) skb_launch_tracked_totals(rv, ra.tracked, ra.n_tracked, table_of(c, tab_prev), ra.tracked_extra, pre_st);
      skb_launch_rank_bounds(rv, t, ra, nkeys != 0, pre_st);
    }
    EV(cudaEventRecord(c->ev_pre[tab], pre_st));
    if (lag2) enqueue_post();  // post(i-1) on the side stream, next to pre(i)

    // ---- stream(i) on the main stream
    EV(cudaStreamWaitEvent(c->stream, c->ev_pre[tab], 0));
    SkbFusedArgs fa{};
    fa.rv = rv; fa.cta_row = c->cta_row.as<uint32_t>(); fa.num_ctas = c->stream_ctas; fa.table = t;
    fa.n_reads = B; fa.cnt_stride = stride; fa.narrow = narrow ? 1 : 0; fa.skip_stream = nkeys == 0; fa.row_base = c->row_base;
    fa.rowbuf = skb_fused_rowbuf(stride, fa.narrow); fa.rowbuf_log2 = fa.rowbuf == 8 ? 3 : 2;
    fa.sums_in = c->sums[s_in].as<unsigned long long>(); fa.sums_out = c->sums[s_out].as<unsigned long long>();
    fa.lb_sum = ra.lb_sum; fa.lb_idx = ra.lb_idx;
    fa.ivl = c->ivl[qs].as<SkbInterval>(); fa.ivl_cap = SKB_IVL_CAP; fa.ivl_total = d_scal + 8 + 4 * qs + 1; fa.abort = d_abort;
    fa.seg_hdr = c->seg_hdr[qs].as<uint4>(); fa.seg_words = c->seg_words[qs].as<uint32_t>(); fa.seg_cap = SKB_SEG_CAP;
    fa.seg_total = d_scal + 8 + 4 * qs + 2;
    fa.dense = dense ? c->dense.p : nullptr; fa.dense_overflow = d_dense_ovf;
    fa.tile_cum = c->tile_cum.as<uint32_t>();
    fa.tpr = std::max(1u, (c->uniform_len + skb_fused_tile() - 1) / skb_fused_tile());
    fa.tpr_magic = (uint32_t)((0x100000000ull + fa.tpr - 1) / fa.tpr);
    { ProfScope ps(c, SKB_K_STREAM, 1); skb_launch_fused(fa, c->stream); }
    EV(cudaEventRecord(c->ev_fused[qs], c->stream));
    have_fused = true;
    if (int rc = check_launch(c, "predict pass")) { rc_final = rc; break; }
    post.on = true; post.dense = dense; post.ra = ra; post.fused_ev = qs;
    if (dense) {
      SkbDenseArgs da{};
      da.dense = c->dense.p; da.wide = narrow ? 0 : 1; da.cnt_stride = stride; da.n_rows = c->n_rows; da.n_reads = B;
      da.row_base = c->row_base; da.top = top;
      // row groups: enough CTAs (64 reads x one group each) to fill the GPU a few times over, at least 64 rows per group
      da.groups = std::max<uint32_t>(1, std::min<uint32_t>(groups_max, (uint32_t)((8ull * c->num_sms * 64 + B - 1) / B)));
      da.sums_in = fa.sums_in; da.sums_out = fa.sums_out;
      da.part_idx = c->part_idx.as<uint32_t>(); da.part_sum = c->part_sum.as<unsigned long long>();
      da.anchor_groups = kAnchorGroups;
      da.anchor_part_idx = c->anch_part_idx.as<uint32_t>(); da.anchor_part_sum = c->anch_part_sum.as<unsigned long long>();
      da.anchor_idx = c->anch_idx.as<uint32_t>(); da.anchor_sum = c->anch_sum.as<unsigned long long>();
      post.da = da;
    }
    recs.push_back({r, B, s_in, c->tracked_cur, full});
    c->st_passes += 1;
    c->sums_cur = s_out;
    tracked_prev = c->tracked_cur;
    c->tracked_cur = x_out;
    c->tab_cur = tab;
    prev_pipelined = steady;
    r += B;
    ++seq;
    if (steady && recs.size() < kBatch && r < R) continue;

    // ---- checkpoint: did any pass since the last one overflow?
    enqueue_post();
    if ((e = join_streams()) != cudaSuccess ||
        (e = cudaMemcpyAsync(h_total, d_abort, 32, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(h_total + 8, d_dense_ovf, 4, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(c->stream)) != cudaSuccess) {
      rc_final = fail(c, SKB_ERR_CUDA, "predict pass: %s", cudaGetErrorString(e));
      break;
    }
    prev_pipelined = false;  // everything is finished: the next pass may use S_{i-1} and U_{i-1} directly
    if (c->trace_passes)
      fprintf(stderr, "[skb] checkpoint after %zu pass(es), last %s at read %u with %u reads: fullest bucket %u of %u, intervals %u, segment records %u%s\n",
              recs.size(), dense ? "dense" : "sparse", recs.back().r, recs.back().B, h_total[2], c->cand_cap, h_total[3], h_total[4],
              h_total[0] || h_total[8] ? " OVERFLOW" : "");
    if (h_total[8] != 0) {  // a 16-bit prefix sum overflowed in the dense pass just checked (dense passes are checked one by one)
      const PassRec pr = recs.back();
      if ((e = cudaMemsetAsync(d_dense_ovf, 0, 4, c->stream)) != cudaSuccess) { rc_final = fail(c, SKB_ERR_CUDA, "predict pass: %s", cudaGetErrorString(e)); break; }
      r = pr.r; c->sums_cur = pr.sums_in; c->tracked_cur = pr.tracked_in; tracked_prev = pr.tracked_in;
      c->st_passes -= 1;
      dense_until = std::max(dense_until, pr.r + pr.B);
      dense_max_reads = 256;  // 255 * 256 < 2^16: cannot overflow
      recs.clear();
      EV(cudaEventRecord(c->ev_join, c->stream));
      EV(cudaStreamWaitEvent(c->side, c->ev_join, 0));
      EV(cudaStreamWaitEvent(c->tabs, c->ev_join, 0));
      continue;
    }
    if (h_total[0] != 0) {  // more contenders than a bucket / a record list holds, in the sparse pass numbered h_total[1]
      const PassRec pr = recs[recs.size() - (seq - h_total[1])];
      if (c->trace_passes)
        fprintf(stderr, "[skb] pass at read %u with %u reads overflowed (fullest bucket %u of %u, intervals %u of %u, segment records %u of %u): redo dense\n", pr.r, pr.B,
                h_total[2], c->cand_cap, h_total[3], (unsigned)SKB_IVL_CAP, h_total[4], (unsigned)SKB_SEG_CAP);
      c->st_passes -= (seq - h_total[1]) - 1;  // the passes behind the failed one did nothing (the failed one did stream)
      if ((e = cudaMemsetAsync(d_abort, 0, 32, c->stream)) != cudaSuccess ||
          (e = cudaMemsetAsync(d_scal + 8, 0, 32, c->stream)) != cudaSuccess) {
        rc_final = fail(c, SKB_ERR_CUDA, "predict pass: %s", cudaGetErrorString(e));
        break;
      }
      // redo from that pass, densely. Its input sums and the tracked rows it started from are intact: the passes
      // behind it left everything alone.
      r = pr.r; c->sums_cur = pr.sums_in; c->tracked_cur = pr.tracked_in; tracked_prev = pr.tracked_in;
      dense_until = std::max(dense_until, pr.r + pr.B);
      c->dense_left = 1;  // and the pass after it: its bounds would be as loose
      c->pass_proven = false;
      recs.clear();
      EV(cudaEventRecord(c->ev_join, c->stream));
      EV(cudaStreamWaitEvent(c->side, c->ev_join, 0));
      EV(cudaStreamWaitEvent(c->tabs, c->ev_join, 0));
      continue;
    }
    if (dense) {
      c->tracked_top = top;  // the dense pass proposed tracked rows: bounds exist from here on
      if (!need_init && recs.back().r >= dense_until && c->dense_left > 0) c->dense_left -= 1;
    } else if (recs.back().full) {
      c->pass_proven = true;  // full-size sparse passes fit: from here on they are enqueued in batches
    }
    recs.clear();
  }
#undef EV
  if (!rc_final && ev_err != cudaSuccess) rc_final = fail(c, SKB_ERR_CUDA, "predict pass (stream ordering): %s", cudaGetErrorString(ev_err));
  if (rc_final) {  // leave nothing in flight behind an error
    cudaStreamSynchronize(c->tabs);
    cudaStreamSynchronize(c->side);
    cudaStreamSynchronize(c->stream);
    return rc_final;
  }
  unsigned long long cands[2] = {0, 0};
  CU(c, cudaMemcpyAsync(cands, d_cand_stat, 16, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->st_cands = cands[0];
  c->st_members = cands[1];
  return SKB_OK;
}

#define NC(c, call)                                                                                         \
  do {                                                                                                      \
    ncclResult_t r__ = (call);                                                                              \
    if (r__ != ncclSuccess) return fail((c), SKB_ERR_COMM, "%s: %s", #call, g_nccl.GetErrorString(r__));   \
  } while (0)

void dist_range(uint64_t n, int rank, int world, uint64_t* begin, uint64_t* count) {
  const uint64_t per = world > 0 ? (n + world - 1) / world : n;
  const uint64_t b = std::min<uint64_t>(n, per * (uint64_t)rank), e = std::min<uint64_t>(n, per * ((uint64_t)rank + 1));
  if (begin) *begin = b;
  if (count) *count = e - b;
}

// Streaming predict over a reference sharded across the ranks of the communicator. Rank r holds the reads
// dist_range(reads_total, r, world) of the call in `b` (it packed and staged only those) and hashes only those; the
// per-read query lists are exchanged (all-gather of the counts, then of the hashes), every rank runs the passes of ALL
// reads against its row shard, the per-rank top-N lists (global row indices) are all-gathered and merged by
// (sum desc, index asc). Exact: every member of the global top-N is in its shard's local top-N, and a read's query
// list only matters below a shard's largest reference hash, so the list cut at the largest hash of ANY shard serves
// them all. The merged ranking of every read ends up in d_out_* on every rank.
int predict_dist_device(skb_ctx* c, skb_batch* b, uint64_t reads_total, uint32_t k, uint32_t s_query, uint64_t seed,
                        uint32_t top, uint32_t* d_out_idx, uint64_t* d_out_sum) {
  if (!c->comm || c->world <= 1) {
    if (reads_total != b->g_first.size()) return fail(c, SKB_ERR_INVALID_ARG, "no communicator: the batch must hold all %llu reads", (unsigned long long)reads_total);
    return predict_device(c, b, k, s_query, seed, top, 0, d_out_idx, d_out_sum);
  }
  if (int rc = predict_checks(c, k, s_query, top, 1)) return rc;
  if (reads_total > 0xFFFFFFF0ull) return fail(c, SKB_ERR_INVALID_ARG, "too many reads in one call");
  if (int rc = use_batch(b)) return rc;
  const int W = c->world;
  const uint32_t R = (uint32_t)reads_total, R_loc = (uint32_t)b->g_first.size();
  uint64_t my_begin = 0, my_count = 0;
  dist_range(R, c->rank, W, &my_begin, &my_count);
  if (R_loc != my_count)
    return fail(c, SKB_ERR_INVALID_ARG, "rank %d of %d must hold reads [%llu, %llu) of the call, the batch has %u", c->rank, W,
                (unsigned long long)my_begin, (unsigned long long)(my_begin + my_count), R_loc);
  c->st_passes = 0; c->st_qhashes = 0; c->st_cands = 0; c->st_ref_bytes = c->ref_len * 8;
  if (R == 0) return SKB_OK;
  const uint32_t Rmax = (R + W - 1) / W;
  // ---- the largest reference hash of any shard: exchanged once per (communicator, upload) — uploading a shard is
  // collective in this sense: every rank uploads before the next collective predict
  if (!c->tau_all_valid) {
    CU(c, c->hmax_all.ensure((size_t)W * 8));
    unsigned long long* d_hmax = c->hmax_all.as<unsigned long long>();
    unsigned long long my_hmax = c->n_rows ? c->hmax : 0ull;
    CU(c, cudaMemcpyAsync(d_hmax + c->rank, &my_hmax, 8, cudaMemcpyHostToDevice, c->stream));
    NC(c, g_nccl.AllGather(d_hmax + c->rank, d_hmax, 1, ncclUint64, c->comm, c->stream));
    std::vector<unsigned long long> hm(W);
    CU(c, cudaMemcpyAsync(hm.data(), d_hmax, (size_t)W * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->tau_all = *std::max_element(hm.begin(), hm.end());
    c->tau_all_valid = true;
  }
  const uint64_t tau = c->tau_all;
  // ---- this rank's reads: query lists
  std::vector<uint32_t> qn_loc;
  std::vector<uint64_t> kmers;
  SelectPlan plan;
  if (int rc = run_hash_select(c, b, k, s_query, seed, true, tau, nullptr, nullptr, qn_loc, kmers, plan)) return rc;
  // ---- everybody's counts
  CU(c, c->qn_all.ensure((size_t)W * Rmax * 4));
  uint32_t* d_qn = c->qn_all.as<uint32_t>();
  CU(c, cudaMemsetAsync(d_qn + (size_t)c->rank * Rmax, 0, (size_t)Rmax * 4, c->stream));
  if (R_loc) CU(c, cudaMemcpyAsync(d_qn + (size_t)c->rank * Rmax, c->g_outn.p, (size_t)R_loc * 4, cudaMemcpyDeviceToDevice, c->stream));
  NC(c, g_nccl.AllGather(d_qn + (size_t)c->rank * Rmax, d_qn, Rmax, ncclUint32, c->comm, c->stream));
  std::vector<uint32_t> qn(R);
  CU(c, c->h_outn.ensure((size_t)W * Rmax * 4 + 64, 0));   // page-locked staging, as in the single-GPU path
  CU(c, c->h_off.ensure(((size_t)R + 1) * 8, 0));
  const uint32_t* qn_pad = c->h_outn.as<uint32_t>();
  uint64_t* off = c->h_off.as<uint64_t>();
  CU(c, cudaMemcpyAsync(c->h_outn.p, d_qn, (size_t)W * Rmax * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  off[0] = 0;
  for (int g = 0; g < W; ++g) {
    uint64_t gb, gc;
    dist_range(R, g, W, &gb, &gc);
    for (uint64_t i = 0; i < gc; ++i) qn[gb + i] = qn_pad[(size_t)g * Rmax + i];
  }
  for (uint32_t r = 0; r < R; ++r) off[r + 1] = off[r] + qn[r];
  const uint64_t QN = off[R];
  CU(c, c->qh.ensure(std::max<uint64_t>(QN, 1) * 8));
  CU(c, c->q_off.ensure(((size_t)R + 1) * 8));
  CU(c, cudaMemcpyAsync(c->q_off.p, off, ((size_t)R + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  if (R_loc) {
    ProfScope ps(c, SKB_K_SELECT, 1);
    skb_launch_compact_queries(c->cand_pool.as<uint64_t>(), c->g_base.as<uint64_t>(), c->g_outn.as<uint32_t>(),
                               c->q_off.as<uint64_t>() + my_begin, R_loc, c->qh.as<uint64_t>(), c->stream);
  }
  if (int rc = check_launch(c, "compact")) return rc;
  // ---- everybody's hashes: rank g broadcasts its slice of the flat list in place
  NC(c, g_nccl.GroupStart());
  for (int g = 0; g < W; ++g) {
    uint64_t gb, gc;
    dist_range(R, g, W, &gb, &gc);
    const uint64_t n = off[gb + gc] - off[gb];
    if (n == 0) continue;
    uint64_t* p = c->qh.as<uint64_t>() + off[gb];
    ncclResult_t r = g_nccl.Broadcast(p, p, n, ncclUint64, g, c->comm, c->stream);
    if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail(c, SKB_ERR_COMM, "ncclBroadcast: %s", g_nccl.GetErrorString(r)); }
  }
  NC(c, g_nccl.GroupEnd());
  QuerySet qs;   // (the offsets are on the device already; the passes are stream-ordered behind the exchange)
  if (int rc = finish_query_set(c, qn, qs, true)) return rc;
  // ---- passes of every read against this rank's rows, then the exchange of the local top-N lists
  CU(c, c->out_idx.ensure(std::max<size_t>(4, (size_t)R * top * 4)));
  CU(c, c->out_sum.ensure(std::max<size_t>(8, (size_t)R * top * 8)));
  CU(c, c->gath_idx.ensure((size_t)W * R * top * 4));
  CU(c, c->gath_sum.ensure((size_t)W * R * top * 8));
  if (c->n_rows == 0) {
    CU(c, cudaMemsetAsync(c->out_idx.p, 0xFF, (size_t)R * top * 4, c->stream));
    CU(c, cudaMemsetAsync(c->out_sum.p, 0, (size_t)R * top * 8, c->stream));
  } else if (int rc = run_passes_reported(c, qs, R, top, c->out_idx.as<uint32_t>(), c->out_sum.as<uint64_t>())) {
    return rc;
  }
  NC(c, g_nccl.AllGather(c->out_idx.p, c->gath_idx.p, (size_t)R * top, ncclUint32, c->comm, c->stream));
  NC(c, g_nccl.AllGather(c->out_sum.p, c->gath_sum.p, (size_t)R * top, ncclUint64, c->comm, c->stream));
  { ProfScope ps(c, SKB_K_MERGE, 1);
    skb_launch_merge_topn(c->gath_idx.as<uint32_t>(), c->gath_sum.as<unsigned long long>(), (uint32_t)W, R, top, d_out_idx,
                          reinterpret_cast<unsigned long long*>(d_out_sum), c->stream); }
  if (int rc = check_launch(c, "merge")) return rc;
  CU(c, cudaStreamSynchronize(c->stream));
  return SKB_OK;
}

}  // namespace

// =========================================================================================================
// exported C ABI
// =========================================================================================================
extern "C" {

const char* skb_version(void) { return SKB_VERSION_STR; }

int skb_create(int device, skb_ctx** out) {
  if (!out) return SKB_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return SKB_ERR_NO_DEVICE;
  if (device < 0 || device >= n) return SKB_ERR_INVALID_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return SKB_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SKB_ERR_CUDA;
  if (prop.major != 10) return SKB_ERR_NO_DEVICE;  // built for sm_100a only
  skb_ctx* c = new skb_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->stream_ctas = c->num_sms;
  if (const char* e = getenv("SKB_STREAM_CTAS")) c->stream_ctas = std::max(1, std::min(c->num_sms, atoi(e)));  // experiments
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->tabs, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < SKB_NTAB; ++i) ok = cudaEventCreateWithFlags(&c->ev_pre[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < SKB_NTAB; ++i) ok = cudaEventCreateWithFlags(&c->ev_tab[i], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->ev_post, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < 2; ++i) ok = cudaEventCreateWithFlags(&c->ev_fused[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { delete c; return SKB_ERR_CUDA; }
  // environment switches are read once, here (tests and experiments; never inside the pass loop)
  if (const char* e = getenv("SKB_CAND_BUDGET")) c->cand_budget = std::max<uint64_t>(64, strtoull(e, nullptr, 10));
  if (const char* e = getenv("SKB_TRACE_PASSES")) c->trace_passes = e[0] != '0';
  if (const char* e = getenv("SKB_DENSE_AFTER_RESET")) c->dense_after_reset = (uint32_t)atoi(e);
  if (const char* e = getenv("SKB_PIPELINE")) c->pipeline = e[0] != '0';  // experiments: 0 = every pass waits for the one before
  if (cudaHostAlloc((void**)&c->h_scal, 64, cudaHostAllocDefault) != cudaSuccess) {
    cudaStreamDestroy(c->stream); delete c; return SKB_ERR_OOM;
  }
  *out = c;
  return SKB_OK;
}

void skb_destroy(skb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->side) cudaStreamSynchronize(c->side);
  if (c->tabs) cudaStreamSynchronize(c->tabs);
  if (c->comm) { g_nccl.CommDestroy(c->comm); c->comm = nullptr; }
  for (auto& e : c->pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  std::vector<DevBuf*> bufs = {&c->ref, &c->row_start, &c->row_len, &c->cta_row, &c->tile_cum, &c->memb, &c->tprefix, &c->textra,
                               &c->g_tau, &c->g_cap, &c->g_base, &c->g_cnt, &c->g_kmers, &c->g_active, &c->g_outn, &c->g_status, &c->g_tiles,
                               &c->cand_pool, &c->sk_hashes, &c->sk_counts, &c->q_off, &c->qh, &c->qread, &c->counts, &c->scal,
                               &c->out_idx, &c->out_sum, &c->misc, &c->dense, &c->part_idx, &c->part_sum, &c->anch_part_idx, &c->anch_part_sum, &c->anch_idx, &c->anch_sum, &c->piece_idx, &c->piece_sum, &c->piece_row, &c->qn_all, &c->gath_idx, &c->gath_sum, &c->hmax_all};
  for (auto& x : c->sums) bufs.push_back(&x);
  for (auto& x : c->tracked) bufs.push_back(&x);
  for (int i = 0; i < SKB_NTAB; ++i)
    for (DevBuf* x : {&c->lb_sum[i], &c->lb_idx[i], &c->t_slots[i], &c->t_fill[i], &c->t_reads[i], &c->t_slot[i], &c->t_bloom[i]}) bufs.push_back(x);
  for (int i = 0; i < 2; ++i)
    for (DevBuf* x : {&c->cand[i], &c->cand_cnt[i], &c->ivl[i], &c->seg_hdr[i], &c->seg_words[i]}) bufs.push_back(x);
  for (DevBuf* b : bufs) b->release();
  c->h_outn.release(); c->h_off.release();
  if (c->h_scal) cudaFreeHost(c->h_scal);
  cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->tabs) cudaStreamDestroy(c->tabs);
  for (int i = 0; i < SKB_NTAB; ++i) if (c->ev_tab[i]) cudaEventDestroy(c->ev_tab[i]);
  if (c->ev_post) cudaEventDestroy(c->ev_post);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  for (cudaEvent_t e : c->ev_pre) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_fused) if (e) cudaEventDestroy(e);
  delete c;
}

const char* skb_last_error(const skb_ctx* c) { return c ? c->err.c_str() : "null context"; }
void* skb_stream(skb_ctx* c) { return c ? (void*)c->stream : nullptr; }
int skb_synchronize(skb_ctx* c) {
  if (!c) return SKB_ERR_INVALID_ARG;
  CU(c, cudaStreamSynchronize(c->side));
  CU(c, cudaStreamSynchronize(c->stream));
  return SKB_OK;
}

// ---- batch ------------------------------------------------------------------------------------------------
int skb_batch_create(skb_ctx* c, skb_batch** out) {
  if (!c || !out) return SKB_ERR_INVALID_ARG;
  skb_batch* b = new skb_batch();
  b->ctx = c;
  *out = b;
  return SKB_OK;
}

void skb_batch_destroy(skb_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  if (b->ev_staged) { cudaEventSynchronize(b->ev_staged); cudaEventDestroy(b->ev_staged); }
  b->codes.release(); b->nmask.release();
  b->h_seg_group.release(); b->h_seg_chunk0.release(); b->h_seg_n.release(); b->h_g_len.release();
  b->d_codes.release(); b->d_nmask.release(); b->d_seg_group.release(); b->d_seg_chunk0.release(); b->d_seg_n.release(); b->d_g_len.release();
  delete b;
}

int skb_batch_clear(skb_batch* b) {
  if (!b) return SKB_ERR_INVALID_ARG;
  if (b->staged && b->ev_staged) {  // the page-locked buffers are about to be rewritten: their copies must be over
    cudaSetDevice(b->ctx->device);
    cudaEventSynchronize(b->ev_staged);
  }
  b->cur = 0; b->rec_pos.clear(); b->rec_len.clear();
  b->g_first.clear(); b->g_end.clear(); b->g_raw.clear(); b->g_packed.clear();
  b->total_raw = 0; b->staged = false; b->nseg = 0;
  return SKB_OK;
}

namespace {
// records given one by one (pointer + length each): what both entry points below come down to
int batch_add_impl(skb_batch* b, const uint8_t* const* recs, const uint64_t* lens, const uint32_t* groups, uint64_t n,
                   uint32_t nthreads) {
  skb_ctx* c = b->ctx;
  if (groups) {
    const uint64_t ng = b->g_first.size();
    if (ng && groups[0] + 1 < ng) return fail(c, SKB_ERR_INVALID_ARG, "groups must continue from the last group");
    for (uint64_t r = 1; r < n; ++r)
      if (groups[r] < groups[r - 1]) return fail(c, SKB_ERR_INVALID_ARG, "groups must be non-decreasing");
  }
  const uint64_t first_rec = b->rec_pos.size();
  const uint64_t first_group_of_call = b->g_first.size();  // (groups == NULL: record r opens group first_group_of_call + r)
  uint64_t cur = b->cur;
  std::vector<uint64_t> before(n + 1, 0);  // bytes of the call's records in front of record r (the threads' shares go by bytes)
  for (uint64_t r = 0; r < n; ++r) {
    const uint64_t len = lens[r];
    before[r + 1] = before[r] + len;
    const uint64_t next = round_up(cur + len + 1, 32);
    const uint64_t gid = groups ? groups[r] : b->g_first.size();
    while (b->g_first.size() <= gid) {  // open new (possibly empty) groups
      b->g_first.push_back(cur / 32); b->g_end.push_back(cur / 32);
      b->g_raw.push_back(0); b->g_packed.push_back(0);
    }
    b->g_end[gid] = next / 32;
    b->g_raw[gid] += len; b->g_packed[gid] += len;
    b->rec_pos.push_back(cur); b->rec_len.push_back(len);
    b->total_raw += len;
    cur = next;
  }
  if (cur / 32 >= 0xFFFFFFF0ull) return fail(c, SKB_ERR_INVALID_ARG, "batch too large (> 2^37 bases)");
  CU(c, b->codes.ensure((cur / 16 + 4) * 4, (b->cur / 16) * 4));
  CU(c, b->nmask.ensure((cur / 32 + 2) * 4, (b->cur / 32) * 4));
  uint32_t* codes = b->codes.as<uint32_t>();
  uint32_t* nmask = b->nmask.as<uint32_t>();
  uint32_t T = nthreads ? nthreads : std::max(1u, std::thread::hardware_concurrency());
  const uint64_t total_bytes = before[n];
  if (total_bytes < (1u << 20) || n < 2) T = 1;
  T = (uint32_t)std::min<uint64_t>(T, n);
  std::vector<uint64_t> kept(b->base_count == SKB_BASES_STRIPPED ? n : 0);
  auto work = [&](uint64_t r0, uint64_t r1) {
    for (uint64_t r = r0; r < r1; ++r) {
      const uint64_t P = b->rec_pos[first_rec + r];
      const uint64_t Pn = (first_rec + r + 1 < b->rec_pos.size()) ? b->rec_pos[first_rec + r + 1] : cur;
      const uint64_t kp = pack_record(recs[r], lens[r], P, Pn, codes, nmask);
      if (!kept.empty()) kept[r] = kp;
    }
  };
  if (T <= 1) {
    work(0, n);
  } else {
    std::vector<std::thread> pool;
    uint64_t r0 = 0;
    for (uint32_t t = 0; t < T; ++t) {
      // split by bytes so long and short records balance
      const uint64_t target = total_bytes * (t + 1) / T;
      uint64_t r1 = (t + 1 == T) ? n : (uint64_t)(std::upper_bound(before.begin() + r0, before.begin() + n, target) - before.begin());
      if (r1 > n) r1 = n;
      if (r1 < r0) r1 = r0;
      pool.emplace_back(work, r0, r1);
      r0 = r1;
    }
    for (auto& th : pool) th.join();
  }
  if (!kept.empty()) {  // total_bases without the removed bytes: take the difference back out of the groups' counts
    uint64_t gid = b->g_first.size() - 1;
    for (uint64_t r = n; r-- > 0;) {
      if (groups) gid = groups[r];
      else gid = first_group_of_call + r;
      const uint64_t removed = lens[r] - kept[r];
      b->g_raw[gid] -= removed;
      b->total_raw -= removed;
    }
  }
  b->cur = cur;
  return SKB_OK;
}
}  // namespace

int skb_batch_add(skb_batch* b, const uint8_t* blob, const uint64_t* offsets, const uint32_t* groups, uint64_t n,
                  uint32_t nthreads) {
  if (!b) return SKB_ERR_INVALID_ARG;
  skb_ctx* c = b->ctx;
  cudaSetDevice(c->device);  // the pinned staging buffers belong to this context's device (callers may use any thread)
  if (b->staged) return fail(c, SKB_ERR_STATE, "batch already staged; clear it before adding");
  if (n == 0) return SKB_OK;
  if (!offsets || (!blob && offsets[n] != offsets[0])) return fail(c, SKB_ERR_INVALID_ARG, "null blob/offsets");
  for (uint64_t r = 0; r < n; ++r)
    if (offsets[r + 1] < offsets[r]) return fail(c, SKB_ERR_INVALID_ARG, "record offsets must be non-decreasing");
  std::vector<const uint8_t*> recs(n);
  std::vector<uint64_t> lens(n);
  for (uint64_t r = 0; r < n; ++r) { recs[r] = blob + offsets[r]; lens[r] = offsets[r + 1] - offsets[r]; }
  return batch_add_impl(b, recs.data(), lens.data(), groups, n, nthreads);
}

int skb_batch_add_records(skb_batch* b, const uint8_t* const* recs, const uint64_t* lens, const uint32_t* groups,
                          uint64_t n, uint32_t nthreads) {
  if (!b) return SKB_ERR_INVALID_ARG;
  skb_ctx* c = b->ctx;
  cudaSetDevice(c->device);
  if (b->staged) return fail(c, SKB_ERR_STATE, "batch already staged; clear it before adding");
  if (n == 0) return SKB_OK;
  if (!recs || !lens) return fail(c, SKB_ERR_INVALID_ARG, "null record pointers/lengths");
  for (uint64_t r = 0; r < n; ++r)
    if (!recs[r] && lens[r]) return fail(c, SKB_ERR_INVALID_ARG, "null record with a non-zero length");
  return batch_add_impl(b, recs, lens, groups, n, nthreads);
}

int skb_batch_set_base_count(skb_batch* b, int mode) {
  if (!b || (mode != SKB_BASES_RAW && mode != SKB_BASES_STRIPPED)) return SKB_ERR_INVALID_ARG;
  if (!b->rec_pos.empty()) return fail(b->ctx, SKB_ERR_STATE, "set the base count mode on an empty batch");
  b->base_count = mode;
  return SKB_OK;
}

uint32_t skb_batch_num_groups(const skb_batch* b) { return b ? (uint32_t)b->g_first.size() : 0; }
uint64_t skb_batch_num_records(const skb_batch* b) { return b ? b->rec_pos.size() : 0; }
uint64_t skb_batch_num_bases(const skb_batch* b) { return b ? b->total_raw : 0; }
uint64_t skb_batch_packed_len(const skb_batch* b) { return b ? b->cur : 0; }
int skb_batch_record_start(const skb_batch* b, uint64_t record, uint64_t* packed_pos, uint64_t* packed_len) {
  if (!b || record >= b->rec_pos.size()) return SKB_ERR_INVALID_ARG;
  if (packed_pos) *packed_pos = b->rec_pos[record];
  if (packed_len) *packed_len = b->rec_len[record];
  return SKB_OK;
}

int skb_batch_stage(skb_batch* b) {
  if (!b) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(b->ctx->device);
  return stage_batch(b);
}

// ---- sketch -----------------------------------------------------------------------------------------------
int skb_sketch(skb_ctx* c, skb_batch* b, uint32_t k, uint32_t s, uint64_t seed, uint64_t* out_hashes,
               uint32_t* out_counts, uint32_t* out_n, uint64_t* out_bases, uint64_t* out_kmers) {
  if (!c || !b || b->ctx != c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (k < 1 || k > SKB_MAX_K) return fail(c, SKB_ERR_UNSUPPORTED_K, "k=%u unsupported (1..%d)", k, SKB_MAX_K);
  if (s < 1) return fail(c, SKB_ERR_INVALID_ARG, "sketch size must be >= 1");
  const uint32_t G = (uint32_t)b->g_first.size();
  if (G == 0) return SKB_OK;
  if (!out_hashes || !out_n) return fail(c, SKB_ERR_INVALID_ARG, "null output");
  if (int rc = use_batch(b)) return rc;
  CU(c, c->sk_hashes.ensure((size_t)G * s * 8));
  CU(c, c->sk_counts.ensure((size_t)G * s * 4));
  std::vector<uint32_t> n;
  std::vector<uint64_t> kmers;
  SelectPlan plan;
  if (int rc = run_hash_select(c, b, k, s, seed, false, 0, c->sk_hashes.as<uint64_t>(), c->sk_counts.as<uint32_t>(), n,
                               kmers, plan))
    return rc;
  CU(c, cudaMemcpyAsync(out_hashes, c->sk_hashes.p, (size_t)G * s * 8, cudaMemcpyDeviceToHost, c->stream));
  if (out_counts) CU(c, cudaMemcpyAsync(out_counts, c->sk_counts.p, (size_t)G * s * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (uint32_t g = 0; g < G; ++g) {
    out_n[g] = n[g];
    if (out_bases) out_bases[g] = b->g_raw[g];
    if (out_kmers) out_kmers[g] = kmers[g];
  }
  return SKB_OK;
}

// ---- reference --------------------------------------------------------------------------------------------
int skb_ref_upload(skb_ctx* c, const uint64_t* hashes, const uint64_t* off, uint32_t n_rows, uint32_t global_row_base) {
  if (!c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  return upload_ref_common(c, hashes, false, off, n_rows, global_row_base);
}
int skb_ref_upload_device(skb_ctx* c, const uint64_t* d_hashes, const uint64_t* off, uint32_t n_rows,
                          uint32_t global_row_base) {
  if (!c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  return upload_ref_common(c, d_hashes, true, off, n_rows, global_row_base);
}
uint32_t skb_ref_rows(const skb_ctx* c) { return c && c->has_ref ? c->n_rows : 0; }

// ---- predict ----------------------------------------------------------------------------------------------
int skb_predict_stream_device(skb_ctx* c, skb_batch* b, uint32_t k, uint32_t s_query, uint64_t seed, uint32_t top,
                              int pad, uint32_t* d_out_idx, uint64_t* d_out_sum) {
  if (!c || !b || b->ctx != c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (b->g_first.size() && (!d_out_idx || !d_out_sum)) return fail(c, SKB_ERR_INVALID_ARG, "null output");
  return predict_device(c, b, k, s_query, seed, top, pad, d_out_idx, d_out_sum);
}

int skb_predict_stream(skb_ctx* c, skb_batch* b, uint32_t k, uint32_t s_query, uint64_t seed, uint32_t top, int pad,
                       uint32_t* out_idx, uint64_t* out_sum) {
  if (!c || !b || b->ctx != c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  const size_t R = b->g_first.size();
  if (R && (!out_idx || !out_sum)) return fail(c, SKB_ERR_INVALID_ARG, "null output");
  if (top < 1 || top > SKB_MAX_TOP) return fail(c, SKB_ERR_INVALID_ARG, "top=%u unsupported (1..%d)", top, SKB_MAX_TOP);
  CU(c, c->out_idx.ensure(std::max<size_t>(4, R * top * 4)));
  CU(c, c->out_sum.ensure(std::max<size_t>(8, R * top * 8)));
  if (int rc = predict_device(c, b, k, s_query, seed, top, pad, c->out_idx.as<uint32_t>(), c->out_sum.as<uint64_t>()))
    return rc;
  if (R) {
    CU(c, cudaMemcpyAsync(out_idx, c->out_idx.p, R * top * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(out_sum, c->out_sum.p, R * top * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  return SKB_OK;
}

// ---- multi-GPU ----------------------------------------------------------------------------------------------
int skb_comm_unique_id(uint8_t* id) {
  if (!id) return SKB_ERR_INVALID_ARG;
  std::string err;
  if (!g_nccl.load(err)) return SKB_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == SKB_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  if (g_nccl.GetUniqueId(&u) != ncclSuccess) return SKB_ERR_COMM;
  std::memcpy(id, &u, sizeof u);
  return SKB_OK;
}

int skb_comm_init(skb_ctx* c, const uint8_t* id, int rank, int world) {
  if (!c || !id || world < 1 || rank < 0 || rank >= world) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (c->comm) return fail(c, SKB_ERR_STATE, "communicator already initialised");
  if (!g_nccl.load(c->err)) return SKB_ERR_COMM;
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof u);
  NC(c, g_nccl.CommInitRank(&c->comm, world, u, rank));
  c->rank = rank; c->world = world;
  c->tau_all_valid = false;
  return SKB_OK;
}

int skb_comm_destroy(skb_ctx* c) {
  if (!c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (c->comm) {
    cudaStreamSynchronize(c->stream);
    g_nccl.CommDestroy(c->comm);
    c->comm = nullptr; c->rank = 0; c->world = 1;
  }
  return SKB_OK;
}

int skb_comm_rank(const skb_ctx* c) { return c ? c->rank : 0; }
int skb_comm_world(const skb_ctx* c) { return c ? c->world : 1; }

void skb_dist_range(uint64_t n, int rank, int world, uint64_t* begin, uint64_t* count) { dist_range(n, rank, world, begin, count); }

int skb_comm_allgather_host(skb_ctx* c, const void* send, void* recv, uint64_t bytes) {
  if (!c || (bytes && (!send || !recv))) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (bytes == 0) return SKB_OK;
  if (!c->comm || c->world <= 1) { std::memcpy(recv, send, bytes); return SKB_OK; }
  CU(c, c->misc.ensure((size_t)bytes * c->world));
  uint8_t* d = c->misc.as<uint8_t>();
  CU(c, cudaMemcpyAsync(d + (size_t)bytes * c->rank, send, bytes, cudaMemcpyHostToDevice, c->stream));
  NC(c, g_nccl.AllGather(d + (size_t)bytes * c->rank, d, bytes, ncclUint8, c->comm, c->stream));
  CU(c, cudaMemcpyAsync(recv, d, (size_t)bytes * c->world, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SKB_OK;
}

int skb_predict_stream_dist_device(skb_ctx* c, skb_batch* b, uint64_t reads_total, uint32_t k, uint32_t s_query,
                                   uint64_t seed, uint32_t top, uint32_t* d_out_idx, uint64_t* d_out_sum) {
  if (!c || !b || b->ctx != c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (reads_total && (!d_out_idx || !d_out_sum)) return fail(c, SKB_ERR_INVALID_ARG, "null output");
  return predict_dist_device(c, b, reads_total, k, s_query, seed, top, d_out_idx, d_out_sum);
}

int skb_predict_stream_dist(skb_ctx* c, skb_batch* b, uint64_t reads_total, uint32_t k, uint32_t s_query, uint64_t seed,
                            uint32_t top, uint32_t* out_idx, uint64_t* out_sum) {
  if (!c || !b || b->ctx != c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (top < 1 || top > SKB_MAX_TOP) return fail(c, SKB_ERR_INVALID_ARG, "top=%u unsupported (1..%d)", top, SKB_MAX_TOP);
  CU(c, c->misc.ensure(std::max<size_t>(16, (size_t)reads_total * top * 12)));
  uint64_t* d_sum = c->misc.as<uint64_t>();
  uint32_t* d_idx = reinterpret_cast<uint32_t*>(d_sum + reads_total * top);
  if (int rc = predict_dist_device(c, b, reads_total, k, s_query, seed, top, d_idx, d_sum)) return rc;
  if (reads_total && out_idx && out_sum) {  // (a rank that does not report may pass NULL)
    CU(c, cudaMemcpyAsync(out_idx, d_idx, (size_t)reads_total * top * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(out_sum, d_sum, (size_t)reads_total * top * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  return SKB_OK;
}

int skb_sums_reset(skb_ctx* c) {
  if (!c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (!c->has_ref) return fail(c, SKB_ERR_NO_REFERENCE, "no reference uploaded");
  CU(c, cudaMemsetAsync(c->sums[c->sums_cur].p, 0, std::max<size_t>(8, (size_t)c->n_rows * 8), c->stream));
  c->tracked_top = 0;
  c->pass_proven = false;
  c->dense_left = c->dense_after_reset;  // passes after the first (always dense) one that are ranked densely too
  return SKB_OK;
}

int skb_sums_download(skb_ctx* c, uint64_t* out) {
  if (!c || !out) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (!c->has_ref) return fail(c, SKB_ERR_NO_REFERENCE, "no reference uploaded");
  if (c->n_rows) CU(c, cudaMemcpyAsync(out, c->sums[c->sums_cur].p, (size_t)c->n_rows * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SKB_OK;
}

int skb_sums_upload(skb_ctx* c, const uint64_t* in) {
  if (!c || !in) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (!c->has_ref) return fail(c, SKB_ERR_NO_REFERENCE, "no reference uploaded");
  if (c->n_rows) CU(c, cudaMemcpyAsync(c->sums[c->sums_cur].p, in, (size_t)c->n_rows * 8, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->tracked_top = 0;
  c->pass_proven = false;
  c->dense_left = c->dense_after_reset;
  return SKB_OK;
}

int skb_set_rank_mode(skb_ctx* c, int mode) {
  if (!c || mode < 0 || mode > 2) return SKB_ERR_INVALID_ARG;
  c->rank_mode = mode;
  c->pass_proven = false;
  return SKB_OK;
}

int skb_set_pass_reads(skb_ctx* c, uint32_t m) {
  if (!c) return SKB_ERR_INVALID_ARG;
  c->pass_user = m;
  choose_pass_max(c);
  c->pass_proven = false;
  return SKB_OK;
}
uint32_t skb_pass_reads(const skb_ctx* c) { return c ? c->pass_max : 0; }

// ---- shared / rank ----------------------------------------------------------------------------------------
int skb_shared_counts(skb_ctx* c, const uint64_t* q_hashes, const uint64_t* q_off, uint32_t Q, uint64_t* out) {
  if (!c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (!c->has_ref) return fail(c, SKB_ERR_NO_REFERENCE, "no reference uploaded");
  if (Q == 0 || c->n_rows == 0) return SKB_OK;
  if (!q_off || !out) return fail(c, SKB_ERR_INVALID_ARG, "null argument");
  const uint64_t qlen = q_off[Q];
  if (qlen && !q_hashes) return fail(c, SKB_ERR_INVALID_ARG, "null query hashes");
  for (uint32_t j = 0; j < Q; ++j) {
    if (q_off[j + 1] < q_off[j]) return fail(c, SKB_ERR_INVALID_ARG, "query offsets must be non-decreasing");
    for (uint64_t x = q_off[j]; x + 1 < q_off[j + 1]; ++x)
      if (!(q_hashes[x] < q_hashes[x + 1])) return fail(c, SKB_ERR_REF_NOT_SORTED, "query sketch %u is not strictly increasing", j);
  }
  const size_t pairs = (size_t)c->n_rows * Q;
  DevBuf dq, dqo, dout;
  int rc = SKB_OK;
  cudaError_t e;
  if ((e = dq.ensure(std::max<uint64_t>(qlen, 1) * 8)) != cudaSuccess || (e = dqo.ensure(((size_t)Q + 1) * 8)) != cudaSuccess ||
      (e = dout.ensure(pairs * 8)) != cudaSuccess) {
    rc = fail(c, SKB_ERR_OOM, "shared buffers: %s", cudaGetErrorString(e));
  } else {
    if (qlen) e = cudaMemcpyAsync(dq.p, q_hashes, qlen * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dqo.p, q_off, ((size_t)Q + 1) * 8, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) rc = fail(c, SKB_ERR_CUDA, "shared: %s", cudaGetErrorString(e));
    if (!rc) {
      { ProfScope ps(c, SKB_K_SHARED, 1);
        skb_launch_shared(ref_view(c), dq.as<uint64_t>(), dqo.as<uint64_t>(), Q, dout.as<unsigned long long>(), c->stream); }
      rc = check_launch(c, "shared");
    }
    if (!rc) {
      e = cudaMemcpyAsync(out, dout.p, pairs * 8, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) rc = fail(c, SKB_ERR_CUDA, "shared: %s", cudaGetErrorString(e));
    }
  }
  dq.release(); dqo.release(); dout.release();
  return rc;
}

int skb_rank_counts(skb_ctx* c, const uint64_t* counts, uint32_t n, uint32_t top, uint32_t* out_idx, uint64_t* out_sum) {
  if (!c || !counts || !out_idx || !out_sum) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (top > n) return fail(c, SKB_ERR_TOP_GT_N, "top (%u) exceeds the number of reference sketches (%u)", top, n);
  if (top < 1 || top > SKB_MAX_TOP) return fail(c, SKB_ERR_INVALID_ARG, "top=%u unsupported (1..%d)", top, SKB_MAX_TOP);
  CU(c, c->misc.ensure((size_t)n * 8 + SKB_MAX_TOP * 16));
  unsigned long long* dv = c->misc.as<unsigned long long>();
  unsigned long long* dos = dv + n;
  uint32_t* doi = reinterpret_cast<uint32_t*>(dos + SKB_MAX_TOP);
  CU(c, cudaMemcpyAsync(dv, counts, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  { ProfScope ps(c, SKB_K_RANK, 1); skb_launch_rank_full(dv, n, top, 0, doi, dos, nullptr, c->stream); }
  if (int rc = check_launch(c, "rank_full")) return rc;
  CU(c, cudaMemcpyAsync(out_idx, doi, (size_t)top * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(out_sum, dos, (size_t)top * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return SKB_OK;
}

int skb_merge_topn_device(skb_ctx* c, const uint32_t* d_idx_parts, const uint64_t* d_sum_parts, uint32_t n_parts,
                          uint64_t n_reads, uint32_t top, uint32_t* d_out_idx, uint64_t* d_out_sum) {
  if (!c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (n_reads == 0) return SKB_OK;
  if (!d_idx_parts || !d_sum_parts || !d_out_idx || !d_out_sum || n_parts == 0 || top == 0)
    return fail(c, SKB_ERR_INVALID_ARG, "bad merge arguments");
  { ProfScope ps(c, SKB_K_MERGE, 1);
    skb_launch_merge_topn(d_idx_parts, reinterpret_cast<const unsigned long long*>(d_sum_parts), n_parts, n_reads, top,
                          d_out_idx, reinterpret_cast<unsigned long long*>(d_out_sum), c->stream); }
  if (int rc = check_launch(c, "merge")) return rc;
  CU(c, cudaStreamSynchronize(c->stream));
  return SKB_OK;
}

// ---- measurement ------------------------------------------------------------------------------------------
int skb_prof_enable(skb_ctx* c, int on) {
  if (!c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  prof_resolve(c);
  c->prof_on = on != 0;
  return SKB_OK;
}
int skb_prof_reset(skb_ctx* c) {
  if (!c) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  prof_resolve(c);
  for (int i = 0; i < SKB_K_COUNT; ++i) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
  return SKB_OK;
}
int skb_prof_get(skb_ctx* c, int id, double* total_ms, uint64_t* launches) {
  if (!c || id < 0 || id >= SKB_K_COUNT) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  prof_resolve(c);
  if (total_ms) *total_ms = c->prof_ms[id];
  if (launches) *launches = c->prof_n[id];
  return SKB_OK;
}
uint64_t skb_launch_count(const skb_ctx* c) { return c ? c->launches : 0; }
int skb_last_predict_stats(const skb_ctx* c, uint64_t* ref_bytes_per_pass, uint64_t* passes, uint64_t* query_hashes,
                           uint64_t* candidates) {
  if (!c) return SKB_ERR_INVALID_ARG;
  if (ref_bytes_per_pass) *ref_bytes_per_pass = c->st_ref_bytes;
  if (passes) *passes = c->st_passes;
  if (query_hashes) *query_hashes = c->st_qhashes;
  if (candidates) *candidates = c->st_cands;
  return SKB_OK;
}

uint64_t skb_last_predict_member_hashes(const skb_ctx* c) {
  if (!c) return 0;
  return c->memb_log2 ? c->st_members : c->st_qhashes;
}

// ---- debug ------------------------------------------------------------------------------------------------
int skb_debug_set(skb_ctx* c, const char* key, uint64_t value) {
  if (!c || !key) return SKB_ERR_INVALID_ARG;
  const std::string k(key);
  if (k == "cand_budget") c->cand_budget = value ? std::max<uint64_t>(64, value) : SKB_CAND_BUDGET;
  else if (k == "trace_passes") c->trace_passes = value != 0;
  else if (k == "dense_after_reset") c->dense_after_reset = (uint32_t)value;
  else if (k == "pipeline") c->pipeline = value != 0;
  else if (k == "stream_ctas") c->stream_ctas = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c->num_sms, value ? value : (uint64_t)c->num_sms));
  else return fail(c, SKB_ERR_INVALID_ARG, "unknown debug key '%s'", key);
  c->pass_proven = false;
  return SKB_OK;
}

int skb_debug_kmer_hashes(skb_ctx* c, skb_batch* b, uint32_t k, uint64_t seed, uint64_t* out_hash, uint8_t* out_valid) {
  if (!c || !b || b->ctx != c || !out_hash || !out_valid) return SKB_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  if (k < 1 || k > SKB_MAX_K) return fail(c, SKB_ERR_UNSUPPORTED_K, "k=%u unsupported (1..%d)", k, SKB_MAX_K);
  if (int rc = use_batch(b)) return rc;
  const uint64_t n = b->cur;
  if (n == 0) return SKB_OK;
  DevBuf dh, dv;
  cudaError_t e;
  int rc = SKB_OK;
  if ((e = dh.ensure(n * 8)) != cudaSuccess || (e = dv.ensure(n)) != cudaSuccess) {
    rc = fail(c, SKB_ERR_OOM, "debug buffers: %s", cudaGetErrorString(e));
  } else {
    e = cudaMemsetAsync(dh.p, 0, n * 8, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(dv.p, 0, n, c->stream);
    if (e != cudaSuccess) { dh.release(); dv.release(); return fail(c, SKB_ERR_CUDA, "hash(dump): %s", cudaGetErrorString(e)); }
    SkbHashArgs ha{};
    ha.pv = view_of(b); ha.k = k; ha.seed = seed;
    ha.dump_hash = dh.as<uint64_t>(); ha.dump_valid = dv.as<uint8_t>();
    { ProfScope ps(c, SKB_K_HASH, 1); skb_launch_hash(ha, c->stream); }
    rc = check_launch(c, "hash(dump)");
    if (!rc) {
      e = cudaMemcpyAsync(out_hash, dh.p, n * 8, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(out_valid, dv.p, n, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) rc = fail(c, SKB_ERR_CUDA, "hash(dump): %s", cudaGetErrorString(e));
    }
  }
  dh.release(); dv.release();
  return rc;
}

}  // extern "C"
