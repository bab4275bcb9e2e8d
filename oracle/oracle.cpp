// oracle.cpp — CPU restatement of the sketchy 0.6.0 MinHash hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it. The product (libsketchy_b200.so) never links, loads or
// falls back to anything in this directory.
//
// PARITY STATUS: "parity unpinned" for everything that lives in the un-vendored crates
// (finch 0.4.1, needletail 0.4.1, murmurhash3 0.0.5 — Cargo.lock:221-233, 361-372, 342-345 of the reference):
// the reference ships no tests, fixtures or golden vectors and cannot be compiled here
// (no cargo/rustc). What IS pinned:
//   * MurmurHash3_x64_128 against the published SMHasher vectors (tests/golden/murmur3_kat.json);
//   * the in-tree code (src/sketchy.rs) is restated line by line with file:line citations below;
//   * the doc invariant "a sketch shares s hashes with itself" (docs/index.md:148-149).
// The crate behaviour is restated from the published algorithm of those crates and each
// such function is tagged [RECALLED]. Every choice that could differ from the real crates is isolated in one
// function so a later check against real sketchy output is a one-line fix.
//
// Build: see oracle/Makefile (g++ -O3 -march=native -shared -fPIC).

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <queue>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------------------
// MurmurHash3_x64_128  — crate murmurhash3 0.0.5 `murmurhash3_x64_128(bytes, seed) -> (u64, u64)` [RECALLED];
// identical to Austin Appleby's public-domain SMHasher reference for seeds < 2^32 (the Rust crate takes a u64
// seed and sets h1 = h2 = seed). SURVEY.md Appendix A.3.
// ---------------------------------------------------------------------------------------------------------
inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

inline uint64_t fmix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

inline uint64_t load_le64(const uint8_t* p) {
  uint64_t v = 0;
  for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
  return v;
}

void murmur3_x64_128(const uint8_t* data, uint64_t len, uint64_t seed, uint64_t out[2]) {
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = seed, h2 = seed;
  const uint64_t nblocks = len / 16;
  for (uint64_t b = 0; b < nblocks; ++b) {
    uint64_t k1 = load_le64(data + 16 * b), k2 = load_le64(data + 16 * b + 8);
    k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ULL;
    k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ULL;
  }
  const uint8_t* tail = data + 16 * nblocks;
  uint64_t k1 = 0, k2 = 0;
  const unsigned t = static_cast<unsigned>(len & 15);
  for (unsigned i = t; i > 8; --i) k2 ^= static_cast<uint64_t>(tail[i - 1]) << (8 * (i - 1 - 8));
  if (t > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
  for (unsigned i = (t > 8 ? 8 : t); i > 0; --i) k1 ^= static_cast<uint64_t>(tail[i - 1]) << (8 * (i - 1));
  if (t > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  h1 ^= len; h2 ^= len;
  h1 += h2; h2 += h1;
  h1 = fmix64(h1); h2 = fmix64(h2);
  h1 += h2; h2 += h1;
  out[0] = h1; out[1] = h2;
}

// finch 0.4.1 sketch_schemes/hashing.rs `hash_f(item, seed) = murmurhash3_x64_128(item, seed).0` [RECALLED].
inline uint64_t hash_f(const uint8_t* kmer, unsigned k, uint64_t seed) {
  uint64_t o[2];
  murmur3_x64_128(kmer, k, seed, o);
  return o[0];
}

// ---------------------------------------------------------------------------------------------------------
// needletail 0.4.1 sequence.rs `normalize(seq, allow_iupac=false)` [RECALLED] — SURVEY.md Appendix A.1.
// Returns 0 for bytes that are removed (whitespace / line endings).
// ---------------------------------------------------------------------------------------------------------
inline uint8_t normalize_byte(uint8_t c) {
  switch (c) {
    case 'A': case 'C': case 'G': case 'T': case 'N': case '-': return c;
    case 'a': return 'A';
    case 'c': return 'C';
    case 'g': return 'G';
    case 't': return 'T';
    case 'u': case 'U': return 'T';
    case '.': case '~': return '-';
    case ' ': case '\t': case '\r': case '\n': return 0;
    default: return 'N';
  }
}

void normalize(const uint8_t* in, uint64_t n, std::vector<uint8_t>& out) {
  out.clear();
  out.reserve(n);
  for (uint64_t i = 0; i < n; ++i) {
    uint8_t c = normalize_byte(in[i]);
    if (c) out.push_back(c);
  }
}

// needletail 0.4.1 sequence.rs `complement` / `reverse_complement` restricted to the normalised alphabet.
inline uint8_t complement(uint8_t c) {
  switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return c;
  }
}

inline bool is_good_base(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// needletail 0.4.1 kmer.rs `CanonicalKmers` iterator [RECALLED] — SURVEY.md Appendix A.2.
// Calls f(pos, kmer_ptr, is_rc) for every window of k good bases; canonical = fwd if fwd < rc
// (byte-lexicographic) else the reverse-complement slice (tie => rc slice).
template <class F>
void canonical_kmers(const std::vector<uint8_t>& buf, unsigned k, F&& f) {
  const uint64_t n = buf.size();
  if (k == 0 || n < k) return;
  std::vector<uint8_t> rc(n);
  for (uint64_t i = 0; i < n; ++i) rc[i] = complement(buf[n - 1 - i]);
  uint64_t pos = 0;
  while (pos + k <= n) {
    // update_position: advance past any window that contains a bad base
    bool ok = true;
    for (uint64_t j = pos + k; j-- > pos;) {
      if (!is_good_base(buf[j])) { pos = j + 1; ok = false; break; }
    }
    if (!ok) continue;
    const uint8_t* fwd = buf.data() + pos;
    const uint8_t* rcv = rc.data() + (n - pos - k);
    if (std::memcmp(fwd, rcv, k) < 0) f(pos, fwd, false);
    else f(pos, rcv, true);
    ++pos;
  }
}

// ---------------------------------------------------------------------------------------------------------
// finch 0.4.1 sketch_schemes/mash.rs `MashSketcher` [RECALLED] — SURVEY.md Appendix A.4. Streaming form
// (max-heap of size s + hash -> (count, extra_count) map), created by SketchParams::create_sketcher()
// at reference src/sketchy.rs:291, 331, 473.
// ---------------------------------------------------------------------------------------------------------
struct MashSketcher {
  uint32_t size;
  unsigned k;
  uint64_t seed;
  uint64_t total_bases = 0, total_kmers = 0;
  std::priority_queue<uint64_t> heap;  // max-heap on hash (finch orders HashedItem by hash only)
  std::unordered_map<uint64_t, std::pair<uint32_t, uint32_t>> counts;
  std::vector<uint8_t> norm;

  MashSketcher(uint32_t s, unsigned kk, uint64_t sd) : size(s), k(kk), seed(sd) {}

  void push(const uint8_t* kmer, bool is_rc) {
    total_kmers += 1;
    const uint64_t h = hash_f(kmer, k, seed);
    const bool add = heap.empty() || h <= heap.top() || heap.size() < size;
    if (!add) return;
    auto it = counts.find(h);
    if (it != counts.end()) {
      it->second.first += 1;
      it->second.second += is_rc ? 1u : 0u;
    } else {
      heap.push(h);
      counts.emplace(h, std::make_pair(1u, is_rc ? 1u : 0u));
      if (heap.size() > size) {
        const uint64_t top = heap.top();
        heap.pop();
        counts.erase(top);
      }
    }
  }

  // `process(record)`: total_bases += record.sequence().len() (raw, before normalisation — `raw_len`
  // lets the caller state what the parser's raw slice length was), normalize(false),
  // reverse complement, canonical k-mers, push.
  void process(const uint8_t* seq, uint64_t len, uint64_t raw_len) {
    total_bases += raw_len;
    normalize(seq, len, norm);
    canonical_kmers(norm, k, [&](uint64_t, const uint8_t* kmer, bool is_rc) { push(kmer, is_rc); });
  }

  // `to_vec()`: heap contents ascending by hash with their counts.
  void to_vec(std::vector<uint64_t>& hashes, std::vector<uint32_t>& cnt) const {
    auto copy = heap;
    hashes.resize(copy.size());
    for (size_t i = copy.size(); i-- > 0;) { hashes[i] = copy.top(); copy.pop(); }
    cnt.resize(hashes.size());
    for (size_t i = 0; i < hashes.size(); ++i) cnt[i] = counts.at(hashes[i]).first;
  }
};

// ---------------------------------------------------------------------------------------------------------
// reference src/sketchy.rs:419-459 `_common_hashes` [IN-TREE]: two-pointer merge; the min_scale tail loops
// (:441-457) only advance indices and never change `common`.
// ---------------------------------------------------------------------------------------------------------
uint64_t common_hashes(const uint64_t* ref, uint64_t nr, const uint64_t* qry, uint64_t nq, double min_scale) {
  uint64_t i = 0, j = 0, common = 0;
  while (i < nq && j < nr) {
    if (qry[i] < ref[j]) i += 1;
    else if (qry[i] > ref[j]) j += 1;
    else { common += 1; i += 1; j += 1; }
  }
  if (min_scale > 0.) {
    const uint64_t max_hash = UINT64_MAX / static_cast<uint64_t>(1.0 / min_scale);
    while (i < nq && qry[i] < max_hash) i += 1;
    while (j < nr && ref[j] < max_hash) j += 1;
  }
  return common;
}

// reference src/sketchy.rs:310, 348: `result_vec.sort_by(|a, b| b.1.cmp(&a.1))` — stable, descending by count;
// :371/:391 slice `[..top]`. Writes the first `top` (index, sum) pairs.
void rank_top(const std::vector<uint64_t>& sums, uint32_t top, uint32_t* out_idx, uint64_t* out_sum) {
  std::vector<uint32_t> order(sums.size());
  for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return sums[a] > sums[b]; });
  for (uint32_t t = 0; t < top; ++t) { out_idx[t] = order[t]; out_sum[t] = sums[order[t]]; }
}

}  // namespace

// =========================================================================================================
// C interface used by the test-suite / bench checker (ctypes). Plain pointers and sizes only.
// =========================================================================================================
extern "C" {

void orc_murmur3_x64_128(const uint8_t* data, uint64_t len, uint64_t seed, uint64_t* out2) {
  murmur3_x64_128(data, len, seed, out2);
}

// normalize(false); returns the output length (out must hold n bytes).
uint64_t orc_normalize(const uint8_t* in, uint64_t n, uint8_t* out) {
  std::vector<uint8_t> v;
  normalize(in, n, v);
  if (!v.empty()) std::memcpy(out, v.data(), v.size());
  return v.size();
}

// All canonical k-mer hashes of one record in emission order; returns the count (out may be null to count).
uint64_t orc_kmer_hashes(const uint8_t* seq, uint64_t len, uint32_t k, uint64_t seed, uint64_t* out,
                         uint64_t cap) {
  std::vector<uint8_t> norm;
  normalize(seq, len, norm);
  uint64_t n = 0;
  canonical_kmers(norm, k, [&](uint64_t, const uint8_t* kmer, bool) {
    if (out && n < cap) out[n] = hash_f(kmer, k, seed);
    ++n;
  });
  return n;
}

uint64_t orc_common_hashes(const uint64_t* ref, uint64_t nr, const uint64_t* qry, uint64_t nq,
                           double min_scale) {
  return common_hashes(ref, nr, qry, nq, min_scale);
}

// Streaming sketcher object (finch MashSketcher).
void* orc_sketcher_new(uint32_t s, uint32_t k, uint64_t seed) { return new MashSketcher(s, k, seed); }
void orc_sketcher_free(void* p) { delete static_cast<MashSketcher*>(p); }
void orc_sketcher_process(void* p, const uint8_t* seq, uint64_t len) {
  static_cast<MashSketcher*>(p)->process(seq, len, len);
}
uint64_t orc_sketcher_to_vec(void* p, uint64_t* hashes, uint32_t* counts, uint64_t cap) {
  std::vector<uint64_t> h;
  std::vector<uint32_t> c;
  static_cast<MashSketcher*>(p)->to_vec(h, c);
  const uint64_t n = std::min<uint64_t>(h.size(), cap);
  for (uint64_t i = 0; i < n; ++i) { hashes[i] = h[i]; if (counts) counts[i] = c[i]; }
  return h.size();
}
void orc_sketcher_totals(void* p, uint64_t* bases, uint64_t* kmers) {
  *bases = static_cast<MashSketcher*>(p)->total_bases;
  *kmers = static_cast<MashSketcher*>(p)->total_kmers;
}

// reference src/sketchy.rs:465-494 `_sketch_files`: one sketcher per file, every record of the file fed to it
// in order, `to_vec()`, totals; files processed by a thread pool (rayon par_iter :470-472), output in input
// order. Records: blob + rec_off[nrec+1]; rec_group[nrec] = file index (non-decreasing).
// out_hashes/out_counts are [ngroups * s]; out_n/out_bases/out_kmers are [ngroups].
int orc_sketch_groups(const uint8_t* blob, const uint64_t* rec_off, const uint32_t* rec_group, uint64_t nrec,
                      uint32_t ngroups, uint32_t k, uint32_t s, uint64_t seed, uint64_t* out_hashes,
                      uint32_t* out_counts, uint32_t* out_n, uint64_t* out_bases, uint64_t* out_kmers,
                      uint32_t nthreads) {
  std::vector<uint64_t> first(ngroups + 1, nrec);
  for (uint64_t r = nrec; r-- > 0;) first[rec_group[r]] = r;
  std::atomic<uint32_t> next{0};
  auto worker = [&]() {
    for (;;) {
      const uint32_t g = next.fetch_add(1);
      if (g >= ngroups) return;
      MashSketcher sk(s, k, seed);
      for (uint64_t r = first[g]; r < nrec && rec_group[r] == g; ++r)
        sk.process(blob + rec_off[r], rec_off[r + 1] - rec_off[r], rec_off[r + 1] - rec_off[r]);
      std::vector<uint64_t> h;
      std::vector<uint32_t> c;
      sk.to_vec(h, c);
      out_n[g] = static_cast<uint32_t>(h.size());
      for (size_t i = 0; i < h.size(); ++i) {
        out_hashes[static_cast<uint64_t>(g) * s + i] = h[i];
        if (out_counts) out_counts[static_cast<uint64_t>(g) * s + i] = c[i];
      }
      out_bases[g] = sk.total_bases;
      out_kmers[g] = sk.total_kmers;
    }
  };
  if (nthreads <= 1) { worker(); return 0; }
  std::vector<std::thread> pool;
  for (uint32_t t = 0; t < nthreads; ++t) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
  return 0;
}

// reference src/sketchy.rs:317-356 `_sum_of_shared_hashes` (streaming predict): for every read a fresh
// sketcher (:331), process (:333), to_vec (:335), `_common_hashes` against every reference (:337-339),
// `sum[i] += shared` (:341), stable sort (:348), first `top` rows (:391). `limit` as at :350-353.
// ref: flat hashes + ref_off[N+1]. reads: blob + read_off[nreads+1]. sums[N] is read-modify-write so a
// caller can continue a stream. Returns the number of reads processed; out_idx/out_sum are [nreads * top].
int64_t orc_predict_stream(const uint64_t* ref, const uint64_t* ref_off, uint32_t N, const uint8_t* blob,
                           const uint64_t* read_off, uint64_t nreads, uint32_t k, uint32_t s_query,
                           uint64_t seed, uint32_t top, uint64_t limit, uint64_t* sums, uint32_t* out_idx,
                           uint64_t* out_sum) {
  if (top > N) return -1;  // the reference panics on result_vec[..top] (src/sketchy.rs:391)
  std::vector<uint64_t> sum(sums, sums + N);
  std::vector<uint64_t> qh;
  std::vector<uint32_t> qc;
  uint64_t read = 1, done = 0;
  for (uint64_t r = 0; r < nreads; ++r) {
    MashSketcher sk(s_query, k, seed);
    sk.process(blob + read_off[r], read_off[r + 1] - read_off[r], read_off[r + 1] - read_off[r]);
    sk.to_vec(qh, qc);
    for (uint32_t i = 0; i < N; ++i) {
      const uint64_t shared =
          common_hashes(ref + ref_off[i], ref_off[i + 1] - ref_off[i], qh.data(), qh.size(), 0.);
      sum[i] += shared;
    }
    rank_top(sum, top, out_idx + r * top, out_sum + r * top);
    ++done;
    read += 1;
    if (read == limit + 1) break;
  }
  std::copy(sum.begin(), sum.end(), sums);
  return static_cast<int64_t>(done);
}

// The same loop with the N merges of every read spread over `nthreads` host threads (contiguous row ranges). NOT what
// the reference does — its predict loop is single-threaded (src/sketchy.rs:328-355) — and only used by bench.py to
// report, next to the faithful single-thread number, what all host cores could do on this path. Results are identical.
int64_t orc_predict_stream_mt(const uint64_t* ref, const uint64_t* ref_off, uint32_t N, const uint8_t* blob,
                              const uint64_t* read_off, uint64_t nreads, uint32_t k, uint32_t s_query,
                              uint64_t seed, uint32_t top, uint64_t limit, uint64_t* sums, uint32_t* out_idx,
                              uint64_t* out_sum, uint32_t nthreads) {
  if (top > N) return -1;
  if (nthreads < 1) nthreads = 1;
  std::vector<uint64_t> sum(sums, sums + N);
  std::vector<uint64_t> qh;
  std::vector<uint32_t> qc;
  uint64_t read = 1, done = 0;
  for (uint64_t r = 0; r < nreads; ++r) {
    MashSketcher sk(s_query, k, seed);
    sk.process(blob + read_off[r], read_off[r + 1] - read_off[r], read_off[r + 1] - read_off[r]);
    sk.to_vec(qh, qc);
    auto part = [&](uint32_t t) {
      const uint32_t lo = static_cast<uint32_t>(static_cast<uint64_t>(N) * t / nthreads);
      const uint32_t hi = static_cast<uint32_t>(static_cast<uint64_t>(N) * (t + 1) / nthreads);
      for (uint32_t i = lo; i < hi; ++i)
        sum[i] += common_hashes(ref + ref_off[i], ref_off[i + 1] - ref_off[i], qh.data(), qh.size(), 0.);
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < nthreads; ++t) pool.emplace_back(part, t);
    part(0);
    for (auto& th : pool) th.join();
    rank_top(sum, top, out_idx + r * top, out_sum + r * top);
    ++done;
    read += 1;
    if (read == limit + 1) break;
  }
  std::copy(sum.begin(), sum.end(), sums);
  return static_cast<int64_t>(done);
}

// reference src/sketchy.rs:281-315 `_shared_hashes` (read-set predict): ONE sketcher over all reads
// (limit check at :297), one to_vec (:302), shared vs every reference (:305-309), stable sort (:310).
// Returns the number of reads consumed (the `read` printed at :312). out_idx/out_sum are [top];
// shared_all (optional) receives all N counts.
int64_t orc_predict_readset(const uint64_t* ref, const uint64_t* ref_off, uint32_t N, const uint8_t* blob,
                            const uint64_t* read_off, uint64_t nreads, uint32_t k, uint32_t s_query,
                            uint64_t seed, uint32_t top, uint64_t limit, uint32_t* out_idx, uint64_t* out_sum,
                            uint64_t* shared_all) {
  if (top > N) return -1;
  MashSketcher sk(s_query, k, seed);
  uint64_t read = 0;
  for (uint64_t r = 0; r < nreads; ++r) {
    sk.process(blob + read_off[r], read_off[r + 1] - read_off[r], read_off[r + 1] - read_off[r]);
    read += 1;
    if (read == limit) break;
  }
  std::vector<uint64_t> qh;
  std::vector<uint32_t> qc;
  sk.to_vec(qh, qc);
  std::vector<uint64_t> shared(N);
  for (uint32_t i = 0; i < N; ++i)
    shared[i] = common_hashes(ref + ref_off[i], ref_off[i + 1] - ref_off[i], qh.data(), qh.size(), 0.);
  rank_top(shared, top, out_idx, out_sum);
  if (shared_all) std::copy(shared.begin(), shared.end(), shared_all);
  return static_cast<int64_t>(read);
}

// reference src/sketchy.rs:238-279 `shared`: all ref x query pairs -> `_common_hashes`; out is [N * Q]
// in the print order of :251-252 (reference outer, query inner).
void orc_shared_matrix(const uint64_t* ref, const uint64_t* ref_off, uint32_t N, const uint64_t* qry,
                       const uint64_t* qry_off, uint32_t Q, uint64_t* out) {
  for (uint32_t i = 0; i < N; ++i)
    for (uint32_t j = 0; j < Q; ++j)
      out[static_cast<uint64_t>(i) * Q + j] = common_hashes(ref + ref_off[i], ref_off[i + 1] - ref_off[i],
                                                            qry + qry_off[j], qry_off[j + 1] - qry_off[j], 0.);
}

}  // extern "C"
