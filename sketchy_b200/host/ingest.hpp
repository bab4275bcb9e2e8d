// ingest.hpp — host-side record batches for the C ABI and a parallel file reader for `sketchy sketch`.
// The reference sketches its input files on a rayon pool, one finch sketcher per file (src/sketchy.rs:465-494); here the
// hashing is one GPU call for all files, so what is left to spread over the host cores is reading, decompressing and
// splitting the files into records. Files are read in windows (bounded memory), every file of a window on its own
// thread, and handed on in file order: the batch the GPU sees is the same as with one reader.
#pragma once
#include <sys/stat.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "fastx.hpp"

namespace ingest {

// allocator that leaves new elements uninitialised: a blob of a gigabyte is sized in one step and then filled by many
// threads (a zero-filling resize would touch every page on one thread first)
template <class T>
struct DefaultInit : std::allocator<T> {
  template <class U> struct rebind { using other = DefaultInit<U>; };
  DefaultInit() = default;
  template <class U> DefaultInit(const DefaultInit<U>&) {}
  template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
using Bytes = std::vector<uint8_t, DefaultInit<uint8_t>>;

// records of a batch as skb_batch_add takes them: one blob, offsets[n + 1], group of every record
struct Blob {
  Bytes bytes;
  std::vector<uint64_t> off{0};
  std::vector<uint32_t> grp;
  void add(const std::string& s, uint32_t g) {
    bytes.insert(bytes.end(), s.begin(), s.end());
    off.push_back(bytes.size());
    grp.push_back(g);
  }
  void append(const Blob& o) {
    const uint64_t base = bytes.size();
    bytes.insert(bytes.end(), o.bytes.begin(), o.bytes.end());
    for (size_t i = 1; i < o.off.size(); ++i) off.push_back(base + o.off[i]);
    grp.insert(grp.end(), o.grp.begin(), o.grp.end());
  }
  void clear() { bytes.clear(); off.assign(1, 0); grp.clear(); }
  size_t n() const { return grp.size(); }
};

inline uint64_t file_size_or_zero(const std::string& path) {
  struct stat st;
  return ::stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode) ? (uint64_t)st.st_size : 0;
}

// end (exclusive) of the window of files that starts at g0: at least one file, at most `budget` bytes on disk
inline size_t window_end(const std::vector<std::string>& files, size_t g0, uint64_t budget) {
  uint64_t sum = 0;
  size_t g = g0;
  while (g < files.size()) {
    const uint64_t sz = file_size_or_zero(files[g]);
    if (g > g0 && sum + sz > budget) break;
    sum += sz;
    ++g;
  }
  return g;
}

// every record of files[g0, g1), record groups = the file's index in `files`; files are read concurrently (up to
// `nthreads`, 0 = all host threads) and concatenated in file order. An unreadable file throws the reader's error; with
// several bad files the one that comes first in `files` is reported.
inline Blob read_files(const std::vector<std::string>& files, size_t g0, size_t g1, unsigned nthreads = 0) {
  const size_t n = g1 - g0;
  std::vector<Blob> parts(n);
  std::vector<std::exception_ptr> errs(n);
  unsigned T = nthreads ? nthreads : std::max(1u, std::thread::hardware_concurrency());
  T = (unsigned)std::min<size_t>(T, std::max<size_t>(n, 1));
  std::atomic<size_t> next{0};
  auto work = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= n) return;
      try {
        fastx::Reader rd(files[g0 + i]);
        fastx::Record r;
        while (rd.next(r)) parts[i].add(r.seq, (uint32_t)(g0 + i));
      } catch (...) {
        errs[i] = std::current_exception();
      }
    }
  };
  if (T <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; ++t) pool.emplace_back(work);
    for (auto& th : pool) th.join();
  }
  for (size_t i = 0; i < n; ++i)
    if (errs[i]) std::rethrow_exception(errs[i]);
  // concatenation in file order: offsets first, then the bytes of the parts are copied by the same number of threads
  // into a blob that is sized without being touched
  Blob out;
  uint64_t total = 0;
  size_t recs = 0;
  std::vector<uint64_t> at(n + 1, 0);
  for (size_t i = 0; i < n; ++i) { at[i + 1] = at[i] + parts[i].bytes.size(); recs += parts[i].n(); }
  total = at[n];
  out.bytes.resize(total);
  out.off.reserve(recs + 1);
  out.grp.reserve(recs);
  for (size_t i = 0; i < n; ++i) {
    for (size_t j = 1; j < parts[i].off.size(); ++j) out.off.push_back(at[i] + parts[i].off[j]);
    out.grp.insert(out.grp.end(), parts[i].grp.begin(), parts[i].grp.end());
  }
  std::atomic<size_t> nextc{0};
  auto copy = [&]() {
    for (;;) {
      const size_t i = nextc.fetch_add(1);
      if (i >= n) return;
      if (!parts[i].bytes.empty()) std::memcpy(out.bytes.data() + at[i], parts[i].bytes.data(), parts[i].bytes.size());
      Bytes().swap(parts[i].bytes);  // release as we go
    }
  };
  if (T <= 1) {
    copy();
  } else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; ++t) pool.emplace_back(copy);
    for (auto& th : pool) th.join();
  }
  return out;
}

}  // namespace ingest
