// ingest.hpp — the host side of `sketchy sketch` and `sketchy predict` between the input files and the C ABI.
// The reference sketches its input files on a rayon pool, one finch sketcher per file (src/sketchy.rs:465-494), and
// predicts read after read (:317-356); here the hashing is one GPU call per window of files / chunk of reads, so what
// is left to the host cores is reading, decompressing, splitting into records and 2-bit packing. This header holds:
//   Channel      a bounded queue between the stages of those pipelines (reader -> packer -> GPU caller -> printer)
//   Files        a window of files held in memory, records as slices of it (load_files; the `sketch` path)
//   ChunkReader  a stream of reads decoded block by block, records as slices of the chunk's buffer (the `predict` path)
//   Blob         the copying form of a window (read_files, through the streaming fastx::Reader): the reference the
//                tests hold the two slice forms against
#pragma once
#include <sys/mman.h>
#include <sys/stat.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
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

// records of a batch as skb_batch_add takes them: one blob, offsets[n + 1], group of every record (the copying form:
// read_files below builds it through the streaming reader; the CLI itself works on slices — Files / Chunk further down —
// and the tests hold the two against each other)
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

// ---- a bounded queue between the stages of a host pipeline (reader -> packer -> GPU caller -> printer) -----------------
// push blocks while the queue is full, pop while it is empty; close() = no more items (pop drains what is queued, then
// returns false); abort() = stop now (both ends return false at once; used when a stage failed).
template <class T>
class Channel {
 public:
  explicit Channel(size_t cap) : cap_(cap) {}
  bool push(T v) {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return q_.size() < cap_ || closed_; });
    if (closed_) return false;
    q_.push_back(std::move(v));
    cv_.notify_all();
    return true;
  }
  bool pop(T& v) {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return !q_.empty() || closed_; });
    if (q_.empty()) return false;
    v = std::move(q_.front());
    q_.pop_front();
    cv_.notify_all();
    return true;
  }
  void close() { std::lock_guard<std::mutex> l(m_); closed_ = true; cv_.notify_all(); }
  void abort() { std::lock_guard<std::mutex> l(m_); closed_ = true; q_.clear(); cv_.notify_all(); }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<T> q_;
  size_t cap_;
  bool closed_ = false;
};

// ---- files held in memory, records as slices (the `sketch` path) ------------------------------------------------------
// A window of files as they are on disk: plain files are read, every file by one read loop, straight into one arena
// that is reused from window to window (no page faults after the first window); compressed files are decoded into a
// buffer of their own. Records are slices of those buffers (fastx::parse_in_place), handed to skb_batch_add_records as
// they lie: the only copy of a plain FASTA file between the page cache and the 2-bit packer is the read itself.
struct Files {
  Bytes arena;                  // plain files back to back
  std::vector<Bytes> decoded;   // file i of the window when it is compressed, not a regular file, or grew while being read
  std::vector<const uint8_t*> rec;
  std::vector<uint64_t> len;
  std::vector<uint32_t> grp;    // the file's index in the window
  size_t n() const { return rec.size(); }
  void clear() { rec.clear(); len.clear(); grp.clear(); decoded.clear(); }
};

// a hint for a large buffer that is about to be touched for the first time (ignored where transparent huge pages are off)
inline void huge_pages(void* p, size_t n) {
  const uintptr_t lo = ((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
  const uintptr_t hi = ((uintptr_t)p + n) & ~(uintptr_t)((2u << 20) - 1);
  if (hi > lo) ::madvise((void*)lo, hi - lo, MADV_HUGEPAGE);
}

inline void decode_all(fastx::RawInput& in, Bytes& out, bool plain = false) {
  std::unique_ptr<fastx::Decoder> dec = plain ? std::unique_ptr<fastx::Decoder>(new fastx::PlainDecoder(in)) : fastx::sniff_decoder(in);
  size_t have = out.size();
  for (;;) {
    if (out.size() < have + (4u << 20)) out.resize(std::max<size_t>(out.size() * 2, have + (4u << 20)));
    const size_t got = dec->read(out.data() + have, out.size() - have);
    if (got == 0) break;
    have += got;
  }
  out.resize(have);
}

// files[g0, g1) into `w` (whose arena is reused); up to `nthreads` files are read and split into records at a time
// (0 = all host threads). Errors as read_files reports them.
inline void load_files(const std::vector<std::string>& files, size_t g0, size_t g1, unsigned nthreads, Files& w) {
  const size_t n = g1 - g0;
  w.clear();
  w.decoded.resize(n);
  std::vector<uint64_t> at(n + 1, 0);
  std::vector<uint64_t> size(n, 0);
  for (size_t i = 0; i < n; ++i) { size[i] = file_size_or_zero(files[g0 + i]); at[i + 1] = at[i] + size[i]; }
  if (w.arena.size() < at[n]) {
    Bytes().swap(w.arena);               // nothing of the last window is kept: no copy into the larger arena
    w.arena.resize(at[n] + at[n] / 8);   // some room: windows of one budget differ by a file's size at most
    // first touch of a fresh arena is the dearest part of the first window: ask for huge pages (a hint; ignored where
    // transparent huge pages are off)
    huge_pages(w.arena.data(), w.arena.size());
  }
  std::vector<std::vector<fastx::Slice>> slices(n);
  std::vector<const uint8_t*> base(n, nullptr);
  std::vector<std::exception_ptr> errs(n);
  unsigned T = nthreads ? nthreads : std::max(1u, std::thread::hardware_concurrency());
  T = (unsigned)std::min<size_t>(T, std::max<size_t>(n, 1));
  std::atomic<size_t> next{0};
  auto work = [&]() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= n) return;
      try {
        const std::string& path = files[g0 + i];
        const int fd = path == "-" ? 0 : ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw fastx::open_error();
        uint8_t magic[6];
        ssize_t got = size[i] ? ::pread(fd, magic, 6, 0) : -1;  // size 0: not a regular file (or empty): the streaming way
        uint64_t have = 0;
        bool in_arena = got >= 0 && fastx::sniff(magic, (size_t)got) == fastx::Packing::Plain;
        if (in_arena) {
          uint8_t* dst = w.arena.data() + at[i];
          while (have < size[i]) {
            const ssize_t r = ::read(fd, dst + have, size[i] - have);
            if (r < 0) { if (errno == EINTR) continue; ::close(fd); throw fastx::open_error(); }
            if (r == 0) break;  // shorter than stat said: what is there
            have += (uint64_t)r;
          }
          uint8_t extra;
          ssize_t r;
          do r = ::read(fd, &extra, 1); while (r < 0 && errno == EINTR);
          if (r == 1) {  // the file grew since stat: the rest goes behind a copy of what was read
            in_arena = false;
            Bytes& d = w.decoded[i];
            d.resize(have + 1);
            std::memcpy(d.data(), dst, have);
            d[have] = extra;
            fastx::RawInput in(fd, path != "-");
            decode_all(in, d, true);  // plain bytes from here on (the magic was sniffed above)
          } else if (path != "-") {
            ::close(fd);
          }
          if (in_arena) { base[i] = dst; fastx::parse_in_place(dst, have, slices[i]); }
        } else {
          fastx::RawInput in(fd, path != "-");
          decode_all(in, w.decoded[i]);
        }
        if (!in_arena) { base[i] = w.decoded[i].data(); fastx::parse_in_place(base[i], w.decoded[i].size(), slices[i]); }
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
  size_t recs = 0;
  for (size_t i = 0; i < n; ++i) recs += slices[i].size();
  w.rec.reserve(recs); w.len.reserve(recs); w.grp.reserve(recs);
  for (size_t i = 0; i < n; ++i)
    for (const fastx::Slice& sl : slices[i]) {
      w.rec.push_back(base[i] + sl.start);
      w.len.push_back(sl.len);
      w.grp.push_back((uint32_t)i);
    }
}

// ---- a stream of reads in chunks, records as slices (the streaming `predict` path) ----------------------------------------
// The input (file, stdin, pipe; plain or compressed) is decoded block by block straight into a chunk's buffer and split
// into records where it lies. A chunk ends after max_reads reads or max_seq_bytes sequence bytes — a rule of the content
// alone, so that the ranks of a multi-GPU run cut the same file into the same chunks — or, on request, when a live input
// pauses with nothing half-read (a sequencer writing to a pipe: what has arrived is predicted at once, as the
// reference prints a row per read as it comes, src/sketchy.rs:328-355). The unfinished tail moves to the next chunk.
struct Chunk {
  Bytes buf;
  std::vector<const uint8_t*> rec;
  std::vector<uint64_t> len;
  size_t n() const { return rec.size(); }
};

class ChunkReader {
 public:
  explicit ChunkReader(const std::string& path, size_t block = 1u << 20) : block_(block) {
    int fd = 0;
    if (path != "-") {
      fd = ::open(path.c_str(), O_RDONLY);
      if (fd < 0) throw fastx::open_error();
    }
    in_.reset(new fastx::RawInput(fd, path != "-"));
    dec_ = fastx::sniff_decoder(*in_);
  }
  // false: the input has ended and the chunk is empty
  // A malformed record ends the stream the way it does in the reference, which handles record after record: the reads
  // in front of it still come out (as a last, short chunk), the call after that throws the reader's error.
  bool next(Chunk& c, size_t max_reads, uint64_t max_seq_bytes, bool stop_when_idle) {
    c.rec.clear(); c.len.clear();
    slices_.clear();
    if (pending_) std::rethrow_exception(pending_);
    size_t have = carry_.size(), pos = 0;
    // a chunk's buffer is sized once for the usual case (FASTQ: two bytes of input per base) and reused as it circulates
    const size_t want = have + block_ + (size_t)std::min<uint64_t>(2 * max_seq_bytes, 192ull << 20);
    if (c.buf.size() < want) {
      Bytes().swap(c.buf);  // nothing in it is live: no copy into the larger buffer
      c.buf.resize(want);
      huge_pages(c.buf.data(), c.buf.size());
    }
    if (have) std::memcpy(c.buf.data(), carry_.data(), have);
    carry_.clear();
    uint64_t seq = 0;
    size_t counted = 0;
    for (;;) {
      if (max_reads > slices_.size() && seq < max_seq_bytes) {
        try {
          fastx::parse_some(c.buf.data(), have, eof_, st_, pos, slices_, max_reads - slices_.size(), max_seq_bytes - seq);
        } catch (...) {
          if (slices_.empty()) throw;
          pending_ = std::current_exception();
          pos = have;  // nothing behind the bad record is used
          break;
        }
      }
      for (; counted < slices_.size(); ++counted) seq += slices_[counted].len;
      if (slices_.size() >= max_reads || seq >= max_seq_bytes || eof_) break;
      if (stop_when_idle && !slices_.empty() && pos == have && in_->would_block()) break;
      if (c.buf.size() - have < block_) {  // grow keeping what is there (a record longer than the buffer, or a long chunk)
        Bytes bigger;
        bigger.resize(std::max(c.buf.size() * 2, have + block_));
        std::memcpy(bigger.data(), c.buf.data(), have);
        bigger.swap(c.buf);
      }
      const size_t got = dec_->read(c.buf.data() + have, block_);
      if (got == 0) eof_ = true;
      have += got;
    }
    carry_.assign(c.buf.data() + pos, c.buf.data() + have);
    c.rec.reserve(slices_.size()); c.len.reserve(slices_.size());
    for (const fastx::Slice& sl : slices_) { c.rec.push_back(c.buf.data() + sl.start); c.len.push_back(sl.len); }
    return !slices_.empty();
  }

 private:
  size_t block_;
  std::unique_ptr<fastx::RawInput> in_;
  std::unique_ptr<fastx::Decoder> dec_;
  fastx::ParseState st_;
  std::vector<fastx::Slice> slices_;
  Bytes carry_;
  bool eof_ = false;
  std::exception_ptr pending_;
};

}  // namespace ingest
