// fastx.hpp — FASTA/FASTQ record reader for the host side, plain or gzip / bzip2 / xz compressed (replaces
// needletail's parse_fastx_file / parse_fastx_stdin at reference src/sketchy.rs:89-92, 474; the reference CLI documents
// "Fast{a,q}.{gz,xz,bz}, stdin if not present", src/cli.rs:26,96). Records keep their RAW sequence slice: for
// multi-line FASTA the interior line breaks are part of it (needletail `raw_seq`; the library strips whitespace while
// packing, and finch counts total_bases on the raw slice — SURVEY.md Appendix F-3).
//
// The compression is sniffed from the first bytes, as needletail does (not from the file name), for files and stdin
// alike. gzip goes through zlib's inflate (headers are in the image). This image has no bzlib.h / lzma.h, only the
// runtime libraries, so libbz2.so.1.0 and liblzma.so.5 are loaded with dlopen on first use and driven through their
// stable C ABI declared below; a box without them gets a clear error for such a file instead of a link failure.
#pragma once
#include <dlfcn.h>
#include <fcntl.h>
#include <poll.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace fastx {

struct Record {
  std::string id;
  std::string seq;  // raw slice, interior newlines kept, trailing line ending dropped
};

inline std::runtime_error open_error() {
  return std::runtime_error("failed to open Fastx file or record with Needletail");
}

// ---- compressed byte sources -------------------------------------------------------------------------------------
class RawInput {  // buffered reads from a file descriptor
 public:
  explicit RawInput(int fd, bool own) : fd_(fd), own_(own), buf_(1 << 20) {}
  ~RawInput() { if (own_ && fd_ >= 0) ::close(fd_); }
  RawInput(const RawInput&) = delete;
  // makes at least one byte available unless the input is exhausted; returns the bytes available
  size_t ensure() {
    if (pos_ < end_) return end_ - pos_;
    if (eof_) return 0;
    pos_ = end_ = 0;
    for (;;) {
      const ssize_t n = ::read(fd_, buf_.data(), buf_.size());
      if (n < 0) { if (errno == EINTR) continue; throw open_error(); }
      if (n == 0) eof_ = true;
      end_ = (size_t)n;
      break;
    }
    return end_;
  }
  // the first n bytes of the input without consuming them (called before anything is consumed; fewer at EOF)
  size_t peek(uint8_t* out, size_t n) {
    while (end_ - pos_ < n && !eof_) {
      const ssize_t r = ::read(fd_, buf_.data() + end_, buf_.size() - end_);
      if (r < 0) { if (errno == EINTR) continue; throw open_error(); }
      if (r == 0) { eof_ = true; break; }
      end_ += (size_t)r;
    }
    const size_t m = std::min(n, end_ - pos_);
    std::memcpy(out, buf_.data() + pos_, m);
    return m;
  }
  const uint8_t* data() const { return buf_.data() + pos_; }
  void consume(size_t n) { pos_ += n; }
  // up to cap bytes straight into `out`: what is buffered first, otherwise one read of the descriptor without the
  // stop in this object's buffer (0 = end of input)
  size_t read_into(uint8_t* out, size_t cap) {
    if (pos_ < end_) {
      const size_t n = std::min(cap, end_ - pos_);
      std::memcpy(out, buf_.data() + pos_, n);
      pos_ += n;
      return n;
    }
    if (eof_) return 0;
    for (;;) {
      const ssize_t n = ::read(fd_, out, cap);
      if (n < 0) { if (errno == EINTR) continue; throw open_error(); }
      if (n == 0) eof_ = true;
      return (size_t)n;
    }
  }
  // true when a read would block right now: nothing buffered, not at the end, and the descriptor has nothing to give
  // (a live pipe whose writer is pausing; a regular file is always readable)
  bool would_block() const {
    if (pos_ < end_ || eof_) return false;
    struct pollfd pfd = {fd_, POLLIN, 0};
    return ::poll(&pfd, 1, 0) == 0;
  }

 private:
  int fd_;
  bool own_, eof_ = false;
  std::vector<uint8_t> buf_;
  size_t pos_ = 0, end_ = 0;
};

class Decoder {
 public:
  virtual ~Decoder() {}
  virtual size_t read(uint8_t* out, size_t cap) = 0;  // 0 = end of data
};

class PlainDecoder : public Decoder {
 public:
  explicit PlainDecoder(RawInput& in) : in_(in) {}
  size_t read(uint8_t* out, size_t cap) override { return in_.read_into(out, cap); }

 private:
  RawInput& in_;
};

class GzDecoder : public Decoder {  // multi-member gzip, like gzread / flate2's MultiGzDecoder
 public:
  explicit GzDecoder(RawInput& in) : in_(in) {
    std::memset(&z_, 0, sizeof z_);
    if (inflateInit2(&z_, 16 + MAX_WBITS) != Z_OK) throw open_error();
  }
  ~GzDecoder() override { inflateEnd(&z_); }
  size_t read(uint8_t* out, size_t cap) override {
    z_.next_out = out;
    z_.avail_out = (uInt)cap;
    while (z_.avail_out == cap) {
      const size_t have = in_.ensure();
      if (have == 0) {
        if (!member_done_) throw open_error();  // truncated stream
        break;
      }
      if (member_done_) {  // another member follows
        if (inflateReset(&z_) != Z_OK) throw open_error();
        member_done_ = false;
      }
      z_.next_in = const_cast<Bytef*>(in_.data());
      z_.avail_in = (uInt)have;
      const int rc = inflate(&z_, Z_NO_FLUSH);
      in_.consume(have - z_.avail_in);
      if (rc == Z_STREAM_END) member_done_ = true;
      else if (rc != Z_OK && rc != Z_BUF_ERROR) throw open_error();
    }
    return cap - z_.avail_out;
  }

 private:
  RawInput& in_;
  z_stream z_;
  bool member_done_ = false;
};

// libbz2's public stream ABI (bzlib.h, unchanged since 1.0)
struct BzStream {
  char* next_in; unsigned int avail_in; unsigned int total_in_lo32; unsigned int total_in_hi32;
  char* next_out; unsigned int avail_out; unsigned int total_out_lo32; unsigned int total_out_hi32;
  void* state;
  void* (*bzalloc)(void*, int, int); void (*bzfree)(void*, void*); void* opaque;
};

class Bz2Decoder : public Decoder {
 public:
  explicit Bz2Decoder(RawInput& in) : in_(in) {
    lib_ = dlopen("libbz2.so.1.0", RTLD_NOW);
    if (!lib_) lib_ = dlopen("libbz2.so.1", RTLD_NOW);
    if (!lib_) throw std::runtime_error("bzip2 input needs libbz2.so.1.0, which is not on this machine");
    init_ = (int (*)(BzStream*, int, int))dlsym(lib_, "BZ2_bzDecompressInit");
    run_ = (int (*)(BzStream*))dlsym(lib_, "BZ2_bzDecompress");
    end_ = (int (*)(BzStream*))dlsym(lib_, "BZ2_bzDecompressEnd");
    if (!init_ || !run_ || !end_) throw std::runtime_error("libbz2 lacks the BZ2_bzDecompress entry points");
    std::memset(&s_, 0, sizeof s_);
    if (init_(&s_, 0, 0) != 0) throw open_error();
    live_ = true;
  }
  ~Bz2Decoder() override {
    if (live_) end_(&s_);
    if (lib_) dlclose(lib_);
  }
  size_t read(uint8_t* out, size_t cap) override {
    s_.next_out = reinterpret_cast<char*>(out);
    s_.avail_out = (unsigned)cap;
    while (s_.avail_out == cap) {
      const size_t have = in_.ensure();
      if (have == 0) {
        if (live_) throw open_error();  // truncated stream
        break;
      }
      if (!live_) {  // concatenated streams (bzip2 -c a b, pbzip2)
        std::memset(&s_, 0, sizeof s_);
        if (init_(&s_, 0, 0) != 0) throw open_error();
        s_.next_out = reinterpret_cast<char*>(out);
        s_.avail_out = (unsigned)cap;
        live_ = true;
      }
      s_.next_in = const_cast<char*>(reinterpret_cast<const char*>(in_.data()));
      s_.avail_in = (unsigned)have;
      const int rc = run_(&s_);
      in_.consume(have - s_.avail_in);
      if (rc == 4) {  // BZ_STREAM_END
        const size_t got = cap - s_.avail_out;
        end_(&s_);
        live_ = false;
        if (got) return got;
        s_.avail_out = (unsigned)cap;  // nothing produced yet: look for a following stream
      } else if (rc != 0) {
        throw open_error();
      }
    }
    return cap - s_.avail_out;
  }

 private:
  RawInput& in_;
  void* lib_ = nullptr;
  int (*init_)(BzStream*, int, int) = nullptr;
  int (*run_)(BzStream*) = nullptr;
  int (*end_)(BzStream*) = nullptr;
  BzStream s_;
  bool live_ = false;
};

// liblzma's public stream ABI (lzma/base.h, liblzma.so.5)
struct LzmaStream {
  const uint8_t* next_in; size_t avail_in; uint64_t total_in;
  uint8_t* next_out; size_t avail_out; uint64_t total_out;
  const void* allocator; void* internal;
  void* reserved_ptr1; void* reserved_ptr2; void* reserved_ptr3; void* reserved_ptr4;
  uint64_t reserved_int1; uint64_t reserved_int2; size_t reserved_int3; size_t reserved_int4;
  int reserved_enum1; int reserved_enum2;
};

class XzDecoder : public Decoder {
 public:
  explicit XzDecoder(RawInput& in) : in_(in) {
    lib_ = dlopen("liblzma.so.5", RTLD_NOW);
    if (!lib_) throw std::runtime_error("xz input needs liblzma.so.5, which is not on this machine");
    auto init = (int (*)(LzmaStream*, uint64_t, uint32_t))dlsym(lib_, "lzma_stream_decoder");
    run_ = (int (*)(LzmaStream*, int))dlsym(lib_, "lzma_code");
    end_ = (void (*)(LzmaStream*))dlsym(lib_, "lzma_end");
    if (!init || !run_ || !end_) throw std::runtime_error("liblzma lacks the stream decoder entry points");
    std::memset(&s_, 0, sizeof s_);  // LZMA_STREAM_INIT
    if (init(&s_, UINT64_MAX, 0x08u /* LZMA_CONCATENATED */) != 0) throw open_error();
    live_ = true;
  }
  ~XzDecoder() override {
    if (live_) end_(&s_);
    if (lib_) dlclose(lib_);
  }
  size_t read(uint8_t* out, size_t cap) override {
    if (done_) return 0;
    s_.next_out = out;
    s_.avail_out = cap;
    while (s_.avail_out == cap) {
      const size_t have = in_.ensure();
      s_.next_in = in_.data();
      s_.avail_in = have;
      const int rc = run_(&s_, have == 0 ? 3 /* LZMA_FINISH */ : 0 /* LZMA_RUN */);
      in_.consume(have - s_.avail_in);
      if (rc == 1) { done_ = true; break; }  // LZMA_STREAM_END
      if (rc != 0) throw open_error();       // includes LZMA_BUF_ERROR on a truncated file
    }
    return cap - s_.avail_out;
  }

 private:
  RawInput& in_;
  void* lib_ = nullptr;
  int (*run_)(LzmaStream*, int) = nullptr;
  void (*end_)(LzmaStream*) = nullptr;
  LzmaStream s_;
  bool live_ = false, done_ = false;
};

// what the first bytes of a file say about its compression (needletail sniffs the same way, not by file name)
enum class Packing { Plain, Gzip, Bzip2, Xz };
inline Packing sniff(const uint8_t* magic, size_t got) {
  if (got >= 2 && magic[0] == 0x1F && magic[1] == 0x8B) return Packing::Gzip;
  if (got >= 3 && magic[0] == 'B' && magic[1] == 'Z' && magic[2] == 'h') return Packing::Bzip2;
  if (got >= 6 && std::memcmp(magic, "\xFD" "7zXZ\0", 6) == 0) return Packing::Xz;
  return Packing::Plain;
}
inline std::unique_ptr<Decoder> sniff_decoder(RawInput& in) {
  uint8_t magic[6] = {0, 0, 0, 0, 0, 0};
  const size_t got = in.peek(magic, 6);
  switch (sniff(magic, got)) {
    case Packing::Gzip: return std::unique_ptr<Decoder>(new GzDecoder(in));
    case Packing::Bzip2: return std::unique_ptr<Decoder>(new Bz2Decoder(in));
    case Packing::Xz: return std::unique_ptr<Decoder>(new XzDecoder(in));
    default: return std::unique_ptr<Decoder>(new PlainDecoder(in));
  }
}

// ---- records -----------------------------------------------------------------------------------------------------
// The records of a whole (decoded) FASTA / FASTQ file that is held in memory, as slices of that memory: the same
// records, in the same order and with the same raw sequence bytes, as Reader::next yields for the file, without copying
// a byte. (`sketchy sketch` reads its files this way; the streaming reader below serves stdin and live streams.)
struct Slice { size_t start, len; };
struct ParseState { bool started = false, fasta = true; };

// Parses complete records of p[pos, n) — a piece of the decoded input that starts where the last call stopped —
// appending their sequence slices (offsets into p) and moving pos behind them. With eof = false a record counts as
// complete only when the bytes behind it prove it (a FASTQ record: its four line ends; a FASTA record: the next line
// that starts with '>'), so that the caller can read on and call again with more data behind the same pos. Stops early
// after max_records records or once the slices of this call hold max_seq_bytes bytes.
inline void parse_some(const uint8_t* p, size_t n, bool eof, ParseState& st, size_t& pos, std::vector<Slice>& out,
                       size_t max_records, uint64_t max_seq_bytes) {
  if (!st.started) {
    while (pos < n && (p[pos] == '\n' || p[pos] == '\r')) ++pos;
    if (pos >= n) return;
    st.fasta = p[pos] == '>';
    if (!st.fasta && p[pos] != '@') throw open_error();
    st.started = true;
  }
  // next line = [b, e): without its '\n' and without one '\r' in front of it. 1 = a line, 0 = the input has ended,
  // -1 = the line's end has not arrived yet
  auto line = [&](size_t& cur, size_t& b, size_t& e) {
    if (cur >= n) return eof ? 0 : -1;
    const uint8_t* nl = static_cast<const uint8_t*>(std::memchr(p + cur, '\n', n - cur));
    if (!nl && !eof) return -1;
    b = cur;
    e = nl ? (size_t)(nl - p) : n;
    cur = nl ? e + 1 : n;
    if (e > b && p[e - 1] == '\r') --e;
    return 1;
  };
  uint64_t bytes = 0;
  const size_t first = out.size();
  while (out.size() - first < max_records && bytes < max_seq_bytes) {
    size_t cur = pos, b = 0, e = 0;
    for (;;) {  // blank lines in front of a header
      const size_t c0 = cur;
      const int r = line(cur, b, e);
      if (r <= 0) { pos = c0; return; }
      if (e != b) break;
    }
    const size_t rec_start = b;
    if (st.fasta) {
      if (p[b] != '>') throw open_error();
      const size_t s0 = cur;  // the raw slice runs to the next line that starts with '>' (or the end), line breaks inside kept
      bool complete = false;
      for (;;) {
        if (cur >= n) { complete = eof; break; }
        if (p[cur] == '>') { complete = true; break; }
        const uint8_t* nl = static_cast<const uint8_t*>(std::memchr(p + cur, '\n', n - cur));
        if (!nl) { cur = n; complete = eof; break; }
        cur = (size_t)(nl - p) + 1;
      }
      if (!complete) { pos = rec_start; return; }
      size_t s1 = cur;
      while (s1 > s0 && (p[s1 - 1] == '\n' || p[s1 - 1] == '\r')) --s1;
      out.push_back({s0, s1 - s0});
      bytes += s1 - s0;
    } else {
      if (p[b] != '@') throw open_error();
      size_t sb = 0, se = 0, pb = 0, pe = 0, qb = 0, qe = 0;
      int r = line(cur, sb, se);
      if (r < 0) { pos = rec_start; return; }
      if (r == 0) throw open_error();
      r = line(cur, pb, pe);
      if (r < 0) { pos = rec_start; return; }
      if (r == 0 || pe == pb || p[pb] != '+') throw open_error();
      r = line(cur, qb, qe);
      if (r < 0) { pos = rec_start; return; }
      if (r == 0) throw open_error();
      if (qe - qb != se - sb) throw open_error();  // needletail rejects a record whose quality length differs
      out.push_back({sb, se - sb});
      bytes += se - sb;
    }
    pos = cur;
  }
}

// The records of a whole (decoded) file held in memory.
inline void parse_in_place(const uint8_t* p, size_t n, std::vector<Slice>& out) {
  ParseState st;
  size_t pos = 0;
  parse_some(p, n, true, st, pos, out, SIZE_MAX, UINT64_MAX);
}

class Reader {
 public:
  explicit Reader(const std::string& path) {
    int fd = 0;
    if (path != "-") {
      fd = ::open(path.c_str(), O_RDONLY);
      if (fd < 0) throw open_error();
    }
    in_.reset(new RawInput(fd, path != "-"));
    dec_ = sniff_decoder(*in_);
    fill();
    while (pos_ < len_ && (buf_[pos_] == '\n' || buf_[pos_] == '\r')) ++pos_;
    if (pos_ < len_) {
      if (buf_[pos_] == '>') fasta_ = true;
      else if (buf_[pos_] == '@') fasta_ = false;
      else throw open_error();
    }
  }
  Reader(const Reader&) = delete;

  // true when every byte received so far has been handed out as records and more input has not arrived yet: a caller
  // that batches records (the streaming predict loop) should work on what it has instead of waiting for a full batch.
  // The reference prints a row as soon as a read arrives (src/sketchy.rs:328-355).
  bool input_idle() const { return pos_ >= len_ && !eof_ && in_->would_block(); }

  bool next(Record& r) {
    std::string line;
    if (!getline(line)) return false;
    while (line.empty()) if (!getline(line)) return false;
    r.id = line.substr(1);
    r.seq.clear();
    if (fasta_) {
      if (line[0] != '>') throw open_error();
      bool first = true;
      while (peek() != -1 && peek() != '>') {  // sequence lines go straight into the record, joined by '\n'
        if (!first) r.seq.push_back('\n');
        append_line(r.seq, true);  // the raw slice, '\r' of a CRLF file included: seq_length counts what needletail's slice holds
        first = false;
      }
      while (!r.seq.empty() && (r.seq.back() == '\n' || r.seq.back() == '\r')) r.seq.pop_back();
    } else {
      if (line[0] != '@') throw open_error();
      if (!getline(r.seq)) throw open_error();
      std::string plus, qual;
      if (!getline(plus) || plus.empty() || plus[0] != '+' || !getline(qual)) throw open_error();
      if (qual.size() != r.seq.size()) throw open_error();  // needletail rejects a record whose quality length differs
    }
    return true;
  }

 private:
  static constexpr size_t kChunk = 1 << 20;
  // decoded bytes not yet handed out: buf_[pos_, len_)
  void fill() {
    if (eof_) return;
    if (pos_ > 0) {
      std::memmove(buf_.data(), buf_.data() + pos_, len_ - pos_);
      len_ -= pos_;
      pos_ = 0;
    }
    if (buf_.size() < len_ + kChunk) buf_.resize(len_ + kChunk);
    const size_t n = dec_->read(reinterpret_cast<uint8_t*>(buf_.data()) + len_, kChunk);
    len_ += n;
    if (n == 0) eof_ = true;
  }
  int peek() {
    if (pos_ >= len_) { fill(); if (pos_ >= len_) return -1; }
    return (unsigned char)buf_[pos_];
  }
  // appends the next line, without its terminator, to `out`; false at the end of the input
  bool append_line(std::string& out, bool keep_cr = false) {
    if (peek() == -1) return false;
    for (;;) {
      const char* p = buf_.data() + pos_;
      const size_t avail = len_ - pos_;
      const char* nl = static_cast<const char*>(std::memchr(p, '\n', avail));
      if (nl) {
        out.append(p, (size_t)(nl - p));
        pos_ += (size_t)(nl - p) + 1;
        break;
      }
      out.append(p, avail);
      pos_ = len_;
      if (eof_) break;
      fill();
      if (pos_ >= len_) break;
    }
    if (!keep_cr && !out.empty() && out.back() == '\r') out.pop_back();
    return true;
  }
  bool getline(std::string& out) {
    out.clear();
    return append_line(out);
  }
  std::unique_ptr<RawInput> in_;
  std::unique_ptr<Decoder> dec_;
  std::vector<char> buf_;
  size_t pos_ = 0, len_ = 0;
  bool eof_ = false, fasta_ = true;
};

}  // namespace fastx
