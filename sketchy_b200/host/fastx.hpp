// fastx.hpp — FASTA/FASTQ (+gzip) record reader for the host side (replaces needletail's parse_fastx_file /
// parse_fastx_stdin at reference src/sketchy.rs:89-92, 474). Records keep their RAW sequence slice: for multi-line FASTA
// the interior line breaks are part of it (needletail `raw_seq`; the library strips whitespace while packing, and
// finch counts total_bases on the raw slice — SURVEY.md Appendix F-3). bz2/xz are not supported in this image.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace fastx {

struct Record {
  std::string id;
  std::string seq;  // raw slice, interior newlines kept, trailing line ending dropped
};

class Reader {
 public:
  explicit Reader(const std::string& path) {
    gz_ = path == "-" ? gzdopen(0, "rb") : gzopen(path.c_str(), "rb");
    if (!gz_) throw std::runtime_error("failed to open Fastx file or record with Needletail");
    gzbuffer(gz_, 1 << 20);
    fill();
    while (pos_ < buf_.size() && (buf_[pos_] == '\n' || buf_[pos_] == '\r')) ++pos_;
    if (pos_ < buf_.size()) {
      if (buf_[pos_] == '>') fasta_ = true;
      else if (buf_[pos_] == '@') fasta_ = false;
      else throw std::runtime_error("failed to open Fastx file or record with Needletail");
    }
  }
  ~Reader() { if (gz_) gzclose(gz_); }
  Reader(const Reader&) = delete;

  bool next(Record& r) {
    std::string line;
    if (!getline(line)) return false;
    while (line.empty()) if (!getline(line)) return false;
    r.id = line.substr(1);
    r.seq.clear();
    if (fasta_) {
      if (line[0] != '>') throw std::runtime_error("failed to open Fastx file or record with Needletail");
      bool first = true;
      while (peek() != -1 && peek() != '>') {
        getline(line);
        if (!first) r.seq.push_back('\n');
        r.seq += line;
        first = false;
      }
      while (!r.seq.empty() && (r.seq.back() == '\n' || r.seq.back() == '\r')) r.seq.pop_back();
    } else {
      if (line[0] != '@') throw std::runtime_error("failed to open Fastx file or record with Needletail");
      if (!getline(r.seq)) throw std::runtime_error("failed to open Fastx file or record with Needletail");
      std::string plus, qual;
      if (!getline(plus) || plus.empty() || plus[0] != '+' || !getline(qual))
        throw std::runtime_error("failed to open Fastx file or record with Needletail");
    }
    return true;
  }

 private:
  void fill() {
    if (eof_) return;
    if (pos_ > 0) { buf_.erase(buf_.begin(), buf_.begin() + pos_); pos_ = 0; }
    const size_t old = buf_.size();
    buf_.resize(old + (1 << 20));
    const int n = gzread(gz_, buf_.data() + old, 1 << 20);
    if (n < 0) throw std::runtime_error("failed to open Fastx file or record with Needletail");
    buf_.resize(old + (size_t)n);
    if (n == 0) eof_ = true;
  }
  int peek() {
    if (pos_ >= buf_.size()) { fill(); if (pos_ >= buf_.size()) return -1; }
    return (unsigned char)buf_[pos_];
  }
  bool getline(std::string& out) {
    out.clear();
    if (peek() == -1) return false;
    for (;;) {
      size_t e = pos_;
      while (e < buf_.size() && buf_[e] != '\n') ++e;
      out.append(buf_.begin() + pos_, buf_.begin() + e);
      if (e < buf_.size()) { pos_ = e + 1; break; }
      pos_ = e;
      if (eof_) break;
      fill();
      if (pos_ >= buf_.size()) break;
    }
    if (!out.empty() && out.back() == '\r') out.pop_back();
    return true;
  }
  gzFile gz_ = nullptr;
  std::vector<char> buf_;
  size_t pos_ = 0;
  bool eof_ = false, fasta_ = true;
};

}  // namespace fastx
