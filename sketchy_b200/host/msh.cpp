#include "msh.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>

#include "capnp_lite.hpp"

namespace msh {
using namespace capnp_lite;

static File decode_bytes(const uint8_t* p, size_t n) {
  const Message m = parse_stream(p, n);
  const Reader r(m);
  const StructView root = as_struct(r, r.follow(0, 0));
  if (root.null) throw std::runtime_error("empty Mash file");
  File f;
  f.kmer_size = root.u32(0);
  f.sketch_size = root.u32(64);
  f.hash_seed = root.u32(160, 42);
  Loc rl = root.ptr(3);                 // referenceList
  if (rl.null) rl = root.ptr(0);        // referenceListOld (files written by old Mash versions)
  const StructView list = as_struct(r, rl);
  const ListView refs = as_list(r, list.ptr(0));
  f.sketches.reserve(refs.count);
  for (uint32_t i = 0; i < refs.count; ++i) {
    const StructView e = list_struct(refs, i);
    Sketch s;
    s.name = read_text(r, e.ptr(2));
    s.comment = read_text(r, e.ptr(3));
    s.seq_length = e.u64(64);
    if (s.seq_length == 0) s.seq_length = e.u32(0);
    s.num_valid_kmers = e.u64(128);
    s.hashes = read_prims<uint64_t>(r, e.ptr(5), 5);
    if (s.hashes.empty()) {
      const std::vector<uint32_t> h32 = read_prims<uint32_t>(r, e.ptr(4), 4);
      s.hashes.assign(h32.begin(), h32.end());
    }
    s.counts = read_prims<uint32_t>(r, e.ptr(6), 4);
    f.sketches.push_back(std::move(s));
  }
  return f;
}

File decode(const std::vector<uint8_t>& bytes) { return decode_bytes(bytes.data(), bytes.size()); }

// ---- writer ---------------------------------------------------------------------------------------------------------
// The message is laid out first and streamed out second: nothing but the small head (the MinHash struct, the composite
// list of Reference structs with their pointers) is built in memory, the hash and count lists go from the sketches'
// vectors straight to the sink. Up to `seg_words` words everything lives in ONE segment, in the order
//   root | alphabet | ReferenceList | Reference[n] | per sketch: name, comment, hashes64, counts32
// (what a small file has always looked like here). A larger message cannot: a pointer reaches 2^29 words (4 GiB), so a
// reference of the C3 size (40,000 x 10,000 hashes = 4.8 GB) needs far pointers, as the Cap'n Proto builders of finch /
// Mash produce them. Then segment 0 holds the head and the texts, and the lists follow in data segments of at most
// `seg_words` words (a longer list gets a segment of its own), each list behind a one-word landing pad that the far
// pointer in its Reference struct names. The reader above (and tests/capnp_py.py) follows both forms.
namespace {
using Sink = std::function<void(const void*, size_t)>;

constexpr uint32_t kRefData = 3, kRefPtrs = 7, kRefWords = kRefData + kRefPtrs;

uint64_t text_words(const std::string& s) { return (s.size() + 1 + 7) / 8; }

void put_words(const Sink& sink, const std::vector<uint64_t>& w) { if (!w.empty()) sink(w.data(), w.size() * 8); }
void put_padded(const Sink& sink, const void* p, size_t bytes) {  // list content, zero-padded to a word
  static const uint8_t zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (bytes) sink(p, bytes);
  if (bytes % 8) sink(zeros, 8 - bytes % 8);
}
void put_text(const Sink& sink, const std::string& s) {
  put_padded(sink, s.c_str(), s.size() + 1);  // with its NUL
}

void write_message(const File& f, uint64_t seg_words, const Sink& sink) {
  const uint64_t n = f.sketches.size();
  if (n * kRefWords >= (1ull << 29)) throw std::runtime_error("too many sketches for one Cap'n Proto list");
  for (const Sketch& s : f.sketches)
    if (s.hashes.size() >= (1ull << 29) || s.counts.size() >= (1ull << 29) || s.name.size() >= (1u << 29) || s.comment.size() >= (1u << 29))
      throw std::runtime_error("sketch too large for a Cap'n Proto list");
  // ---- layout of segment 0's head
  const uint64_t root = 1;                      // word 0: root pointer
  const uint64_t alphabet = root + 3 + 4;       // "ACGT\0": one word
  const uint64_t plist = alphabet + 1;          // ReferenceList: 0 data words, 1 pointer
  const uint64_t tag = plist + 1;               // composite list: tag word + n elements
  const uint64_t head_words = tag + 1 + n * kRefWords;
  uint64_t total = head_words;
  for (const Sketch& s : f.sketches)
    total += text_words(s.name) + text_words(s.comment) + s.hashes.size() + (s.counts.size() * 4 + 7) / 8;
  const bool single = total <= seg_words;
  if (!single) {
    uint64_t seg0 = head_words;
    for (const Sketch& s : f.sketches) seg0 += text_words(s.name) + text_words(s.comment);
    if (seg0 >= (1ull << 29)) throw std::runtime_error("sketch names too long for one Cap'n Proto segment");
  }
  std::vector<uint64_t> head(head_words, 0);
  head[0] = struct_ptr_word(0, root, 3, 4);
  // the field set finch's write_mash_file fills [RECALLED]: kmerSize, windowSize = k, minHashesPerWindow, concatenated,
  // alphabet "ACGT", hashSeed (stored XOR its default 42); error, noncanonical, preserveCase stay at their defaults
  head[root + 0] = (uint64_t)f.kmer_size | ((uint64_t)f.kmer_size << 32);        // kmerSize, windowSize
  head[root + 1] = (uint64_t)f.sketch_size | (1ull << 32);                       // minHashesPerWindow, concatenated = true (bit 96)
  head[root + 2] = ((uint64_t)((uint32_t)f.hash_seed ^ 42u)) << 32;              // error 0.0, hashSeed XOR default
  head[root + 3 + 2] = list_ptr_word(root + 3 + 2, alphabet, 2, 5);
  std::memcpy(&head[alphabet], "ACGT", 4);
  head[root + 3 + 3] = struct_ptr_word(root + 3 + 3, plist, 0, 1);
  head[plist] = list_ptr_word(plist, tag, 7, (uint32_t)(n * kRefWords));
  head[tag] = ((uint64_t)((uint32_t)n << 2)) | ((uint64_t)kRefData << 32) | ((uint64_t)kRefPtrs << 48);
  // ---- where everything behind the head goes
  struct DataSeg { uint64_t words = 0; };
  std::vector<DataSeg> data;            // segments 1.. (multi-segment form)
  uint64_t at0 = head_words;            // next free word of segment 0
  auto place_far = [&](uint64_t list_words) {  // -> (segment id, pad word) of a list in a data segment
    if (data.empty() || (data.back().words && data.back().words + 1 + list_words > seg_words)) data.emplace_back();
    const uint64_t pad = data.back().words;
    data.back().words += 1 + list_words;
    return std::make_pair((uint32_t)data.size(), pad);
  };
  for (uint64_t i = 0; i < n; ++i) {
    const Sketch& s = f.sketches[i];
    const uint64_t e = tag + 1 + i * kRefWords;
    head[e + 0] = 0;             // Reference.length (u32, pre-length64 Mash): left 0 as finch does
    head[e + 1] = s.seq_length;  // length64
    head[e + 2] = s.num_valid_kmers;
    const uint64_t hw = s.hashes.size(), cw = (s.counts.size() * 4 + 7) / 8;
    head[e + 3 + 2] = list_ptr_word(e + 3 + 2, at0, 2, (uint32_t)s.name.size() + 1);
    at0 += text_words(s.name);
    head[e + 3 + 3] = list_ptr_word(e + 3 + 3, at0, 2, (uint32_t)s.comment.size() + 1);
    at0 += text_words(s.comment);
    if (single) {
      head[e + 3 + 5] = list_ptr_word(e + 3 + 5, at0, 5, (uint32_t)s.hashes.size());
      at0 += hw;
      head[e + 3 + 6] = list_ptr_word(e + 3 + 6, at0, 4, (uint32_t)s.counts.size());
      at0 += cw;
    } else {
      const auto h = place_far(hw);
      head[e + 3 + 5] = far_ptr_word(h.first, h.second);
      const auto c = place_far(cw);
      head[e + 3 + 6] = far_ptr_word(c.first, c.second);
    }
  }
  for (const DataSeg& d : data)
    if (d.words >= (1ull << 32)) throw std::runtime_error("sketch too large for a Cap'n Proto segment");
  if (at0 >= (1ull << 32)) throw std::runtime_error("sketch file too large for a Cap'n Proto segment");
  // ---- stream framing: segment count - 1, the segments' sizes in words, padded to a word
  std::vector<uint32_t> frame;
  frame.push_back((uint32_t)data.size());
  frame.push_back((uint32_t)at0);
  for (const DataSeg& d : data) frame.push_back((uint32_t)d.words);
  if (frame.size() % 2) frame.push_back(0);
  sink(frame.data(), frame.size() * 4);
  // ---- segment 0
  put_words(sink, head);
  for (const Sketch& s : f.sketches) {
    put_text(sink, s.name);
    put_text(sink, s.comment);
    if (single) {
      put_padded(sink, s.hashes.data(), s.hashes.size() * 8);
      put_padded(sink, s.counts.data(), s.counts.size() * 4);
    }
  }
  // ---- data segments: the same walk as the layout above, so every pad lands where its far pointer says
  if (!single) {
    for (const Sketch& s : f.sketches) {
      const uint64_t pad_h = list_ptr_word(0, 1, 5, (uint32_t)s.hashes.size());  // the content follows its pad at once
      sink(&pad_h, 8);
      put_padded(sink, s.hashes.data(), s.hashes.size() * 8);
      const uint64_t pad_c = list_ptr_word(0, 1, 4, (uint32_t)s.counts.size());
      sink(&pad_c, 8);
      put_padded(sink, s.counts.data(), s.counts.size() * 4);
    }
  }
}

uint64_t segment_words() {
  // 2 GiB segments by default; SKETCHY_B200_MSH_SEGMENT_WORDS lets the tests see the multi-segment form on small files
  if (const char* e = getenv("SKETCHY_B200_MSH_SEGMENT_WORDS")) return std::max<uint64_t>(1, strtoull(e, nullptr, 10));
  return 1ull << 28;
}
}  // namespace

std::vector<uint8_t> encode(const File& f) {
  std::vector<uint8_t> out;
  write_message(f, segment_words(), [&](const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    out.insert(out.end(), b, b + n);
  });
  return out;
}

File read_file(const std::string& path) {
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) throw std::runtime_error("failed to open file");
  // one read into a buffer of the file's size (no growth, no second copy: the decoder works on views of it)
  struct stat st;
  size_t cap = (fstat(fileno(fp), &st) == 0 && S_ISREG(st.st_mode)) ? (size_t)st.st_size + 1 : (1u << 20);
  std::unique_ptr<uint8_t[]> buf(new uint8_t[cap]);
  size_t have = 0;
  for (;;) {
    if (have == cap) {  // not a regular file, or it grew
      std::unique_ptr<uint8_t[]> bigger(new uint8_t[cap * 2]);
      std::memcpy(bigger.get(), buf.get(), have);
      buf.swap(bigger);
      cap *= 2;
    }
    const size_t n = std::fread(buf.get() + have, 1, cap - have, fp);
    if (n == 0) break;
    have += n;
  }
  std::fclose(fp);
  return decode_bytes(buf.get(), have);
}

void write_file(const std::string& path, const File& f) {
  FILE* fp = std::fopen(path.c_str(), "wb");
  if (!fp) throw std::runtime_error("failed to open file");
  std::setvbuf(fp, nullptr, _IOFBF, 4u << 20);
  bool ok = true;
  try {
    write_message(f, segment_words(), [&](const void* p, size_t n) { ok = ok && std::fwrite(p, 1, n, fp) == n; });
  } catch (...) {
    std::fclose(fp);
    throw;
  }
  ok = std::fclose(fp) == 0 && ok;
  if (!ok) throw std::runtime_error("failed to write file");
}

}  // namespace msh
