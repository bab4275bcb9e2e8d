// capnp_lite.hpp — the subset of the Cap'n Proto encoding needed for Mash sketch files (.msh): unpacked stream
// framing, struct / list / far / double-far pointers on the read side; a single-segment bump writer on the write side.
// Written from the published encoding specification (capnproto.org/encoding.html); capnp 0.14.3 (the crate the
// reference links, Cargo.lock:107-110) is not available offline.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace capnp_lite {

struct Message {
  std::vector<std::vector<uint64_t>> segs;
};

inline Message parse_stream(const std::vector<uint8_t>& buf) {
  auto rd32 = [&](size_t at) -> uint32_t {
    if (at + 4 > buf.size()) throw std::runtime_error("truncated Cap'n Proto header");
    uint32_t v;
    std::memcpy(&v, buf.data() + at, 4);
    return v;
  };
  const uint32_t nseg = rd32(0) + 1;
  if (nseg > (1u << 20)) throw std::runtime_error("implausible segment count");
  size_t at = 4;
  std::vector<uint32_t> sizes(nseg);
  for (uint32_t i = 0; i < nseg; ++i, at += 4) sizes[i] = rd32(at);
  if (at % 8) at += 4;
  Message m;
  m.segs.resize(nseg);
  for (uint32_t i = 0; i < nseg; ++i) {
    const size_t bytes = (size_t)sizes[i] * 8;
    if (at + bytes > buf.size()) throw std::runtime_error("truncated Cap'n Proto segment");
    m.segs[i].resize(sizes[i]);
    if (bytes) std::memcpy(m.segs[i].data(), buf.data() + at, bytes);
    at += bytes;
  }
  return m;
}

// A resolved object location: segment + word index of the content, plus the pointer word that describes it.
struct Loc {
  uint32_t seg = 0;
  uint64_t word = 0;   // index of the first content word
  uint64_t desc = 0;   // struct or list pointer word describing the content (offset bits meaningless)
  bool null = true;
};

class Reader {
 public:
  explicit Reader(const Message& m) : m_(m) {}
  const Message& msg() const { return m_; }

  uint64_t word(uint32_t seg, uint64_t w) const {
    if (seg >= m_.segs.size() || w >= m_.segs[seg].size()) throw std::runtime_error("Cap'n Proto pointer out of bounds");
    return m_.segs[seg][w];
  }

  // follow the pointer stored at (seg, w)
  Loc follow(uint32_t seg, uint64_t w) const {
    Loc r;
    uint64_t p = word(seg, w);
    if (p == 0) return r;
    r.null = false;
    const unsigned kind = p & 3;
    if (kind == 2) {  // far pointer
      const bool dbl = (p >> 2) & 1;
      const uint64_t off = (p >> 3) & 0x1FFFFFFF;
      const uint32_t tseg = (uint32_t)(p >> 32);
      if (!dbl) {
        const uint64_t pad = word(tseg, off);
        if ((pad & 3) == 2) throw std::runtime_error("far pointer landing pad is a far pointer");
        const int32_t o = (int32_t)((uint32_t)pad) >> 2;
        r.seg = tseg; r.word = off + 1 + o; r.desc = pad;
      } else {
        const uint64_t far2 = word(tseg, off), tag = word(tseg, off + 1);
        if ((far2 & 3) != 2 || ((far2 >> 2) & 1)) throw std::runtime_error("bad double-far landing pad");
        r.seg = (uint32_t)(far2 >> 32); r.word = (far2 >> 3) & 0x1FFFFFFF; r.desc = tag;
      }
    } else if (kind <= 1) {
      const int32_t o = (int32_t)((uint32_t)p) >> 2;
      r.seg = seg; r.word = w + 1 + o; r.desc = p;
    } else {
      throw std::runtime_error("capability pointers are not supported");
    }
    return r;
  }

 private:
  const Message& m_;
};

struct StructView {
  const Reader* r = nullptr;
  uint32_t seg = 0;
  uint64_t data = 0;
  uint32_t dwords = 0, pwords = 0;
  bool null = true;
  uint64_t data_word(uint32_t i) const { return (null || i >= dwords) ? 0 : r->word(seg, data + i); }
  uint32_t u32(uint32_t bit_off, uint32_t dflt = 0) const {
    return (uint32_t)(data_word(bit_off / 64) >> (bit_off % 64)) ^ dflt;
  }
  uint64_t u64(uint32_t bit_off) const { return data_word(bit_off / 64); }
  Loc ptr(uint32_t i) const {
    if (null || i >= pwords) return Loc();
    return r->follow(seg, data + dwords + i);
  }
};

inline StructView as_struct(const Reader& r, const Loc& l) {
  StructView s;
  if (l.null) return s;
  if ((l.desc & 3) != 0) throw std::runtime_error("expected a struct pointer");
  s.r = &r; s.seg = l.seg; s.data = l.word; s.null = false;
  s.dwords = (uint32_t)((l.desc >> 32) & 0xFFFF);
  s.pwords = (uint32_t)((l.desc >> 48) & 0xFFFF);
  return s;
}

struct ListView {
  const Reader* r = nullptr;
  uint32_t seg = 0;
  uint64_t first = 0;    // first element word
  uint32_t esize = 0;    // element size tag
  uint32_t count = 0;    // elements
  uint32_t dwords = 0, pwords = 0;  // composite element shape
  bool null = true;
};

inline ListView as_list(const Reader& r, const Loc& l) {
  ListView v;
  if (l.null) return v;
  if ((l.desc & 3) != 1) throw std::runtime_error("expected a list pointer");
  v.r = &r; v.seg = l.seg; v.null = false;
  v.esize = (uint32_t)((l.desc >> 32) & 7);
  const uint32_t n = (uint32_t)(l.desc >> 35);
  if (v.esize == 7) {
    const uint64_t tag = r.word(l.seg, l.word);
    v.count = (uint32_t)((uint32_t)tag >> 2);
    v.dwords = (uint32_t)((tag >> 32) & 0xFFFF);
    v.pwords = (uint32_t)((tag >> 48) & 0xFFFF);
    v.first = l.word + 1;
    if ((uint64_t)v.count * (v.dwords + v.pwords) > n) throw std::runtime_error("composite list overruns its words");
  } else {
    v.count = n;
    v.first = l.word;
  }
  // the whole list must lie inside its segment BEFORE anybody sizes a buffer from `count` (a truncated or crafted file
  // would otherwise ask for gigabytes)
  static const uint32_t kBits[8] = {0, 1, 8, 16, 32, 64, 64, 0};
  const uint64_t span = v.esize == 7 ? (uint64_t)v.count * (v.dwords + v.pwords) : ((uint64_t)v.count * kBits[v.esize] + 63) / 64;
  if (v.seg >= r.msg().segs.size() || v.first + span > r.msg().segs[v.seg].size())
    throw std::runtime_error("Cap'n Proto pointer out of bounds");
  return v;
}

inline StructView list_struct(const ListView& v, uint32_t i) {
  StructView s;
  if (v.null || v.esize != 7 || i >= v.count) return s;
  s.r = v.r; s.seg = v.seg; s.null = false; s.dwords = v.dwords; s.pwords = v.pwords;
  s.data = v.first + (uint64_t)i * (v.dwords + v.pwords);
  return s;
}

inline std::string read_text(const Reader& r, const Loc& l) {
  const ListView v = as_list(r, l);
  if (v.null || v.count == 0) return std::string();
  if (v.esize != 2) throw std::runtime_error("Text is not a byte list");
  std::string out(v.count - 1, '\0');  // drop the NUL terminator
  for (uint32_t i = 0; i + 1 < v.count; ++i) out[i] = (char)((r.word(v.seg, v.first + i / 8) >> (8 * (i % 8))) & 0xFF);
  return out;
}

template <class T>
inline std::vector<T> read_prims(const Reader& r, const Loc& l, uint32_t expect_esize) {
  const ListView v = as_list(r, l);
  std::vector<T> out;
  if (v.null) return out;
  if (v.esize != expect_esize) throw std::runtime_error("unexpected list element size");
  out.resize(v.count);
  const uint32_t per = 8 / sizeof(T);
  for (uint32_t i = 0; i < v.count; ++i) {
    const uint64_t w = r.word(v.seg, v.first + i / per);
    out[i] = (T)(w >> (8 * sizeof(T) * (i % per)));
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// single-segment writer
// ---------------------------------------------------------------------------------------------------------------
class Writer {
 public:
  Writer() { w_.push_back(0); }  // word 0: root pointer
  uint64_t alloc(uint64_t nwords) {
    const uint64_t at = w_.size();
    w_.resize(at + nwords, 0);
    return at;
  }
  uint64_t& at(uint64_t i) { return w_[i]; }
  void set_struct_ptr(uint64_t ptr_word, uint64_t target, uint32_t dwords, uint32_t pwords) {
    const int64_t off = (int64_t)target - (int64_t)ptr_word - 1;
    w_[ptr_word] = ((uint64_t)((uint32_t)(off << 2))) | ((uint64_t)dwords << 32) | ((uint64_t)pwords << 48);
  }
  void set_list_ptr(uint64_t ptr_word, uint64_t target, uint32_t esize, uint32_t count) {
    const int64_t off = (int64_t)target - (int64_t)ptr_word - 1;
    w_[ptr_word] = ((uint64_t)((uint32_t)(off << 2) | 1u)) | ((uint64_t)esize << 32) | ((uint64_t)count << 35);
  }
  void write_text(uint64_t ptr_word, const std::string& s) {
    const uint32_t n = (uint32_t)s.size() + 1;
    const uint64_t t = alloc((n + 7) / 8);
    std::memcpy(reinterpret_cast<uint8_t*>(&w_[t]), s.data(), s.size());
    set_list_ptr(ptr_word, t, 2, n);
  }
  template <class T>
  void write_prims(uint64_t ptr_word, const std::vector<T>& v, uint32_t esize) {
    const uint64_t bytes = v.size() * sizeof(T);
    const uint64_t t = alloc((bytes + 7) / 8);
    if (bytes) std::memcpy(reinterpret_cast<uint8_t*>(&w_[t]), v.data(), bytes);
    set_list_ptr(ptr_word, t, esize, (uint32_t)v.size());
  }
  std::vector<uint8_t> to_stream() const {
    std::vector<uint8_t> out(8 + w_.size() * 8);
    const uint32_t zero = 0, words = (uint32_t)w_.size();
    std::memcpy(out.data(), &zero, 4);
    std::memcpy(out.data() + 4, &words, 4);
    std::memcpy(out.data() + 8, w_.data(), w_.size() * 8);
    return out;
  }

 private:
  std::vector<uint64_t> w_;
};

}  // namespace capnp_lite
