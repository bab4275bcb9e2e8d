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

// Segments are views into the caller's buffer (a reference file of gigabytes is not copied a second time); words are
// read with memcpy, so the buffer needs no alignment.
struct Seg {
  const uint8_t* p = nullptr;
  uint64_t words = 0;
  uint64_t size() const { return words; }
};
struct Message {
  std::vector<Seg> segs;
};

inline Message parse_stream(const uint8_t* buf, size_t n) {
  auto rd32 = [&](size_t at) -> uint32_t {
    if (at + 4 > n) throw std::runtime_error("truncated Cap'n Proto header");
    uint32_t v;
    std::memcpy(&v, buf + at, 4);
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
    if (at + bytes > n) throw std::runtime_error("truncated Cap'n Proto segment");
    m.segs[i].p = buf + at;
    m.segs[i].words = sizes[i];
    at += bytes;
  }
  return m;
}
inline Message parse_stream(const std::vector<uint8_t>& buf) { return parse_stream(buf.data(), buf.size()); }

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
    uint64_t v;
    std::memcpy(&v, m_.segs[seg].p + w * 8, 8);
    return v;
  }
  // the bytes of `nwords` words from (seg, w) on, for bulk copies of list contents (bounds checked like word())
  const uint8_t* bytes(uint32_t seg, uint64_t w, uint64_t nwords) const {
    if (seg >= m_.segs.size() || w + nwords > m_.segs[seg].size()) throw std::runtime_error("Cap'n Proto pointer out of bounds");
    return m_.segs[seg].p + w * 8;
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
  // (little-endian host, as everywhere in this code base: list bytes are element bytes)
  return std::string(reinterpret_cast<const char*>(r.bytes(v.seg, v.first, ((uint64_t)v.count + 7) / 8)), v.count - 1);  // without the NUL
}

template <class T>
inline std::vector<T> read_prims(const Reader& r, const Loc& l, uint32_t expect_esize) {
  const ListView v = as_list(r, l);
  std::vector<T> out;
  if (v.null) return out;
  if (v.esize != expect_esize) throw std::runtime_error("unexpected list element size");
  const uint64_t nbytes = (uint64_t)v.count * sizeof(T);
  const uint8_t* src = r.bytes(v.seg, v.first, (nbytes + 7) / 8);
  out.resize(v.count);
  if (nbytes) std::memcpy(out.data(), src, nbytes);
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// pointer words for a writer (msh.cpp plans the layout of a message and streams it out)
// ---------------------------------------------------------------------------------------------------------------
inline uint64_t struct_ptr_word(uint64_t ptr_word, uint64_t target, uint32_t dwords, uint32_t pwords) {
  const int64_t off = (int64_t)target - (int64_t)ptr_word - 1;
  return ((uint64_t)((uint32_t)(off << 2))) | ((uint64_t)dwords << 32) | ((uint64_t)pwords << 48);
}
inline uint64_t list_ptr_word(uint64_t ptr_word, uint64_t target, uint32_t esize, uint32_t count) {
  const int64_t off = (int64_t)target - (int64_t)ptr_word - 1;
  return ((uint64_t)((uint32_t)(off << 2) | 1u)) | ((uint64_t)esize << 32) | ((uint64_t)count << 35);
}
// far pointer to a one-word landing pad at word `pad` of segment `seg`
inline uint64_t far_ptr_word(uint32_t seg, uint64_t pad) { return 2u | (pad << 3) | ((uint64_t)seg << 32); }

}  // namespace capnp_lite
