#include "msh.hpp"

#include <algorithm>
#include <cstdio>
#include <stdexcept>

#include "capnp_lite.hpp"

namespace msh {
using namespace capnp_lite;

File decode(const std::vector<uint8_t>& bytes) {
  const Message m = parse_stream(bytes);
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

std::vector<uint8_t> encode(const File& f) {
  Writer w;
  const uint64_t root = w.alloc(3 + 4);
  w.set_struct_ptr(0, root, 3, 4);
  // the field set finch's write_mash_file fills [RECALLED]: kmerSize, windowSize = k, minHashesPerWindow, concatenated,
  // alphabet "ACGT", hashSeed (stored XOR its default 42); error, noncanonical, preserveCase stay at their defaults
  w.at(root + 0) = (uint64_t)f.kmer_size | ((uint64_t)f.kmer_size << 32);        // kmerSize, windowSize
  w.at(root + 1) = (uint64_t)f.sketch_size | (1ull << 32);                       // minHashesPerWindow, concatenated = true (bit 96)
  w.at(root + 2) = ((uint64_t)((uint32_t)f.hash_seed ^ 42u)) << 32;              // error 0.0, hashSeed XOR default
  w.write_text(root + 3 + 2, "ACGT");                                            // alphabet
  const uint64_t plist = w.alloc(1);  // ReferenceList: 0 data words, 1 pointer
  w.set_struct_ptr(root + 3 + 3, plist, 0, 1);
  const uint32_t n = (uint32_t)f.sketches.size();
  const uint32_t ew = 3 + 7;
  const uint64_t tag = w.alloc(1 + (uint64_t)n * ew);
  w.at(tag) = ((uint64_t)(n << 2)) | ((uint64_t)3 << 32) | ((uint64_t)7 << 48);
  w.set_list_ptr(plist, tag, 7, n * ew);
  for (uint32_t i = 0; i < n; ++i) {
    const Sketch& s = f.sketches[i];
    const uint64_t e = tag + 1 + (uint64_t)i * ew;
    w.at(e + 0) = 0;             // Reference.length (u32, pre-length64 Mash): left 0 as finch does
    w.at(e + 1) = s.seq_length;  // length64
    w.at(e + 2) = s.num_valid_kmers;
    w.write_text(e + 3 + 2, s.name);
    w.write_text(e + 3 + 3, s.comment);
    w.write_prims<uint64_t>(e + 3 + 5, s.hashes, 5);
    w.write_prims<uint32_t>(e + 3 + 6, s.counts, 4);
  }
  return w.to_stream();
}

File read_file(const std::string& path) {
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) throw std::runtime_error("failed to open file");
  std::vector<uint8_t> buf;
  uint8_t tmp[1 << 16];
  size_t n;
  while ((n = std::fread(tmp, 1, sizeof tmp, fp)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  std::fclose(fp);
  return decode(buf);
}

void write_file(const std::string& path, const File& f) {
  const std::vector<uint8_t> bytes = encode(f);
  FILE* fp = std::fopen(path.c_str(), "wb");
  if (!fp) throw std::runtime_error("failed to open file");
  std::fwrite(bytes.data(), 1, bytes.size(), fp);
  std::fclose(fp);
}

}  // namespace msh
