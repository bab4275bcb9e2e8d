// msh.hpp — Mash sketch files (.msh, Cap'n Proto `MinHash` message), the reference sketch format that must stay
// unchanged (reference src/sketchy.rs:158 write_mash_file, :511 read_mash_file). Field layout derived from Mash's
// MinHash.capnp by the Cap'n Proto struct-layout rules [RECALLED: no real file or schema copy is
// available offline — see DESIGN.md §2]:
//   MinHash   : data 3 words {kmerSize u32@0, windowSize u32@32, minHashesPerWindow u32@64, concatenated bit 96,
//               noncanonical bit 97, preserveCase bit 98, error f32@128, hashSeed u32@160 (default 42, stored XOR 42)},
//               pointers {referenceListOld, locusList, alphabet, referenceList}
//   Reference : data 3 words {length u32@0, length64 u64@64, numValidKmers u64@128},
//               pointers {sequence, quality, name, comment, hashes32, hashes64, counts32}
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace msh {

struct Sketch {
  std::string name, comment;
  uint64_t seq_length = 0, num_valid_kmers = 0;
  std::vector<uint64_t> hashes;
  std::vector<uint32_t> counts;
};

struct File {
  uint32_t kmer_size = 16;
  uint32_t sketch_size = 0;  // minHashesPerWindow
  uint64_t hash_seed = 0;
  std::vector<Sketch> sketches;
};

File read_file(const std::string& path);
void write_file(const std::string& path, const File& f);
std::vector<uint8_t> encode(const File& f);
File decode(const std::vector<uint8_t>& bytes);

}  // namespace msh
