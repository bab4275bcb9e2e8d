"""Host-side 2-bit packer block paths (sketchy_b200/csrc/pack_avx2.cpp, pack_avx512.cpp) against a plain restatement of
the rule they implement — needletail `normalize(false)` folded into the packing (reference: reached through
`sketcher.process(record)`, src/sketchy.rs:296, 333, 477): ACGT / acgt / Uu -> codes 0..3, blank / tab / CR / LF
removed (a block that holds one is left to the byte-wise path: the block functions must stop in front of it), every
other byte kept as an invalid position. CPU only: the two files are compiled into a small harness; a path the host CPU
lacks is skipped. The same functions run inside skb_batch_add on the GPU box, where the `-m gpu` tests compare every
k-mer hash with the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "sketchy_b200", "csrc")
HARNESS = r'''
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <random>
#include <vector>
extern "C" uint64_t skb_pack_blocks_avx2(const uint8_t*, uint64_t, uint32_t*, uint32_t*);
extern "C" uint64_t skb_pack_blocks_avx512(const uint8_t*, uint64_t, uint32_t*, uint32_t*);
typedef uint64_t (*fn)(const uint8_t*, uint64_t, uint32_t*, uint32_t*);
static uint8_t cls[256];
// the rule, 32 bytes at a time
static uint64_t plain(const uint8_t* s, uint64_t n, uint32_t* codes, uint32_t* nmask) {
  uint64_t i = 0;
  for (; i + 32 <= n; i += 32) {
    uint32_t c[2] = {0, 0}, m = 0;
    for (int j = 0; j < 32; ++j) {
      const uint8_t v = cls[s[i + j]];
      if (v == 5) return i;
      if (v < 4) c[j >> 4] |= (uint32_t)v << (2 * (j & 15)); else m |= 1u << j;
    }
    codes[i >> 4] = c[0]; codes[(i >> 4) + 1] = c[1]; nmask[i >> 5] = m;
  }
  return i;
}
int main(int argc, char** argv) {
  memset(cls, 4, 256);
  cls['A'] = cls['a'] = 0; cls['C'] = cls['c'] = 1; cls['G'] = cls['g'] = 2; cls['T'] = cls['t'] = cls['U'] = cls['u'] = 3;
  cls[' '] = cls['\t'] = cls['\r'] = cls['\n'] = 5;
  std::vector<std::pair<const char*, fn>> paths;
  if (__builtin_cpu_supports("avx2")) paths.push_back({"avx2", skb_pack_blocks_avx2});
  if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi"))
    paths.push_back({"avx512", skb_pack_blocks_avx512});
  std::mt19937_64 rng(7);
  for (int t = 0; t < 6000; ++t) {
    const size_t n = t < 300 ? (size_t)t : rng() % 1500;
    std::vector<uint8_t> s(n + 1);
    const int mode = t % 5;
    for (size_t i = 0; i < n; ++i) {
      const uint64_t r = rng();
      uint8_t b = (uint8_t)"ACGTacgtUuNnRYKM-.*"[r % 19];
      if (mode == 1 && r % 97 == 0) b = (uint8_t)(r >> 8);                 // any byte, rarely (incl. >= 0x80)
      if (mode == 2 && r % 211 == 0) b = (uint8_t)"\n\r \t"[(r >> 8) % 4];  // removed bytes, rarely
      if (mode == 3) b = (uint8_t)(r >> 16);                               // noise
      if (mode == 4) b = (uint8_t)"ACGT"[r & 3];                           // clean
      s[i] = b;
    }
    std::vector<uint32_t> c0(n / 16 + 8, 0xA5A5A5A5u), m0(n / 32 + 8, 0xA5A5A5A5u);
    const uint64_t a = plain(s.data(), n, c0.data(), m0.data());
    for (auto& p : paths) {
      std::vector<uint32_t> c1(n / 16 + 8, 0xA5A5A5A5u), m1(n / 32 + 8, 0xA5A5A5A5u);
      const uint64_t b = p.second(s.data(), n, c1.data(), m1.data());
      // the same bytes consumed, the same words for them, nothing written behind them
      if (a != b || c0 != c1 || m0 != m1) { printf("%s differs: case %d, %zu bytes, consumed %llu vs %llu\n", p.first, t, n, (unsigned long long)b, (unsigned long long)a); return 1; }
    }
  }
  for (auto& p : paths) printf("%s ok\n", p.first);
  return 0;
}
'''


def test_block_packers_equal_the_rule(tmp_path):
    (tmp_path / "h.cpp").write_text(HARNESS)
    objs = []
    for name, flags in (("pack_avx2", ["-mavx2"]), ("pack_avx512", ["-mavx512f", "-mavx512bw", "-mavx512vbmi"])):
        o = str(tmp_path / f"{name}.o")
        subprocess.check_call(["g++", "-O3", *flags, "-std=c++17", "-Wall", "-Werror", "-c", os.path.join(CSRC, f"{name}.cpp"), "-o", o])
        objs.append(o)
    exe = str(tmp_path / "h")
    subprocess.check_call(["g++", "-O1", "-std=c++17", str(tmp_path / "h.cpp"), *objs, "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    if not p.stdout.strip():
        pytest.skip("this CPU has neither AVX2 nor AVX-512 VBMI: only the byte-wise packer would run")


def test_avx512_object_runs_nothing_at_load_time(tmp_path):
    """The AVX-512 file is linked into a library that also loads on CPUs without AVX-512: it must hold no static
    initialiser (its table is a literal)."""
    o = str(tmp_path / "p.o")
    subprocess.check_call(["g++", "-O3", "-mavx512f", "-mavx512bw", "-mavx512vbmi", "-std=c++17", "-c", os.path.join(CSRC, "pack_avx512.cpp"), "-o", o])
    sections = subprocess.run(["objdump", "-h", o], capture_output=True, text=True).stdout
    assert ".init_array" not in sections and ".ctors" not in sections
