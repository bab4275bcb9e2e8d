// pack_avx2.cpp — AVX2 fast path of the host-side 2-bit packer (compiled with -mavx2; api.cu calls it only after
// __builtin_cpu_supports("avx2")). Same result as the byte-wise packer in api.cu: needletail normalize(false)
// (reference: needletail 0.4 `normalize`, used by src/sketchy.rs:296,330,482) folded into the packing:
//   ACGT / acgt / Uu -> codes 0..3 (U is T), every other byte that is kept -> invalid (breaks k-mers, code bits 0),
//   blank, tab, CR, LF are removed from the sequence — a block holding one of those is left to the byte-wise path.
#include <immintrin.h>
#include <stdint.h>

extern "C" uint64_t skb_pack_blocks_avx2(const uint8_t* s, uint64_t nbytes, uint32_t* codes, uint32_t* nmask) {
  const __m256i up = _mm256_set1_epi8((char)0xDF);
  const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G');
  const __m256i cT = _mm256_set1_epi8('T'), cU = _mm256_set1_epi8('U');
  const __m256i w0 = _mm256_set1_epi8(' '), w1 = _mm256_set1_epi8('\t'), w2 = _mm256_set1_epi8('\r'), w3 = _mm256_set1_epi8('\n');
  const __m256i three = _mm256_set1_epi8(3);
  const __m256i m1 = _mm256_set1_epi16(0x0401);      // byte pair  -> b0 + 4 * b1
  const __m256i m2 = _mm256_set1_epi32(0x00100001);  // word pair  -> w0 + 16 * w1
  const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                        0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
  uint64_t i = 0;
  for (; i + 32 <= nbytes; i += 32) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
    const __m256i ws = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, w0), _mm256_cmpeq_epi8(v, w1)),
                                       _mm256_or_si256(_mm256_cmpeq_epi8(v, w2), _mm256_cmpeq_epi8(v, w3)));
    if (!_mm256_testz_si256(ws, ws)) break;  // a removed byte shifts everything after it: byte-wise path
    const __m256i u = _mm256_and_si256(v, up);
    const __m256i valid = _mm256_or_si256(
        _mm256_or_si256(_mm256_cmpeq_epi8(u, cA), _mm256_cmpeq_epi8(u, cC)),
        _mm256_or_si256(_mm256_cmpeq_epi8(u, cG), _mm256_or_si256(_mm256_cmpeq_epi8(u, cT), _mm256_cmpeq_epi8(u, cU))));
    // ((c >> 1) ^ (c >> 2)) & 3 maps A,C,G,T(U) to 0,1,2,3 in either case
    __m256i x = _mm256_xor_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 2));
    x = _mm256_and_si256(_mm256_and_si256(x, three), valid);
    const __m256i y = _mm256_maddubs_epi16(x, m1);
    const __m256i z = _mm256_madd_epi16(y, m2);
    const __m256i w = _mm256_shuffle_epi8(z, pick);
    codes[(i >> 4)] = (uint32_t)_mm256_extract_epi32(w, 0);
    codes[(i >> 4) + 1] = (uint32_t)_mm256_extract_epi32(w, 4);
    nmask[i >> 5] = ~(uint32_t)_mm256_movemask_epi8(valid);
  }
  return i;
}
