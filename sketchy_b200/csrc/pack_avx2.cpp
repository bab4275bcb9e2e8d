// pack_avx2.cpp — AVX2 fast path of the host-side 2-bit packer (compiled with -mavx2; api.cu calls it only after
// __builtin_cpu_supports("avx2")). Same result as the byte-wise packer in api.cu: needletail normalize(false)
// (reference: needletail 0.4 `normalize`, used by src/sketchy.rs:296,330,482) folded into the packing:
//   ACGT / acgt / Uu -> codes 0..3 (U is T), every other byte that is kept -> invalid (breaks k-mers, code bits 0),
//   blank, tab, CR, LF are removed from the sequence — a block holding one of those is left to the byte-wise path.
#include <immintrin.h>
#include <stdint.h>
// classification by two nibble look-ups: bit 0 = {A,C,G}-type low nibble with high nibble 4/6, bit 1 = {T,U}-type with
// high nibble 5/7, bit 2 = blank (0x20), bit 3 = tab/LF/CR (0x09, 0x0A, 0x0D)
extern "C" uint64_t skb_pack_blocks_avx2(const uint8_t* s, uint64_t nbytes, uint32_t* codes, uint32_t* nmask) {
  const __m256i lo_lut = _mm256_setr_epi8(4, 1, 0, 1, 2, 2, 0, 1, 0, 8, 8, 0, 0, 8, 0, 0, 4, 1, 0, 1, 2, 2, 0, 1, 0, 8, 8, 0, 0, 8, 0, 0);
  const __m256i hi_lut = _mm256_setr_epi8(8, 0, 4, 0, 1, 2, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0, 8, 0, 4, 0, 1, 2, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0);
  const __m256i nib = _mm256_set1_epi8(0x0F);
  const __m256i three = _mm256_set1_epi8(3), ws_bits = _mm256_set1_epi8(12);
  const __m256i m1 = _mm256_set1_epi16(0x0401);
  const __m256i m2 = _mm256_set1_epi32(0x00100001);
  const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                        0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
  const __m256i zero = _mm256_setzero_si256();
  uint64_t i = 0;
  for (; i + 32 <= nbytes; i += 32) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
    const __m256i cl = _mm256_shuffle_epi8(lo_lut, _mm256_and_si256(v, nib));
    // a byte >= 0x80 must classify as "other": pshufb zeroes lanes whose index has bit 7 set, so feed it (v >> 4) with bit 7 kept
    const __m256i hidx = _mm256_or_si256(_mm256_and_si256(_mm256_srli_epi16(v, 4), nib), _mm256_and_si256(v, _mm256_set1_epi8((char)0x80)));
    const __m256i ch = _mm256_shuffle_epi8(hi_lut, hidx);
    const __m256i cls = _mm256_and_si256(cl, ch);
    if (!_mm256_testz_si256(cls, ws_bits)) break;  // a removed byte shifts everything after it: byte-wise path
    const __m256i invalid = _mm256_cmpeq_epi8(cls, zero);
    __m256i x = _mm256_xor_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 2));
    x = _mm256_andnot_si256(invalid, _mm256_and_si256(x, three));
    const __m256i y = _mm256_maddubs_epi16(x, m1);
    const __m256i z = _mm256_madd_epi16(y, m2);
    const __m256i w = _mm256_shuffle_epi8(z, pick);
    codes[(i >> 4)] = (uint32_t)_mm256_cvtsi256_si32(w);
    codes[(i >> 4) + 1] = (uint32_t)_mm256_extract_epi32(w, 4);
    nmask[i >> 5] = (uint32_t)_mm256_movemask_epi8(invalid);
  }
  return i;
}
