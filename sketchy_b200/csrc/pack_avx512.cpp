// pack_avx512.cpp — AVX-512 (F + BW + VBMI) fast path of the host-side 2-bit packer: 64 bases per step. Compiled with
// -mavx512f -mavx512bw -mavx512vbmi; api.cu calls it only after the CPU reports those features. Same result as the
// byte-wise packer in api.cu and the AVX2 path in pack_avx2.cpp (needletail normalize(false) folded into the packing:
// reference needletail 0.4 `normalize`, reached through src/sketchy.rs:296, 333, 477):
//   ACGT / acgt / Uu -> codes 0..3 (U is T), every other byte that is kept -> invalid (breaks k-mers, code bits 0),
//   blank, tab, CR, LF are removed from the sequence — a block holding one of those is left to the narrower paths.
// One VPERMI2B classifies all 64 bytes through a 128-entry table (bytes >= 0x80 are invalid by their sign bit), two
// multiply-adds and a VPMOVDB squeeze the 2-bit codes into 16 bytes, the invalid mask is a compare's k-register.
#include <immintrin.h>
#include <stdint.h>

extern "C" uint64_t skb_pack_blocks_avx2(const uint8_t* s, uint64_t nbytes, uint32_t* codes, uint32_t* nmask);

namespace {
// class of a byte < 0x80: 0..3 = the base's code (A a, C c, G g, T t U u), 4 = kept but not a base, 5 = removed (blank,
// tab, CR, LF). A literal table: this file is built with AVX-512 flags, so it must not run any code at load time on
// a CPU that has not been asked about them.
alignas(64) const uint8_t kLut[128] = {
    4, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 4, 4, 5, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    5, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
};
}  // namespace

extern "C" uint64_t skb_pack_blocks_avx512(const uint8_t* s, uint64_t nbytes, uint32_t* codes, uint32_t* nmask) {
  const __m512i l0 = _mm512_load_si512(kLut), l1 = _mm512_load_si512(kLut + 64);
  const __m512i m1 = _mm512_set1_epi16(0x0401);
  const __m512i m2 = _mm512_set1_epi32(0x00100001);
  const __m512i four = _mm512_set1_epi8(4), five = _mm512_set1_epi8(5);
  uint64_t i = 0;
  for (; i + 64 <= nbytes; i += 64) {
    const __m512i v = _mm512_loadu_si512(s + i);
    const __m512i cls = _mm512_permutex2var_epi8(l0, v, l1);  // index = the byte's low 7 bits
    const __mmask64 high = _mm512_movepi8_mask(v);             // bytes >= 0x80: kept, not a base
    const __mmask64 removed = _mm512_cmpeq_epi8_mask(cls, five) & ~high;
    if (removed) break;  // a removed byte shifts everything after it
    const __mmask64 invalid = _mm512_cmpeq_epi8_mask(cls, four) | high;
    const __m512i x = _mm512_maskz_mov_epi8(~invalid, cls);
    const __m512i y = _mm512_maddubs_epi16(x, m1);   // base pairs: b0 + 4 b1
    const __m512i z = _mm512_madd_epi16(y, m2);      // four bases in the low byte of every 32-bit lane
    _mm_storeu_si128(reinterpret_cast<__m128i*>(codes + (i >> 4)), _mm512_cvtepi32_epi8(z));
    nmask[i >> 5] = (uint32_t)invalid;
    nmask[(i >> 5) + 1] = (uint32_t)(invalid >> 32);
  }
  // what is left (a 32-byte block, or the clean half in front of a removed byte) goes the narrower way
  return i + skb_pack_blocks_avx2(s + i, nbytes - i, codes + (i >> 4), nmask + (i >> 5));
}
