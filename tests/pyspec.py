"""Independent pure-Python restatement of SURVEY.md Appendix A (closed forms), used ONLY by the tests to
cross-check oracle/oracle.cpp (a second implementation written from the same spec, different code shape:
set-based bottom-s instead of heap+map, set intersection instead of a merge, key sort instead of a stable sort).
Small inputs only.
"""
from __future__ import annotations

M64 = (1 << 64) - 1
C1 = 0x87C37B91114253D5
C2 = 0x4CF5AD432745937F


def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & M64


def _fmix(x):
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & M64
    x ^= x >> 33
    x = (x * 0xC4CEB9FE1A85EC53) & M64
    x ^= x >> 33
    return x


def murmur3_x64_128(data: bytes, seed: int = 0) -> tuple[int, int]:
    h1 = h2 = seed & M64
    n = len(data)
    for b in range(n // 16):
        k1 = int.from_bytes(data[16 * b:16 * b + 8], "little")
        k2 = int.from_bytes(data[16 * b + 8:16 * b + 16], "little")
        k1 = (k1 * C1) & M64; k1 = _rotl(k1, 31); k1 = (k1 * C2) & M64; h1 ^= k1
        h1 = _rotl(h1, 27); h1 = (h1 + h2) & M64; h1 = (h1 * 5 + 0x52DCE729) & M64
        k2 = (k2 * C2) & M64; k2 = _rotl(k2, 33); k2 = (k2 * C1) & M64; h2 ^= k2
        h2 = _rotl(h2, 31); h2 = (h2 + h1) & M64; h2 = (h2 * 5 + 0x38495AB5) & M64
    tail = data[16 * (n // 16):]
    if len(tail) > 8:
        k2 = int.from_bytes(tail[8:], "little")
        k2 = (k2 * C2) & M64; k2 = _rotl(k2, 33); k2 = (k2 * C1) & M64; h2 ^= k2
    if len(tail) > 0:
        k1 = int.from_bytes(tail[:8], "little")
        k1 = (k1 * C1) & M64; k1 = _rotl(k1, 31); k1 = (k1 * C2) & M64; h1 ^= k1
    h1 ^= n; h2 ^= n
    h1 = (h1 + h2) & M64; h2 = (h2 + h1) & M64
    h1 = _fmix(h1); h2 = _fmix(h2)
    h1 = (h1 + h2) & M64; h2 = (h2 + h1) & M64
    return h1, h2


_KEEP = {ord(c): ord(c) for c in "ACGTN-"}
_KEEP.update({ord("a"): ord("A"), ord("c"): ord("C"), ord("g"): ord("G"), ord("t"): ord("T"),
              ord("u"): ord("T"), ord("U"): ord("T"), ord("."): ord("-"), ord("~"): ord("-")})
_DROP = {ord(" "), ord("\t"), ord("\r"), ord("\n")}


def normalize(seq: bytes) -> bytes:
    return bytes(_KEEP.get(c, ord("N")) for c in seq if c not in _DROP)


_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def canonical_kmer_hashes(seq: bytes, k: int, seed: int = 0) -> list[int]:
    s = normalize(seq)
    out = []
    for p in range(0, len(s) - k + 1):
        w = s[p:p + k]
        if any(c not in b"ACGT" for c in w):
            continue
        rc = w.translate(_COMP)[::-1]
        out.append(murmur3_x64_128(min(w, rc), seed)[0])
    return out


def bottom_s(records: list[bytes], k: int, s: int, seed: int = 0):
    """Closed form of finch MashSketcher: the s smallest DISTINCT hashes with exact occurrence counts."""
    cnt: dict[int, int] = {}
    bases = kmers = 0
    for r in records:
        bases += len(r)
        for h in canonical_kmer_hashes(r, k, seed):
            kmers += 1
            cnt[h] = cnt.get(h, 0) + 1
    keys = sorted(cnt)[:s]
    return keys, [cnt[h] for h in keys], bases, kmers


def predict_stream(ref_rows: list[list[int]], reads: list[bytes], k: int, s_query: int, seed: int, top: int):
    sums = [0] * len(ref_rows)
    sets = [set(r) for r in ref_rows]
    out = []
    for rd in reads:
        q, _, _, _ = bottom_s([rd], k, s_query, seed)
        qs = set(q)
        for i, rs in enumerate(sets):
            sums[i] += len(qs & rs)
        order = sorted(range(len(sums)), key=lambda i: (-sums[i], i))[:top]
        out.append([(i, sums[i]) for i in order])
    return out, sums
