"""Deterministic synthetic genomes / reads / reference matrices of the BASELINE.json shapes (numpy, host side).

Used by the tests and by bench.py; never on the product path. Seeds are explicit so every result is reproducible.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP[a] = b


def random_genome(length: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]


def mutate(genome: np.ndarray, rate: float, seed: int) -> np.ndarray:
    """iid SNPs at `rate` (lineage member = lineage base with SNPs)."""
    rng = np.random.default_rng(seed)
    g = genome.copy()
    n = rng.binomial(g.size, rate)
    pos = rng.integers(0, g.size, size=n)
    g[pos] = _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]
    return g


def revcomp(seq: np.ndarray) -> np.ndarray:
    return _COMP[seq[::-1]]


def sample_reads(genomes: list[np.ndarray], n_reads: int, read_len: int, seed: int, sub: float = 0.03,
                 ins: float = 0.02, dele: float = 0.02) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """ONT-like reads: uniform (genome, offset, strand), iid substitution / insertion / deletion errors.
    Returns (blob u8, offsets u64[n+1], source genome index[n])."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, len(genomes), size=n_reads)
    take = int(read_len * 1.1) + 16
    out, lens = [], []
    for r in range(n_reads):
        g = genomes[src[r]]
        p = int(rng.integers(0, max(1, g.size - take)))
        frag = g[p:p + take]
        if rng.random() < 0.5:
            frag = revcomp(frag)
        u = rng.random(frag.size)
        keep = u >= dele                      # deletions
        is_sub = (u >= dele) & (u < dele + sub)
        frag = frag.copy()
        frag[is_sub] = _ACGT[rng.integers(0, 4, size=int(is_sub.sum()), dtype=np.uint8)]
        reps = keep.astype(np.int64)
        is_ins = rng.random(frag.size) < ins  # insertion of one random base after the position
        reps = reps + (is_ins & keep)
        seq = np.repeat(frag, reps)
        ins_pos = np.flatnonzero(np.repeat(is_ins & keep, reps))[1::2] if is_ins.any() else np.zeros(0, np.int64)
        if ins_pos.size:
            seq[ins_pos] = _ACGT[rng.integers(0, 4, size=ins_pos.size, dtype=np.uint8)]
        seq = seq[:read_len]
        out.append(seq)
        lens.append(seq.size)
    off = np.zeros(n_reads + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens, dtype=np.uint64)
    return np.concatenate(out), off, src


def expand_reference(base_rows: list[np.ndarray], n_rows: int, replace_frac: float, seed: int) -> np.ndarray:
    """C3 reference matrix: row g of lineage l = base row l with `replace_frac` of its entries replaced
    by fresh uniform draws below the row maximum, re-sorted, strictly increasing; lineages interleaved in file order.
    All base rows must have the same length s. Returns [n_rows, s] uint64."""
    s = base_rows[0].size
    L = len(base_rows)
    out = np.empty((n_rows, s), dtype=np.uint64)
    for g in range(n_rows):
        rng = np.random.default_rng(seed + g)
        row = base_rows[g % L].copy()
        m = int(round(s * replace_frac))
        for _ in range(8):
            pos = rng.choice(s, size=m, replace=False)
            row2 = row.copy()
            row2[pos] = rng.integers(0, int(row.max()), size=m, dtype=np.uint64)
            row2.sort()
            if (row2[1:] > row2[:-1]).all():
                row = row2
                break
        out[g] = row
    return out
