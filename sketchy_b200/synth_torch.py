"""Synthetic C3-shaped workload generated with torch (GPU when present, CPU otherwise): the 40k x 10k reference
hash matrix and 100k ONT-like 5 kb reads. Data plumbing for bench.py only — never the product path.
Every draw comes from an explicitly seeded generator; row blocks and read blocks are seeded independently of how
the rows are sharded, so every rank (and the reference arm) sees the same data.
"""
from __future__ import annotations

import numpy as np
import torch

ROW_BLOCK = 1000


def expand_reference_block(base_rows: torch.Tensor, row0: int, n_rows: int, replace_frac: float, seed: int,
                           device) -> torch.Tensor:
    """Rows [row0, row0+n_rows) of the matrix: row g = base row (g % L) with `replace_frac` of its entries replaced by
    uniform draws below the row maximum, re-sorted. int64 (all hashes of a bottom-s sketch are < 2^63)."""
    L, s = base_rows.shape
    assert row0 % ROW_BLOCK == 0
    out = torch.empty((n_rows, s), dtype=torch.int64, device=device)
    for b0 in range(0, n_rows, ROW_BLOCK):
        nb = min(ROW_BLOCK, n_rows - b0)
        gen = torch.Generator(device=device)
        gen.manual_seed(seed + (row0 + b0) // ROW_BLOCK)
        gidx = torch.arange(row0 + b0, row0 + b0 + nb, device=device) % L
        rows = base_rows.to(device)[gidx].clone()
        rmax = rows[:, -1:].to(torch.float64)
        mask = torch.rand((nb, s), generator=gen, device=device) < replace_frac
        draw = (torch.rand((nb, s), generator=gen, device=device, dtype=torch.float64) * rmax).to(torch.int64)
        rows = torch.where(mask, draw, rows)
        rows, _ = torch.sort(rows, dim=1)
        # strictly increasing is a precondition of the upload; nudge the (astronomically rare) collisions
        dup = rows[:, 1:] <= rows[:, :-1]
        if bool(dup.any()):
            fix = torch.cumsum(torch.cat([torch.zeros((nb, 1), dtype=torch.int64, device=device), dup.to(torch.int64)], 1), 1)
            rows = rows + fix
        out[b0:b0 + nb] = rows
    return out


_COMP = torch.full((256,), ord("N"), dtype=torch.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b
_ACGT = torch.tensor(list(b"ACGT"), dtype=torch.uint8)


def sample_reads(genomes: torch.Tensor, n_reads: int, read_len: int, seed: int, sub: float = 0.03, ins: float = 0.02,
                 dele: float = 0.02, block: int = 8192) -> np.ndarray:
    """[n_reads, read_len] uint8 ASCII reads sampled uniformly (genome, offset, strand) from `genomes`
    ([L, glen] uint8 ASCII) with iid substitutions, insertions and deletions."""
    device = genomes.device
    L, glen = genomes.shape
    take = int(read_len * 1.12) + 64
    comp = _COMP.to(device)
    acgt = _ACGT.to(device)
    out = np.empty((n_reads, read_len), dtype=np.uint8)
    for r0 in range(0, n_reads, block):
        nb = min(block, n_reads - r0)
        gen = torch.Generator(device=device)
        gen.manual_seed(seed + r0 // block)
        src = torch.randint(0, L, (nb,), generator=gen, device=device)
        pos = torch.randint(0, glen - take, (nb,), generator=gen, device=device)
        strand = torch.rand((nb,), generator=gen, device=device) < 0.5
        idx = pos[:, None] + torch.arange(take, device=device)[None, :]
        frag = genomes[src[:, None], idx]
        rc = comp[frag.flip(1).long()]
        frag = torch.where(strand[:, None], rc, frag)
        u = torch.rand((nb, take), generator=gen, device=device)
        keep = u >= dele
        is_sub = keep & (u < dele + sub)
        rnd = acgt[torch.randint(0, 4, (nb, take), generator=gen, device=device)]
        frag = torch.where(is_sub, rnd, frag)
        is_ins = keep & (torch.rand((nb, take), generator=gen, device=device) < ins)
        rnd2 = acgt[torch.randint(0, 4, (nb, take), generator=gen, device=device)]
        reps = keep.to(torch.int64) + is_ins.to(torch.int64)
        dest = torch.cumsum(reps, 1) - reps
        total = reps.sum(1)
        assert int(total.min()) >= read_len, "fragment too short after deletions"
        buf = torch.zeros((nb, read_len + 2), dtype=torch.uint8, device=device)
        d1 = torch.where(keep & (dest < read_len), dest, torch.full_like(dest, read_len))
        buf.scatter_(1, d1, frag)
        d2 = torch.where(is_ins & (dest + 1 < read_len), dest + 1, torch.full_like(dest, read_len + 1))
        buf.scatter_(1, d2, rnd2)
        out[r0:r0 + nb] = buf[:, :read_len].cpu().numpy()
    return out


def random_genomes(n: int, length: int, seed: int, device) -> torch.Tensor:
    acgt = _ACGT.to(device)
    out = torch.empty((n, length), dtype=torch.uint8, device=device)
    for i in range(n):
        gen = torch.Generator(device=device)
        gen.manual_seed(seed + i)
        out[i] = acgt[torch.randint(0, 4, (length,), generator=gen, device=device)]
    return out
