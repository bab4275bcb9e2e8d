"""Multi-GPU plumbing for streaming predict (one process per GPU, torch.distributed for the exchange).

The reference is single-process; the only exchange step the sharded path needs is: every rank holds
a contiguous range of reference rows, computes its local top-N per read with GLOBAL row indices, the per-rank lists
are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) and merged by (sum desc, index asc). Exact because
every member of the global top-N is in its shard's local top-N and the key is a total order.
"""
from __future__ import annotations

import numpy as np


def shard_rows(n_rows: int, rank: int, world: int, block: int = 1) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`; boundaries are multiples of `block` (except the last)."""
    lo = (n_rows * rank // world) // block * block
    hi = n_rows if rank == world - 1 else (n_rows * (rank + 1) // world) // block * block
    return lo, hi


def all_gather_topn(idx, sums, group=None):
    """idx [R, top] (int32 view of u32), sums [R, top] (int64 view of u64) torch tensors on this rank's device ->
    ([W, R, top], [W, R, top]). Uses all_gather_into_tensor (NCCL / gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    g_idx = torch.empty((world,) + tuple(idx.shape), dtype=idx.dtype, device=idx.device)
    g_sum = torch.empty((world,) + tuple(sums.shape), dtype=sums.dtype, device=sums.device)
    if idx.is_cuda:
        dist.all_gather_into_tensor(g_idx.view(-1), idx.contiguous().view(-1), group=group)
        dist.all_gather_into_tensor(g_sum.view(-1), sums.contiguous().view(-1), group=group)
    else:  # gloo has no all_gather_into_tensor on every build: use the list form
        li = [torch.empty_like(idx) for _ in range(world)]
        ls = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(li, idx.contiguous(), group=group)
        dist.all_gather(ls, sums.contiguous(), group=group)
        g_idx, g_sum = torch.stack(li), torch.stack(ls)
    return g_idx, g_sum
