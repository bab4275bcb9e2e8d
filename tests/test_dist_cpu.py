"""world_size-2 gloo test (CPU) of the multi-GPU protocol: contiguous row shards, local top-N with global indices
(computed here by the oracle), all-gather through torch.distributed, merge by (sum desc, index asc) == the unsharded
ranking. The CUDA merge kernel itself is covered by tests/test_gpu_parity.py::test_sharded_predict_and_merge."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from sketchy_b200 import dist as skd
from sketchy_b200 import synth


def merge_reference(g_idx: np.ndarray, g_sum: np.ndarray, top: int):
    """numpy checker of the merge: per read, sort the W*top entries by (sum desc, idx asc), keep `top`."""
    W, R, T = g_idx.shape
    out_i = np.zeros((R, top), dtype=np.uint32)
    out_s = np.zeros((R, top), dtype=np.uint64)
    for r in range(R):
        ents = [(int(g_sum[w, r, t]), int(g_idx[w, r, t])) for w in range(W) for t in range(T)]
        ents.sort(key=lambda e: (-e[0], e[1]))
        for t in range(top):
            out_s[r, t], out_i[r, t] = ents[t]
    return out_i, out_s


def _world(seed=5):
    base = [synth.random_genome(15_000, seed * 100 + l) for l in range(4)]
    genomes = [synth.mutate(base[g % 4], 0.002, seed * 7 + g) for g in range(23)]
    sk, _, _ = oracle.sketch_groups([g.tobytes() for g in genomes], list(range(23)), 23, 16, 200, 0)
    rows = [h for h, _ in sk]
    off = np.zeros(24, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    blob, roff, _ = synth.sample_reads(base, 60, 1200, seed)
    return np.concatenate(rows), off, blob, roff


def _rank_main(rank, world, port, top, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref, off, blob, roff = _world()
    N = off.size - 1
    lo, hi = skd.shard_rows(N, rank, world)
    sub_off = off[lo:hi + 1] - off[lo]
    sub_ref = ref[int(off[lo]):int(off[hi])]
    ltop = min(top, hi - lo)
    li, ls, _ = oracle.predict_stream(sub_ref, sub_off, (blob, roff), 16, 200, 0, ltop)
    R = li.shape[0]
    idx = np.full((R, top), 0xFFFFFFFF, dtype=np.uint32)   # pad entries sort last (sum 0, idx max)
    sm = np.zeros((R, top), dtype=np.uint64)
    idx[:, :ltop] = li + np.uint32(lo)                      # GLOBAL row indices
    sm[:, :ltop] = ls
    g_idx, g_sum = skd.all_gather_topn(torch.from_numpy(idx.view(np.int32)), torch.from_numpy(sm.view(np.int64)))
    mi, ms = merge_reference(g_idx.numpy().view(np.uint32), g_sum.numpy().view(np.uint64), top)
    ei, es, _ = oracle.predict_stream(ref, off, (blob, roff), 16, 200, 0, top)
    ok = bool((mi == ei).all() and (ms == es).all())
    ret[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("top", [5, 14])
def test_two_rank_gather_and_merge_equals_unsharded(top):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rank_main, args=(2, port, top, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_shard_rows_cover_everything():
    for n in (0, 1, 7, 40_000, 1_000_003):
        for w in (1, 2, 3, 8):
            for blk in (1, 1000):
                rs = [skd.shard_rows(n, r, w, blk) for r in range(w)]
                assert rs[0][0] == 0 and rs[-1][1] == n
                assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
                assert all(lo <= hi for lo, hi in rs)
