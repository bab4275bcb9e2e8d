"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun). Every rank uploads its contiguous range of
reference rows, joins the library's NCCL communicator (the id travels over a gloo broadcast), holds only its slice of
the reads, and calls the collective skb_predict_stream_dist; rank 0 compares the merged ranking with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    from sketchy_b200 import synth
    from sketchy_b200._lib import Context, dist_range
    import oracle

    ctx = Context(local)
    uid = torch.from_numpy(ctx.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init(uid.numpy(), rank, world)

    k, s, seed = 16, 400, 0
    base = [synth.random_genome(60_000, 300 + l) for l in range(5)]
    sk, _, _ = oracle.sketch_groups([g.tobytes() for g in base], list(range(5)), 5, k, s, seed)
    rng = np.random.default_rng(17)
    rows = []
    for g in range(1203):                      # not a multiple of the world size; lineages interleaved: ties cross shards
        row = sk[g % 5][0].copy()
        pos = rng.choice(row.size, size=8, replace=False)
        row[pos] = rng.integers(0, int(row.max()), size=8, dtype=np.uint64)
        rows.append(np.unique(row))
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ref = np.concatenate(rows)
    n_reads = 5001
    blob, roff, _ = synth.sample_reads(base, n_reads, 1500, 23)

    lo, cnt = dist_range(len(rows), rank, world)
    ctx.ref_upload(ref[int(off[lo]):int(off[lo + cnt])], off[lo:lo + cnt + 1] - off[lo], row_base=lo)
    ok = True
    for top, chunks, mode in ((10, 1, 0), (3, 3, 1), (10, 2, 2)):
        ctx.set_rank_mode(mode)
        ctx.sums_reset()
        got_i, got_s = [], []
        bounds = np.linspace(0, n_reads, chunks + 1).astype(int)
        for c in range(chunks):
            q_lo, q_hi = int(bounds[c]), int(bounds[c + 1])
            b0, bc = dist_range(q_hi - q_lo, rank, world)
            sub = slice(q_lo + b0, q_lo + b0 + bc)
            b = ctx.batch()
            if bc:
                b.add(blob[int(roff[sub.start]):int(roff[sub.stop])], roff[sub.start:sub.stop + 1] - roff[sub.start])
            gi, gs = ctx.predict_stream_dist(b, q_hi - q_lo, k, s, seed, top)
            b.close()
            got_i.append(gi.copy()); got_s.append(gs.copy())
        gi, gs = np.concatenate(got_i), np.concatenate(got_s)
        if rank == 0:
            ei, es, _ = oracle.predict_stream(ref, off, (blob, roff), k, s, seed, top, nthreads=os.cpu_count() or 1)
            good = bool((gi == ei).all() and (gs == es).all())
            print(f"[dist_worker] world {world} top {top} chunks {chunks} mode {mode}: {'ok' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
        # every rank must hold the same merged answer
        t = torch.from_numpy(gi.astype(np.int64).copy())
        ref_t = t.clone()
        dist.broadcast(ref_t, 0)
        ok = ok and bool((t == ref_t).all())
    ctx.comm_destroy()
    ctx.close()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
