#!/usr/bin/env python
"""bench.py — predict reads/s of the B200 MinHash hot path on the BASELINE.json C3 workload
(100,000 synthetic 5 kb ONT-like reads vs a 40,000-genome k=16 s=10,000 reference, --top 10).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --steps K --warmup W    # the CPU oracle port, timed on the host cores

One JSON line on stdout (rank 0). A step = streaming predict of ALL reads (running sums reset at step start) =
ceil(reads / reads_per_pass) passes over the HBM-resident reference matrix. `value` times it with the packed
reads already in HBM; `e2e` times the same job from host ASCII buffers through the C ABI (2-bit packing into
pinned memory, H2D, kernels, D2H of the top-N). Multi-GPU: reference rows are sharded by contiguous range, every
rank sees every read, local top-N are all-gathered over NCCL and merged by (sum desc, index asc).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, SEED = 16, 0
GENOME_LEN = 2_800_000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--refs", type=int, default=40_000)
    p.add_argument("--sketch-size", type=int, default=10_000)
    p.add_argument("--reads", type=int, default=100_000)
    p.add_argument("--read-len", type=int, default=5_000)
    p.add_argument("--lineages", type=int, default=40)
    p.add_argument("--top", type=int, default=10)
    p.add_argument("--pass-reads", type=int, default=0, help="reads per streaming pass (0 = library default)")
    p.add_argument("--cpu-sample", type=int, default=12, help="reads in the bounded CPU-baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--verify", action="store_true", help="N>1: check the merged ranking against one GPU holding all rows")
    p.add_argument("--sketch-genomes", type=int, default=64, help="genomes in the bounded sketch-throughput sample (0 = skip)")
    return p.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.STDOUT, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 15:   # nvidia-smi takes a second or two to print its first row
                time.sleep(0.05)
        except Exception as e:  # nvidia-smi missing: report it, do not fake clocks
            log("clock sampler unavailable:", e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler_unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm and self.rows:
            log("clock sampler rows unparsed:", self.rows[:2])
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic_bytes() -> float | None:
    """DRAM bytes (read + write) of one fused_kernel launch from the committed `ncu --set full` capture of this
    workload (profiles/r01_fused_kernel_ncu.md); None when the summary is missing."""
    try:
        rd = wr = None
        for line in open(os.path.join(ROOT, "profiles", "r01_fused_kernel_ncu.md")):
            f = [x.strip() for x in line.split("|")]
            if len(f) >= 4 and f[1] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[3]]
                rd, wr = (v, wr) if f[1].endswith("read.sum") else (rd, v)
        return rd + wr if rd is not None and wr is not None else None
    except Exception:
        return None


def workload_name(N: int, s: int, R: int, top: int) -> str:
    """config.workload: the same string for this repo's arm and the reference arm."""
    if (N, s, R, top) == (40000, 10000, 100000, 10):
        return "C3: predict 100,000 synthetic ONT reads vs 40,000-genome reference, k=16, s=10000, --top 10"
    return f"predict {R} reads vs {N} x s={s}, k=16, --top {top}"


def measured_peak_gbs() -> tuple[float, str]:
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def make_workload(args, device, rank, world, need_host_matrix):
    """Synthetic C3 data. Returns dict with base genomes (torch), local reference rows (torch int64 on `device`),
    reads ([R, read_len] uint8 numpy) and offsets."""
    import torch
    from sketchy_b200 import synth_torch as st
    t0 = time.time()
    genomes = st.random_genomes(args.lineages, GENOME_LEN, 3000, device)
    log(f"[rank {rank}] genomes {time.time() - t0:.1f}s")
    return genomes


def sketch_base_rows_gpu(ctx, genomes, s):
    rows = []
    for l in range(genomes.shape[0]):
        b = ctx.batch().add_records([genomes[l].cpu().numpy()])
        sk, _, _ = ctx.sketch(b, K, s, SEED)
        b.close()
        rows.append(sk[0][0])
    return rows


def run_b200(args):
    import torch
    import torch.distributed as dist
    from sketchy_b200 import synth_torch as st
    from sketchy_b200._lib import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ctx = Context(local)  # raises without a B200: no CPU fallback
    if args.pass_reads:
        ctx.set_pass_reads(args.pass_reads)
    N, s, R, top = args.refs, args.sketch_size, args.reads, args.top

    # ---------------- synthetic data (untimed) ----------------
    t0 = time.time()
    genomes = st.random_genomes(args.lineages, GENOME_LEN, 3000, device)
    base_rows = sketch_base_rows_gpu(ctx, genomes, s)
    assert all(r.size == s for r in base_rows)
    base_t = torch.from_numpy(np.stack(base_rows).astype(np.int64))
    assert int(base_t.min()) >= 0
    from sketchy_b200.dist import shard_rows
    lo, hi = shard_rows(N, rank, world, st.ROW_BLOCK)
    ref = st.expand_reference_block(base_t.to(device), lo, hi - lo, 0.02, 4000, device)
    off = np.arange(hi - lo + 1, dtype=np.uint64) * np.uint64(s)
    ctx.ref_upload_device(ref.data_ptr(), off, row_base=lo)
    reads = st.sample_reads(genomes, R, args.read_len, 777)
    roff = np.arange(R + 1, dtype=np.uint64) * np.uint64(args.read_len)
    blob_pinned = torch.from_numpy(reads.reshape(-1)).pin_memory()
    blob = blob_pinned.numpy()
    log(f"[rank {rank}] data ready in {time.time() - t0:.1f}s: rows [{lo},{hi}) x {s}, {R} reads x {args.read_len}")

    # ---------------- CPU baseline on a bounded sample (rank 0, N=1 only) + parity gate ----------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle  # the checker / baseline, never the thing measured as `value`
        ref_host = ref.cpu().numpy().view(np.uint64).reshape(-1)
        n_s = min(args.cpu_sample, R)
        sub = (blob[:n_s * args.read_len], roff[:n_s + 1])
        t1 = time.time()
        ei, es, _ = oracle.predict_stream(ref_host, off, sub, K, s, SEED, top)
        dt = time.time() - t1
        cpu_baseline = {"value": n_s / dt, "unit": "reads/s", "cores": 1, "kind": "port",
                        "sample": f"first {n_s} of the {R} reads vs the full {N}x{s} matrix, {dt:.1f}s "
                                  f"(C++ restatement of sketchy 0.6.0, not the Rust binary; the reference's predict "
                                  f"loop is single-threaded, src/sketchy.rs:328-355)"}
        # labelled extra: the same sample with every read's N merges spread over all host cores (the reference does
        # not do this; it shows what the box's CPU could do on the path, not what sketchy 0.6.0 does)
        ncores = os.cpu_count() or 1
        t1 = time.time()
        ai, as_, _ = oracle.predict_stream(ref_host, off, sub, K, s, SEED, top, nthreads=ncores)
        dt_all = time.time() - t1
        assert (ai == ei).all() and (as_ == es).all()
        cpu_baseline["all_cores"] = {"value": n_s / dt_all, "unit": "reads/s", "cores": ncores,
                                     "note": "not the reference's behaviour (its predict loop is single-threaded): "
                                             "the oracle's N merges per read spread over all host threads"}
        b = ctx.batch().add(sub[0], sub[1])
        gi, gs = ctx.predict_stream(b, K, s, SEED, top)
        b.close()
        assert (gi == ei).all() and (gs == es).all(), "GPU predict differs from the oracle on the CPU sample"
        log(f"[rank 0] parity gate ok on {n_s} reads; cpu {n_s / dt:.2f} reads/s")
        del ref_host
    del ref
    torch.cuda.empty_cache()

    # ---------------- resident batch ----------------
    batch = ctx.batch().add(blob, roff)
    batch.stage()
    d_idx = torch.zeros((R, top), dtype=torch.int32, device=device)
    d_sum = torch.zeros((R, top), dtype=torch.int64, device=device)
    if world > 1:
        g_idx = torch.zeros((world, R, top), dtype=torch.int32, device=device)
        g_sum = torch.zeros((world, R, top), dtype=torch.int64, device=device)
        m_idx = torch.zeros((R, top), dtype=torch.int32, device=device)
        m_sum = torch.zeros((R, top), dtype=torch.int64, device=device)
    ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=device)

    def step_resident():
        ctx.sums_reset()
        ctx.predict_stream_device(batch, K, s, SEED, top, d_idx.data_ptr(), d_sum.data_ptr(), pad=world > 1)
        if world > 1:
            dist.all_gather_into_tensor(g_idx.view(-1), d_idx.view(-1))
            dist.all_gather_into_tensor(g_sum.view(-1), d_sum.view(-1))
            torch.cuda.current_stream().synchronize()
            ctx.merge_topn_device(g_idx.data_ptr(), g_sum.data_ptr(), world, R, top, m_idx.data_ptr(), m_sum.data_ptr())

    def sync_all():
        torch.cuda.synchronize()
        ctx.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sync_all()
    if world > 1 and args.verify:
        # parity gate of the sharded path: the merged ranking of the first reads == one GPU holding every row
        n_v = min(4096, R)
        ok = torch.ones(1, device=device)
        if rank == 0:
            full = st.expand_reference_block(base_t.to(device), 0, N, 0.02, 4000, device)
            c2 = Context(local)
            c2.ref_upload_device(full.data_ptr(), np.arange(N + 1, dtype=np.uint64) * np.uint64(s), row_base=0)
            del full
            vb = c2.batch().add(blob[:n_v * args.read_len], roff[:n_v + 1])
            vi, vs = c2.predict_stream(vb, K, s, SEED, top)
            vb.close(); c2.close()
            mi = m_idx[:n_v].cpu().numpy().view(np.uint32)
            ms = m_sum[:n_v].cpu().numpy().view(np.uint64)
            good = bool((mi == vi).all() and (ms == vs).all())
            log(f"[rank 0] sharded ({world} GPUs) vs unsharded on {n_v} reads: {'ok' if good else 'MISMATCH'}")
            ok[0] = 1.0 if good else 0.0
        dist.broadcast(ok, 0)
        assert ok.item() == 1.0, "sharded predict differs from unsharded"
    ctx.prof_reset()
    ctx.prof_enable(True)
    launches0 = ctx.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(args.steps):
        step_resident()
    e1.record(ext)
    sync_all()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0
    stream_ms, stream_n = ctx.prof_get("stream")
    prof = {k: ctx.prof_get(k) for k in ("hash", "select", "table", "stream", "rank", "merge")}
    ctx.prof_enable(False)
    stats = ctx.last_predict_stats()
    # checksum of the whole job's answer (every read's top-N rows and sums): equal across builds, pass sizes and GPU
    # counts when the results are identical, so kernel variants and shardings can be compared at full size
    import zlib
    r_idx, r_sum = (m_idx, m_sum) if world > 1 else (d_idx, d_sum)
    result_crc = zlib.crc32(r_sum.cpu().numpy().tobytes(), zlib.crc32(r_idx.cpu().numpy().tobytes()))
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---------------- end to end through the C ABI with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        hb = ctx.batch()
        # caller-owned page-locked result arrays: the D2H of every chunk's top-N is a DMA straight into them
        oi = torch.zeros((R, top), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        os_ = torch.zeros((R, top), dtype=torch.int64).pin_memory().numpy().view(np.uint64)

        # a streaming caller feeds reads in chunks: chunk i+1 is normalised + 2-bit packed into pinned memory by the
        # library's host threads while the GPU works on chunk i (two batches, double buffered)
        # chunk sizes grow so that packing + copying chunk i+1 always fits inside the GPU time of chunk i, and they are
        # multiples of the library's pass sizes (ramp, then 4096 per pass): no partial passes
        b0 = 128   # the library's first pass size after a reset (largest pass whose buckets hold the whole shard) ...
        while b0 * 2 <= (24 << 20) // max(hi - lo, 64) and b0 * 2 <= 4096:
            b0 *= 2
        ramp = sum(b0 << i for i in range(8) if (b0 << i) < 4096)   # ... and the reads of the passes that ramp up to 4096
        sizes, c_lo = [ramp if ramp else 4_096, 4 * 4_096], 0
        chunks = []
        while c_lo < R:
            n_c = sizes[len(chunks)] if len(chunks) < len(sizes) else 10 * 4_096
            chunks.append((c_lo, min(c_lo + n_c, R)))
            c_lo = chunks[-1][1]
        pack_s = [0.0]
        hbs = [hb, ctx.batch()]
        # every rank packs every read: share the host cores between the ranks of the box instead of oversubscribing them
        pack_threads = max(1, (os.cpu_count() or 1) // world)

        def pack(j, p_lo, p_hi):
            t0 = time.perf_counter()
            hbs[j].clear()
            hbs[j].add(blob[p_lo * args.read_len:p_hi * args.read_len], roff[p_lo:p_hi + 1] - roff[p_lo],
                       nthreads=pack_threads)
            pack_s[0] += time.perf_counter() - t0
            hbs[j].stage()   # H2D on the copy stream, overlapping the kernels of the previous chunk

        def step_e2e():
            ctx.sums_reset()
            pack(0, *chunks[0])
            for ci, (q_lo, q_hi) in enumerate(chunks):
                th = None
                if ci + 1 < len(chunks):
                    th = threading.Thread(target=pack, args=((ci + 1) & 1, *chunks[ci + 1]))
                    th.start()
                b = hbs[ci & 1]
                if world > 1:
                    ctx.predict_stream_device(b, K, s, SEED, top, d_idx[q_lo:q_hi].data_ptr(), d_sum[q_lo:q_hi].data_ptr(), pad=True)
                else:
                    ctx.predict_stream(b, K, s, SEED, top, out=(oi[q_lo:q_hi], os_[q_lo:q_hi]))   # H2D + kernels + D2H of the top-N
                if th is not None:
                    th.join()
            if world > 1:
                dist.all_gather_into_tensor(g_idx.view(-1), d_idx.view(-1))
                dist.all_gather_into_tensor(g_sum.view(-1), d_sum.view(-1))
                torch.cuda.current_stream().synchronize()
                ctx.merge_topn_device(g_idx.data_ptr(), g_sum.data_ptr(), world, R, top, m_idx.data_ptr(),
                                      m_sum.data_ptr())
                oi[:] = m_idx.cpu().numpy().view(np.uint32)
                os_[:] = m_sum.cpu().numpy().view(np.uint64)

        for _ in range(2):
            step_e2e()
        sync_all()
        t1 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 3))
        pack_s[0] = 0.0
        for _ in range(n_e2e):
            step_e2e()
        sync_all()
        dt = (time.perf_counter() - t1) / n_e2e
        if world > 1:
            t = torch.tensor([dt], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        packed = sum(-(-((b1 - b0) * (args.read_len + 1)) // 32) * 32 for b0, b1 in chunks)
        h2d = packed // 4 + packed // 8 + (packed // 1024 + R) * 9
        e2e = {"value": R / dt, "unit": "reads/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(R * top * 12), "ms_per_step": dt * 1e3,
               "includes": "host normalise+2-bit pack into pinned memory (a first chunk covering the ramp passes, 16384, then chunks of 40960 reads, packed and copied "
                           "to the device while the GPU works on the previous chunk), all kernels, D2H of top-N into page-locked host arrays",
               "host_threads": pack_threads, "host_pack_ms_per_step": pack_s[0] / n_e2e * 1e3}
        # the e2e result must equal the resident result
        if world == 1:
            assert (oi == d_idx.cpu().numpy().view(np.uint32)).all() and (os_ == d_sum.cpu().numpy().view(np.uint64)).all()
        for x in hbs:
            x.close()

    # ---------------- sketch throughput on a bounded C2 sample (k=16, s=1000), rank 0 at N=1 ----------------
    sketch_info = None
    if rank == 0 and world == 1 and args.sketch_genomes > 0:
        ng = args.sketch_genomes
        recs = [st.random_genomes(1, GENOME_LEN, 9000 + g, device)[0].cpu().numpy() for g in range(min(ng, 8))]
        recs = [recs[g % len(recs)] for g in range(ng)]           # 8 distinct genomes, repeated (hashing cost is data independent)
        sb = ctx.batch().add_records(recs)
        sb.stage()
        ctx.sketch(sb, K, 1000, SEED)                               # warm-up
        ctx.prof_reset(); ctx.prof_enable(True)
        ctx.synchronize()
        t1 = time.perf_counter()
        n_it = 3
        for _ in range(n_it):
            sk_out, _, _ = ctx.sketch(sb, K, 1000, SEED)
        ctx.synchronize()
        dt = (time.perf_counter() - t1) / n_it
        hash_ms, _ = ctx.prof_get("hash")
        sel_ms, _ = ctx.prof_get("select")
        ctx.prof_enable(False)
        gbp = ng * GENOME_LEN / 1e9
        sketch_info = {"workload": f"C2 sample: {ng} x 2.8 Mbp assemblies, k=16, s=1000, packed batch resident in HBM",
                       "gbp_per_s": gbp / dt, "kernel_gbp_per_s": gbp / ((hash_ms + sel_ms) / n_it * 1e-3),
                       "hash_kernel_ms": hash_ms / n_it, "select_kernel_ms": sel_ms / n_it, "call_ms": dt * 1e3,
                       "bound": "integer pipes (about 100 instructions per k-mer, ~0.4 B/base of traffic)"}
        assert all(h.size == 1000 for h, _ in sk_out)
        sb.close()

    if rank == 0:
        peak, how = measured_peak_gbs()
        rows_local = hi - lo
        passes = stats["passes"]
        # query hashes that actually enter a pass (the membership prefilter drops the ones no reference row holds)
        keys_per_pass = stats.get("member_hashes", stats["query_hashes"]) / max(passes, 1)
        bytes_per_launch = rows_local * s * 8 + keys_per_pass * 8
        avg_ms = stream_ms / max(stream_n, 1)
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        ms_per_step = ms / args.steps
        out = {
            "metric": "predict reads/s (streaming, 100k 5kb reads vs 40k x s=10000 reference, top 10)",
            "value": R / (ms_per_step * 1e-3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(N, s, R, top), "refs": N, "sketch_size": s, "reads": R,
                       "read_len": args.read_len, "k": K, "top": top, "lineages": args.lineages,
                       "passes_per_step": passes, "reads_per_pass_max": args.pass_reads or 4096,
                       "l2": "reference matrix (%.2f GB per GPU) is larger than L2; streamed from HBM every pass"
                             % (rows_local * s * 8 / 1e9),
                       "parallelism": f"reference rows sharded over {world} GPU(s); NCCL all-gather of local top-N"
                       if world > 1 else "1 GPU"},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "fused_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": f"of {how}",
                         # ncu capture of a full single-GPU pass: comparable with bytes_per_launch at N=1 only
                         "traffic": ncu_traffic_bytes() if world == 1 and (N, s) == (40000, 10000) else None,
                         "traffic_source": "profiles/r01_fused_kernel_ncu.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch)",
                         "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
                         "launches": stream_n,
                         "stream_share_of_step": stream_ms / ms if ms > 0 else None},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "predict_stats": stats,
            "result_crc32": f"{result_crc:08x}",
            "cpu_baseline": cpu_baseline,
            "sketch": sketch_info,
        }
        print(json.dumps(out), flush=True)
    batch.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU implementation of the path = the oracle port (the Rust crate cannot be built here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import oracle
    from sketchy_b200 import synth_torch as st
    device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    N, s, R, top = args.refs, args.sketch_size, args.reads, args.top
    t0 = time.time()
    genomes = st.random_genomes(args.lineages, GENOME_LEN, 3000, device)
    gh = [genomes[l].cpu().numpy().tobytes() for l in range(args.lineages)]
    ncores = os.cpu_count() or 1
    sk, _, _ = oracle.sketch_groups(gh, list(range(args.lineages)), args.lineages, K, s, SEED, nthreads=ncores)
    base_t = torch.from_numpy(np.stack([h for h, _ in sk]).astype(np.int64))
    ref = st.expand_reference_block(base_t.to(device), 0, N, 0.02, 4000, device).cpu().numpy().view(np.uint64).reshape(-1)
    off = np.arange(N + 1, dtype=np.uint64) * np.uint64(s)
    n_s = min(args.cpu_sample, R)
    total_reads = n_s * (max(args.warmup, 1) + args.steps)
    reads = st.sample_reads(genomes, max(total_reads, 8192), args.read_len, 777)[:total_reads]
    log(f"[reference] data ready in {time.time() - t0:.1f}s")
    roff = np.arange(n_s + 1, dtype=np.uint64) * np.uint64(args.read_len)
    times = []
    for it in range(max(args.warmup, 1) + args.steps):
        blob = np.ascontiguousarray(reads[it * n_s:(it + 1) * n_s].reshape(-1))
        t1 = time.perf_counter()
        oracle.predict_stream(ref, off, (blob, roff), K, s, SEED, top)
        dt = time.perf_counter() - t1
        if it >= max(args.warmup, 1):
            times.append(dt)
    per_step = float(np.mean(times))
    v = n_s / per_step
    t1 = time.perf_counter()   # labelled extra: one step with the merges spread over all host threads
    oracle.predict_stream(ref, off, (blob, roff), K, s, SEED, top, nthreads=ncores)
    v_all = n_s / (time.perf_counter() - t1)
    out = {"impl": "reference", "metric": "predict reads/s (streaming, 100k 5kb reads vs 40k x s=10000 reference, top 10)",
           "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
           "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "u64", "data": "synthetic",
           "config": {"workload": workload_name(N, s, R, top), "refs": N, "sketch_size": s, "reads": R,
                      "read_len": args.read_len, "k": K, "top": top, "lineages": args.lineages,
                      "sample": f"each step = a bounded sample of {n_s} consecutive reads of the workload against the "
                                f"full {N} x {s} matrix"},
           "cpu_baseline": {"value": v, "unit": "reads/s", "cores": 1, "kind": "port",
                            "sample": f"{n_s} reads per step vs the full {N}x{s} matrix; C++ restatement of sketchy "
                                      f"0.6.0 (oracle/oracle.cpp), single thread like the reference's predict loop",
                            "all_cores": {"value": v_all, "unit": "reads/s", "cores": ncores,
                                          "note": "not the reference's behaviour: the N merges per read spread over "
                                                  "all host threads"}},
           "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
