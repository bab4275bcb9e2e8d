#!/usr/bin/env python
"""bench.py — predict reads/s of the B200 MinHash hot path on the BASELINE.json workloads.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one process per GPU (torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W    # the CPU oracle port, timed on the host cores
  python bench.py --config c4|c5|c3i ...                    # the other BASELINE configs (records kept under profiles/)

Default = C3 (100,000 synthetic 5 kb ONT-like reads vs a 40,000-genome k=16 s=10,000 reference, --top 10), the
configuration BASELINE.json's metric is quoted on. One JSON line on stdout (rank 0).

A step = streaming predict of ALL reads (running sums reset at step start) = ceil(reads / reads_per_pass) passes over the
HBM-resident reference matrix. `value` times it with the packed reads already in HBM; `e2e` times the same job from host
ASCII buffers through the C ABI (2-bit packing into pinned memory, H2D, kernels, D2H of the top-N). Multi-GPU: the
reference rows are sharded by contiguous range over the ranks, every rank packs / copies / hashes 1/N of the reads, the
per-read query-hash lists and the local top-N lists are exchanged over the library's own NCCL communicator
(skb_predict_stream_dist) and merged by (sum desc, index asc).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, SEED = 16, 0
GENOME_LEN = 2_800_000
PASS_READS = 4096  # the library's full pass (u8 counters)

CONFIGS = {
    # name: (refs, sketch size, reads, top, lineages, row distribution, consensus)
    "c3": (40_000, 10_000, 100_000, 10, 40, "lineage", False),
    "c3i": (40_000, 10_000, 100_000, 10, 40, "independent", False),
    "c4": (40_000, 1_000, 1_000_000, 5, 40, "lineage", True),
    "c5": (1_000_000, 10_000, 1_000_000, 10, 1000, "lineage", False),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    p.add_argument("--refs", type=int, default=0)
    p.add_argument("--sketch-size", type=int, default=0)
    p.add_argument("--reads", type=int, default=0)
    p.add_argument("--read-len", type=int, default=5_000)
    p.add_argument("--lineages", type=int, default=0)
    p.add_argument("--top", type=int, default=0)
    p.add_argument("--pass-reads", type=int, default=0, help="reads per streaming pass (0 = library default)")
    p.add_argument("--rank-mode", type=int, default=0, help="0 automatic, 1 candidate lists wherever possible, 2 brute force always")
    p.add_argument("--cpu-sample", type=int, default=12, help="reads in the bounded CPU-baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the zero-hit variant and the sketch leg")
    p.add_argument("--sketch-genomes", type=int, default=128, help="genomes in the bounded sketch-throughput sample (0 = skip)")
    p.add_argument("--sketch-only", action="store_true", help="run the sketch leg alone and print its record (profiling aid)")
    a = p.parse_args(argv)
    refs, s, reads, top, lin, dist_kind, cons = CONFIGS[a.config]
    a.refs = a.refs or refs
    a.sketch_size = a.sketch_size or s
    a.reads = a.reads or reads
    a.top = a.top or top
    a.lineages = a.lineages or lin
    a.row_dist, a.consensus = dist_kind, cons
    return a


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


def ncu_traffic_bytes() -> tuple[float | None, str]:
    """DRAM bytes (read + write) of one fused_kernel launch from the committed `ncu --set full` capture of the C3
    workload; None when the summary is missing."""
    for name in ("r02_fused_kernel_ncu.md", "r01_fused_kernel_ncu.md"):
        try:
            rd = wr = None
            for line in open(os.path.join(ROOT, "profiles", name)):
                f = [x.strip() for x in line.split("|")]
                if len(f) >= 4 and f[1] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[3]]
                    rd, wr = (v, wr) if f[1].endswith("read.sum") else (rd, v)
            if rd is not None and wr is not None:
                return rd + wr, f"profiles/{name} (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
        except Exception:
            pass
    return None, "no ncu summary under profiles/"


def workload_name(cfg: str, N: int, s: int, R: int, top: int) -> str:
    """config.workload: the same string for this repo's arm and the reference arm."""
    if (cfg, N, s, R, top) == ("c3", 40000, 10000, 100000, 10):
        return "C3: predict 100,000 synthetic ONT reads vs 40,000-genome reference, k=16, s=10000, --top 10"
    if (cfg, N, s, top) == ("c4", 40000, 1000, 5):
        return f"C4: streaming predict + genotype consensus, {R} synthetic ONT reads vs 40,000-genome s=1000 reference, --top 5 --consensus"
    if (cfg, N, s, top) == ("c5", 1000000, 10000, 10):
        return f"C5: scale-out, {R} reads vs 1,000,000-genome s=10000 reference sharded over the GPUs, --top 10"
    if cfg == "c3i":
        return f"C3 zero-hit variant: {R} reads vs {N} independent rows x s={s}, k=16, --top {top}"
    return f"predict {R} reads vs {N} x s={s}, k=16, --top {top}"


def measured_peak_gbs() -> tuple[float, str]:
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def sketch_base_rows_gpu(ctx, genomes, s):
    rows = []
    for l in range(genomes.shape[0]):
        b = ctx.batch().add_records([genomes[l].cpu().numpy()])
        sk, _, _ = ctx.sketch(b, K, s, SEED)
        b.close()
        rows.append(sk[0][0])
    return rows


def independent_rows(n_rows: int, s: int, row0: int, device):
    """SURVEY §8d second distribution: every row = s fresh sorted uniform draws in [0, 2^64 * s / 2.8e6): no read ever
    hits a row (the zero-hit extreme: the stream and the filter probe alone)."""
    import torch
    from sketchy_b200 import synth_torch as st
    top = float(2 ** 63) * 2.0 * s / GENOME_LEN
    out = torch.empty((n_rows, s), dtype=torch.int64, device=device)
    for b0 in range(0, n_rows, st.ROW_BLOCK):
        nb = min(st.ROW_BLOCK, n_rows - b0)
        gen = torch.Generator(device=device)
        gen.manual_seed(5000 + (row0 + b0) // st.ROW_BLOCK)
        rows = (torch.rand((nb, s), generator=gen, device=device, dtype=torch.float64) * top).to(torch.int64)
        rows, _ = torch.sort(rows, dim=1)
        dup = rows[:, 1:] <= rows[:, :-1]
        if bool(dup.any()):
            fix = torch.cumsum(torch.cat([torch.zeros((nb, 1), dtype=torch.int64, device=device), dup.to(torch.int64)], 1), 1)
            rows = rows + fix
        out[b0:b0 + nb] = rows
    return out


def genotype_table(n_rows: int, lineages: int) -> np.ndarray:
    """C4 genotype index: 6 categorical columns per reference (an MLST-like value tied to the lineage, 5 binary R/S)."""
    rng = np.random.default_rng(99)
    lin = np.arange(n_rows) % lineages
    cols = [lin.astype(np.int32)]
    for c in range(5):
        per_lineage = rng.integers(0, 2, size=lineages)
        flip = rng.random(n_rows) < 0.05
        cols.append(np.where(flip, 1 - per_lineage[lin], per_lineage[lin]).astype(np.int32))
    return np.stack(cols, axis=1)   # [N, 6]


def consensus_calls(idx: np.ndarray, table: np.ndarray) -> np.ndarray:
    """Per read and genotype column the most frequent value among the read's top rows; ties go to the value that comes
    first in rank order (api.consensus_value; the reference's HashMap max_by is nondeterministic on ties,
    src/sketchy.rs:380-387, 408). idx [R, top] -> [R, columns]."""
    v = table[idx.astype(np.int64)]                              # [R, top, C]
    same = np.zeros(v.shape, dtype=np.int8)                      # occurrences of the value at each rank position
    for t in range(v.shape[1]):
        same += (v == v[:, t:t + 1, :])
    first_best = same.argmax(axis=1)                             # first rank position holding the most frequent value
    return np.take_along_axis(v, first_best[:, None, :], axis=1)[:, 0, :]


def run_b200(args):
    import torch
    import torch.distributed as dist
    from sketchy_b200 import synth_torch as st
    from sketchy_b200._lib import Context, dist_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ctx = Context(local)  # raises without a B200: no CPU fallback
    if world > 1:  # the library's own communicator; its id travels over the launcher's process group
        uid = torch.from_numpy(ctx.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(device)
        dist.broadcast(uid, 0)
        ctx.comm_init(uid.cpu().numpy(), rank, world)
    if args.sketch_only:
        print(json.dumps({"sketch": sketch_leg(ctx, args, device)}), flush=True)
        ctx.close()
        return
    if args.pass_reads:
        ctx.set_pass_reads(args.pass_reads)
    if args.rank_mode:
        ctx.set_rank_mode(args.rank_mode)
    pass_reads_eff = None
    N, s, R, top = args.refs, args.sketch_size, args.reads, args.top
    cfg = args.config

    # ---------------- synthetic data (untimed) ----------------
    t0 = time.time()
    n_base = args.lineages
    genomes = st.random_genomes(n_base, GENOME_LEN, 3000, device)
    lo, cnt = dist_range(N, rank, world)
    lo_al = lo // st.ROW_BLOCK * st.ROW_BLOCK           # row blocks are seeded independently of the sharding
    hi = lo + cnt
    if args.row_dist == "independent":
        ref = independent_rows(hi - lo_al, s, lo_al, device)[lo - lo_al:]
        base_t = None
    else:
        base_rows = sketch_base_rows_gpu(ctx, genomes, s)
        assert all(r.size == s for r in base_rows)
        base_t = torch.from_numpy(np.stack(base_rows).astype(np.int64))
        assert int(base_t.min()) >= 0
        ref = st.expand_reference_block(base_t.to(device), lo_al, hi - lo_al, 0.02, 4000, device)[lo - lo_al:]
    off = np.arange(cnt + 1, dtype=np.uint64) * np.uint64(s)
    ctx.ref_upload_device(ref.data_ptr(), off, row_base=lo)
    pass_reads_eff = ctx.pass_reads   # automatic: by the size of this rank's shard (or --pass-reads)
    reads = st.sample_reads(genomes, R, args.read_len, 777)
    roff = np.arange(R + 1, dtype=np.uint64) * np.uint64(args.read_len)
    blob_pinned = torch.from_numpy(reads.reshape(-1)).pin_memory()
    blob = blob_pinned.numpy()
    log(f"[rank {rank}] data ready in {time.time() - t0:.1f}s: rows [{lo},{hi}) x {s}, {R} reads x {args.read_len}")

    # ---------------- CPU baseline on a bounded sample (rank 0, N=1 only) + parity gates ----------------
    cpu_baseline, parity = None, None
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
        # second gate: full-size passes. 2 x 4096 + 300 reads (two full passes and a short one) against every 32nd row
        # of the matrix, every read checked against the oracle (its merges spread over the host threads)
        n_g = min(R, 2 * PASS_READS + 300)
        sel = np.arange(0, N, 32)
        sub_ref = np.ascontiguousarray(ref_host.reshape(N, s)[sel]).reshape(-1)
        sub_off = np.arange(sel.size + 1, dtype=np.uint64) * np.uint64(s)
        t1 = time.time()
        ei2, es2, _ = oracle.predict_stream(sub_ref, sub_off, (blob[:n_g * args.read_len], roff[:n_g + 1]), K, s, SEED, top,
                                            nthreads=ncores)
        c2 = Context(local)
        gate_modes = {}
        for mode in (1, 2):
            c2.set_rank_mode(mode)
            c2.ref_upload(sub_ref, sub_off)
            b2 = c2.batch().add(blob[:n_g * args.read_len], roff[:n_g + 1])
            gi2, gs2 = c2.predict_stream(b2, K, s, SEED, top)
            b2.close()
            assert (gi2 == ei2).all() and (gs2 == es2).all(), f"GPU predict (rank mode {mode}) differs from the oracle on full passes"
            gate_modes[mode] = c2.last_predict_stats()["passes"]
        c2.close()
        parity = {"sample_reads_vs_full_matrix": n_s, "full_pass_reads": n_g, "full_pass_rows": int(sel.size),
                  "rank_modes_checked": sorted(gate_modes), "seconds": round(time.time() - t1, 1)}
        log(f"[rank 0] parity gates ok: {n_s} reads vs the full matrix, {n_g} reads vs {sel.size} rows in both rank modes; "
            f"cpu {n_s / dt:.2f} reads/s")
        del ref_host, sub_ref
    del ref
    torch.cuda.empty_cache()

    # ---------------- resident batch: this rank's slice of the reads ----------------
    r_lo, r_cnt = dist_range(R, rank, world)
    batch = ctx.batch()
    if r_cnt:
        batch.add(blob[r_lo * args.read_len:(r_lo + r_cnt) * args.read_len], roff[r_lo:r_lo + r_cnt + 1] - roff[r_lo])
    batch.stage()
    d_idx = torch.zeros((R, top), dtype=torch.int32, device=device)
    d_sum = torch.zeros((R, top), dtype=torch.int64, device=device)
    ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=device)

    def step_resident():
        ctx.sums_reset()
        ctx.predict_stream_dist_device(batch, R, K, s, SEED, top, d_idx.data_ptr(), d_sum.data_ptr())

    def sync_all():
        torch.cuda.synchronize()
        ctx.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sync_all()
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
    # checksum of the whole job's answer (every read's top-N rows and sums): equal across builds, pass sizes, ranking
    # modes and GPU counts when the results are identical
    result_crc = zlib.crc32(d_sum.cpu().numpy().tobytes(), zlib.crc32(d_idx.cpu().numpy().tobytes()))
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---------------- end to end through the C ABI with host buffers ----------------
    e2e = None
    consensus_info = None
    if not args.no_e2e:
        # caller-owned page-locked result arrays: the D2H of every chunk's top-N is a DMA straight into them
        oi = torch.zeros((R, top), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        os_ = torch.zeros((R, top), dtype=torch.int64).pin_memory().numpy().view(np.uint64)
        # a streaming caller feeds reads in chunks of whole passes: chunk i+1 is normalised + 2-bit packed into pinned
        # memory by the library's host threads and copied to the device while the GPU works on chunk i (two batches,
        # double buffered). With N ranks every rank packs, copies and hashes 1/N of each chunk.
        # Chunks double in size (1, 2, 4, 8 passes, then 10 at most): the GPU starts after 4 % of the packing, and the
        # later calls are few (a call costs ~0.4 ms of host round trips on top of its passes).
        P = ctx.pass_reads
        chunks, c_lo, n_pass = [], 0, 1
        while c_lo < R:
            c_hi = min(R, c_lo + n_pass * P)
            if R - c_hi < P // 2:        # do not leave a sliver for a call of its own
                c_hi = R
            chunks.append((c_lo, c_hi))
            c_lo = c_hi
            n_pass = min(2 * n_pass, 10)
        chunk_reads = max(b1 - b0 for b0, b1 in chunks)
        pack_s = [0.0]
        NB = 3                                  # ring of batches: one being packed, one being copied, one being read by the kernels
        hbs = [ctx.batch() for _ in range(NB)]
        # the ranks of one box share its host cores; two are left to the thread that drives the GPU and to the reader
        pack_threads = max(1, ((os.cpu_count() or 1) - 2 * world) // world) if not os.environ.get("SKB_BENCH_PACK_ALL") else max(1, (os.cpu_count() or 1) // world)
        gt = genotype_table(N, args.lineages).astype(np.int16) if args.consensus else None
        calls = [np.zeros((R, gt.shape[1]), dtype=np.int16) if gt is not None else None]
        from concurrent.futures import ThreadPoolExecutor
        import queue
        pool = ThreadPoolExecutor(max_workers=4)   # C4: the genotype consensus of a chunk, formed on the host while the GPU works on the next

        def consensus_chunk(q_lo, q_hi):
            calls[0][q_lo:q_hi] = consensus_calls(oi[q_lo:q_hi], gt)

        def pack(j, p_lo, p_hi):
            t0 = time.perf_counter()
            hbs[j].clear()                      # (waits for the batch's previous copies: they are long over)
            m_lo, m_cnt = dist_range(p_hi - p_lo, rank, world)
            a, b = p_lo + m_lo, p_lo + m_lo + m_cnt
            if m_cnt:
                hbs[j].add(blob[a * args.read_len:b * args.read_len], roff[a:b + 1] - roff[a], nthreads=pack_threads)
            hbs[j].stage()                      # enqueues the H2D on the copy stream and returns: the packer goes on
            pack_s[0] += time.perf_counter() - t0

        trace = {"first_pack_ms": 0.0, "call_ms": [0.0] * len(chunks), "wait_pack_ms": [0.0] * len(chunks)}

        def step_e2e():
            ctx.sums_reset()
            free = threading.Semaphore(NB)
            ready = queue.Queue()

            def producer():                     # the caller's reader thread: packs chunk after chunk, NB - 1 ahead at most
                for ci, (p_lo, p_hi) in enumerate(chunks):
                    free.acquire()
                    pack(ci % NB, p_lo, p_hi)
                    ready.put(ci)

            th = threading.Thread(target=producer)
            t_a = time.perf_counter()
            th.start()
            pending = []
            for ci, (q_lo, q_hi) in enumerate(chunks):
                t_w = time.perf_counter()
                got = ready.get()
                assert got == ci
                t_a2 = time.perf_counter()
                if ci == 0:
                    trace["first_pack_ms"] += (t_a2 - t_a) * 1e3
                else:
                    trace["wait_pack_ms"][ci] += (t_a2 - t_w) * 1e3
                ctx.predict_stream_dist(hbs[ci % NB], q_hi - q_lo, K, s, SEED, top, out=(oi[q_lo:q_hi], os_[q_lo:q_hi]),
                                        report=rank == 0)   # H2D wait + kernels (+ exchange) + D2H of the merged top-N
                trace["call_ms"][ci] += (time.perf_counter() - t_a2) * 1e3
                free.release()
                if gt is not None and rank == 0:
                    pending.append(pool.submit(consensus_chunk, q_lo, q_hi))
            th.join()
            for f in pending:
                f.result()

        for _ in range(2):
            step_e2e()
        sync_all()
        t1 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 3))
        pack_s[0] = 0.0
        trace["first_pack_ms"] = 0.0
        trace["call_ms"] = [0.0] * len(chunks)
        trace["wait_pack_ms"] = [0.0] * len(chunks)
        for _ in range(n_e2e):
            step_e2e()
        sync_all()
        dt = (time.perf_counter() - t1) / n_e2e
        if world > 1:
            t = torch.tensor([dt], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        packed = sum(-(-((b1 - b0) * (args.read_len + 1)) // 32) * 32 for b0, b1 in chunks)
        h2d = (packed // 4 + packed // 8 + (packed // 1024 + R) * 9 + R * 8) // world
        e2e = {"value": R / dt, "unit": "reads/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(R * top * 12), "ms_per_step": dt * 1e3,
               "includes": f"host normalise + 2-bit pack into pinned memory (chunks of {P}, {2 * P}, {4 * P} ... up to {chunk_reads} reads, each rank its 1/{world} "
                           f"of a chunk; a reader thread packs up to {NB - 1} chunks ahead, copies run on the copy stream), all kernels, "
                           + ("the NCCL exchanges, " if world > 1 else "") + "D2H of the top-N into page-locked host arrays"
                           + (", the per-read genotype consensus on the host" if gt is not None else ""),
               "host_threads": pack_threads, "host_pack_ms_per_step": pack_s[0] / n_e2e * 1e3,
               "h2d_bytes_note": "per rank",
               "timeline_ms": {"chunk_reads": [b1 - b0 for b0, b1 in chunks],
                               "first_pack_and_copy": round(trace["first_pack_ms"] / n_e2e, 3),
                               "library_call_per_chunk": [round(x / n_e2e, 3) for x in trace["call_ms"]],
                               "wait_for_chunk_packed": [round(x / n_e2e, 3) for x in trace["wait_pack_ms"]]}}
        if rank == 0:
            # the e2e result must equal the resident result
            assert (oi == d_idx.cpu().numpy().view(np.uint32)).all() and (os_ == d_sum.cpu().numpy().view(np.uint64)).all()
            if gt is not None:
                from sketchy_b200.api import consensus_value
                chk = np.linspace(0, R - 1, 2000).astype(int)
                for r_ in chk:   # the vectorised consensus == the per-read rule of the host mirror
                    for c_ in range(gt.shape[1]):
                        assert str(calls[0][r_, c_]) == consensus_value([str(x) for x in gt[oi[r_].astype(np.int64), c_]])
                consensus_info = {"columns": int(gt.shape[1]), "reads": R, "calls_crc32": f"{zlib.crc32(calls[0].tobytes()):08x}",
                                  "checked_against_host_mirror": int(chk.size)}
        pool.shutdown()
        for x in hbs:
            x.close()

    # ---------------- extras (rank 0 at N=1): the zero-hit variant of C3 and the sketch leg ----------------
    zero_hit, sketch_info = None, None
    peak, how = measured_peak_gbs()
    # (extras never take the predict line down with them: a failure is reported in their place)
    if rank == 0 and world == 1 and not args.no_extras and cfg == "c3":
        try:
            zero_hit = zero_hit_variant(ctx, args, device, blob, roff, peak, how)
        except Exception as e:
            zero_hit = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_extras and args.sketch_genomes > 0:
        try:
            sketch_info = sketch_leg(ctx, args, device)
        except Exception as e:
            sketch_info = {"error": repr(e)}

    if rank == 0:
        passes = stats["passes"]
        # query hashes that actually enter a pass (the membership prefilter drops the ones no reference row holds)
        keys_per_pass = stats.get("member_hashes", stats["query_hashes"]) / max(passes, 1)
        bytes_per_launch = cnt * s * 8 + keys_per_pass * 8
        avg_ms = stream_ms / max(stream_n, 1)
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        ms_per_step = ms / args.steps
        traffic, traffic_src = ncu_traffic_bytes()
        out = {
            "metric": f"predict reads/s (streaming, {R // 1000}k 5kb reads vs {N // 1000}k x s={s} reference, top {top})",
            "value": R / (ms_per_step * 1e-3), "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if cfg == "c5" else "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": config_of(cfg, N, s, R, top, args, world),
            "passes": {"per_step": passes, "reads_per_pass_max": pass_reads_eff},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "fused_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": f"of {how}",
                         # ncu capture of a full single-GPU C3 pass: comparable with bytes_per_launch there only
                         "traffic": traffic if world == 1 and (N, s) == (40000, 10000) else None,
                         "traffic_source": traffic_src,
                         "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
                         "launches": stream_n, "reads_per_launch": R / max(passes, 1),
                         "stream_share_of_step": stream_ms / ms if ms > 0 else None},
            "roofline_zero_hit": zero_hit,
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "predict_stats": stats,
            "result_crc32": f"{result_crc:08x}",
            "parity": parity,
            "cpu_baseline": cpu_baseline,
            "consensus": consensus_info,
            "sketch": sketch_info,
        }
        print(json.dumps(out), flush=True)
    batch.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def zero_hit_variant(ctx, args, device, blob, roff, peak, how):
    """The same reads against 40,000 independent rows: no read hits any row, so a pass is the stream and
    the filter probe alone — the upper bracket of the streaming kernel next to the hit-heavy lineage matrix. (The
    reference-membership prefilter drops nearly every query hash here; the few false positives keep the passes honest:
    they stream the whole matrix against a nearly empty filter.)"""
    import torch
    from sketchy_b200._lib import Context
    N, s, R, top = args.refs, args.sketch_size, args.reads, args.top
    c2 = Context(ctx.device)
    ref = independent_rows(N, s, 0, device)
    c2.ref_upload_device(ref.data_ptr(), np.arange(N + 1, dtype=np.uint64) * np.uint64(s), row_base=0)
    del ref
    torch.cuda.empty_cache()
    b2 = c2.batch().add(blob, roff)
    b2.stage()
    d_i = torch.zeros((R, top), dtype=torch.int32, device=device)
    d_s = torch.zeros((R, top), dtype=torch.int64, device=device)

    def step():
        c2.sums_reset()
        c2.predict_stream_device(b2, K, s, SEED, top, d_i.data_ptr(), d_s.data_ptr())

    for _ in range(2):
        step()
    c2.synchronize()
    c2.prof_reset(); c2.prof_enable(True)
    t1 = time.perf_counter()
    n_it = 3
    for _ in range(n_it):
        step()
    c2.synchronize()
    dt = (time.perf_counter() - t1) / n_it
    sm, sn = c2.prof_get("stream")
    c2.prof_enable(False)
    st_ = c2.last_predict_stats()
    # with all sums zero the ranking is the first `top` rows for every read
    assert (d_i.cpu().numpy() == np.arange(top)[None, :]).all() and int(d_s.abs().sum()) == 0
    b2.close(); c2.close()
    avg_ms = sm / max(sn, 1)
    bytes_per_launch = N * s * 8 + st_.get("member_hashes", 0) / max(st_["passes"], 1) * 8
    ach = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    return {"bound": "hbm", "kernel": "fused_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "peak_source": f"of {how}", "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms, "launches": sn,
            "value": R / dt, "value_unit": "reads/s", "passes_per_step": st_["passes"],
            "member_hashes_per_step": st_.get("member_hashes"),
            "workload": workload_name("c3i", N, s, R, top)}


def config_of(cfg, N, s, R, top, args, world):
    """The `config` object of a bench line: what defines the workload and its placement, and nothing that a run
    measures — both arms (`--impl b200`, `--impl reference`) print the same dict for the same command line."""
    per_gpu = (N + world - 1) // world
    return {"workload": workload_name(cfg, N, s, R, top), "refs": N, "sketch_size": s, "reads": R,
            "read_len": args.read_len, "k": K, "top": top, "lineages": args.lineages, "rows": args.row_dist,
            "l2": "reference matrix (%.2f GB per GPU) is larger than L2; streamed from HBM every pass" % (per_gpu * s * 8 / 1e9),
            "parallelism": (f"reference rows sharded over {world} GPUs, reads hashed 1/{world} per rank; NCCL all-gather of "
                            "the query lists and of the local top-N (library communicator)") if world > 1 else "1 GPU"}


def sketch_leg(ctx, args, device):
    """C2 on a bounded sample (k=16, s=1000, 2.8 Mbp assemblies): kernel-only Gbp/s from a packed batch resident in HBM,
    Gbp/s per library call from host ASCII, FASTA on a RAM disk -> .msh through the `sketchy sketch` binary, and the
    oracle's `_sketch_files` port on all host cores (one file per thread, as the reference's rayon pool does,
    src/sketchy.rs:470-472) on a smaller sample."""
    import shutil
    import tempfile
    import oracle
    from sketchy_b200 import synth_torch as st
    from sketchy_b200 import build as skb_build
    ng = args.sketch_genomes
    distinct = [st.random_genomes(1, GENOME_LEN, 9000 + g, device)[0].cpu().numpy() for g in range(min(ng, 8))]
    recs = [distinct[g % len(distinct)] for g in range(ng)]      # 8 distinct genomes, repeated (hashing cost is data independent)
    gbp = ng * GENOME_LEN / 1e9
    sb = ctx.batch().add_records(recs)
    sb.stage()
    ctx.sketch(sb, K, 1000, SEED)                                  # warm-up
    ctx.prof_reset(); ctx.prof_enable(True)
    ctx.synchronize()
    t1 = time.perf_counter()
    n_it = 3
    for _ in range(n_it):
        sk_out, _, _ = ctx.sketch(sb, K, 1000, SEED)
    ctx.synchronize()
    dt_res = (time.perf_counter() - t1) / n_it
    hash_ms, _ = ctx.prof_get("hash")
    sel_ms, _ = ctx.prof_get("select")
    ctx.prof_enable(False)
    assert all(h.size == 1000 for h, _ in sk_out)
    sb.close()
    # per call from host ASCII: normalise + pack into pinned memory, H2D, kernels, D2H of the sketches. The first call of a
    # batch also allocates its page-locked buffers; a pipelined caller reuses its batches, so the second call is the figure
    t1 = time.perf_counter()
    hb = ctx.batch().add_records(recs)
    sk2, _, _ = ctx.sketch(hb, K, 1000, SEED)
    dt_host_first = time.perf_counter() - t1
    t1 = time.perf_counter()
    hb.clear()
    hb.add_records(recs)
    sk2, _, _ = ctx.sketch(hb, K, 1000, SEED)
    dt_host = time.perf_counter() - t1
    hb.close()
    assert all((a[0] == b[0]).all() for a, b in zip(sk2, sk_out))
    # FASTA on a RAM disk -> .msh through the C++ host (process start and CUDA context creation included)
    cli = None
    try:
        tmp = tempfile.mkdtemp(prefix="skb_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        paths = []
        n_cli = max(ng, 512)   # enough files for the binary's window pipeline to reach its steady state behind the start-up
        gbp_cli = n_cli * GENOME_LEN / 1e9
        for g in range(n_cli):
            pth = os.path.join(tmp, f"g{g:05d}.fa")
            with open(pth, "wb") as f:
                f.write(b">g%d\n" % g); f.write(recs[g % ng].tobytes()); f.write(b"\n")
            paths.append(pth)
        exe = skb_build.CLI
        t1 = time.perf_counter()
        r = subprocess.run([exe, "sketch", "-k", "16", "-s", "1000", "-o", os.path.join(tmp, "ref.msh"), "-i", *paths],
                           capture_output=True, text=True, timeout=600, env=dict(os.environ, SKB_TRACE_SKETCH="1"))
        dt_cli = time.perf_counter() - t1
        import re
        marks = {}   # the binary's own timeline (ms after main): context ready, all windows sketched, .msh written
        for key, pat in (("context_ready_ms", r"context ready ([0-9.]+) ms"), ("windows_done_ms", r"all windows done ([0-9.]+) ms"),
                         ("msh_written_ms", r"\.msh written ([0-9.]+) ms")):
            m = re.search(pat, r.stderr)
            if m:
                marks[key] = float(m.group(1))
        if r.returncode == 0:
            t1 = time.perf_counter()   # the same binary on one file: what of the wall time is process start + context creation
            subprocess.run([exe, "sketch", "-k", "16", "-s", "1000", "-o", os.path.join(tmp, "one.msh"), "-i", paths[0]],
                           capture_output=True, text=True, timeout=600)
            dt_one = time.perf_counter() - t1
            cli = {"files": n_cli, "gbp": gbp_cli, "gbp_per_s": gbp_cli / dt_cli, "wall_s": dt_cli, "one_file_wall_s": dt_one,
                   # (CUDA start-up varies by hundreds of ms from run to run: no figure when the difference drowns in it)
                   "gbp_per_s_beyond_start_up": gbp_cli * (n_cli - 1) / n_cli / (dt_cli - dt_one) if dt_cli - dt_one > 0.05 else None,
                   "in_process_timeline": marks,
                   "gbp_per_s_after_context_ready": (gbp_cli / ((marks["windows_done_ms"] - marks["context_ready_ms"]) * 1e-3)
                                                     if "windows_done_ms" in marks and "context_ready_ms" in marks else None),
                   "after_context_ready_note": "the first windows are read while the context comes up; packing, copies, kernels and the remaining reads follow it",
                   "host_threads": os.cpu_count(),
                   "msh_bytes": os.path.getsize(os.path.join(tmp, "ref.msh"))}
        else:
            cli = {"error": r.stderr[-300:]}
        shutil.rmtree(tmp, ignore_errors=True)
    except Exception as e:   # the sketch leg is an extra: never take the predict line down with it
        cli = {"error": repr(e)}
    # CPU: the oracle's sketch driver, one file per host thread
    ncores = os.cpu_count() or 1
    n_cpu = min(ng, 2 * ncores)
    t1 = time.perf_counter()
    osk, _, _ = oracle.sketch_groups([r_.tobytes() for r_ in recs[:n_cpu]], list(range(n_cpu)), n_cpu, K, 1000, SEED, nthreads=ncores)
    dt_cpu = time.perf_counter() - t1
    assert all((osk[g][0] == sk_out[g][0]).all() for g in range(n_cpu)), "GPU sketches differ from the oracle"
    kern_gbp = gbp / ((hash_ms + sel_ms) / n_it * 1e-3)
    bound = 600.0   # ~30 IMAD + ~25 ALU per k-mer at 64 lanes/clk/SM x 148 SMs x 1.965 GHz
    inst_per_kmer = 86.0   # ncu: smsp__inst_executed.sum x 32 / k-mers (profiles/r02_hash_kernel_ncu.md); SASS body: 75
    issue_bound = 148 * 4 * 1.965 * 32 / inst_per_kmer   # Gbp/s at one warp instruction per cycle per scheduler
    return {"workload": f"C2 sample: {ng} x 2.8 Mbp assemblies, k=16, s=1000",
            "kernel_gbp_per_s": kern_gbp, "hash_kernel_ms": hash_ms / n_it, "select_kernel_ms": sel_ms / n_it,
            "resident_call_gbp_per_s": gbp / dt_res, "resident_call_ms": dt_res * 1e3,
            "host_call_gbp_per_s": gbp / dt_host, "host_call_ms": dt_host * 1e3, "host_call_first_ms": dt_host_first * 1e3,
            "host_call_includes": "normalise + 2-bit pack into pinned memory (all host threads), H2D, kernels, D2H of the sketches; "
                                  "a reused batch (the first call of a batch, which allocates its page-locked buffers: host_call_first_ms)",
            "cli_fasta_to_msh": cli,
            "roofline": {"bound": "integer pipes", "achieved": kern_gbp, "peak": bound, "unit": "Gbp/s", "frac": kern_gbp / bound,
                         "peak_source": "estimate (SURVEY App. D.2: 55 instructions per k-mer at the 1965 MHz maximum clock)",
                         "issue_bound_gbp_per_s": issue_bound, "frac_of_issue_bound": kern_gbp / issue_bound,
                         "issue_bound_source": f"{inst_per_kmer:.0f} warp instructions per 32 k-mers executed (ncu), one per cycle per "
                                               "scheduler, 148 SMs x 4 schedulers x 1.965 GHz; pipe utilisation (ALU 65 %, FMA 40 %, "
                                               "issue 71 %) in profiles/r02_hash_kernel_ncu.md"},
            "cpu_baseline": {"value": n_cpu * GENOME_LEN / 1e9 / dt_cpu, "unit": "Gbp/s", "cores": ncores, "kind": "port",
                             "sample": f"{n_cpu} of the assemblies, one file per host thread as rayon does "
                                       f"(src/sketchy.rs:470-472), {dt_cpu:.1f}s; sketches equal the GPU's"}}


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
    if args.config == "c5":   # bounded sample of the 80 GB reference: the rows one of eight GPUs would hold
        N = N // 8
    t0 = time.time()
    n_base = args.lineages
    genomes = st.random_genomes(n_base, GENOME_LEN, 3000, device)
    ncores = os.cpu_count() or 1
    if args.row_dist == "independent":
        ref = independent_rows(N, s, 0, device).cpu().numpy().view(np.uint64).reshape(-1)
    else:
        gh = [genomes[l].cpu().numpy().tobytes() for l in range(n_base)]
        sk, _, _ = oracle.sketch_groups(gh, list(range(n_base)), n_base, K, s, SEED, nthreads=ncores)
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
    R_, N_ = args.reads, args.refs
    out = {"impl": "reference", "metric": f"predict reads/s (streaming, {R_ // 1000}k 5kb reads vs {N_ // 1000}k x s={s} reference, top {top})",
           "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
           "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak" if args.config == "c5" else "strong",
           "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": config_of(args.config, N_, s, R_, top, args, args.gpus),   # the same dict as the GPU arm's line at this N
           "sample": f"each step = a bounded sample of {n_s} consecutive reads of the workload against "
                     + (f"the full {N} x {s} matrix" if N == N_ else f"{N} x {s} rows (one GPU's shard of the {N_}-row reference)"),
           "cpu_baseline": {"value": v, "unit": "reads/s", "cores": 1, "kind": "port",
                            "sample": f"{n_s} reads per step vs {N}x{s} rows; C++ restatement of sketchy "
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
