#!/usr/bin/env python
"""Instruction budget of fused_kernel per 256-hash chunk, by stage, from an `ncu --set full --import-source on` report.

usage: tools/inst_budget.py <report.ncu-rep> <chunks-in-the-profiled-launch> > profiles/<round>_fused_kernel_inst_budget.md
(C3 on one GPU: 40,000 rows x 10,000 hashes / 256 = 1562500 chunks.) Stages are located by marker strings in
kernels_predict.cu, so the table follows the source as it moves; code inlined from CUDA's own headers (ballots,
__syncwarp, reductions) is its own row."""
import os
import subprocess
import sys
import tempfile

rep, chunks = sys.argv[1], float(sys.argv[2])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cu = os.path.join(root, "sketchy_b200", "csrc", "kernels_predict.cu")
tsv = tempfile.mktemp(suffix=".tsv")
subprocess.run([sys.executable, os.path.join(root, "tools", "prof_lines.py"), rep, "fused_kernel", cu, "1"],
               env=dict(os.environ, DUMP_TSV=tsv), check=True, stdout=subprocess.DEVNULL)
rows = []
for l in open(tsv):
    fn, ln, s_, i_ = l.rstrip("\n").split("\t")
    rows.append((fn, int(ln), int(s_), int(i_)))
src = open(cu).read().split("\n")


def at(txt, nth=0):
    hits = [i + 1 for i, l in enumerate(src) if txt in l]
    if not hits:
        raise KeyError(txt)
    return hits[nth]


def fn_range(first, last):  # [line of `first`, line of `last`]
    return at(first), at(last)


M = "kernels_predict.cu"
R = {
    "probe_fn": fn_range("uint32_t bloom_word(uint32_t lo)", "return r & 1u;"),
    "table_home": fn_range("uint32_t table_home(uint64_t h", "return (uint32_t)((h * 0x9E3779B97F4A7C15ull)"),
    "load_slot": fn_range("SkbSlot load_slot(const SkbSlot* p)", "s.meta = ((unsigned long long)v.w << 32) | v.z;"),
    "table_lookup": fn_range("bool table_lookup(const SkbTable& t", "slot = (slot + 1) & (t.cap - 1);"),
    # an asm statement is attributed to its LAST line: take whole functions (start .. next function's start - 1)
    "wait_free": (at("void mbar_wait_free(uint64_t* bar"), at("void mbar_wait_sleepy(uint64_t* bar") - 1),
    "mbar_try": (at("bool mbar_try(uint64_t* bar, uint32_t parity) {"), at("void mbar_wait(uint64_t* bar, uint32_t parity) {") - 1),
    "sleepy": (at("void mbar_wait_sleepy(uint64_t* bar"), at("void bulk_load(void* smem_dst") - 1),
    "mbar_other_a": (at("void mbar_init(uint64_t* bar"), at("bool mbar_try(uint64_t* bar, uint32_t parity) {") - 1),
    "mbar_wait": (at("void mbar_wait(uint64_t* bar, uint32_t parity) {"), at("void mbar_wait_free(uint64_t* bar") - 1),
    "bulk": (at("void bulk_load(void* smem_dst"), at("void count_hit(uint32_t* cbuf, uint32_t rd)") - 1),
    "hit": fn_range("void count_hit(uint32_t* cbuf, uint32_t rd)", "for (uint32_t j = 0; j < c; ++j) count_hit<CPW>(cbuf, t.reads[st + j]);"),
    "subiter": fn_range("struct SubIter {", "settle(a, r1);"),
    "try_close": (at("auto try_close = [&]"), at("auto finish_batch = [&]") - 1),
    "finish": (at("auto finish_batch = [&]"), at("auto start_batch = [&]") - 1),
    "start": (at("auto start_batch = [&]"), at("uint64_t policy;") - 1),
    "issue_copy": (at("auto issue_copy = [&]"), at("SubIter it, pre;") - 1),
    "row_ctl": (at("SubIter it, pre;"), at("for (uint32_t ch = 0; ch < (uint32_t)FS_CHUNKS; ++ch) {") - 1),
    "chunk_load": (at("for (uint32_t ch = 0; ch < (uint32_t)FS_CHUNKS; ++ch) {"), at("uint32_t pm = 0;") - 1),
    "probe_loop": (at("uint32_t pm = 0;"), at("const uint32_t anyb = __ballot_sync") - 1),
    "compact": (at("const uint32_t anyb = __ballot_sync"), at("const uint32_t tail = (qhead + qn) & (FS_QCAP - 1);") - 1),
    "push": (at("const uint32_t tail = (qhead + qn) & (FS_QCAP - 1);"), at("// lookups: a batch is consumed one chunk after it was issued") - 1),
    "sched": (at("// lookups: a batch is consumed one chunk after it was issued"), at("===== rank warps: cumulative sums") - 1),
    "rank": (at("===== rank warps: cumulative sums"), at("// per-read counts of the tracked rows (the rows that define the bounds)") - 1),
}


def inr(r, *names):
    return r[0] == M and any(R[n][0] <= r[1] <= R[n][1] for n in names)


CATS = [
    ("filter probe: `bloom_probe` + the `pm` loop (8 hashes per lane)", lambda r: inr(r, "probe_fn", "probe_loop")),
    ("staged chunk loads (4 x LDS.128), ring refill, sub-tile iterator, bulk-copy issue", lambda r: inr(r, "chunk_load", "subiter", "issue_copy", "bulk")),
    ("compaction of the passers: ballots, prefix, room check", lambda r: inr(r, "compact")),
    ("FIFO push: 8 predicated 8-byte stores + tags + bookkeeping", lambda r: inr(r, "push")),
    ("`start_batch` (FIFO read, home slot, 16-byte table load) + `table_home`", lambda r: inr(r, "start", "table_home")),
    ("`finish_batch` (key compare, re-queue, outstanding counts) + `try_close`", lambda r: inr(r, "finish", "try_close", "load_slot")),
    ("`apply_hit` / `count_hit` (shared-memory atomics per read of the key)", lambda r: inr(r, "hit")),
    ("lookup scheduling, row open / close, loop control", lambda r: inr(r, "sched", "row_ctl")),
    ("consumer warps spinning for a row buffer (`mbar_wait_free`)", lambda r: inr(r, "wait_free")),
    ("rank warps polling for a finished row (`mbar_try` + nanosleep)", lambda r: inr(r, "mbar_try", "sleepy")),
    ("other barrier traffic: staging-ring waits, arrives, expect_tx", lambda r: inr(r, "mbar_other_a", "mbar_wait")),
    ("rank warps: row totals, prefix scan, candidate walk, buffer clear", lambda r: inr(r, "rank")),
    ("warp intrinsics inlined from CUDA headers (ballot, syncwarp, shuffles, reductions)", lambda r: r[0] != M and r[0] != ""),
]
tot_i = sum(r[3] for r in rows)
tot_s = sum(r[2] for r in rows)
print(f"| where (`kernels_predict.cu`) | warp instr / chunk | share | stall samples |\n|---|---:|---:|---:|")
seen = 0
for name, pred in CATS:
    sel = [r for r in rows if pred(r)]
    i_ = sum(r[3] for r in sel)
    s_ = sum(r[2] for r in sel)
    seen += i_
    print(f"| {name} | {i_ / chunks:.1f} | {100 * i_ / tot_i:.1f} % | {100 * s_ / tot_s:.1f} % |")
print(f"| everything else (prologue, synchronous fallbacks, unattributed) | {(tot_i - seen) / chunks:.1f} | {100 * (tot_i - seen) / tot_i:.1f} % | |")
print(f"| **total** ({tot_i / 1e6:.1f} M warp instructions) | **{tot_i / chunks:.1f}** | 100 % | |")
