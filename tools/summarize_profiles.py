#!/usr/bin/env python
"""Turn the ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.

usage: tools/summarize_profiles.py <round-tag> <launches.csv> <fused_full.ncu-rep> [<hash_full.ncu-rep>]
"""
import collections
import csv
import subprocess
import sys

tag, launches, rep = sys.argv[1:4]
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
mine_total = 0.0
for r in rows[hi + 2:]:
    if len(r) <= iv or "unnamed" not in r[ik] or "at::" in r[ik] or "at_cuda_detail" in r[ik]:
        continue  # torch data-generation kernels of the untimed set-up are not ours
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    name = r[ik].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    agg.setdefault(name, []).append(v)
    mine_total += v
with open(f"profiles/{tag}_launches_summary.md", "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 2` (gpu__time_duration.sum, --clock-control none)\n\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes. Only this repo's\n"
            "kernels are listed (torch kernels of the untimed synthetic-data set-up are dropped). Includes the set-up\n"
            "sketching of the 40 base genomes (hash_kernel / select_kernel launches with long durations).\n\n")
    f.write("| kernel | launches | total ms | avg us | share of our GPU time |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
        f.write(f"| {k} | {len(v)} | {sum(v) / 1e3:.2f} | {sum(v) / len(v):.1f} | {100 * sum(v) / mine_total:.1f}% |\n")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(out.split("\n")))
h, u, r = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
with open(f"profiles/{tag}_fused_kernel_ncu.md", "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none` of fused_kernel (one steady-state pass of `python bench.py`)\n\n")
    f.write("| metric | value | unit |\n|---|---:|---|\n")
    for w in want:
        if w in h:
            i = h.index(w)
            f.write(f"| {w} | {r[i]} | {u[i]} |\n")
if len(sys.argv) > 4:
    out = subprocess.run(["ncu", "-i", sys.argv[4], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.split("\n")))
    h, u, r = rr[0], rr[1], rr[2]
    want_h = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
              "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum",
              "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    want_h += [x for x in h if "issue_stalled" in x and x.endswith("per_issue_active.ratio") and "not_issued" not in x]
    with open(f"profiles/{tag}_hash_kernel_ncu.md", "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of hash_kernel<K16, seed 0> (sketch leg: 64 x 2.8 Mbp assemblies = 179.2 M k-mers, k=16)\n\n")
        f.write("instructions per k-mer = smsp__inst_executed.sum x 32 / 179.2 M\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w in want_h:
            if w in h:
                i = h.index(w)
                f.write(f"| {w} | {r[i]} | {u[i]} |\n")
print("written")
