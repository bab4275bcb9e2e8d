#!/usr/bin/env python
"""Attribute ncu warp-stall samples / executed instructions of one kernel to CUDA source lines.

usage: tools/prof_lines.py <report.ncu-rep> <kernel-name-substring> <source.cu> [top_n]
Needs the library built with -lineinfo (it is) and nvdisasm / cuobjdump / ncu on PATH. ncu's SASS page and
nvdisasm list a function's instructions in the same order, so row i of one is row i of the other.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kname, srcfile = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "sketchy_b200", "libsketchy_b200.so")
tmp = tempfile.mkdtemp()
stem = os.path.splitext(os.path.basename(srcfile))[0]
subprocess.run(["cuobjdump", "-xelf", stem, so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]


def func_lines(start):
    cur, seq = None, []
    for l in dis[start + 1:]:
        if l.startswith("//---------------------"):
            break
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            seq.append(cur)
    return seq


# several template instances may match the name: take the one whose instruction count equals the profiled one
cands = [func_lines(i) for i, l in enumerate(dis) if l.startswith(".text.") and kname in l]
seq = min(cands, key=lambda q: abs(len(q) - len(data)))
assert len(seq) == len(data), (len(seq), len(data), [len(c) for c in cands])
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
srcs = {}


def text_of(key):
    """source text of (file name, line); the kernel inlines code from the headers next to it"""
    if not key:
        return ""
    fn, ln = key
    if fn not in srcs:
        path = os.path.join(os.path.dirname(os.path.abspath(srcfile)), fn)
        srcs[fn] = open(path).read().split("\n") if os.path.exists(path) else []
    lines = srcs[fn]
    return lines[ln - 1].strip()[:95] if 0 < ln <= len(lines) else ""


agg = {}
for i, r in enumerate(data):
    a = agg.setdefault(seq[i], [0, 0])
    a[0] += int(r[iS]); a[1] += int(r[iI])
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print(f"kernel {kname}: {len(seq)} SASS instrs, {toti} executed, {tot} samples")
key_i = 1 if os.environ.get("BY_INST") else 0
if os.environ.get("DUMP_TSV"):  # machine-readable: file, line, samples, executed instructions (tools/inst_budget.py)
    with open(os.environ["DUMP_TSV"], "w") as f:
        for key, (s_, i_) in agg.items():
            fn, ln = key if key else ("", 0)
            f.write(f"{fn}\t{ln}\t{s_}\t{i_}\n")
main = os.path.basename(srcfile)
for key, (s_, i_) in sorted(agg.items(), key=lambda x: -x[1][key_i])[:top_n]:
    fn, ln = key if key else ("", 0)
    where = f"{ln:5d}" if fn == main else f"{fn}:{ln}"
    print(f"{where:>5s} {s_:7d} {100 * s_ / max(tot, 1):5.1f}% inst={i_:11d}  {text_of(key)}")
