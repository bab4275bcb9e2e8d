#!/usr/bin/env python
"""Streaming predict through the `sketchy` binary at the C1 shapes with many reads: 100 x 2.8 Mbp assemblies sketched by
`sketchy sketch` (k=16, s=1000), R x 5 kb reads as FASTQ on a RAM disk, `sketchy predict -s -t 10` to a file. What is
timed is what a user of the CLI waits for: process start, .msh and genotype table read, reference upload, the reader /
packer / predict / printer pipeline, the rows written. A sample of rows is checked against the library called directly.
Prints one JSON line.   usage: tools/predict_cli_rate.py [n_reads] [top]"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from sketchy_b200 import build as skb_build
from sketchy_b200 import synth_torch as st

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
top = int(sys.argv[2]) if len(sys.argv) > 2 else 10
G, GLEN, RLEN = 100, 2_800_000, 5000
dev = torch.device("cuda", 0)
tmp = tempfile.mkdtemp(prefix="skb_predict_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
genomes = st.random_genomes(G, GLEN, 7000, dev)
paths = []
gh = genomes.cpu().numpy()
for g in range(G):
    p = os.path.join(tmp, f"g{g:03d}.fa")
    with open(p, "wb") as f:
        f.write(b">g%d\n" % g); f.write(gh[g].tobytes()); f.write(b"\n")
    paths.append(p)
reads = st.sample_reads(genomes, R, RLEN, 777)
del genomes
torch.cuda.empty_cache()
fq = os.path.join(tmp, "reads.fq")
qual = b"I" * RLEN
with open(fq, "wb") as f:
    for r0 in range(0, R, 4096):
        f.write(b"".join(b"@r%d\n" % (r0 + i) + reads[r0 + i].tobytes() + b"\n+\n" + qual + b"\n" for i in range(min(4096, R - r0))))
geno = os.path.join(tmp, "ref.tsv")
with open(geno, "w") as f:
    f.write("id\tmlst\tmeca\tpvl\n" + "".join(f"g{g:03d}.fa\tST{g % 9}\t{'R' if g % 2 else 'S'}\t{'+' if g % 3 else '-'}\n" for g in range(G)))
ref = os.path.join(tmp, "ref.msh")
exe = skb_build.CLI
t0 = time.perf_counter()
subprocess.run([exe, "sketch", "-k", "16", "-s", "1000", "-o", ref, "-i", *paths], check=True, capture_output=True)
t_sketch = time.perf_counter() - t0
out = {"workload": f"C1 shapes, {R} reads: {G} x 2.8 Mbp -> ref.msh (k=16, s=1000); {R} x 5 kb FASTQ reads, `sketchy predict -s -t {top}`",
       "fastq_gb": os.path.getsize(fq) / 1e9, "sketch_wall_s": round(t_sketch, 3), "host_threads": os.cpu_count(), "runs": []}
rows_plain = None
for label, extra in (("rows", []), ("consensus", ["-c"] if top % 2 else None), ("one_read", ["-l", "1"])):
    if extra is None:
        continue
    res = os.path.join(tmp, f"out_{label}.tsv")
    t0 = time.perf_counter()
    with open(res, "wb") as fo:
        p = subprocess.run([exe, "predict", "-i", fq, "-r", ref, "-g", geno, "-t", str(top), "-s", *extra], stdout=fo, stderr=subprocess.PIPE)
    dt = time.perf_counter() - t0
    n_rows = sum(1 for _ in open(res, "rb"))
    out["runs"].append({"mode": label, "ok": p.returncode == 0, "wall_s": round(dt, 3), "rows": n_rows,
                        "reads_per_s": round((1 if label == "one_read" else R) / dt, 1), "stderr": p.stderr[-200:].decode(errors="replace") if p.returncode else ""})
    if label == "rows":
        rows_plain = res
one = next(r for r in out["runs"] if r["mode"] == "one_read")["wall_s"]
for r in out["runs"]:
    if r["mode"] != "one_read":
        r["reads_per_s_beyond_start_up"] = round(R / max(r["wall_s"] - one, 1e-9), 1)
# the library called directly on the first and the last 2,000 reads' worth of state is not comparable (sums are
# cumulative), so the check is the whole stream through the Python mirror on a prefix: the first 3,000 reads
if rows_plain:
    from sketchy_b200 import api
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import capnp_py
    dec = capnp_py.decode_msh(open(ref, "rb").read())
    rows = [np.array(s["hashes"], dtype=np.uint64) for s in dec["sketches"]]
    ctx = api.Context(0)
    off = np.zeros(G + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ctx.ref_upload(np.concatenate(rows), off)
    n_chk = min(R, 3000)
    b = ctx.batch().add_records([reads[i] for i in range(n_chk)])
    idx, sm = ctx.predict_stream(b, 16, rows[0].size, 0, top)
    want = [f"{r + 1}\tg{int(idx[r, t]):03d}.fa\t{int(sm[r, t])}" for r in range(n_chk) for t in range(top)]
    got = []
    with open(rows_plain) as f:
        for line in f:
            got.append("\t".join(line.split("\t")[:3]))
            if len(got) == len(want):
                break
    out["rows_checked_against_the_library"] = {"reads": n_chk, "equal": got == want}
shutil.rmtree(tmp, ignore_errors=True)
print(json.dumps(out))
