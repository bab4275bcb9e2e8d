#!/usr/bin/env python
"""C2 through the `sketchy sketch` binary: F FASTA files of 2.8 Mbp on a RAM disk -> .msh, with 1, 2, ... ranks (one process
per GPU, files partitioned over the ranks, no collective on the data path). Prints one JSON line.
usage: tools/sketch_scale.py [n_files] [worlds, e.g. 1,2,4,8]"""
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

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 256
worlds = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2").split(",")]
worlds = [w for w in worlds if w <= torch.cuda.device_count()]
GLEN = 2_800_000
tmp = tempfile.mkdtemp(prefix="skb_sketch_scale_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
dev = torch.device("cuda", 0)
paths = []
t0 = time.time()
for lo in range(0, n_files, 32):
    g = st.random_genomes(min(32, n_files - lo), GLEN, 5000 + lo, dev).cpu().numpy()
    for i in range(g.shape[0]):
        p = os.path.join(tmp, f"g{lo + i:05d}.fa")
        with open(p, "wb") as f:
            f.write(b">g%d\n" % (lo + i)); f.write(g[i].tobytes()); f.write(b"\n")
        paths.append(p)
gen_s = time.time() - t0
torch.cuda.empty_cache()
lst = os.path.join(tmp, "files.txt")
open(lst, "w").write("\n".join(paths) + "\n")
gbp = n_files * GLEN / 1e9
out = {"workload": f"C2: {n_files} x 2.8 Mbp FASTA files on a RAM disk -> .msh (k=16, s=1000), `sketchy sketch`", "gbp": gbp,
       "generate_s": round(gen_s, 1), "runs": []}
ref_bytes = None
for w in worlds:
    comm = os.path.join(tmp, f"comm_{w}")
    msh = os.path.join(tmp, f"out_{w}.msh")
    t1 = time.perf_counter()
    procs = []
    for r in range(w):
        env = dict(os.environ, SKETCHY_B200_RANK=str(r), SKETCHY_B200_WORLD=str(w), SKETCHY_B200_DEVICE=str(r), SKETCHY_B200_COMM_FILE=comm)
        if os.environ.get("SKB_SCALE_CVD"):   # every rank sees its own GPU only (what a job launcher's GPU binding does)
            env.update(CUDA_VISIBLE_DEVICES=str(r), SKETCHY_B200_DEVICE="0")
        procs.append(subprocess.Popen([skb_build.CLI, "sketch", "-k", "16", "-s", "1000", "-o", msh, "-i", *paths], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE))
    outs = [p.communicate(timeout=900) for p in procs]
    dt = time.perf_counter() - t1
    ok = all(p.returncode == 0 for p in procs)
    b = open(msh, "rb").read() if ok else b""
    if ref_bytes is None:
        ref_bytes = b
    out["runs"].append({"ranks": w, "wall_s": round(dt, 3), "gbp_per_s": round(gbp / dt, 3), "ok": ok, "same_msh_as_1_rank": b == ref_bytes,
                        "stderr": "" if ok else outs[0][1][-300:].decode(errors="replace")})
    if os.environ.get("SKB_TRACE_SKETCH"):
        sys.stderr.write(f"--- {w} rank(s) ---\n" + "".join(o[1].decode(errors="replace") for o in outs))
# what of the wall time is process start + CUDA context creation: the same binary on one file
t1 = time.perf_counter()
subprocess.run([skb_build.CLI, "sketch", "-k", "16", "-s", "1000", "-o", os.path.join(tmp, "one.msh"), "-i", paths[0]],
               capture_output=True, timeout=900)
out["one_file_wall_s"] = round(time.perf_counter() - t1, 3)
for r in out["runs"]:
    r["gbp_per_s_beyond_start_up"] = round(gbp / max(r["wall_s"] - out["one_file_wall_s"], 1e-9), 3)
shutil.rmtree(tmp, ignore_errors=True)
print(json.dumps(out))
