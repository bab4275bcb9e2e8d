#!/usr/bin/env python
"""Generates tests/golden/small_world.json: a tiny seeded world (genomes -> sketches, reads -> streaming predict rows)
with the outputs of the CPU oracle. The reference itself ships no fixtures and cannot be built here, so
these vectors pin the ORACLE (and through it the GPU path) against regressions, not against the Rust binary.
Run from the repo root: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from sketchy_b200 import synth  # noqa: E402

K, S, SEED, TOP = 16, 48, 42, 4
base = [synth.random_genome(4000, 11 + i) for i in range(3)]
genomes = [synth.mutate(base[g % 3], 0.003, 50 + g).tobytes().decode() for g in range(9)]
genomes[4] = genomes[4][:1500] + "NNNNNNNNNN" + genomes[4][1500:3000].lower() + "\n" + genomes[4][3000:]
sk, bases, kmers = oracle.sketch_groups([g.encode() for g in genomes], list(range(9)), 9, K, S, SEED)
blob, roff, _ = synth.sample_reads(base, 20, 600, 3)
reads = ["ACGTAC"] + [blob[int(roff[i]):int(roff[i + 1])].tobytes().decode() for i in range(20)]
rows = [h for h, _ in sk]
off = np.zeros(10, dtype=np.uint64)
off[1:] = np.cumsum([r.size for r in rows])
idx, sums, final = oracle.predict_stream(np.concatenate(rows), off, [r.encode() for r in reads], K, S, SEED, TOP)
out = {
    "k": K, "s": S, "seed": SEED, "top": TOP,
    "genomes": genomes,
    "sketches": [{"hashes": [str(int(x)) for x in h], "counts": [int(x) for x in c]} for h, c in sk],
    "seq_length": [int(x) for x in bases], "num_valid_kmers": [int(x) for x in kmers],
    "reads": reads,
    "predict_idx": idx.tolist(), "predict_sum": [[int(v) for v in r] for r in sums.tolist()],
    "final_sums": [int(x) for x in final],
}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_world.json"), "w"))
print("written", len(json.dumps(out)), "bytes")
