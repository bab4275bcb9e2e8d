"""Golden-vector tests: the committed small world (tests/golden/small_world.json, made by make_golden.py) must be
reproduced by the oracle (CPU) and by the CUDA path through the C ABI (GPU)."""
import json
import os

import numpy as np
import pytest

import oracle

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "small_world.json")))


def _ref():
    rows = [np.array([int(x) for x in s["hashes"]], dtype=np.uint64) for s in G["sketches"]]
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    return rows, np.concatenate(rows), off


def test_oracle_reproduces_golden():
    gs = [g.encode() for g in G["genomes"]]
    sk, bases, kmers = oracle.sketch_groups(gs, list(range(len(gs))), len(gs), G["k"], G["s"], G["seed"])
    for (h, c), e in zip(sk, G["sketches"]):
        assert [str(int(x)) for x in h] == e["hashes"] and c.tolist() == e["counts"]
    assert bases.tolist() == G["seq_length"] and kmers.tolist() == G["num_valid_kmers"]
    rows, ref, off = _ref()
    idx, sums, final = oracle.predict_stream(ref, off, [r.encode() for r in G["reads"]], G["k"], G["s"], G["seed"], G["top"])
    assert idx.tolist() == G["predict_idx"] and sums.tolist() == G["predict_sum"]
    assert final.tolist() == G["final_sums"]


@pytest.mark.gpu
def test_gpu_reproduces_golden():
    from sketchy_b200._lib import Context
    ctx = Context(0)
    gs = [g.encode() for g in G["genomes"]]
    b = ctx.batch().add_records(gs)
    sk, bases, kmers = ctx.sketch(b, G["k"], G["s"], G["seed"])
    for (h, c), e in zip(sk, G["sketches"]):
        assert [str(int(x)) for x in h] == e["hashes"] and c.tolist() == e["counts"]
    assert bases.tolist() == G["seq_length"] and kmers.tolist() == G["num_valid_kmers"]
    rows, ref, off = _ref()
    ctx.ref_upload(ref, off)
    rb = ctx.batch().add_records([r.encode() for r in G["reads"]])
    idx, sums = ctx.predict_stream(rb, G["k"], G["s"], G["seed"], G["top"])
    assert idx.tolist() == G["predict_idx"] and sums.tolist() == G["predict_sum"]
    assert ctx.sums_download().tolist() == G["final_sums"]
    ctx.close()
