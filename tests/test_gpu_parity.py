"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.
Bit-exact for every integer result (hashes, counts, totals, shared counts, top-N order with index tie-break)."""
import random

import numpy as np
import pytest

import oracle
from sketchy_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from sketchy_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


def _rand_dna(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(alphabet) for _ in range(n))


DIRTY = [
    b"",
    b"ACGT",
    b"ACGTACGTACGTACGT",
    b"T" * 16,
    b"acgtacgtnnACGTTGCAAGGCTTAACCGGTTAACCGGTTuuUUacgtRYKMACGTTGCATGCATGCATGCAGGG",
    b"ACGTTGCAAGGC\nTTAACCGGTT\r\nAACCGGTTACGTAGCTAGCTAGGATC  CGATCGATCG\tATCGATTTAGC",
    b"N" * 40,
    b"ACGTTGCAAGGCTTAAC-CGGTTAACCGG.TTACGTAGCTAGCT~AGGATCCGATCGATCGATCGATTTAGC",
]


@pytest.mark.parametrize("k", [16, 21, 11, 31, 32, 5, 1, 17])
def test_kmer_hashes_positional(ctx, k):
    """K1: every packed position's canonical k-mer hash == the oracle's emission order, record by record."""
    rng = random.Random(100 + k)
    recs = list(DIRTY) + [_rand_dna(rng, rng.randint(0, 700), b"ACGTACGTACGTN") for _ in range(40)]
    recs += [_rand_dna(rng, 5000)]
    b = ctx.batch().add_records(recs)
    seed = 42 if k % 2 else 0
    h, v = ctx.debug_kmer_hashes(b, k, seed)
    covered = np.zeros(h.size, dtype=bool)
    for r, rec in enumerate(recs):
        pos, raw_len = b.record_start(r)
        assert raw_len == len(rec)
        ln = len(oracle.normalize(rec))
        exp = oracle.kmer_hashes(rec, k, seed)
        span = slice(pos, pos + max(ln, 0))
        got = h[span][v[span] == 1]
        assert got.tolist() == exp.tolist(), (r, rec[:40])
        covered[span] = True
    assert v[~covered].sum() == 0  # padding / separators never emit
    b.close()


def _files(rng):
    big = _rand_dna(rng, 300_000)
    rep = (_rand_dna(rng, 37) * 3000)  # heavy duplicates: forces candidate overflow + retry
    files = [
        [big[:150_000], big[150_000:]],
        [],                                       # empty file
        [b"ACGTACGTAC"],                          # shorter than k
        [_rand_dna(rng, 700)],                    # fewer than s distinct k-mers
        [rep],
        [_rand_dna(rng, 50_000, b"ACGTN"), _rand_dna(rng, 10), _rand_dna(rng, 80_000)],
        [big[1000:200_000].lower()],
    ]
    return files


def test_records_where_they_lie_equal_the_blob(ctx):
    """skb_batch_add_records (pointer + length per record: how `sketchy sketch` hands over the slices of its file buffers)
    against skb_batch_add (one blob + offsets): the same packed batch — every position's hash and validity, the
    records' places, the totals — with and without explicit groups, on dirty and empty records and in both base
    count modes; a second call appends."""
    import ctypes as C
    rng = random.Random(77)
    recs = list(DIRTY) + [_rand_dna(rng, rng.randint(0, 3000), b"ACGTACGTACGTNacgt\n") for _ in range(60)] + [_rand_dna(rng, 70000)]
    groups = np.sort(np.array([rng.randint(0, 9) for _ in recs], dtype=np.uint32))
    groups[0] = 0
    keep = [np.frombuffer(r, dtype=np.uint8) if r else np.zeros(0, dtype=np.uint8) for r in recs]   # scattered buffers
    ptrs = np.array([a.__array_interface__["data"][0] if a.size else 0 for a in keep], dtype=np.uint64)
    lens = np.array([a.size for a in keep], dtype=np.uint64)
    blob = np.concatenate(keep)
    off = np.zeros(len(recs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    for g in (None, groups):
        for stripped in (False, True):
            a = ctx.batch().set_base_count(stripped).add(blob, off, g)
            b = ctx.batch().set_base_count(stripped)
            half = len(recs) // 2
            gp = None if g is None else g.ctypes.data_as(C.c_void_p)
            gp2 = None if g is None else C.c_void_p(g.ctypes.data + 4 * half)
            ctx.check(ctx.lib.skb_batch_add_records(b.h, C.c_void_p(ptrs.ctypes.data), C.c_void_p(lens.ctypes.data), gp, half, 3))
            ctx.check(ctx.lib.skb_batch_add_records(b.h, C.c_void_p(ptrs.ctypes.data + 8 * half), C.c_void_p(lens.ctypes.data + 8 * half),
                                                    gp2, len(recs) - half, 0))
            assert (a.num_groups, a.num_records, a.num_bases, a.packed_len) == (b.num_groups, b.num_records, b.num_bases, b.packed_len)
            assert all(a.record_start(r) == b.record_start(r) for r in range(len(recs)))
            ha, va = ctx.debug_kmer_hashes(a, 16, 42)
            hb, vb = ctx.debug_kmer_hashes(b, 16, 42)
            assert (va == vb).all() and (ha[va == 1] == hb[vb == 1]).all()
            ska, ba, ka = ctx.sketch(a, 16, 200, 42)
            skb_, bb, kb = ctx.sketch(b, 16, 200, 42)
            assert list(ba) == list(bb) and list(ka) == list(kb)
            assert all((x[0] == y[0]).all() and (x[1] == y[1]).all() for x, y in zip(ska, skb_))
            a.close(); b.close()
    # a null record with a length is refused
    b = ctx.batch()
    bad_p = np.array([0], dtype=np.uint64)
    bad_l = np.array([5], dtype=np.uint64)
    assert ctx.lib.skb_batch_add_records(b.h, C.c_void_p(bad_p.ctypes.data), C.c_void_p(bad_l.ctypes.data), None, 1, 1) != 0
    b.close()


@pytest.mark.parametrize("k,s,seed", [(16, 1000, 0), (16, 100, 42), (21, 500, 42), (16, 10000, 0), (31, 64, 7)])
def test_sketch_parity(ctx, k, s, seed):
    """K1+K2 == finch MashSketcher per file: hashes, occurrence counts, total_bases, total_kmers."""
    rng = random.Random(7)
    files = _files(rng)
    recs = [r for f in files for r in f]
    groups = [g for g, f in enumerate(files) for _ in f]
    exp, eb, ek = oracle.sketch_groups(recs, groups, len(files), k, s, seed, nthreads=4)
    b = ctx.batch()
    # last file has records; empty file in the middle is created by the group numbering
    b.add_records(recs, np.asarray(groups, dtype=np.uint32))
    assert b.num_groups == len(files)
    got, gb, gk = ctx.sketch(b, k, s, seed)
    for g in range(len(files)):
        assert got[g][0].tolist() == exp[g][0].tolist(), g
        assert got[g][1].tolist() == exp[g][1].tolist(), g
        assert (int(gb[g]), int(gk[g])) == (int(eb[g]), int(ek[g])), g
    b.close()


def test_total_bases_raw_or_stripped(ctx):
    """total_bases counts the raw sequence bytes by default (line breaks of a multi-line record
    included, as the reference's reader hands them over) or, as a named switch, the bases left after whitespace is
    removed. Hashes and k-mer counts are the same either way."""
    rng = random.Random(5)
    body = _rand_dna(rng, 1000)
    multi = b"\n".join(body[i:i + 60] for i in range(0, len(body), 60)) + b"\r\n"
    recs = [multi, body, b"AC GT\tACGTNACGTACGTACGTTTGACCA"]
    res = {}
    for stripped in (False, True):
        b = ctx.batch().set_base_count(stripped)
        b.add_records(recs, np.asarray([0, 1, 2], dtype=np.uint32))
        sk, bases, kmers = ctx.sketch(b, 16, 50, 0)
        res[stripped] = (sk, bases.tolist(), kmers.tolist())
        b.close()
    assert res[False][1] == [len(r) for r in recs]
    assert res[True][1] == [len(oracle.normalize(r)) for r in recs]
    assert res[False][2] == res[True][2]
    for g in range(3):
        assert res[False][0][g][0].tolist() == res[True][0][g][0].tolist()
    assert res[True][0][0][0].tolist() == res[True][0][1][0].tolist()   # line breaks do not change the sketch


def _world(seed, n_lineages, per_lineage, glen, s, n_reads, rlen, k=16, hseed=0, ragged=False):
    base = [synth.random_genome(glen, seed * 1000 + l) for l in range(n_lineages)]
    genomes = []
    for g in range(n_lineages * per_lineage):   # lineages interleaved in file order
        genomes.append(synth.mutate(base[g % n_lineages], 0.002, seed * 7919 + g))
    if ragged:
        genomes[3] = genomes[3][:s // 2 + k]     # a tiny genome: fewer than s k-mers, huge max hash
        genomes[5] = genomes[5][:k - 1]          # no k-mers at all: empty row
    sk, _, _ = oracle.sketch_groups([g.tobytes() for g in genomes], list(range(len(genomes))), len(genomes), k, s,
                                    hseed, nthreads=8)
    rows = [h for h, _ in sk]
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ref = np.concatenate(rows)
    blob, roff, src = synth.sample_reads(base, n_reads, rlen, seed + 5)
    return ref, off, blob, roff


def _check_predict(ctx, ref, off, blob, roff, k, s_query, hseed, top, pass_reads, chunks=1, modes=(0, 1)):
    """Streaming predict through the C ABI == the oracle, in every ranking mode (0 = automatic: these small worlds
    are ranked by brute force; 1 = candidate lists from per-read bounds wherever possible)."""
    n = roff.size - 1
    ei, es, esums = oracle.predict_stream(ref, off, (blob, roff), k, s_query, hseed, top)
    for mode in modes:
        ctx.set_rank_mode(mode)
        ctx.ref_upload(ref, off)
        ctx.set_pass_reads(pass_reads)
        gi_all, gs_all = [], []
        bounds = np.linspace(0, n, chunks + 1).astype(int)
        for c in range(chunks):
            lo, hi = bounds[c], bounds[c + 1]
            b = ctx.batch()
            sub_off = roff[lo:hi + 1] - roff[lo]
            b.add(blob[int(roff[lo]):int(roff[hi])] if hi > lo else np.zeros(1, np.uint8), sub_off)
            gi, gs = ctx.predict_stream(b, k, s_query, hseed, top)
            gi_all.append(gi)
            gs_all.append(gs)
            b.close()
        gi = np.concatenate(gi_all)
        gs = np.concatenate(gs_all)
        assert gi.shape == ei.shape
        bad = np.flatnonzero((gi != ei).any(axis=1) | (gs != es).any(axis=1))
        assert bad.size == 0, (mode, bad[:5], gi[bad[:2]], ei[bad[:2]], gs[bad[:2]], es[bad[:2]])
        assert (ctx.sums_download() == esums).all(), mode
    ctx.set_rank_mode(0)


@pytest.mark.parametrize("pass_reads", [1, 7, 64, 0])
def test_predict_stream_parity_uniform(ctx, pass_reads):
    ref, off, blob, roff = _world(seed=3, n_lineages=6, per_lineage=20, glen=40_000, s=500, n_reads=300, rlen=1500)
    # first reads shorter than k: all sums zero -> ranking = first `top` references in file order
    blob = np.concatenate([np.frombuffer(b"ACGTACGT", np.uint8), blob])
    roff = np.concatenate([[0], roff + 8]).astype(np.uint64)
    _check_predict(ctx, ref, off, blob, roff, 16, 500, 0, 10, pass_reads)


def test_predict_stream_parity_ragged_and_chunked(ctx):
    ref, off, blob, roff = _world(seed=11, n_lineages=5, per_lineage=9, glen=30_000, s=300, n_reads=200, rlen=2000,
                                  hseed=42, ragged=True)
    assert len(set(np.diff(off).tolist())) > 1
    _check_predict(ctx, ref, off, blob, roff, 16, 300, 42, 5, 32, chunks=3)


def test_predict_stream_k21_top1_and_large_top(ctx):
    ref, off, blob, roff = _world(seed=21, n_lineages=4, per_lineage=40, glen=20_000, s=200, n_reads=120, rlen=1000,
                                  k=21, hseed=42)
    _check_predict(ctx, ref, off, blob, roff, 21, 200, 42, 1, 0)
    _check_predict(ctx, ref, off, blob, roff, 21, 200, 42, 100, 16)


def test_predict_small_squery_truncates(ctx):
    """s_query smaller than the read's k-mer count: the read sketch is its s smallest hashes."""
    ref, off, blob, roff = _world(seed=31, n_lineages=3, per_lineage=5, glen=5_000, s=2000, n_reads=60, rlen=3000)
    _check_predict(ctx, ref, off, blob, roff, 16, 50, 0, 3, 0)


def test_predict_medium_scale(ctx):
    """N=3000 x s=1000 reference, 150 reads of 5 kb, top 10 (oracle finishes in seconds)."""
    ref, off, blob, roff = _world(seed=41, n_lineages=10, per_lineage=300, glen=100_000, s=1000, n_reads=150,
                                  rlen=5000)
    _check_predict(ctx, ref, off, blob, roff, 16, 1000, 0, 10, 0)


def test_predict_more_contenders_than_a_bucket(ctx):
    """4500 references beat the tracked row at once while the candidate budget is forced tiny: buckets overflow and the
    pass is redone with the brute-force ranking."""
    ctx.debug_set("cand_budget", 4096)
    ga = synth.random_genome(20_000, 901)
    gb = synth.random_genome(20_000, 902)
    sk, _, _ = oracle.sketch_groups([ga.tobytes(), gb.tobytes()], [0, 1], 2, 16, 300, 0)
    rows = [sk[0][0]] + [sk[1][0]] * 4500
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ref = np.concatenate(rows)
    blob, roff, _ = synth.sample_reads([gb], 24, 1500, 77, sub=0.0, ins=0.0, dele=0.0)
    try:
        for top in (1, 3):
            _check_predict(ctx, ref, off, blob, roff, 16, 300, 0, top, 8)
    finally:
        ctx.debug_set("cand_budget", 0)


def test_shared_counts_and_rank(ctx):
    ref, off, blob, roff = _world(seed=51, n_lineages=4, per_lineage=6, glen=20_000, s=400, n_reads=5, rlen=500)
    ctx.ref_upload(ref, off)
    got = ctx.shared_counts(ref, off)
    exp = oracle.shared_matrix(ref, off, ref, off)
    assert (got == exp).all()
    assert (np.diag(got) == np.diff(off)).all()     # docs/index.md:148-149: self-shared = s
    col = exp[:, 3].copy()
    gi, gs = ctx.rank_counts(col, 7)
    order = sorted(range(col.size), key=lambda i: (-int(col[i]), i))[:7]
    assert gi.tolist() == order and gs.tolist() == [int(col[i]) for i in order]


def test_readset_mode(ctx):
    from sketchy_b200.api import Sketch, Sketchy
    ref, off, blob, roff = _world(seed=61, n_lineages=4, per_lineage=6, glen=20_000, s=400, n_reads=40, rlen=800)
    sk = Sketchy(0)
    rows = [Sketch(name=f"g{i}", hashes=ref[int(off[i]):int(off[i + 1])]) for i in range(off.size - 1)]
    sk.load_reference(rows)
    reads = [blob[int(roff[i]):int(roff[i + 1])].tobytes() for i in range(roff.size - 1)]
    for limit in (0, 9):
        n, gi, gs, shared = sk.predict_readset_records(reads, 5, limit)
        en, ei, es, eall = oracle.predict_readset(ref, off, reads, 16, 400, 0, 5, limit)
        assert n == en and gi.tolist() == ei.tolist() and gs.tolist() == es.tolist()
        assert shared.tolist() == eall.tolist()
    sk.close()


def test_sharded_predict_and_merge(ctx):
    """Two shards on one GPU (two contexts) + the merge kernel == the unsharded oracle (fake-shard test of the
    multi-GPU path: local top-N with global indices, gather, merge by (sum desc, index asc))."""
    import torch
    from sketchy_b200._lib import Context
    ref, off, blob, roff = _world(seed=71, n_lineages=5, per_lineage=8, glen=20_000, s=300, n_reads=150, rlen=1200)
    N = off.size - 1
    top = 10
    ei, es, _ = oracle.predict_stream(ref, off, (blob, roff), 16, 300, 0, top)
    n = roff.size - 1
    parts_i, parts_s = [], []
    cuts = [0, 7, N]  # uneven shards; the first is smaller than top -> padded entries
    for p in range(2):
        c = Context(0)
        lo, hi = cuts[p], cuts[p + 1]
        c.ref_upload(ref[int(off[lo]):int(off[hi])], off[lo:hi + 1] - off[lo], row_base=lo)
        b = c.batch().add(blob, roff)
        di = torch.zeros((n, top), dtype=torch.int32, device="cuda")
        ds = torch.zeros((n, top), dtype=torch.int64, device="cuda")
        c.predict_stream_device(b, 16, 300, 0, top, di.data_ptr(), ds.data_ptr(), pad=True)
        parts_i.append(di)
        parts_s.append(ds)
        b.close()
        c.close()
    pi = torch.stack(parts_i).contiguous()
    ps = torch.stack(parts_s).contiguous()
    oi = torch.zeros((n, top), dtype=torch.int32, device="cuda")
    os_ = torch.zeros((n, top), dtype=torch.int64, device="cuda")
    ctx.merge_topn_device(pi.data_ptr(), ps.data_ptr(), 2, n, top, oi.data_ptr(), os_.data_ptr())
    torch.cuda.synchronize()
    gi = oi.cpu().numpy().view(np.uint32)
    gs = os_.cpu().numpy().view(np.uint64)
    assert (gi == ei).all() and (gs == es).all()


def test_error_codes(ctx):
    from sketchy_b200._lib import SkbError
    ref = np.array([1, 2, 3, 10, 9, 11], dtype=np.uint64)
    with pytest.raises(SkbError) as e:
        ctx.ref_upload(ref, np.array([0, 3, 6], dtype=np.uint64))
    assert e.value.code == -4
    ctx.ref_upload(np.array([1, 2, 3, 9, 10, 11], dtype=np.uint64), np.array([0, 3, 6], dtype=np.uint64))
    b = ctx.batch().add_records([b"ACGTACGTACGTACGTACGT"])
    with pytest.raises(SkbError) as e:
        ctx.predict_stream(b, 16, 3, 0, 3)       # top > N: the reference panics
    assert e.value.code == -5
    with pytest.raises(SkbError) as e:
        ctx.predict_stream(b, 33, 3, 0, 1)
    assert e.value.code == -7
    gi, gs = ctx.predict_stream(b, 16, 3, 0, 2)
    assert gi.tolist() == [[0, 1]] and gs.tolist() == [[0, 0]]
    b.close()


def _dense_second_opinion(ctx, blob, roff, k, s, reads_to_check, gi, gs, final):
    """Independent GPU path as a second opinion: one sketch per read -> skb_shared_counts (binary-search kernel) ->
    cumulative sums -> numpy ranking on the host."""
    qb = ctx.batch().add(blob, roff)
    qs, _, _ = ctx.sketch(qb, k, s, 0)
    qb.close()
    qoff = np.zeros(len(qs) + 1, dtype=np.uint64)
    qoff[1:] = np.cumsum([h.size for h, _ in qs])
    counts = ctx.shared_counts(np.concatenate([h for h, _ in qs]), qoff)   # [N, R]
    cum = np.cumsum(counts, axis=1)
    assert (cum[:, -1] == final).all()
    idx = np.arange(cum.shape[0])
    for r in reads_to_check:
        order = np.lexsort((idx, -cum[:, r].astype(np.int64)))[:10]
        assert gi[r].tolist() == order.tolist(), r
        assert gs[r].tolist() == cum[order, r].tolist(), r


def _check_predict_large(ctx, ref, off, blob, roff, k, s, top, modes, pass_reads=0):
    """Every read of a large case against the oracle (its N merges per read spread over all host threads: the same
    arithmetic as the single-threaded loop, checked in tests/test_oracle.py), in several ranking modes."""
    import os
    ei, es, esums = oracle.predict_stream(ref, off, (blob, roff), k, s, 0, top, nthreads=os.cpu_count() or 1)
    out = None
    for mode in modes:
        ctx.set_rank_mode(mode)
        ctx.ref_upload(ref, off)
        ctx.set_pass_reads(pass_reads)
        rb = ctx.batch().add(blob, roff)
        gi, gs = ctx.predict_stream(rb, k, s, 0, top)
        final = ctx.sums_download()
        stats = ctx.last_predict_stats()
        rb.close()
        ctx.set_pass_reads(0)
        bad = np.flatnonzero((gi != ei).any(axis=1) | (gs != es).any(axis=1))
        assert bad.size == 0, (mode, bad[:5], gi[bad[:2]], ei[bad[:2]], gs[bad[:2]], es[bad[:2]])
        assert (final == esums).all(), mode
        out = (gi, gs, final, stats)
    ctx.set_rank_mode(0)
    return out


def test_predict_large_scale_against_oracle(ctx):
    """6,000 x s=2,000 reference, 3,000 reads of 5 kb: the streaming path (fused kernel, bounds, candidate lists, and the
    brute-force ranking) against the oracle on every read, plus the independent dense GPU path on a sample."""
    base = [synth.random_genome(200_000, 7000 + l) for l in range(12)]
    b = ctx.batch().add_records([g.tobytes() for g in base])
    sk, _, _ = ctx.sketch(b, 16, 2000, 0)
    b.close()
    rng = np.random.default_rng(5)
    rows = []
    for g in range(6000):
        row = sk[g % 12][0].copy()
        pos = rng.choice(row.size, size=40, replace=False)
        row[pos] = rng.integers(0, int(row.max()), size=40, dtype=np.uint64)
        row = np.unique(row)
        rows.append(row)
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ref = np.concatenate(rows)
    blob, roff, _ = synth.sample_reads(base, 3000, 5000, 99)
    gi, gs, final, _ = _check_predict_large(ctx, ref, off, blob, roff, 16, 2000, 10, modes=(0, 1, 2))
    _dense_second_opinion(ctx, blob, roff, 16, 2000, list(range(0, 64)) + list(range(64, 3000, 97)) + [2999], gi, gs, final)


def test_predict_full_size_passes_against_oracle(ctx):
    """Passes of the maximum size (4096 reads in the default build: u8 counters, every pass-local read id in use, a
    ragged last pass): 20,000 short reads vs 2,000 x s=500, every read against the oracle in all ranking modes, and the
    independent dense GPU path around every multiple of 4096."""
    base = [synth.random_genome(60_000, 7100 + l) for l in range(8)]
    b = ctx.batch().add_records([g.tobytes() for g in base])
    sk, _, _ = ctx.sketch(b, 16, 500, 0)
    b.close()
    rng = np.random.default_rng(6)
    rows = []
    for g in range(2000):
        row = sk[g % 8][0].copy()
        pos = rng.choice(row.size, size=10, replace=False)
        row[pos] = rng.integers(0, int(row.max()), size=10, dtype=np.uint64)
        rows.append(np.unique(row))
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ref = np.concatenate(rows)
    n_reads = 20_000
    blob, roff, _ = synth.sample_reads(base, n_reads, 600, 98)
    gi, gs, final, stats = _check_predict_large(ctx, ref, off, blob, roff, 16, 500, 10, modes=(0, 1, 2))
    assert stats["passes"] <= 5, stats   # 4 x 4096 + 3616
    # the largest pass the kernel holds (8192 reads: 13-bit read ids, four counter buffers)
    _, _, _, stats8 = _check_predict_large(ctx, ref, off, blob, roff, 16, 500, 10, modes=(1, 2), pass_reads=8192)
    assert stats8["passes"] <= 3, stats8
    picks = set(list(range(0, 32)) + list(range(32, n_reads, 131)) + [n_reads - 1])
    for edge in range(4096, n_reads, 4096):
        picks.update(range(edge - 24, edge + 24))
    _dense_second_opinion(ctx, blob, roff, 16, 500, sorted(picks), gi, gs, final)


def test_read_with_more_than_65535_query_hashes(ctx):
    """A contig fed as a read against large sketches keeps > 65535 query hashes under the reference maximum, more than a
    16-bit per-(row, read) counter holds: its query list is cut into pieces that stream as consecutive pass reads (the
    sums are cumulative) and only the ranking after the last piece is reported. Short reads before and after it."""
    base = [synth.random_genome(150_000, 8100 + l) for l in range(3)]
    s = 120_000
    sk, _, _ = oracle.sketch_groups([g.tobytes() for g in base], [0, 1, 2], 3, 16, s, 0)
    rows = [sk[i % 3][0] for i in range(7)]
    assert all(r.size > 70_000 for r in rows)
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ref = np.concatenate(rows)
    short, soff, _ = synth.sample_reads(base, 6, 2000, 3)
    recs = [short[int(soff[i]):int(soff[i + 1])].tobytes() for i in range(6)]
    recs.insert(3, base[1][10_000:140_000].tobytes())      # ~130,000 distinct k-mers, all under the reference maximum
    blob = np.frombuffer(b"".join(recs), dtype=np.uint8)
    roff = np.zeros(len(recs) + 1, dtype=np.uint64)
    roff[1:] = np.cumsum([len(r) for r in recs])
    _check_predict(ctx, ref, off, blob, roff, 16, s, 0, 3, 0, modes=(0, 1))


def test_limits_and_edge_cases(ctx):
    from sketchy_b200._lib import SkbError
    g = synth.random_genome(3000, 1)
    sk, _, _ = oracle.sketch_groups([g.tobytes()] , [0], 1, 16, 50, 0)
    rows = [sk[0][0]] * 200
    off = np.arange(201, dtype=np.uint64) * np.uint64(50)
    ctx.ref_upload(np.concatenate(rows), off)
    ctx.set_pass_reads(0)
    # empty batch and reads without any k-mer
    b = ctx.batch()
    gi, gs = ctx.predict_stream(b, 16, 50, 0, 3)
    assert gi.shape == (0, 3)
    with pytest.raises(SkbError) as e:      # a staged batch is immutable until cleared
        b.add_records([b"ACGT"])
    assert e.value.code == -10
    b.clear()
    b.add_records([b"", b"ACGT", b"NNNNNNNNNNNNNNNNNNNNNNNN"])
    gi, gs = ctx.predict_stream(b, 16, 50, 0, 3)
    assert gi.tolist() == [[0, 1, 2]] * 3 and gs.tolist() == [[0, 0, 0]] * 3
    b.close()
    # top at the supported maximum, equal sums everywhere -> index order
    b = ctx.batch().add_records([g[:500].tobytes()])
    gi, gs = ctx.predict_stream(b, 16, 50, 0, 128)
    assert gi[0].tolist() == list(range(128)) and len(set(gs[0].tolist())) == 1
    with pytest.raises(SkbError) as e:
        ctx.predict_stream(b, 16, 50, 0, 129)
    assert e.value.code == -1
    b.close()
    # unsorted query sketch is rejected (the merge of src/sketchy.rs:419-459 is only a set intersection on sorted input)
    with pytest.raises(SkbError) as e:
        ctx.shared_counts(np.array([5, 3, 9], dtype=np.uint64), np.array([0, 3], dtype=np.uint64))
    assert e.value.code == -4
    # sketch: k = 32 and k = 1 against the oracle
    for k in (32, 1, 2):
        bb = ctx.batch().add_records([g.tobytes()])
        got, gbases, gk = ctx.sketch(bb, k, 40, 7)
        exp, eb, ek = oracle.sketch_groups([g.tobytes()], [0], 1, k, 40, 7)
        assert got[0][0].tolist() == exp[0][0].tolist() and got[0][1].tolist() == exp[0][1].tolist()
        assert int(gk[0]) == int(ek[0])
        bb.close()


def test_membership_prefilter_does_not_change_results(monkeypatch):
    """The reference-membership prefilter only drops query hashes that occur in no reference row: the ranking, the
    sums and the final state must equal a context that was uploaded with the prefilter switched off."""
    import sketchy_b200 as skb
    base = [synth.random_genome(60_000, 400 + l) for l in range(6)]
    sk, _, _ = oracle.sketch_groups([g.tobytes() for g in base], list(range(6)), 6, 16, 800, 0)
    rng = np.random.default_rng(11)
    rows = []
    for g in range(900):
        row = sk[g % 6][0].copy()
        pos = rng.choice(row.size, size=25, replace=False)
        row[pos] = rng.integers(0, int(row.max()), size=25, dtype=np.uint64)
        rows.append(np.unique(row))
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ref = np.concatenate(rows)
    blob, roff, _ = synth.sample_reads(base, 700, 3000, 5)
    out = []
    for off_flag in ("0", "1"):
        monkeypatch.setenv("SKB_NO_PREFILTER", off_flag)
        c = skb.Context(0)
        c.ref_upload(ref, off)
        b = c.batch().add(blob, roff)
        gi, gs = c.predict_stream(b, 16, 800, 0, 10)
        st = c.last_predict_stats()
        out.append((gi.copy(), gs.copy(), c.sums_download().copy(), st))
        b.close()
        c.close()
    assert (out[0][0] == out[1][0]).all() and (out[0][1] == out[1][1]).all() and (out[0][2] == out[1][2]).all()
    assert out[1][3]["member_hashes"] == out[1][3]["query_hashes"]          # prefilter off: every key is kept
    assert 0 < out[0][3]["member_hashes"] < out[0][3]["query_hashes"]       # reads carry errors: novel k-mers are dropped


def test_overflow_inside_a_batch_of_passes_rolls_back(ctx):
    """Steady state (passes enqueued eight at a time, checked on the host afterwards): the stream switches from lineage A
    to lineage B, whose 3,000 identical rows overtake the tracked rows at the same read and overflow every bucket. The
    device marks the failing pass, the passes queued behind it do nothing, the host rolls back to it and redoes it with
    the brute-force ranking. Results must still equal the oracle's."""
    ctx.debug_set("cand_budget", 16 * 512)
    ga = [synth.random_genome(20_000, 910 + i) for i in range(3)]
    gb = synth.random_genome(20_000, 920)
    sk, _, _ = oracle.sketch_groups([g.tobytes() for g in ga] + [gb.tobytes()], [0, 1, 2, 3], 4, 16, 300, 0)
    rows = [sk[i % 3][0] for i in range(30)] + [sk[3][0]] * 3000
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ref = np.concatenate(rows)
    blob_a, roff_a, _ = synth.sample_reads(ga, 200, 1200, 5, sub=0.0, ins=0.0, dele=0.0)
    blob_b, roff_b, _ = synth.sample_reads([gb], 260, 1200, 6, sub=0.0, ins=0.0, dele=0.0)
    blob = np.concatenate([blob_a, blob_b])
    roff = np.concatenate([roff_a, roff_b[1:] + roff_a[-1]])
    try:
        _check_predict(ctx, ref, off, blob, roff, 16, 300, 0, 3, 16, modes=(1,))
        assert ctx.last_predict_stats()["passes"] > (460 + 15) // 16     # a pass streamed twice: once overflowing, once dense
    finally:
        ctx.debug_set("cand_budget", 0)
