"""Host-side tests of the `sketchy` CLI shell: .msh codec (CPU, cross-checked against an independent Python codec,
including a multi-segment file with far / double-far pointers), genotype/info/check behaviour, and (GPU) the full
sketch -> shared -> predict pipeline against the oracle with the reference's row formats."""
import bz2
import gzip
import lzma
import os
import random
import subprocess

import numpy as np
import pytest

import capnp_py
import oracle
from sketchy_b200 import build, synth

CLI = None


def cli():
    global CLI
    if CLI is None:
        build.build()
        CLI = build.CLI
    return CLI


def run(*args, stdin=None, ok=True):
    p = subprocess.run([cli(), *args], input=stdin, capture_output=True, text=True)
    if ok:
        assert p.returncode == 0, p.stderr
    return p


def _file(rng, n=4, s=30):
    f = {"k": 16, "seed": 7, "s": s, "sketches": []}
    for i in range(n):
        hs = sorted(rng.sample(range(1, 2**63), rng.randint(0, s)))
        f["sketches"].append({"name": f"genome_{i}.fa", "comment": "" if i % 2 else "c%d" % i,
                              "seq_length": rng.randint(0, 2**40), "num_valid_kmers": rng.randint(0, 2**33),
                              "hashes": hs, "counts": [rng.randint(1, 9) for _ in hs]})
    return f


def _to_text(f):
    lines = [f"{f['k']} {f['seed']} {f['s']} {len(f['sketches'])}"]
    for s in f["sketches"]:
        lines.append(f"{s['name']}\t{s['comment']}\t{s['seq_length']}\t{s['num_valid_kmers']}")
        lines.append(" ".join(map(str, s["hashes"])))
        lines.append(" ".join(map(str, s["counts"])))
    return "\n".join(lines) + "\n"


def test_msh_writer_decodes_with_independent_reader_and_round_trips(tmp_path):
    f = _file(random.Random(3))
    txt, out = tmp_path / "a.txt", tmp_path / "a.msh"
    txt.write_text(_to_text(f))
    run("msh-from-text", str(txt), str(out))
    dec = capnp_py.decode_msh(out.read_bytes())
    assert dec == f
    assert run("msh-to-text", str(out)).stdout == _to_text(f)


def test_msh_reader_follows_far_and_double_far_pointers(tmp_path):
    f = _file(random.Random(4), n=3)
    out = tmp_path / "multi.msh"
    out.write_bytes(capnp_py.encode_msh_multiseg(f))
    assert capnp_py.decode_msh(out.read_bytes()) == f       # the Python pair agrees with itself ...
    assert run("msh-to-text", str(out)).stdout == _to_text(f)  # ... and the C++ reader with both


def test_info_check_and_errors(tmp_path):
    f = _file(random.Random(5), n=3)
    txt, ref = tmp_path / "r.txt", tmp_path / "r.msh"
    txt.write_text(_to_text(f))
    run("msh-from-text", str(txt), str(ref))
    info = run("info", "-i", str(ref)).stdout.splitlines()
    assert [l.split()[0] for l in info] == [s["name"] for s in f["sketches"]]
    assert [int(l.split()[1]) for l in info] == [s["seq_length"] for s in f["sketches"]]
    p = run("info", "-i", str(ref), "-p").stdout.strip()
    assert p == f"type=mash sketch_size={len(f['sketches'][0]['hashes'])} kmer_size=16 seed=7"
    g = tmp_path / "g.tsv"
    g.write_text("id\tst\tres\n" + "".join(f"{s['name']}\tST{i}\tR\n" for i, s in enumerate(f["sketches"])))
    assert run("check", "-r", str(ref), "-g", str(g)).stdout == "ok\n"
    g.write_text("id\tst\n" + f"{f['sketches'][0]['name']}\tST0\n")
    bad = run("check", "-r", str(ref), "-g", str(g), ok=False)
    assert bad.returncode != 0 and "must have the same length" in bad.stderr
    bad = run("info", "-i", str(tmp_path / "r.txt"), ok=False)
    assert bad.returncode != 0 and "Mash (.msh) or Finch (.fsh) extension" in bad.stderr
    bad = run("predict", "-r", str(ref), "-g", str(g), "-t", "2", "-c", ok=False)
    assert "--top must be an odd number when using --consensus" in bad.stderr


@pytest.mark.gpu
def test_cli_sketch_shared_predict_match_oracle(tmp_path):
    rng = random.Random(9)
    base = [synth.random_genome(30_000, 400 + i) for i in range(3)]
    genomes = [synth.mutate(base[g % 3], 0.002, 500 + g) for g in range(7)]
    paths = []
    for g, seq in enumerate(genomes):
        p = tmp_path / f"g{g}.fa"
        s = seq.tobytes().decode()
        if g == 2:   # multi-record, multi-line, lower-case, N run
            s = s[:9000] + "NNNNN" + s[9000:20000].lower() + s[20000:]
            body = ">c1 first\n" + "\n".join(s[i:i + 70] for i in range(0, 15000, 70)) + "\n>c2\n" + s[15000:] + "\n"
            recs = ["\n".join(s[i:i + 70] for i in range(0, 15000, 70)), s[15000:]]
        else:
            body = f">g{g}\n{s}\n"
            recs = [s]
        p.write_text(body)
        paths.append((str(p), recs))
    k, s_, seed = 16, 200, 42
    ref = tmp_path / "ref.msh"
    run("sketch", "-i", *[p for p, _ in paths], "-o", str(ref), "-s", str(s_), "-k", str(k), "-e", str(seed))
    dec = capnp_py.decode_msh(ref.read_bytes())
    recs = [r.encode() for _, rs in paths for r in rs]
    groups = [g for g, (_, rs) in enumerate(paths) for _ in rs]
    exp, eb, ek = oracle.sketch_groups(recs, groups, len(paths), k, s_, seed)
    assert dec["k"] == k and dec["seed"] == seed
    for g, sk in enumerate(dec["sketches"]):
        assert sk["name"] == f"g{g}.fa"
        assert sk["hashes"] == exp[g][0].tolist() and sk["counts"] == exp[g][1].tolist()
        assert sk["seq_length"] == int(eb[g]) and sk["num_valid_kmers"] == int(ek[g])
    # shared: reference outer, query inner; self-shared = s (docs/index.md:148-149)
    out = run("shared", "-r", str(ref), "-q", str(ref)).stdout.splitlines()
    rows = [np.array(sk["hashes"], dtype=np.uint64) for sk in dec["sketches"]]
    it = iter(out)
    for i in range(7):
        for j in range(7):
            assert next(it) == f"g{i}.fa g{j}.fa {oracle.common_hashes(rows[i], rows[j])}"
    # streaming predict with header, top 3, and consensus; read-set mode
    geno = tmp_path / "ref.tsv"
    geno.write_text("id\tmlst\tmeca\n" + "".join(f"g{g}.fa\tST{g % 3}\t{'R' if g % 2 else 'S'}\n" for g in range(7)))
    blob, roff, _ = synth.sample_reads(base, 40, 900, 31)
    reads = [blob[int(roff[i]):int(roff[i + 1])].tobytes() for i in range(40)]
    fq = tmp_path / "reads.fq"
    fq.write_text("".join(f"@r{i}\n{r.decode()}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads)))
    off = np.zeros(8, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    ei, es, _ = oracle.predict_stream(np.concatenate(rows), off, reads, k, s_, seed, 3)
    gm = {f"g{g}.fa": [f"ST{g % 3}", "R" if g % 2 else "S"] for g in range(7)}
    exp_lines = ["reads\tsketch_id\tshared_hashes\tmlst\tmeca"]
    for r in range(40):
        for t in range(3):
            n = f"g{int(ei[r, t])}.fa"
            exp_lines.append(f"{r + 1}\t{n}\t{int(es[r, t])}\t" + "\t".join(gm[n]))
    got = run("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "3", "-s", "-H").stdout.splitlines()
    assert got == exp_lines
    # compressed reads (src/cli.rs:96 "Fast{a,q}.{gz,xz,bz}"): the content is sniffed, the rows are the same
    for ext, comp in (("gz", gzip.compress), ("bz2", bz2.compress), ("xz", lzma.compress)):
        fz = tmp_path / f"reads.fq.{ext}"
        fz.write_bytes(comp(fq.read_bytes()))
        gz_rows = run("predict", "-i", str(fz), "-r", str(ref), "-g", str(geno), "-t", "3", "-s", "-H").stdout.splitlines()
        assert gz_rows == exp_lines, ext
    lim = run("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "3", "-s", "-l", "5").stdout.splitlines()
    assert lim == exp_lines[1:1 + 15]
    n, ri, rs, _ = oracle.predict_readset(np.concatenate(rows), off, reads, k, s_, seed, 3)
    got = run("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "3").stdout.splitlines()
    assert got == [f"{n}\tg{int(i)}.fa\t{int(s)}\t" + "\t".join(gm[f"g{int(i)}.fa"]) for i, s in zip(ri, rs)]
    cons = run("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "3", "-s", "-c").stdout.splitlines()
    assert len(cons) == 40 and all(l.split("\t")[1:3] == ["-", "-"] for l in cons)
    # stdin file list for sketch (src/sketchy.rs:137-146)
    ref2 = tmp_path / "ref2.msh"
    run("sketch", "-o", str(ref2), "-s", str(s_), "-k", str(k), "-e", str(seed), stdin="\n".join(p for p, _ in paths) + "\n")
    assert capnp_py.decode_msh(ref2.read_bytes()) == dec
