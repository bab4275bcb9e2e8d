"""Host-side tests of the `sketchy` CLI shell: .msh codec (CPU, cross-checked against an independent Python codec,
including a multi-segment file with far / double-far pointers), genotype/info/check behaviour, the host logic of
`sketch` / `predict` against a recording stand-in of the library (CPU: what is handed to the ABI, in which order), and
(GPU) the full sketch -> shared -> predict pipeline against the oracle with the reference's row formats."""
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


def run(*args, stdin=None, ok=True, env=None):
    p = subprocess.run([cli(), *args], input=stdin, capture_output=True, text=True, env=None if env is None else dict(os.environ, **env))
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
    assert capnp_py.decode_msh_header(out.read_bytes()) == {"window": f["k"], "concatenated": True, "noncanonical": False,
                                                            "alphabet": "ACGT"}
    assert run("msh-to-text", str(out)).stdout == _to_text(f)


def test_msh_reader_follows_far_and_double_far_pointers(tmp_path):
    f = _file(random.Random(4), n=3)
    out = tmp_path / "multi.msh"
    out.write_bytes(capnp_py.encode_msh_multiseg(f))
    assert capnp_py.decode_msh(out.read_bytes()) == f       # the Python pair agrees with itself ...
    assert run("msh-to-text", str(out)).stdout == _to_text(f)  # ... and the C++ reader with both


def test_msh_writer_splits_large_files_into_segments(tmp_path):
    """A Cap'n Proto pointer reaches 2^29 words, so a reference of the C3 size (40,000 x 10,000 hashes = 4.8 GB) cannot be
    one segment: past a segment budget the writer puts the hash and count lists into data segments behind far pointers
    (what the builders of finch / Mash do). With the budget lowered the form shows on a small file: the independent
    Python decoder and the C++ reader both get the content back; at the default budget a small file is one segment."""
    import struct
    f = _file(random.Random(13), n=9, s=40)
    txt = tmp_path / "a.txt"
    txt.write_text(_to_text(f))
    one = tmp_path / "one.msh"
    run("msh-from-text", str(txt), str(one))
    assert struct.unpack("<I", one.read_bytes()[:4])[0] == 0
    for budget, min_segments in ((1, 19), (64, 3), (200, 2)):
        out = tmp_path / f"multi_{budget}.msh"
        run("msh-from-text", str(txt), str(out), env={"SKETCHY_B200_MSH_SEGMENT_WORDS": str(budget)})
        b = out.read_bytes()
        assert struct.unpack("<I", b[:4])[0] + 1 >= min_segments
        assert capnp_py.decode_msh(b) == f
        assert run("msh-to-text", str(out)).stdout == _to_text(f)
        assert run("info", "-i", str(out)).stdout == run("info", "-i", str(one)).stdout


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
    # many small chunks through the reader / packer / predict / printer threads (the sums carry over from call to call),
    # with and without -l, and windows of one file in `sketch`: the same rows, the same file
    for extra in ([], ["-l", "17"]):
        small = run("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "3", "-s", "-H", *extra,
                    env={"SKETCHY_B200_CHUNK_READS": "7"}).stdout.splitlines()
        assert small == (exp_lines if not extra else exp_lines[:1 + 17 * 3])
    ref_w = tmp_path / "ref_w.msh"
    run("sketch", "-i", *[p for p, _ in paths], "-o", str(ref_w), "-s", str(s_), "-k", str(k), "-e", str(seed), env={"SKETCHY_B200_WINDOW_BYTES": "1"})
    assert ref_w.read_bytes() == ref.read_bytes()
    n, ri, rs, _ = oracle.predict_readset(np.concatenate(rows), off, reads, k, s_, seed, 3)
    got = run("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "3").stdout.splitlines()
    assert got == [f"{n}\tg{int(i)}.fa\t{int(s)}\t" + "\t".join(gm[f"g{int(i)}.fa"]) for i, s in zip(ri, rs)]
    # consensus rows (src/sketchy.rs:363-389): per genotype column the most frequent value among the read's top rows; the
    # expected values come from the oracle's ranking through the host mirror's rule (ties -> first in rank order; the
    # reference is nondeterministic there, so parity with it is claimed for strict pluralities, which these columns have
    # in most rows and which are counted below)
    from sketchy_b200.api import consensus_value
    cons = run("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "3", "-s", "-c").stdout.splitlines()
    exp_cons, strict = [], 0
    for r in range(40):
        cols = [[gm[f"g{int(ei[r, t])}.fa"][j] for t in range(3)] for j in range(2)]
        strict += sum(max(c.count(v) for v in c) >= 2 for c in cols)
        exp_cons.append(f"{r + 1}\t-\t-\t" + "\t".join(consensus_value(c) for c in cols))
    assert cons == exp_cons
    assert strict >= 60      # most of the 80 (read, column) calls have a strict plurality: those equal the reference's
    # stdin file list for sketch (src/sketchy.rs:137-146)
    ref2 = tmp_path / "ref2.msh"
    run("sketch", "-o", str(ref2), "-s", str(s_), "-k", str(k), "-e", str(seed), stdin="\n".join(p for p, _ in paths) + "\n")
    assert capnp_py.decode_msh(ref2.read_bytes()) == dec


# ---- host logic of the GPU sub-commands against a recording stand-in of the library (no GPU) ---------------------------
def _fnv(b: bytes) -> int:
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.fixture(scope="module")
def mock_cli(tmp_path_factory):
    """The CLI host linked against tests/mock_abi.cpp, which records every ABI call instead of computing."""
    d = tmp_path_factory.mktemp("mockcli")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", os.path.join(root, "tests", "mock_abi.cpp"), "-o",
                           str(d / "libsketchy_b200.so")])
    host = os.path.join(root, "sketchy_b200", "host")
    exe = str(d / "sketchy")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", os.path.join(host, "main.cpp"), os.path.join(host, "msh.cpp"),
                           "-o", exe, "-L" + str(d), "-lsketchy_b200", "-lz", "-ldl", "-Wl,-rpath," + str(d)])

    def run_mock(*args, stdin=None):
        log = d / "calls.log"
        if log.exists():
            log.unlink()
        p = subprocess.run([exe, *args], input=stdin, capture_output=True, env=dict(os.environ, MOCK_ABI_LOG=str(log)))
        return p, (log.read_text().splitlines() if log.exists() else [])
    run_mock.exe = exe
    return run_mock


def test_sketch_host_hands_every_record_to_the_library_in_file_order(mock_cli, tmp_path):
    """`sketch`: files are read on several threads, but the library must see one record per FASTA record, raw slices
    (interior line breaks kept), grouped by file index in file order — for files named on the command line and for a
    file list on stdin (src/sketchy.rs:137-146, 465-494); empty files keep their place in the output."""
    rng = random.Random(5)
    files, expect = [], []
    for g in range(9):
        recs = [] if g == 4 else ["".join(rng.choice("ACGTN") for _ in range(rng.randint(1, 300))) for _ in range(rng.randint(1, 3))]
        body, raw = "", []
        for i, s in enumerate(recs):
            lines = [s[j:j + 60] for j in range(0, len(s), 60)]
            body += f">c{i}\n" + "\n".join(lines) + "\n"
            raw.append("\n".join(lines).encode())
        p = tmp_path / f"f{g}.fa"
        data = body.encode()
        p.write_bytes(gzip.compress(data) if g % 3 == 1 else (bz2.compress(data) if g % 3 == 2 and data else data))
        files.append(str(p))
        expect += [f"  rec group={g} len={len(r)} fnv={_fnv(r)}" for r in raw]
    out = tmp_path / "o.msh"
    for via_stdin in (False, True):
        if via_stdin:
            p, log = mock_cli("sketch", "-o", str(out), "-s", "50", "-k", "21", "-e", "9", stdin=("\n".join(files) + "\n").encode())
        else:
            p, log = mock_cli("sketch", "-i", *files, "-o", str(out), "-s", "50", "-k", "21", "-e", "9")
        assert p.returncode == 0, p.stderr
        assert [l for l in log if l.startswith("  rec")] == expect
        assert [l for l in log if l.startswith("batch_add")] == [f"batch_add n={len(expect)} groups=given"]
        assert log[-1] == f"sketch k=21 s=50 seed=9 groups=9 records={len(expect)}"
        dec = capnp_py.decode_msh(out.read_bytes())
        assert [s["name"] for s in dec["sketches"]] == [f"f{g}.fa" for g in range(9)]
        assert dec["k"] == 21 and dec["seed"] == 9


def test_sketch_host_pipelines_windows_of_files_in_order(mock_cli, tmp_path):
    """`sketch` with many windows (loader, packer and GPU caller on their own threads): every window is cleared, filled
    with its files' records (groups counted from the window's first file) and sketched once, windows in file order;
    an unreadable file in a late window ends the run with the reference's error and no stuck thread."""
    rng = random.Random(8)
    files, per_file = [], []
    for g in range(12):
        recs = ["".join(rng.choice("ACGT") for _ in range(rng.randint(50, 200))) for _ in range(rng.randint(1, 3))]
        p = tmp_path / f"w{g}.fa"
        p.write_text("".join(f">c{i}\n{s}\n" for i, s in enumerate(recs)))
        files.append(str(p))
        per_file.append([f"len={len(r)} fnv={_fnv(r.encode())}" for r in recs])
    out = tmp_path / "o.msh"
    budget = os.path.getsize(files[0]) + os.path.getsize(files[1]) + os.path.getsize(files[2])   # ~3 files per window
    exe = mock_cli.exe
    log = tmp_path / "calls.log"
    p = subprocess.run([exe, "sketch", "-i", *files, "-o", str(out), "-s", "50"], capture_output=True,
                       env=dict(os.environ, MOCK_ABI_LOG=str(log), SKETCHY_B200_WINDOW_BYTES=str(budget)))
    assert p.returncode == 0, p.stderr
    lines = log.read_text().splitlines()
    # the packer and the caller log from two threads: per kind of call the order is fixed
    adds = [l for l in lines if l.startswith("batch_add")]
    sketches = [l for l in lines if l.startswith("sketch ")]
    assert len(adds) == len(sketches) >= 3 and lines.count("batch_clear") == len(adds)
    got, g = [], 0
    recs = [l.split() for l in lines if l.startswith("  rec")]
    i = 0
    for a, sk in zip(adds, sketches):
        n = int(a.split()[1].split("=")[1])
        win = recs[i:i + n]
        i += n
        groups = sorted({int(r[1].split("=")[1]) for r in win})
        assert groups == list(range(len(groups)))                      # groups of a window start at 0
        assert f"groups={len(groups)} records={n}" in sk
        for j in groups:
            assert [" ".join(r[2:]) for r in win if r[1] == f"group={j}"] == per_file[g + j]
        g += len(groups)
    assert g == 12 and i == len(recs)
    dec = capnp_py.decode_msh(out.read_bytes())
    assert [s["name"] for s in dec["sketches"]] == [f"w{g}.fa" for g in range(12)]
    # a bad file in the last window: clean error, no hang
    (tmp_path / "bad.fa").write_bytes(b"no sequence here\n")
    p = subprocess.run([exe, "sketch", "-i", *files, str(tmp_path / "bad.fa"), "-o", str(out), "-s", "50"], capture_output=True, timeout=60,
                       env=dict(os.environ, SKETCHY_B200_WINDOW_BYTES=str(budget)))
    assert p.returncode == 1 and b"failed to open Fastx file or record with Needletail" in p.stderr


def test_predict_host_batches_reads_and_follows_a_live_stream(mock_cli, tmp_path):
    """streaming `predict`: every read is its own group (one sketcher per read, src/sketchy.rs:331), `-l` stops feeding
    reads (:350-353), rows are numbered from 1; when a live stdin pauses, the reads that have arrived are predicted at
    once instead of waiting for a full batch. Read-set mode puts all reads in group 0 (:291)."""
    import time
    ref = tmp_path / "ref.msh"
    rng = random.Random(6)
    f = _file(rng, n=3, s=8)
    for s in f["sketches"]:
        while len(s["hashes"]) < 2:               # a usable reference: s_query comes from sketch #0
            s["hashes"] = sorted(rng.sample(range(1, 2**63), 4)); s["counts"] = [1] * 4
    (tmp_path / "ref.txt").write_text(_to_text(f))
    run("msh-from-text", str(tmp_path / "ref.txt"), str(ref))
    geno = tmp_path / "g.tsv"
    geno.write_text("id\tst\n" + "".join(f"{s['name']}\tST{i}\n" for i, s in enumerate(f["sketches"])))
    reads = ["".join(rng.choice("ACGT") for _ in range(rng.randint(20, 200))) for _ in range(5)]
    fq = tmp_path / "r.fq"
    fq.write_text("".join(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads)))
    names = [s["name"] for s in f["sketches"]]
    s_query = len(f["sketches"][0]["hashes"])
    p, log = mock_cli("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "2", "-s", "-H")
    assert p.returncode == 0, p.stderr
    assert [l for l in log if l.startswith(("batch_add", "predict"))] == [
        "batch_add n=5 groups=null", f"predict_stream k={f['k']} s_query={s_query} seed={f['seed']} top=2 reads_total=5 reads=5"]
    assert [l for l in log if l.startswith("  rec")] == [f"  rec group={i} len={len(r)} fnv={_fnv(r.encode())}" for i, r in enumerate(reads)]
    rows = p.stdout.decode().splitlines()
    assert rows[0] == "reads\tsketch_id\tshared_hashes\tst"
    assert rows[1:] == [f"{r + 1}\t{names[t]}\t7\tST{t}" for r in range(5) for t in range(2)]
    p, log = mock_cli("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "1", "-s", "-l", "3")
    assert [l for l in log if l.startswith("batch_add")] == ["batch_add n=3 groups=null"] and len(p.stdout.splitlines()) == 3
    # consensus row over the top 3 (src/sketchy.rs:363-389); the stand-in ranks rows 0,1,2 -> a three-way tie -> first in rank order
    p, log = mock_cli("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "3", "-s", "-c")
    assert p.stdout.decode().splitlines() == [f"{r + 1}\t-\t-\tST0" for r in range(5)]
    p, log = mock_cli("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "2", "-s", "-c")
    assert p.returncode == 1 and b"--top must be an odd number when using --consensus" in p.stderr
    # read-set mode: one sketcher for all reads, then shared counts and one ranking
    p, log = mock_cli("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "2")
    assert [l for l in log if l.startswith("  rec")] == [f"  rec group=0 len={len(r)} fnv={_fnv(r.encode())}" for r in reads]
    assert [l.split()[0] for l in log if not l.startswith(("  rec", "create", "batch_clear"))] == [
        "ref_upload", "batch_add", "sketch", "shared_counts", "rank_counts"]
    assert p.stdout.decode().splitlines() == [f"5\t{names[t]}\t3\tST{t}" for t in range(2)]
    # live stream: two bursts -> two predict calls, rows of the first burst arrive before the second is written
    log_path = str(tmp_path / "live.log")   # (the fixture's runner waits for exit; this one is driven by hand)
    pr = subprocess.Popen([mock_cli.exe, "predict", "-r", str(ref), "-g", str(geno), "-t", "1", "-s"], stdin=subprocess.PIPE,
                          stdout=subprocess.PIPE, env=dict(os.environ, MOCK_ABI_LOG=log_path))
    burst1 = "".join(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads[:2])).encode()
    pr.stdin.write(burst1); pr.stdin.flush()
    first = [pr.stdout.readline(), pr.stdout.readline()]          # would block forever without the idle check + flush
    time.sleep(0.1)
    pr.stdin.write(f"@r2\n{reads[2]}\n+\n{'I' * len(reads[2])}\n".encode()); pr.stdin.close()
    rest = pr.stdout.read().decode().splitlines()
    assert pr.wait() == 0
    assert [l.decode().split("\t")[0] for l in first] == ["1", "2"] and [l.split("\t")[0] for l in rest] == ["3"]
    calls = [l for l in open(log_path).read().splitlines() if l.startswith("batch_add")]
    assert calls == ["batch_add n=2 groups=null", "batch_add n=1 groups=null"]


def test_streaming_predict_prints_the_reads_in_front_of_a_bad_record(mock_cli, tmp_path):
    """The reference handles record after record (src/sketchy.rs:328-355): a malformed record stops it after the rows of
    every read before it. The chunked pipeline does the same: the chunks in front of the bad record run through, then the
    reader's error ends the run with status 1."""
    rng = random.Random(14)
    f = _file(rng, n=3, s=8)
    for s in f["sketches"]:
        s["hashes"] = sorted(rng.sample(range(1, 2**63), 4)); s["counts"] = [1] * 4
    (tmp_path / "ref.txt").write_text(_to_text(f))
    ref = tmp_path / "ref.msh"
    run("msh-from-text", str(tmp_path / "ref.txt"), str(ref))
    geno = tmp_path / "g.tsv"
    geno.write_text("id\tst\n" + "".join(f"{s['name']}\tST{i}\n" for i, s in enumerate(f["sketches"])))
    good = "".join(f"@r{i}\nACGTACGTACGTACGTACGT\n+\n{'I' * 20}\n" for i in range(11))
    fq = tmp_path / "bad.fq"
    fq.write_text(good + "@r11\nACGTACGT\n+\nIII\n" + "@r12\nACGT\n+\nIIII\n")      # read 12: quality shorter than the sequence
    for chunk_reads in ("65536", "4"):
        p = subprocess.run([mock_cli.exe, "predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "1", "-s"], capture_output=True,
                           timeout=60, env=dict(os.environ, SKETCHY_B200_CHUNK_READS=chunk_reads))
        assert p.returncode == 1 and b"failed to open Fastx file or record with Needletail" in p.stderr
        assert [l.split(b"\t")[0] for l in p.stdout.splitlines()] == [str(i + 1).encode() for i in range(11)]


def test_readset_predict_ranks_long_lists_on_the_host(mock_cli, tmp_path):
    """The device ranks up to SKB_MAX_TOP (128) rows; the reference has no such limit (src/sketchy.rs:310, :391). Read-set
    mode, which ranks once, takes a longer `--top` through a host ranking in the same order (count desc, index asc);
    streaming mode refuses it before anything is printed."""
    rng = random.Random(12)
    f = _file(rng, n=200, s=6)
    for s in f["sketches"]:
        s["hashes"] = sorted(rng.sample(range(1, 2**63), 4)); s["counts"] = [1] * 4
    (tmp_path / "ref.txt").write_text(_to_text(f))
    ref = tmp_path / "ref.msh"
    run("msh-from-text", str(tmp_path / "ref.txt"), str(ref))
    geno = tmp_path / "g.tsv"
    geno.write_text("id\tst\n" + "".join(f"{s['name']}\tST{i % 5}\n" for i, s in enumerate(f["sketches"])))
    fq = tmp_path / "r.fq"
    fq.write_text("@r0\nACGTACGTACGTACGTACGTAA\n+\n" + "I" * 22 + "\n")
    names = [s["name"] for s in f["sketches"]]
    p, log = mock_cli("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "150")
    assert p.returncode == 0, p.stderr
    assert not any(l.startswith("rank_counts") for l in log)
    # the stand-in's shared count of row i is i: the best 150 are rows 199 ... 50
    assert p.stdout.decode().splitlines() == [f"1\t{names[i]}\t{i}\tST{i % 5}" for i in range(199, 49, -1)]
    p, log = mock_cli("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", "150", "-s")
    assert p.returncode == 1 and p.stdout == b"" and b"--top must be between 1 and 128 for streaming predict" in p.stderr


@pytest.mark.gpu
def test_cli_on_the_c1_shapes_matches_oracle(tmp_path):
    """BASELINE.json configs[0] at full size through the C++ host: sketch 100 synthetic 2.8 Mbp assemblies (4 lineages x
    25, SNP rate 0.002, single-line FASTA; k=16, s=1000), then predict 1,000 synthetic 5 kb ONT-like reads with --top 10,
    streaming and read-set mode, every sketch and every row against the oracle."""
    import os
    k, s_, seed, top = 16, 1000, 0, 10
    base = [synth.random_genome(2_800_000, 1000 + l) for l in range(4)]
    genomes = [synth.mutate(base[g % 4], 0.002, 2000 + g) for g in range(100)]
    paths = []
    for g, seq in enumerate(genomes):
        p = tmp_path / f"asm{g:03d}.fasta"
        with open(p, "wb") as f:
            f.write(b">asm%d\n" % g); f.write(seq.tobytes()); f.write(b"\n")
        paths.append(str(p))
    ref = tmp_path / "ref.msh"
    run("sketch", "-i", *paths, "-o", str(ref), "-s", str(s_), "-k", str(k), "-e", str(seed))
    dec = capnp_py.decode_msh(ref.read_bytes())
    exp, eb, ek = oracle.sketch_groups([g.tobytes() for g in genomes], list(range(100)), 100, k, s_, seed, nthreads=os.cpu_count() or 1)
    for g, sk in enumerate(dec["sketches"]):
        assert sk["name"] == f"asm{g:03d}.fasta"
        assert sk["hashes"] == exp[g][0].tolist() and sk["counts"] == exp[g][1].tolist()
        assert sk["seq_length"] == int(eb[g]) and sk["num_valid_kmers"] == int(ek[g])
    geno = tmp_path / "ref.tsv"
    geno.write_text("id\tmlst\tmeca\n" + "".join(f"asm{g:03d}.fasta\tST{g % 4}\t{'R' if g % 2 else 'S'}\n" for g in range(100)))
    blob, roff, _ = synth.sample_reads(base, 1000, 5000, 777)
    reads = [blob[int(roff[i]):int(roff[i + 1])].tobytes() for i in range(1000)]
    fq = tmp_path / "reads.fq"
    fq.write_text("".join(f"@r{i}\n{r.decode()}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads)))
    rows = [np.array(sk["hashes"], dtype=np.uint64) for sk in dec["sketches"]]
    off = np.zeros(101, dtype=np.uint64)
    off[1:] = np.cumsum([r.size for r in rows])
    flat = np.concatenate(rows)
    ei, es, _ = oracle.predict_stream(flat, off, reads, k, s_, seed, top)
    gm = {f"asm{g:03d}.fasta": [f"ST{g % 4}", "R" if g % 2 else "S"] for g in range(100)}
    exp_lines = []
    for r in range(1000):
        for t in range(top):
            n = f"asm{int(ei[r, t]):03d}.fasta"
            exp_lines.append(f"{r + 1}\t{n}\t{int(es[r, t])}\t" + "\t".join(gm[n]))
    got = run("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", str(top), "-s").stdout.splitlines()
    assert got == exp_lines
    n, ri, rs, _ = oracle.predict_readset(flat, off, reads, k, s_, seed, top)
    got = run("predict", "-i", str(fq), "-r", str(ref), "-g", str(geno), "-t", str(top)).stdout.splitlines()
    assert got == [f"{n}\tasm{int(i):03d}.fasta\t{int(s)}\t" + "\t".join(gm[f"asm{int(i):03d}.fasta"]) for i, s in zip(ri, rs)]
