"""Host FASTA/FASTQ reader (sketchy_b200/host/fastx.hpp): plain, gzip (multi-member), bzip2 (concatenated streams) and
xz (concatenated) inputs, from a path and from stdin, must yield the same records; truncated compressed files fail with
the reference's error text. CPU only: a small C++ harness over the header prints id, raw length and a checksum per
record. The reference CLI accepts Fast{a,q}.{gz,xz,bz} (src/cli.rs:26,96; needletail sniffs the first bytes)."""
import bz2
import gzip
import lzma
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = r'''
#include "ingest.hpp"
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
  try {
    if (argc > 2 && std::string(argv[2]) == "slices") {   // the whole file in memory, records as slices of it
      ingest::Files w;
      ingest::load_files({argv[1]}, 0, 1, 1, w);
      for (size_t i = 0; i < w.n(); ++i) {
        unsigned long long h = 1469598103934665603ull;
        for (size_t j = 0; j < w.len[i]; ++j) { h ^= w.rec[i][j]; h *= 1099511628211ull; }
        std::printf("%zu\t%llu\n", (size_t)w.len[i], h);
      }
      return 0;
    }
    if (argc > 2 && std::string(argv[2]) == "chunks") {   // the streaming predict path: chunks of record slices
      ingest::ChunkReader cr(argv[1], (size_t)atol(argv[3]));
      ingest::Chunk c;
      while (cr.next(c, (size_t)atol(argv[4]), (uint64_t)atol(argv[5]), false)) {
        for (size_t i = 0; i < c.n(); ++i) {
          unsigned long long h = 1469598103934665603ull;
          for (size_t j = 0; j < c.len[i]; ++j) { h ^= c.rec[i][j]; h *= 1099511628211ull; }
          std::printf("%zu\t%llu\n", (size_t)c.len[i], h);
        }
        if (argc > 6) std::printf("chunk %zu\n", c.n());
      }
      return 0;
    }
    fastx::Reader rd(argc > 1 ? argv[1] : "-");
    fastx::Record r;
    while (rd.next(r)) {
      unsigned long long h = 1469598103934665603ull;
      for (unsigned char c : r.seq) { h ^= c; h *= 1099511628211ull; }
      if (argc > 2) std::printf("%s\tidle=%d\n", r.id.c_str(), rd.input_idle() ? 1 : 0);   // live-stream probe
      else std::printf("%s\t%zu\t%llu\n", r.id.c_str(), r.seq.size(), h);
      std::fflush(stdout);
    }
  } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
  return 0;
}
'''


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("fastx")
    src = d / "h.cpp"
    src.write_text(HARNESS)
    exe = d / "h"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "sketchy_b200", "host"), str(src), "-o",
                           str(exe), "-lz", "-ldl"])
    return str(exe)


def _fasta(rng, n):
    out = []
    for i in range(n):
        seq = "".join(rng.choice("ACGTN") for _ in range(rng.randint(0, 400)))
        w = rng.choice([60, 70, 10_000])
        eol = rng.choice(["\n", "\r\n"])
        out.append(f">contig_{i} some description{eol}" + eol.join(seq[j:j + w] for j in range(0, len(seq), w)) + eol)
    return "".join(out).encode()


def _fastq(rng, n):
    out = []
    for i in range(n):
        seq = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 3000)))
        out.append(f"@read_{i} ch=5\n{seq}\n+\n{'I' * len(seq)}\n")
    return "".join(out).encode()


def _fnv(b: bytes) -> int:
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _expected(raw: bytes, kind: str) -> bytes:
    """Independent parse: id = header line minus its marker; FASTA sequence = the raw slice between the header line and
    the next record, interior line breaks kept as they are in the file (CR included: the length needletail's slice has,
    SURVEY.md Appendix F-3), trailing line break dropped; FASTQ sequence = line 2 of the record."""
    lines = [l[:-1] if l.endswith(b"\r") else l for l in raw.split(b"\n")]
    out = []
    if kind == "fasta":
        rid, seq = None, []
        for l in raw.split(b"\n") + [b">"]:
            if l.startswith(b">"):
                if rid is not None:
                    body = b"\n".join(seq).rstrip(b"\r\n")
                    out.append(b"%s\t%d\t%d\n" % (rid, len(body), _fnv(body)))
                rid, seq = l[1:].rstrip(b"\r"), []
            elif rid is not None:
                seq.append(l)
    else:
        rec = [l for l in lines if l != b""]
        for i in range(0, len(rec), 4):
            out.append(b"%s\t%d\t%d\n" % (rec[i][1:], len(rec[i + 1]), _fnv(rec[i + 1])))
    return b"".join(out)


def _run(exe, path=None, data=None):
    p = subprocess.run([exe] + ([path] if path else []), input=data, capture_output=True)
    return p.returncode, p.stdout, p.stderr.decode()


@pytest.mark.parametrize("kind", ["fasta", "fastq"])
def test_compressed_inputs_yield_the_same_records(harness, tmp_path, kind):
    rng = random.Random(11)
    raw = _fasta(rng, 300) if kind == "fasta" else _fastq(rng, 900)   # > 1 MiB of FASTQ: several buffer refills
    half = raw.rfind(b"\n>" if kind == "fasta" else b"\n@read", 0, len(raw) // 2) + 1
    forms = {
        "plain": raw,
        "gz": gzip.compress(raw),
        "gz2": gzip.compress(raw[:half]) + gzip.compress(raw[half:]),        # multi-member (bgzip, cat a.gz b.gz)
        "bz2": bz2.compress(raw),
        "bz2x2": bz2.compress(raw[:half]) + bz2.compress(raw[half:]),         # concatenated streams (pbzip2)
        "xz": lzma.compress(raw, format=lzma.FORMAT_XZ),
        "xzx2": lzma.compress(raw[:half], format=lzma.FORMAT_XZ) + lzma.compress(raw[half:], format=lzma.FORMAT_XZ),
    }
    rc, want, err = _run(harness, data=raw)
    assert rc == 0 and want.count(b"\n") == (300 if kind == "fasta" else 900), err
    assert want == _expected(raw, kind)
    for name, blob in forms.items():
        path = tmp_path / f"x.{name}"            # the name says nothing: the content is sniffed
        path.write_bytes(blob)
        rc, got, err = _run(harness, path=str(path))
        assert rc == 0 and got == want, (name, err)
        p = subprocess.run([harness, str(path), "slices"], capture_output=True)   # the `sketch` path: slices of the file in memory
        assert p.returncode == 0 and p.stdout == b"".join(l.split(b"\t", 1)[1] + b"\n" for l in want.splitlines()), (name, "slices")
        for block, reads, nbytes in ((1 << 20, 1 << 16, 1 << 30), (4099, 7, 1 << 30), (64, 1000, 5000)):   # the streaming `predict` path
            p = subprocess.run([harness, str(path), "chunks", str(block), str(reads), str(nbytes)], capture_output=True)
            assert p.returncode == 0 and p.stdout == b"".join(l.split(b"\t", 1)[1] + b"\n" for l in want.splitlines()), (name, "chunks", block)
        p = subprocess.run([harness, "-", "chunks", "1000", "50", "100000"], input=blob, capture_output=True)
        assert p.returncode == 0 and p.stdout == b"".join(l.split(b"\t", 1)[1] + b"\n" for l in want.splitlines()), (name, "chunks", "stdin")
        rc, got, err = _run(harness, data=blob)   # stdin
        assert rc == 0 and got == want, (name, "stdin", err)


def test_truncated_and_garbage_inputs_fail_like_the_reference(harness, tmp_path):
    rng = random.Random(12)
    raw = _fastq(rng, 200)
    for name, blob in {"gz": gzip.compress(raw), "bz2": bz2.compress(raw), "xz": lzma.compress(raw)}.items():
        path = tmp_path / f"t.{name}"
        path.write_bytes(blob[:len(blob) // 2])
        rc, _, err = _run(harness, path=str(path))
        assert rc == 1 and "failed to open Fastx file or record with Needletail" in err, name
    (tmp_path / "g.txt").write_bytes(b"hello\nworld\n")
    rc, _, err = _run(harness, path=str(tmp_path / "g.txt"))
    assert rc == 1 and "failed to open Fastx file" in err
    rc, _, err = _run(harness, path=str(tmp_path / "missing.fa"))
    assert rc == 1 and "failed to open Fastx file" in err
    (tmp_path / "empty.fa").write_bytes(b"")
    rc, out, _ = _run(harness, path=str(tmp_path / "empty.fa"))
    assert rc == 0 and out == b""


def test_live_stream_reports_idle_between_bursts(harness, tmp_path):
    """A pausing writer on stdin (a sequencer): after the last record of a burst the reader says the input is idle, so
    the streaming predict loop works on what has arrived instead of waiting for a full batch (the reference prints a
    row per read as it arrives, src/sketchy.rs:328-355). A regular file is never idle."""
    import time
    p = subprocess.Popen([harness, "-", "idle"], stdin=subprocess.PIPE, stdout=subprocess.PIPE)
    p.stdin.write(b"@a\nACGT\n+\nIIII\n@b\nACGTA\n+\nIIIII\n")
    p.stdin.flush()
    first = [p.stdout.readline(), p.stdout.readline()]
    time.sleep(0.2)
    p.stdin.write(b"@c\nAC\n+\nII\n")
    p.stdin.close()
    rest = p.stdout.read().splitlines()
    assert p.wait() == 0
    assert first == [b"a\tidle=0\n", b"b\tidle=1\n"]
    assert rest == [b"c\tidle=0"]            # the writer hung up: end of input, not a pause
    f = tmp_path / "r.fq"
    f.write_bytes(b"@a\nACGT\n+\nIIII\n@b\nACGTA\n+\nIIIII\n")
    out = subprocess.run([harness, str(f), "idle"], capture_output=True).stdout.splitlines()
    assert out == [b"a\tidle=0", b"b\tidle=0"]


INGEST_HARNESS = r'''
#include "ingest.hpp"
#include <cstdio>
#include <cstdlib>
// usage: h <threads> <window-budget-bytes> file...   -> two lines per window (copying reader, in-place loader):
//        g0 g1 records bytes fnv(bytes) fnv(off) fnv(grp)
static unsigned long long fnv(const void* p, size_t n) {
  unsigned long long h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= ((const unsigned char*)p)[i]; h *= 1099511628211ull; }
  return h;
}
int main(int argc, char** argv) {
  try {
    const unsigned T = (unsigned)atoi(argv[1]);
    const unsigned long long budget = strtoull(argv[2], nullptr, 10);
    std::vector<std::string> files(argv + 3, argv + argc);
    ingest::Files w;  // reused from window to window, as the CLI does
    for (size_t g0 = 0; g0 < files.size();) {
      const size_t g1 = ingest::window_end(files, g0, budget);
      const ingest::Blob b = ingest::read_files(files, g0, g1, T);
      std::printf("%zu %zu %zu %zu %llu %llu %llu\n", g0, g1, b.n(), b.bytes.size(), fnv(b.bytes.data(), b.bytes.size()),
                  fnv(b.off.data(), b.off.size() * 8), fnv(b.grp.data(), b.grp.size() * 4));
      // the same window through the in-place loader (record slices of the files in memory), brought to the same form
      ingest::load_files(files, g0, g1, T, w);
      std::vector<unsigned char> bytes;
      std::vector<uint64_t> off{0};
      std::vector<uint32_t> grp;
      for (size_t r = 0; r < w.n(); ++r) {
        bytes.insert(bytes.end(), w.rec[r], w.rec[r] + w.len[r]);
        off.push_back(bytes.size());
        grp.push_back(w.grp[r] + (uint32_t)g0);
      }
      std::printf("%zu %zu %zu %zu %llu %llu %llu\n", g0, g1, w.n(), bytes.size(), fnv(bytes.data(), bytes.size()),
                  fnv(off.data(), off.size() * 8), fnv(grp.data(), grp.size() * 4));
      g0 = g1;
    }
  } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
  return 0;
}
'''


def test_parallel_file_reader_equals_the_sequential_one(tmp_path):
    """`sketchy sketch` reads its files on all host threads (the reference uses a rayon pool, src/sketchy.rs:470-472):
    the record batch must not depend on the thread count, windows must respect the budget, and the first unreadable
    file decides the error."""
    src = tmp_path / "i.cpp"
    src.write_text(INGEST_HARNESS)
    exe = str(tmp_path / "i")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "sketchy_b200", "host"), str(src),
                           "-o", exe, "-lz", "-ldl"])
    rng = random.Random(21)
    comps = [lambda b: b, gzip.compress, bz2.compress, lambda b: lzma.compress(b, format=lzma.FORMAT_XZ)]
    files = []
    for g in range(14):
        raw = b"" if g == 5 else _fasta(rng, rng.randint(1, 6))
        p = tmp_path / f"g{g}.fa"
        p.write_bytes(comps[g % 4](raw) if raw else raw)
        files.append(str(p))
    outs = {}
    for t in (1, 3, 16):
        p = subprocess.run([exe, str(t), str(1 << 30)] + files, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        outs[t] = p.stdout
    assert outs[1] == outs[3] == outs[16] and outs[1].split()[:2] == ["0", "14"]
    a, b = outs[1].splitlines()
    assert a == b, "the in-place loader and the copying reader disagree"
    # a tiny budget: one window per file, still every file exactly once and in order
    p = subprocess.run([exe, "4", "1"] + files, capture_output=True, text=True)
    both = p.stdout.splitlines()
    assert both[0::2] == both[1::2], "the in-place loader and the copying reader disagree"
    wins = [l.split() for l in both[0::2]]
    assert [(int(w[0]), int(w[1])) for w in wins] == [(g, g + 1) for g in range(14)]
    assert sum(int(w[2]) for w in wins) == int(outs[1].split()[2])
    # unreadable files: the reference's error text
    (tmp_path / "bad.fa").write_bytes(b"not a sequence file\n")
    p = subprocess.run([exe, "8", str(1 << 30)] + files[:3] + [str(tmp_path / "bad.fa")] + files[3:] + [str(tmp_path / "nope.fa")],
                       capture_output=True, text=True)
    assert p.returncode == 1 and "failed to open Fastx file or record with Needletail" in p.stderr


def test_fastq_quality_length_must_match(harness, tmp_path):
    """needletail rejects a FASTQ record whose quality string is not as long as its sequence; so does the host reader."""
    good = b"@a\nACGT\n+\nIIII\n"
    rc, out, _ = _run(harness, data=good + good.replace(b"@a", b"@b"))
    assert rc == 0 and out.count(b"\n") == 2
    rc, out, err = _run(harness, data=good + b"@b\nACGT\n+\nIII\n")
    assert rc != 0


def test_slices_of_a_file_in_memory_equal_the_streamed_records(harness, tmp_path):
    """fastx::parse_in_place (the `sketch` path) against fastx::Reader on the awkward shapes: no final line break, blank
    lines, a header at the end of the file, CRLF, leading blank lines, empty sequences, records that fail."""
    rng = random.Random(5)
    cases = [b">a\nACGT", b">a\nACGT\n", b">a\r\nAC\r\nGT\r\n>b\r\n\r\n", b"\n\n>a\nAC\n\nGT\n\n\n>b\nTT\n>c", b">a\n>b\n>c\nA\n",
             b">a\n\n\n", b">", b">\n", b"\r\n>a x y\nacgtnN\n", b"@r\nACGT\n+\nIIII", b"@r\nACGT\n+\nIIII\n\n\n@s\nAC\n+r\nII\n",
             b"@r\r\nACGT\r\n+\r\nIIII\r\n", b"@r\nACGT\n+\nIII\n", b"@r\nACGT\n", b"@r\nACGT\nIIII\nIIII\n", b"@r\n\n+\n\n",
             b"x\n>a\nAC\n", b">a\nAC\n@r\nAC\n+\nII\n"]
    for _ in range(200):   # random line soups over a small alphabet: whatever the reader does, the slices must do
        marker = rng.choice([b">", b"@"])
        lines = [rng.choice([b"", b"AC", b"ACGTN", marker + b"id", b"+", b"II", b"IIIII", b"\r", b"AC\r"]) for _ in range(rng.randint(0, 12))]
        cases.append(marker + b"x\n" + b"\n".join(lines) + rng.choice([b"", b"\n", b"\r\n"]))
    for i, raw in enumerate(cases):
        f = tmp_path / f"c{i}"
        f.write_bytes(raw)
        a = subprocess.run([harness, str(f)], capture_output=True)
        for mode in (["slices"], ["chunks", "1", "3", "1000000"], ["chunks", "5", "1", "1000000"], ["chunks", "4096", "1000", "6"]):
            b = subprocess.run([harness, str(f)] + mode, capture_output=True)
            assert (a.returncode == 0) == (b.returncode == 0), (raw, mode, a.stderr, b.stderr)
            if a.returncode == 0:
                assert b.stdout == b"".join(l.split(b"\t", 1)[1] + b"\n" for l in a.stdout.splitlines()), (raw, mode)
            else:
                assert a.stderr == b.stderr, (raw, mode)
                if mode[0] == "chunks":   # the records in front of the bad one still come out, as from the streaming reader
                    assert b.stdout == b"".join(l.split(b"\t", 1)[1] + b"\n" for l in a.stdout.splitlines()), (raw, mode)


CHANNEL_HARNESS = r'''
#include "ingest.hpp"
#include <cstdio>
int main() {
  // order and completeness through a queue of 2 with a slow consumer; close() lets the consumer drain, then stop
  ingest::Channel<int> ch(2);
  long long sum = 0; int last = -1; bool ordered = true;
  std::thread prod([&] { for (int i = 0; i < 10000; ++i) if (!ch.push(i)) return; ch.close(); });
  int v;
  while (ch.pop(v)) { ordered = ordered && v == last + 1; last = v; sum += v; }
  prod.join();
  if (!ordered || last != 9999 || sum != 49995000LL) { std::puts("order"); return 1; }
  if (ch.push(1) || ch.pop(v)) { std::puts("closed"); return 1; }          // closed and empty: both ends say no
  // abort() releases a producer blocked on a full queue and a consumer blocked on an empty one, and drops what is queued
  ingest::Channel<int> full(1), empty(1);
  full.push(7);
  bool pushed = true, popped = true;
  std::thread a([&] { pushed = full.push(8); }), b([&] { int x; popped = empty.pop(x); });
  std::this_thread::sleep_for(std::chrono::milliseconds(50));
  full.abort(); empty.abort();
  a.join(); b.join();
  if (pushed || popped || full.pop(v)) { std::puts("abort"); return 1; }
  std::puts("ok");
  return 0;
}
'''


def test_pipeline_channel(tmp_path):
    """ingest::Channel, the bounded queue between the stages of the CLI's host pipelines: order, back-pressure, close
    (drain, then stop) and abort (everybody lets go at once, queued items are dropped)."""
    (tmp_path / "c.cpp").write_text(CHANNEL_HARNESS)
    exe = str(tmp_path / "c")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "sketchy_b200", "host"), str(tmp_path / "c.cpp"),
                           "-o", exe, "-lz", "-ldl"])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0 and p.stdout.strip() == "ok", p.stdout
