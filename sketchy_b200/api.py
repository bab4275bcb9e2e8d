"""Host-side mirror of the reference's `Sketchy` interface (reference src/sketchy.rs:57-601) over the C ABI.

Same method names, argument meaning and error behaviour as the reference so parity tests read like tests of the
reference itself; the compute goes through libsketchy_b200.so (no CPU path here). File formats (.msh, FASTA/FASTQ,
genotype TSV) are handled by the C++ host (``sketchy_b200/host``: the `sketchy` CLI).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from ._lib import Batch, Context, SkbError


class SketchyError(Exception):
    """reference src/sketchy.rs:17-39 (messages kept verbatim)."""

    INVALID_SIZE = "reference sketch and genotype table must have the same length"
    INVALID_EXTENSION = "reference sketch file must have Mash (.msh) or Finch (.fsh) extension"
    INVALID_CONSENSUS_TOP = "--top must be an odd number when using --consensus"
    INVALID_CONSENSUS_GENOTYPE = "consensus genotype could not be computed"


@dataclass
class PredictConfig:
    """reference src/sketchy.rs:43-50."""
    top: int = 1
    limit: int = 0
    stream: bool = False
    consensus: bool = False
    header: bool = False


@dataclass
class Sketch:
    """finch::serialization::Sketch as used by the reference (src/sketchy.rs:483-491)."""
    name: str
    hashes: np.ndarray
    counts: np.ndarray | None = None
    seq_length: int = 0
    num_valid_kmers: int = 0
    comment: str = ""
    kmer_length: int = 16
    hash_seed: int = 0


def records_blob(records) -> tuple[np.ndarray, np.ndarray]:
    arrs = [np.frombuffer(bytes(r), dtype=np.uint8) if not isinstance(r, np.ndarray) else r for r in records]
    off = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        off[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
    blob = np.concatenate(arrs) if arrs and off[-1] else np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(blob), off


def flatten_sketches(sketches) -> tuple[np.ndarray, np.ndarray]:
    rows = [np.asarray(s.hashes if isinstance(s, Sketch) else s, dtype=np.uint64) for s in sketches]
    off = np.zeros(len(rows) + 1, dtype=np.uint64)
    if rows:
        off[1:] = np.cumsum([r.size for r in rows], dtype=np.uint64)
    flat = np.concatenate(rows) if rows and off[-1] else np.zeros(0, dtype=np.uint64)
    return flat, off


def consensus_value(values: list[str]) -> str:
    """reference src/sketchy.rs:404-413: most frequent value. The reference iterates a HashMap with `max_by`, so
    ties are nondeterministic there; here ties go to the value seen first in rank order (documented in DESIGN.md)."""
    if not values:
        raise SketchyError(SketchyError.INVALID_CONSENSUS_GENOTYPE)
    best, best_n = None, 0
    counts: dict[str, int] = {}
    for v in values:
        counts[v] = counts.get(v, 0) + 1
    for v in values:
        if counts[v] > best_n:
            best, best_n = v, counts[v]
    return best


class Sketchy:
    """GPU-backed counterpart of the reference's stateless `Sketchy` struct. One instance owns one context
    (one GPU); the reference sketches uploaded by `load_reference` stay resident in HBM."""

    def __init__(self, device: int = 0, rank: int = 0, world_size: int = 1):
        self.ctx = Context(device)
        self.rank, self.world_size = rank, world_size
        self.names: list[str] = []
        self.s_query = 0
        self.k = 16
        self.seed = 0
        self.n_total = 0
        self.row_base = 0

    # ---- sketch (src/sketchy.rs:128-167, 465-494) ---------------------------------------------------------
    def sketch_files_from_records(self, files: list[tuple[str, list[bytes]]], sketch_size: int, kmer_size: int,
                                  seed: int, nthreads: int = 0) -> list[Sketch]:
        """One sketcher per file; `files` is [(basename, [record sequences])] in input order."""
        b = self.ctx.batch()
        recs, groups = [], []
        for g, (_, rs) in enumerate(files):
            recs.extend(rs)
            groups.extend([g] * len(rs))
        if recs:
            blob, off = records_blob(recs)
            b.add(blob, off, np.asarray(groups, dtype=np.uint32), nthreads)
        out: list[Sketch] = []
        G = len(files)
        res = [(np.zeros(0, np.uint64), np.zeros(0, np.uint32))] * G
        bases = np.zeros(G, np.uint64)
        kmers = np.zeros(G, np.uint64)
        if b.num_groups:
            sk, ob, ok = self.ctx.sketch(b, kmer_size, sketch_size, seed)
            for g in range(len(sk)):
                res[g] = sk[g]
                bases[g], kmers[g] = ob[g], ok[g]
        b.close()
        for g, (name, _) in enumerate(files):
            out.append(Sketch(name=os.path.basename(name), hashes=res[g][0], counts=res[g][1],
                              seq_length=int(bases[g]), num_valid_kmers=int(kmers[g]), kmer_length=kmer_size,
                              hash_seed=seed))
        return out

    # ---- reference residency (src/sketchy.rs:81-82, 497-536) ----------------------------------------------
    def load_reference(self, sketches: list[Sketch]):
        """Shard rows by contiguous index range over world_size ranks; derive the per-read sketcher parameters
        from sketch #0 exactly as the reference does (kmers_to_sketch = len(hashes[0]), k and seed of the file)."""
        if not sketches:
            raise IndexError("reference sketch file holds no sketches")  # the reference panics on sketches[0]
        self.names = [s.name for s in sketches]
        self.n_total = len(sketches)
        self.s_query = int(len(sketches[0].hashes))
        self.k = sketches[0].kmer_length
        self.seed = sketches[0].hash_seed
        lo = self.n_total * self.rank // self.world_size
        hi = self.n_total * (self.rank + 1) // self.world_size
        self.row_base = lo
        flat, off = flatten_sketches(sketches[lo:hi])
        self.ctx.ref_upload(flat, off, lo)

    # ---- streaming predict (src/sketchy.rs:317-356) -------------------------------------------------------
    def predict_stream_records(self, reads, top: int, limit: int = 0):
        """Returns (idx[R, top], sum[R, top]) for the reads processed (limit as at :350-353)."""
        if top > self.n_total:
            raise SkbError(-5, "top exceeds the number of reference sketches (the reference panics, src/sketchy.rs:391)")
        reads = list(reads)
        if limit > 0:
            reads = reads[:limit]
        b = self.ctx.batch()
        if reads:
            blob, off = records_blob(reads)
            b.add(blob, off, None)
        idx, sm = self.ctx.predict_stream(b, self.k, max(self.s_query, 1), self.seed, top, pad=self.world_size > 1)
        b.close()
        return idx, sm

    # ---- read-set predict (src/sketchy.rs:281-315) --------------------------------------------------------
    def predict_readset_records(self, reads, top: int, limit: int = 0):
        """ONE sketch over all reads (limit check at :297: `read == limit` after the increment)."""
        if top > self.n_total:
            raise SkbError(-5, "top exceeds the number of reference sketches")
        reads = list(reads)
        n_used = len(reads)
        if limit > 0 and limit < len(reads):
            n_used = limit
        b = self.ctx.batch()
        if n_used:
            blob, off = records_blob(reads[:n_used])
            b.add(blob, off, np.zeros(n_used, dtype=np.uint32))
            sk, _, _ = self.ctx.sketch(b, self.k, max(self.s_query, 1), self.seed)
            q = sk[0][0]
        else:
            q = np.zeros(0, np.uint64)
        b.close()
        shared = self.ctx.shared_counts(q, np.array([0, q.size], dtype=np.uint64))[:, 0]
        idx, sm = self.ctx.rank_counts(shared, top)
        return n_used, idx + np.uint32(self.row_base), sm, shared

    # ---- shared (src/sketchy.rs:238-279) ------------------------------------------------------------------
    def shared_matrix(self, query_sketches: list[Sketch]) -> np.ndarray:
        flat, off = flatten_sketches(query_sketches)
        return self.ctx.shared_counts(flat, off)

    # ---- output rows (src/sketchy.rs:358-402) -------------------------------------------------------------
    def format_rows(self, read: int, idx_row, sum_row, genotypes: dict[str, list[str]], consensus: bool) -> list[str]:
        if consensus:
            cols = None
            for i in idx_row:
                g = genotypes[self.names[int(i)]]
                if cols is None:
                    cols = [[] for _ in g]
                for j, v in enumerate(g):
                    cols[j].append(v)
            return [f"{read}\t-\t-\t" + "\t".join(consensus_value(c) for c in (cols or []))]
        rows = []
        for i, s in zip(idx_row, sum_row):
            name = self.names[int(i)]
            rows.append(f"{read}\t{name}\t{int(s)}\t" + "\t".join(genotypes[name]))
        return rows

    def close(self):
        self.ctx.close()
