"""ctypes binding of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, ``__graft_entry__.smoke()`` and bench.py's
``cpu_baseline`` / ``--impl reference`` legs. Nothing under ``sketchy_b200/`` may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "liboracle.so")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    """Compile oracle.cpp with g++ (see oracle/Makefile)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "build/liboracle.so"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_murmur3_x64_128.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _u64p]
        L.orc_murmur3_x64_128.restype = None
        L.orc_normalize.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_normalize.restype = C.c_uint64
        L.orc_kmer_hashes.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_kmer_hashes.restype = C.c_uint64
        L.orc_common_hashes.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_double]
        L.orc_common_hashes.restype = C.c_uint64
        L.orc_sketcher_new.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_sketcher_new.restype = C.c_void_p
        L.orc_sketcher_free.argtypes = [C.c_void_p]
        L.orc_sketcher_process.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_sketcher_to_vec.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_sketcher_to_vec.restype = C.c_uint64
        L.orc_sketcher_totals.argtypes = [C.c_void_p, _u64p, _u64p]
        L.orc_sketch_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                        C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_uint32]
        L.orc_sketch_groups.restype = C.c_int
        L.orc_predict_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.orc_predict_stream.restype = C.c_int64
        L.orc_predict_stream_mt.argtypes = L.orc_predict_stream.argtypes + [C.c_uint32]
        L.orc_predict_stream_mt.restype = C.c_int64
        L.orc_predict_readset.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        L.orc_predict_readset.restype = C.c_int64
        L.orc_shared_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                        C.c_void_p]
        L.orc_shared_matrix.restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _bytes(seq) -> np.ndarray:
    if isinstance(seq, np.ndarray):
        return np.ascontiguousarray(seq, dtype=np.uint8)
    if isinstance(seq, str):
        seq = seq.encode()
    return np.frombuffer(bytes(seq), dtype=np.uint8)


def murmur3_x64_128(data, seed: int = 0) -> tuple[int, int]:
    b = _bytes(data)
    out = (C.c_uint64 * 2)()
    lib().orc_murmur3_x64_128(_p(b) if b.size else None, b.size, seed, out)
    return int(out[0]), int(out[1])


def normalize(seq) -> bytes:
    b = _bytes(seq)
    out = np.empty(max(b.size, 1), dtype=np.uint8)
    n = lib().orc_normalize(_p(b) if b.size else None, b.size, _p(out))
    return out[:n].tobytes()


def kmer_hashes(seq, k: int, seed: int = 0) -> np.ndarray:
    b = _bytes(seq)
    cap = max(int(b.size), 1)
    out = np.empty(cap, dtype=np.uint64)
    n = lib().orc_kmer_hashes(_p(b) if b.size else None, b.size, k, seed, _p(out), cap)
    return out[:n].copy()


def common_hashes(ref: np.ndarray, qry: np.ndarray, min_scale: float = 0.0) -> int:
    ref = np.ascontiguousarray(ref, dtype=np.uint64)
    qry = np.ascontiguousarray(qry, dtype=np.uint64)
    return int(lib().orc_common_hashes(_p(ref), ref.size, _p(qry), qry.size, min_scale))


class Sketcher:
    """finch MashSketcher (streaming form): create -> process(record)* -> to_vec() / totals()."""

    def __init__(self, s: int, k: int, seed: int = 0):
        self.s, self.k, self.seed = s, k, seed
        self._h = lib().orc_sketcher_new(s, k, seed)

    def process(self, seq) -> None:
        b = _bytes(seq)
        lib().orc_sketcher_process(self._h, _p(b) if b.size else None, b.size)

    def to_vec(self) -> tuple[np.ndarray, np.ndarray]:
        h = np.empty(max(self.s, 1), dtype=np.uint64)
        c = np.empty(max(self.s, 1), dtype=np.uint32)
        n = lib().orc_sketcher_to_vec(self._h, _p(h), _p(c), self.s)
        return h[:n].copy(), c[:n].copy()

    def totals(self) -> tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        lib().orc_sketcher_totals(self._h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_sketcher_free(self._h)
            self._h = None


def pack_records(records) -> tuple[np.ndarray, np.ndarray]:
    """list of bytes-like -> (blob u8, offsets u64[n+1])."""
    arrs = [_bytes(r) for r in records]
    off = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        off[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
    blob = np.concatenate(arrs) if arrs and off[-1] else np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(blob), off


def sketch_groups(records, groups, ngroups: int, k: int, s: int, seed: int = 0, nthreads: int = 1):
    """reference `_sketch_files`: one sketcher per group (file). Returns list of (hashes, counts), bases, kmers."""
    blob, off = pack_records(records)
    grp = np.ascontiguousarray(groups, dtype=np.uint32)
    oh = np.zeros(max(ngroups * s, 1), dtype=np.uint64)
    oc = np.zeros(max(ngroups * s, 1), dtype=np.uint32)
    on = np.zeros(max(ngroups, 1), dtype=np.uint32)
    ob = np.zeros(max(ngroups, 1), dtype=np.uint64)
    ok = np.zeros(max(ngroups, 1), dtype=np.uint64)
    lib().orc_sketch_groups(_p(blob), _p(off), _p(grp), len(records), ngroups, k, s, seed, _p(oh), _p(oc), _p(on),
                            _p(ob), _p(ok), nthreads)
    out = [(oh[g * s:g * s + on[g]].copy(), oc[g * s:g * s + on[g]].copy()) for g in range(ngroups)]
    return out, ob[:ngroups], ok[:ngroups]


def predict_stream(ref: np.ndarray, ref_off: np.ndarray, reads, k: int, s_query: int, seed: int, top: int,
                   limit: int = 0, sums: np.ndarray | None = None, nthreads: int = 1):
    """reference `_sum_of_shared_hashes`. Returns (idx[nproc, top], sum[nproc, top], sums[N]).
    nthreads > 1 spreads every read's N merges over host threads: same results, but not how the reference runs (its
    predict loop is single-threaded); bench.py reports that number only as a labelled extra."""
    ref = np.ascontiguousarray(ref, dtype=np.uint64)
    ref_off = np.ascontiguousarray(ref_off, dtype=np.uint64)
    N = ref_off.size - 1
    blob, off = reads if isinstance(reads, tuple) else pack_records(reads)
    n = off.size - 1
    sums = np.zeros(N, dtype=np.uint64) if sums is None else np.ascontiguousarray(sums, dtype=np.uint64).copy()
    oi = np.zeros((max(n, 1), top), dtype=np.uint32)
    os_ = np.zeros((max(n, 1), top), dtype=np.uint64)
    if nthreads > 1:
        done = lib().orc_predict_stream_mt(_p(ref), _p(ref_off), N, _p(blob), _p(off), n, k, s_query, seed, top,
                                           limit, _p(sums), _p(oi), _p(os_), nthreads)
    else:
        done = lib().orc_predict_stream(_p(ref), _p(ref_off), N, _p(blob), _p(off), n, k, s_query, seed, top, limit,
                                        _p(sums), _p(oi), _p(os_))
    if done < 0:
        raise ValueError("top > number of reference sketches (reference panics, src/sketchy.rs:391)")
    return oi[:done], os_[:done], sums


def predict_readset(ref: np.ndarray, ref_off: np.ndarray, reads, k: int, s_query: int, seed: int, top: int,
                    limit: int = 0):
    """reference `_shared_hashes`. Returns (n_reads, idx[top], shared[top], shared_all[N])."""
    ref = np.ascontiguousarray(ref, dtype=np.uint64)
    ref_off = np.ascontiguousarray(ref_off, dtype=np.uint64)
    N = ref_off.size - 1
    blob, off = reads if isinstance(reads, tuple) else pack_records(reads)
    oi = np.zeros(top, dtype=np.uint32)
    os_ = np.zeros(top, dtype=np.uint64)
    allc = np.zeros(N, dtype=np.uint64)
    done = lib().orc_predict_readset(_p(ref), _p(ref_off), N, _p(blob), _p(off), off.size - 1, k, s_query, seed,
                                     top, limit, _p(oi), _p(os_), _p(allc))
    if done < 0:
        raise ValueError("top > number of reference sketches")
    return int(done), oi, os_, allc


def shared_matrix(ref, ref_off, qry, qry_off) -> np.ndarray:
    ref = np.ascontiguousarray(ref, dtype=np.uint64)
    ref_off = np.ascontiguousarray(ref_off, dtype=np.uint64)
    qry = np.ascontiguousarray(qry, dtype=np.uint64)
    qry_off = np.ascontiguousarray(qry_off, dtype=np.uint64)
    N, Q = ref_off.size - 1, qry_off.size - 1
    out = np.zeros((N, Q), dtype=np.uint64)
    lib().orc_shared_matrix(_p(ref), _p(ref_off), N, _p(qry), _p(qry_off), Q, _p(out))
    return out
