"""ctypes binding of libsketchy_b200.so (C ABI in include/sketchy_b200.h).

Fails loudly when the CUDA extension is missing or no B200 is present — there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SKB_LIB") or os.path.join(_HERE, "libsketchy_b200.so")  # SKB_LIB: another build of the same library (kernel A/B runs)

ERRORS = {0: "SKB_OK", -1: "SKB_ERR_INVALID_ARG", -2: "SKB_ERR_CUDA", -3: "SKB_ERR_NO_DEVICE",
          -4: "SKB_ERR_REF_NOT_SORTED", -5: "SKB_ERR_TOP_GT_N", -6: "SKB_ERR_NO_REFERENCE",
          -7: "SKB_ERR_UNSUPPORTED_K", -8: "SKB_ERR_OOM", -9: "SKB_ERR_INTERNAL", -10: "SKB_ERR_STATE",
          -11: "SKB_ERR_COMM"}
KERNEL_IDS = {"hash": 0, "select": 1, "table": 2, "stream": 3, "rank": 4, "merge": 5, "shared": 6, "misc": 7}
MAX_TOP = 128

# every symbol include/sketchy_b200.h declares: (name, restype, argtypes)
_vp, _u32, _u64, _i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
SYMBOLS = [
    ("skb_create", _i, [_i, C.POINTER(_vp)]),
    ("skb_destroy", None, [_vp]),
    ("skb_last_error", C.c_char_p, [_vp]),
    ("skb_version", C.c_char_p, []),
    ("skb_stream", _vp, [_vp]),
    ("skb_synchronize", _i, [_vp]),
    ("skb_batch_create", _i, [_vp, C.POINTER(_vp)]),
    ("skb_batch_destroy", None, [_vp]),
    ("skb_batch_clear", _i, [_vp]),
    ("skb_batch_add", _i, [_vp, _vp, _vp, _vp, _u64, _u32]),
    ("skb_batch_add_records", _i, [_vp, _vp, _vp, _vp, _u64, _u32]),
    ("skb_batch_num_groups", _u32, [_vp]),
    ("skb_batch_num_records", _u64, [_vp]),
    ("skb_batch_num_bases", _u64, [_vp]),
    ("skb_batch_stage", _i, [_vp]),
    ("skb_sketch", _i, [_vp, _vp, _u32, _u32, _u64, _vp, _vp, _vp, _vp, _vp]),
    ("skb_ref_upload", _i, [_vp, _vp, _vp, _u32, _u32]),
    ("skb_ref_upload_device", _i, [_vp, _vp, _vp, _u32, _u32]),
    ("skb_ref_rows", _u32, [_vp]),
    ("skb_predict_stream", _i, [_vp, _vp, _u32, _u32, _u64, _u32, _i, _vp, _vp]),
    ("skb_predict_stream_device", _i, [_vp, _vp, _u32, _u32, _u64, _u32, _i, _vp, _vp]),
    ("skb_sums_reset", _i, [_vp]),
    ("skb_sums_download", _i, [_vp, _vp]),
    ("skb_sums_upload", _i, [_vp, _vp]),
    ("skb_batch_set_base_count", _i, [_vp, _i]),
    ("skb_debug_set", _i, [_vp, C.c_char_p, _u64]),
    ("skb_set_pass_reads", _i, [_vp, _u32]),
    ("skb_pass_reads", _u32, [_vp]),
    ("skb_set_rank_mode", _i, [_vp, _i]),
    ("skb_comm_unique_id", _i, [_vp]),
    ("skb_comm_init", _i, [_vp, _vp, _i, _i]),
    ("skb_comm_destroy", _i, [_vp]),
    ("skb_comm_rank", _i, [_vp]),
    ("skb_comm_world", _i, [_vp]),
    ("skb_comm_allgather_host", _i, [_vp, _vp, _vp, _u64]),
    ("skb_dist_range", None, [_u64, _i, _i, C.POINTER(_u64), C.POINTER(_u64)]),
    ("skb_predict_stream_dist", _i, [_vp, _vp, _u64, _u32, _u32, _u64, _u32, _vp, _vp]),
    ("skb_predict_stream_dist_device", _i, [_vp, _vp, _u64, _u32, _u32, _u64, _u32, _vp, _vp]),
    ("skb_shared_counts", _i, [_vp, _vp, _vp, _u32, _vp]),
    ("skb_rank_counts", _i, [_vp, _vp, _u32, _u32, _vp, _vp]),
    ("skb_merge_topn_device", _i, [_vp, _vp, _vp, _u32, _u64, _u32, _vp, _vp]),
    ("skb_prof_enable", _i, [_vp, _i]),
    ("skb_prof_reset", _i, [_vp]),
    ("skb_prof_get", _i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(_u64)]),
    ("skb_launch_count", _u64, [_vp]),
    ("skb_last_predict_stats", _i, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    ("skb_last_predict_member_hashes", _u64, [_vp]),
    ("skb_batch_packed_len", _u64, [_vp]),
    ("skb_batch_record_start", _i, [_vp, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    ("skb_debug_kmer_hashes", _i, [_vp, _vp, _u32, _u64, _vp, _vp]),
]


def dist_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, begin + count) of `n` items that belong to `rank` (the split every collective call assumes)."""
    b, c = _u64(), _u64()
    load_library().skb_dist_range(n, rank, world, C.byref(b), C.byref(c))
    return int(b.value), int(c.value)


class SkbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """Load libsketchy_b200.so; raises if it is missing and cannot be built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{SO_PATH} is missing: run `python -m sketchy_b200.build`")
        from . import build as _b
        _b.build()
    lib = C.CDLL(SO_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError here = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw (device) address


class Context:
    """One context per GPU/process (skb_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.skb_create(device, C.byref(h))
        if rc != 0:
            raise SkbError(rc, "skb_create failed (a B200 / sm_100 device is required; there is no CPU fallback)")
        self.h = h
        self.device = device

    def check(self, rc: int):
        if rc != 0:
            raise SkbError(rc, self.lib.skb_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.skb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- ingest
    def batch(self) -> "Batch":
        return Batch(self)

    # -- sketch
    def sketch(self, batch: "Batch", k: int, s: int, seed: int = 0):
        G = batch.num_groups
        oh = np.zeros(max(G * s, 1), dtype=np.uint64)
        oc = np.zeros(max(G * s, 1), dtype=np.uint32)
        on = np.zeros(max(G, 1), dtype=np.uint32)
        ob = np.zeros(max(G, 1), dtype=np.uint64)
        ok = np.zeros(max(G, 1), dtype=np.uint64)
        self.check(self.lib.skb_sketch(self.h, batch.h, k, s, seed, _ptr(oh), _ptr(oc), _ptr(on), _ptr(ob), _ptr(ok)))
        sk = [(oh[g * s:g * s + on[g]].copy(), oc[g * s:g * s + on[g]].copy()) for g in range(G)]
        return sk, ob[:G], ok[:G]

    # -- reference
    def ref_upload(self, hashes: np.ndarray, off: np.ndarray, row_base: int = 0):
        hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        self.check(self.lib.skb_ref_upload(self.h, _ptr(hashes) if hashes.size else None, _ptr(off), off.size - 1,
                                           row_base))

    def ref_upload_device(self, d_hashes_ptr: int, off: np.ndarray, row_base: int = 0):
        off = np.ascontiguousarray(off, dtype=np.uint64)
        self.check(self.lib.skb_ref_upload_device(self.h, _ptr(d_hashes_ptr), _ptr(off), off.size - 1, row_base))

    @property
    def ref_rows(self) -> int:
        return int(self.lib.skb_ref_rows(self.h))

    # -- predict
    def predict_stream(self, batch: "Batch", k: int, s_query: int, seed: int, top: int, pad: bool = False, out=None):
        """`out` = (idx uint32 [R, top], sum uint64 [R, top]) caller-owned C-contiguous host arrays to fill in place
        (page-locked ones make the device-to-host copy a plain DMA); fresh arrays otherwise."""
        R = batch.num_groups
        if out is not None:
            oi, os_ = out
            if not (oi.dtype == np.uint32 and os_.dtype == np.uint64 and oi.shape == (R, top) and os_.shape == (R, top)
                    and oi.flags.c_contiguous and os_.flags.c_contiguous):
                raise ValueError("out must be C-contiguous (uint32 [R, top], uint64 [R, top])")
            self.check(self.lib.skb_predict_stream(self.h, batch.h, k, s_query, seed, top, int(pad), _ptr(oi), _ptr(os_)))
            return oi, os_
        oi = np.zeros((max(R, 1), top), dtype=np.uint32)
        os_ = np.zeros((max(R, 1), top), dtype=np.uint64)
        self.check(self.lib.skb_predict_stream(self.h, batch.h, k, s_query, seed, top, int(pad), _ptr(oi), _ptr(os_)))
        return oi[:R], os_[:R]

    def predict_stream_device(self, batch: "Batch", k: int, s_query: int, seed: int, top: int, d_idx: int,
                              d_sum: int, pad: bool = False):
        self.check(self.lib.skb_predict_stream_device(self.h, batch.h, k, s_query, seed, top, int(pad), _ptr(d_idx),
                                                      _ptr(d_sum)))

    # ---- multi-GPU (one process per GPU; see include/sketchy_b200.h) ----
    def comm_unique_id(self) -> np.ndarray:
        """On one rank: the communicator's id (128 bytes) to hand to every rank."""
        uid = np.zeros(128, dtype=np.uint8)
        rc = self.lib.skb_comm_unique_id(_ptr(uid))
        if rc != 0:
            raise SkbError(rc, "skb_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return uid

    def comm_init(self, uid: np.ndarray, rank: int, world: int):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        assert uid.size == 128
        self.check(self.lib.skb_comm_init(self.h, _ptr(uid), rank, world))

    def comm_destroy(self):
        self.check(self.lib.skb_comm_destroy(self.h))

    def predict_stream_dist(self, batch: "Batch", reads_total: int, k: int, s_query: int, seed: int, top: int,
                            out=None, report: bool = True):
        """Collective: `batch` holds this rank's reads (dist_range(reads_total, rank, world)); returns the merged
        ranking of every read (report=False: this rank takes part but does not copy the result to the host)."""
        if not report:
            self.check(self.lib.skb_predict_stream_dist(self.h, batch.h, reads_total, k, s_query, seed, top, None, None))
            return None, None
        if out is None:
            out = (np.zeros((max(reads_total, 1), top), dtype=np.uint32), np.zeros((max(reads_total, 1), top), dtype=np.uint64))
        oi, os_ = out
        self.check(self.lib.skb_predict_stream_dist(self.h, batch.h, reads_total, k, s_query, seed, top, _ptr(oi), _ptr(os_)))
        return oi[:reads_total], os_[:reads_total]

    def predict_stream_dist_device(self, batch: "Batch", reads_total: int, k: int, s_query: int, seed: int, top: int,
                                   d_idx: int, d_sum: int):
        self.check(self.lib.skb_predict_stream_dist_device(self.h, batch.h, reads_total, k, s_query, seed, top,
                                                           _ptr(d_idx), _ptr(d_sum)))

    def sums_reset(self):
        self.check(self.lib.skb_sums_reset(self.h))

    def sums_download(self) -> np.ndarray:
        out = np.zeros(max(self.ref_rows, 1), dtype=np.uint64)
        self.check(self.lib.skb_sums_download(self.h, _ptr(out)))
        return out[:self.ref_rows]

    def sums_upload(self, sums: np.ndarray):
        sums = np.ascontiguousarray(sums, dtype=np.uint64)
        assert sums.size == self.ref_rows
        self.check(self.lib.skb_sums_upload(self.h, _ptr(sums)))

    def set_pass_reads(self, n: int):
        self.check(self.lib.skb_set_pass_reads(self.h, n))

    @property
    def pass_reads(self) -> int:
        """Reads per streaming pass in effect (automatic choice or set_pass_reads)."""
        return int(self.lib.skb_pass_reads(self.h))

    def set_rank_mode(self, mode: int):
        """0 = automatic, 1 = candidate lists wherever possible, 2 = brute-force ranking always (same results)."""
        self.check(self.lib.skb_set_rank_mode(self.h, mode))

    def shared_counts(self, q_hashes: np.ndarray, q_off: np.ndarray) -> np.ndarray:
        q_hashes = np.ascontiguousarray(q_hashes, dtype=np.uint64)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
        Q = q_off.size - 1
        out = np.zeros((max(self.ref_rows, 1), max(Q, 1)), dtype=np.uint64)
        self.check(self.lib.skb_shared_counts(self.h, _ptr(q_hashes) if q_hashes.size else None, _ptr(q_off), Q,
                                              _ptr(out)))
        return out[:self.ref_rows, :Q]

    def rank_counts(self, counts: np.ndarray, top: int):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        oi = np.zeros(max(top, 1), dtype=np.uint32)
        os_ = np.zeros(max(top, 1), dtype=np.uint64)
        self.check(self.lib.skb_rank_counts(self.h, _ptr(counts), counts.size, top, _ptr(oi), _ptr(os_)))
        return oi[:top], os_[:top]

    def merge_topn_device(self, d_idx_parts: int, d_sum_parts: int, n_parts: int, n_reads: int, top: int,
                          d_out_idx: int, d_out_sum: int):
        self.check(self.lib.skb_merge_topn_device(self.h, _ptr(d_idx_parts), _ptr(d_sum_parts), n_parts, n_reads, top,
                                                  _ptr(d_out_idx), _ptr(d_out_sum)))

    # -- measurement
    @property
    def stream_ptr(self) -> int:
        return int(self.lib.skb_stream(self.h) or 0)

    def synchronize(self):
        self.check(self.lib.skb_synchronize(self.h))

    def prof_enable(self, on: bool = True):
        self.check(self.lib.skb_prof_enable(self.h, int(on)))

    def prof_reset(self):
        self.check(self.lib.skb_prof_reset(self.h))

    def prof_get(self, kernel: str) -> tuple[float, int]:
        ms, n = C.c_double(), C.c_uint64()
        self.check(self.lib.skb_prof_get(self.h, KERNEL_IDS[kernel], C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    @property
    def launch_count(self) -> int:
        return int(self.lib.skb_launch_count(self.h))

    def last_predict_stats(self) -> dict:
        a, b, c_, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.check(self.lib.skb_last_predict_stats(self.h, C.byref(a), C.byref(b), C.byref(c_), C.byref(d)))
        return {"ref_bytes_per_pass": int(a.value), "passes": int(b.value), "query_hashes": int(c_.value),
                "candidates": int(d.value), "member_hashes": int(self.lib.skb_last_predict_member_hashes(self.h))}

    # -- debug
    def debug_set(self, key: str, value: int):
        """Internal knobs for tests / experiments (results never depend on them); value 0 restores the default of
        "cand_budget" and "stream_ctas"."""
        self.check(self.lib.skb_debug_set(self.h, key.encode(), int(value)))

    def debug_kmer_hashes(self, batch: "Batch", k: int, seed: int = 0):
        n = batch.packed_len
        oh = np.zeros(max(n, 1), dtype=np.uint64)
        ov = np.zeros(max(n, 1), dtype=np.uint8)
        self.check(self.lib.skb_debug_kmer_hashes(self.h, batch.h, k, seed, _ptr(oh), _ptr(ov)))
        return oh[:n], ov[:n]


class Batch:
    """Pinned 2-bit packed record batch (skb_batch)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(ctx.lib.skb_batch_create(ctx.h, C.byref(h)))
        self.h = h

    def clear(self):
        self.ctx.check(self.ctx.lib.skb_batch_clear(self.h))

    def set_base_count(self, stripped: bool):
        """total_bases of a group: raw sequence bytes (default, what the reference's reader passes on) or the bases
        left after whitespace is removed. Call on an empty batch."""
        self.ctx.check(self.ctx.lib.skb_batch_set_base_count(self.h, 1 if stripped else 0))
        return self

    def add(self, blob: np.ndarray, offsets: np.ndarray, groups=None, nthreads: int = 0):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        g = None if groups is None else np.ascontiguousarray(groups, dtype=np.uint32)
        n = offsets.size - 1
        if g is not None:
            assert g.size == n
        self.ctx.check(self.ctx.lib.skb_batch_add(self.h, _ptr(blob), _ptr(offsets), _ptr(g), n, nthreads))
        return self

    def add_records(self, records, groups=None, nthreads: int = 0):
        arrs = [np.frombuffer(bytes(r), dtype=np.uint8) if not isinstance(r, np.ndarray) else r for r in records]
        if arrs and sum(a.size for a in arrs) >= 65536 * len(arrs):
            # few long records (assemblies): handed over where they lie (skb_batch_add_records), no concatenation
            arrs = [np.ascontiguousarray(a, dtype=np.uint8) for a in arrs]
            ptrs = np.array([a.__array_interface__["data"][0] for a in arrs], dtype=np.uint64)
            lens = np.array([a.size for a in arrs], dtype=np.uint64)
            g = None if groups is None else np.ascontiguousarray(groups, dtype=np.uint32)
            self.ctx.check(self.ctx.lib.skb_batch_add_records(self.h, _ptr(ptrs), _ptr(lens), _ptr(g), len(arrs), nthreads))
            return self
        off = np.zeros(len(arrs) + 1, dtype=np.uint64)
        if arrs:
            off[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
        blob = np.concatenate(arrs) if arrs and off[-1] else np.zeros(1, dtype=np.uint8)
        return self.add(blob, off, groups, nthreads)

    def stage(self):
        self.ctx.check(self.ctx.lib.skb_batch_stage(self.h))
        return self

    @property
    def num_groups(self) -> int:
        return int(self.ctx.lib.skb_batch_num_groups(self.h))

    @property
    def num_records(self) -> int:
        return int(self.ctx.lib.skb_batch_num_records(self.h))

    @property
    def num_bases(self) -> int:
        return int(self.ctx.lib.skb_batch_num_bases(self.h))

    @property
    def packed_len(self) -> int:
        return int(self.ctx.lib.skb_batch_packed_len(self.h))

    def record_start(self, r: int) -> tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        self.ctx.check(self.ctx.lib.skb_batch_record_start(self.h, r, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.skb_batch_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
