"""ctypes binding of libhica.so (the C ABI declared in include/hica.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device
is present, construction of a :class:`Handle` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhica.so")

T_NAMES = ("upload", "distance", "linkage", "leaforder", "rank", "breakpoints", "filter", "p2_score", "rank_presort")

# every symbol include/hica.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = (
    "hica_version", "hica_create", "hica_destroy", "hica_last_error", "hica_launch_count", "hica_get_timers",
    "hica_linkage_stats", "hica_debug_stats",
    "hica_p1_alloc", "hica_p1_upload_rows", "hica_p1_mirror_lower", "hica_p1_device_buffer", "hica_p1_rank_rows", "hica_p1_rank_commit", "hica_p1_rank_presort",
    "hica_p1_upload", "hica_p1_upload_symmetric", "hica_p1_upload_device", "hica_p1_set_row_sums", "hica_p1_distance", "hica_p1_get_distance",
    "hica_p1_get_pairwise_row_sums", "hica_p1_get_row_sums", "hica_p1_upload_distance", "hica_p1_linkage",
    "hica_p1_set_order", "hica_p1_get_reordered", "hica_p1_rank", "hica_p1_upload_similarity", "hica_p1_get_rank", "hica_p1_set_rank",
    "hica_p1_window_sig", "hica_p1_breakpoints", "hica_p1_filter", "hica_p1_tie_report", "hica_hypergeom_less",
    "hica_p2_upload", "hica_p2_use_p1_matrix", "hica_p2_select", "hica_p2_set_segments", "hica_p2_total",
    "hica_p2_score_candidates", "hica_p2_score_window", "hica_p2_score_orders", "hica_p2_order_groups",
    "hica_s32_alloc", "hica_s32_buffer", "hica_s32_row_sums", "hica_s32_build_w", "hica_s32_ipc_export", "hica_s32_ipc_open",
    "hica_s32_ipc_close", "hica_s32_linkage",
    "hica_io_read_matrix", "hica_io_last_error", "hica_io_log_transform", "hica_io_read_loci", "hica_io_loci_fetch",
    "hica_p1_order_stats", "hica_order_stats_host",
)

_lib = None


class HicaError(RuntimeError):
    pass


def load_library():
    """dlopen libhica.so; raises HicaError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HicaError(
            "libhica.so not found at %s - build it with `python -m hic_genome_assembler_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    L.hica_version.restype = C.c_int
    L.hica_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.hica_destroy.argtypes = [vp]
    L.hica_last_error.argtypes = [vp]
    L.hica_last_error.restype = C.c_char_p
    L.hica_launch_count.argtypes = [vp]
    L.hica_launch_count.restype = i64
    L.hica_get_timers.argtypes = [vp, vp]
    L.hica_p1_upload.argtypes = [vp, vp, i64]
    L.hica_p1_upload_device.argtypes = [vp, vp, i64]
    L.hica_p1_upload_symmetric.argtypes = [vp, vp, i64]
    L.hica_p1_alloc.argtypes = [vp, i64]
    L.hica_p1_upload_rows.argtypes = [vp, vp, i64, i64, C.c_int]
    L.hica_p1_mirror_lower.argtypes = [vp]
    L.hica_p1_device_buffer.argtypes = [vp, C.c_int, vp, vp, vp]
    L.hica_p1_rank_rows.argtypes = [vp, i64, i64]
    L.hica_p1_rank_commit.argtypes = [vp]
    L.hica_p1_rank_presort.argtypes = [vp]
    L.hica_p1_set_row_sums.argtypes = [vp, vp]
    L.hica_p1_distance.argtypes = [vp]
    L.hica_p1_get_distance.argtypes = [vp, vp]
    L.hica_p1_get_pairwise_row_sums.argtypes = [vp, vp]
    L.hica_p1_get_row_sums.argtypes = [vp, vp]
    L.hica_p1_upload_distance.argtypes = [vp, vp, i64]
    L.hica_p1_linkage.argtypes = [vp, vp, vp]
    L.hica_p1_set_order.argtypes = [vp, vp]
    L.hica_p1_get_reordered.argtypes = [vp, C.c_int, vp]
    L.hica_p1_rank.argtypes = [vp]
    L.hica_p1_upload_similarity.argtypes = [vp, vp, i64]
    L.hica_p1_get_rank.argtypes = [vp, vp]
    L.hica_p1_set_rank.argtypes = [vp, vp, i64]
    L.hica_p1_window_sig.argtypes = [vp, i64, i64, i64, dbl, C.c_int, vp, vp, vp]
    L.hica_p1_breakpoints.argtypes = [vp, i64, dbl, vp, i64, vp, vp]
    L.hica_p1_filter.argtypes = [vp, vp, i64, dbl, vp, vp, vp]
    L.hica_p1_tie_report.argtypes = [vp, vp]
    L.hica_hypergeom_less.argtypes = [vp, vp, vp, vp, vp, i64, dbl, vp]
    L.hica_p2_upload.argtypes = [vp, vp, i64]
    L.hica_p2_use_p1_matrix.argtypes = [vp]
    L.hica_p2_select.argtypes = [vp, vp, i32]
    L.hica_p2_set_segments.argtypes = [vp, vp, vp, i32]
    L.hica_p2_total.argtypes = [vp, vp, i32, vp]
    L.hica_p2_score_candidates.argtypes = [vp, vp, i64, i32, dbl, dbl, vp, vp, vp]
    L.hica_p2_score_window.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp, i32, i64, i64, dbl, dbl, vp, vp, vp]
    L.hica_p2_score_orders.argtypes = [vp, vp, i64, i32, dbl, dbl, vp, vp, vp]
    L.hica_p2_order_groups.argtypes = [vp, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, i32, vp, i32, vp, vp, vp, vp]
    L.hica_s32_alloc.argtypes = [vp, i64, i64, i64, C.c_int]
    L.hica_s32_buffer.argtypes = [vp, C.c_int, vp, vp]
    L.hica_s32_row_sums.argtypes = [vp]
    L.hica_s32_build_w.argtypes = [vp, vp, i64, i64]
    L.hica_s32_ipc_export.argtypes = [vp, C.c_int, vp]
    L.hica_s32_ipc_open.argtypes = [vp, vp, vp]
    L.hica_s32_ipc_close.argtypes = [vp, vp]
    L.hica_s32_linkage.argtypes = [vp, vp, vp]
    L.hica_io_read_matrix.argtypes = [C.c_char_p, vp, i64, i64, vp, C.c_int, vp, vp]
    L.hica_io_log_transform.argtypes = [vp, vp, i64, dbl, dbl, C.c_int, C.c_int]
    L.hica_io_read_loci.argtypes = [C.c_char_p, C.c_char_p, vp, i64, vp, vp]
    L.hica_io_loci_fetch.argtypes = [vp, vp, vp, vp, vp, vp]
    L.hica_p1_order_stats.argtypes = [vp, C.c_int, C.c_int32, vp, vp, vp]
    L.hica_order_stats_host.argtypes = [vp, vp, i64, C.c_int32, vp, vp, vp]
    L.hica_io_last_error.argtypes = []
    L.hica_io_last_error.restype = C.c_char_p
    _lib = L
    return L


def io_read_matrix(path: str, index_of_id: np.ndarray, n: int, out: Optional[np.ndarray] = None, threads: int = 0):
    """hica_io_read_matrix: HiC-Pro triplet text -> dense symmetric n x n fp64 (host side, no GPU needed).
    ``index_of_id[bin id]`` = matrix row or -1; ``out`` may be a preallocated (pinned) n x n float64 array.
    Returns (matrix, lines parsed, lines applied)."""
    L = load_library()
    idx = np.ascontiguousarray(index_of_id, dtype=np.int32)
    if out is None:
        out = np.empty((n, n), dtype=np.float64)
    if out.shape != (n, n) or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array of shape (%d, %d)" % (n, n))
    lines, edges = C.c_int64(0), C.c_int64(0)
    rc = L.hica_io_read_matrix(os.fsencode(path), _ptr(idx), len(idx) - 1, n, _ptr(out), int(threads),
                               C.byref(lines), C.byref(edges))
    if rc:
        raise HicaError(L.hica_io_last_error().decode())
    return out, int(lines.value), int(edges.value)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class Handle:
    """One GPU context of libhica (one per device, not re-entrant)."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.hica_create(int(device), C.byref(h))
        if rc != 0:
            raise HicaError("hica_create failed: %s" % self._L.hica_last_error(None).decode())
        self._h = h
        self.device = device
        self.n = 0

    # -- plumbing ---------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.hica_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise HicaError(self._L.hica_last_error(self._h).decode() or ("libhica error %d" % rc))

    @property
    def launches(self) -> int:
        return int(self._L.hica_launch_count(self._h))

    def timers(self) -> dict:
        ms = np.zeros(len(T_NAMES))
        self._check(self._L.hica_get_timers(self._h, _ptr(ms)))
        return dict(zip(T_NAMES, ms.tolist()))

    def debug_stats(self) -> list:
        out = np.zeros(16)
        self._L.hica_debug_stats.argtypes = [C.c_void_p, C.c_void_p]
        self._check(self._L.hica_debug_stats(self._h, _ptr(out)))
        return [round(float(v), 3) for v in out[:6]]

    def linkage_stats(self) -> dict:
        out = np.zeros(8, dtype=np.int64)
        self._L.hica_linkage_stats.argtypes = [C.c_void_p, C.c_void_p]
        self._check(self._L.hica_linkage_stats(self._h, _ptr(out)))
        return {"searches": int(out[0]), "search_cycles": int(out[1]), "merge_cycles": int(out[2]),
                "load_cycles": int(out[3]), "reduce_cycles": int(out[4]), "exchange_cycles": int(out[5]),
                "cached_searches": int(out[6]), "cache_slow_paths": int(out[7])}

    # -- Part 1 -----------------------------------------------------------
    def p1_upload(self, A: np.ndarray):
        A = _c(A, np.float64)
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError("square matrix expected")
        self.n = A.shape[0]
        self._check(self._L.hica_p1_upload(self._h, _ptr(A), self.n))

    def p1_upload_symmetric(self, A: np.ndarray):
        """Upload a symmetric matrix: only its upper triangle crosses PCIe (the lower one is never read)."""
        A = _c(A, np.float64)
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError("square matrix expected")
        self.n = A.shape[0]
        self._check(self._L.hica_p1_upload_symmetric(self._h, _ptr(A), self.n))

    # row-block sharding: see hica.h and dist.run_part1_sharded
    def p1_alloc(self, n: int):
        self.n = int(n)
        self._check(self._L.hica_p1_alloc(self._h, self.n))

    def p1_upload_rows(self, rows: np.ndarray, r0: int, r1: int, symmetric: bool = False):
        """rows = rows [r0, r1) of the n x n matrix (a C-contiguous (r1-r0) x n view of the host matrix)."""
        rows = _c(rows, np.float64)
        if rows.shape != (r1 - r0, self.n):
            raise ValueError("row block of shape %s expected, got %s" % ((r1 - r0, self.n), rows.shape))
        self._check(self._L.hica_p1_upload_rows(self._h, _ptr(rows) if r1 > r0 else None, int(r0), int(r1),
                                                1 if symmetric else 0))

    def p1_mirror_lower(self):
        self._check(self._L.hica_p1_mirror_lower(self._h))

    def p1_device_buffer(self, which: int) -> Tuple[int, int, int]:
        """(device pointer, bytes, element bytes) of the matrix (which = 0) or the rank matrix (which = 1)."""
        ptr, nbytes, eb = C.c_void_p(0), C.c_int64(0), C.c_int32(0)
        self._check(self._L.hica_p1_device_buffer(self._h, int(which), C.byref(ptr), C.byref(nbytes), C.byref(eb)))
        return int(ptr.value or 0), int(nbytes.value), int(eb.value)

    def p1_rank_rows(self, r0: int, r1: int):
        self._check(self._L.hica_p1_rank_rows(self._h, int(r0), int(r1)))

    def p1_rank_presort(self):
        """Start the value sort of the rank step asynchronously (see hica.h); returns at once."""
        self._check(self._L.hica_p1_rank_presort(self._h))

    def p1_rank_commit(self):
        self._check(self._L.hica_p1_rank_commit(self._h))

    def p1_upload_device(self, dev_ptr: int, n: int):
        self.n = int(n)
        self._check(self._L.hica_p1_upload_device(self._h, C.c_void_p(dev_ptr), self.n))

    def p1_upload_distance(self, D: np.ndarray):
        D = _c(D, np.float64)
        self.n = D.shape[0]
        self._check(self._L.hica_p1_upload_distance(self._h, _ptr(D), self.n))

    def p1_upload_similarity(self, S: np.ndarray):
        S = _c(S, np.float64)
        self.n = S.shape[0]
        self._check(self._L.hica_p1_upload_similarity(self._h, _ptr(S), self.n))

    def p1_set_row_sums(self, rs: Optional[np.ndarray] = None):
        rs = None if rs is None else _c(rs, np.float64)
        self._check(self._L.hica_p1_set_row_sums(self._h, _ptr(rs)))

    def p1_distance(self):
        self._check(self._L.hica_p1_distance(self._h))

    def p1_get_distance(self) -> np.ndarray:
        D = np.empty((self.n, self.n))
        self._check(self._L.hica_p1_get_distance(self._h, _ptr(D)))
        return D

    def p1_get_pairwise_row_sums(self) -> np.ndarray:
        s = np.empty(self.n)
        self._check(self._L.hica_p1_get_pairwise_row_sums(self._h, _ptr(s)))
        return s

    def p1_get_row_sums(self) -> np.ndarray:
        s = np.empty(self.n)
        self._check(self._L.hica_p1_get_row_sums(self._h, _ptr(s)))
        return s

    def p1_linkage(self, want_Z: bool = True) -> Tuple[Optional[np.ndarray], np.ndarray]:
        Z = np.empty((self.n - 1, 4)) if want_Z else None
        leaves = np.empty(self.n, dtype=np.int32)
        self._check(self._L.hica_p1_linkage(self._h, _ptr(Z), _ptr(leaves)))
        return Z, leaves

    def p1_set_order(self, order: Sequence[int]):
        o = _c(order, np.int32)
        if o.size != self.n:
            raise ValueError("order must have n entries")
        self._check(self._L.hica_p1_set_order(self._h, _ptr(o)))

    def p1_get_reordered(self, what: int) -> np.ndarray:
        out = np.empty((self.n, self.n))
        self._check(self._L.hica_p1_get_reordered(self._h, int(what), _ptr(out)))
        return out

    def p1_order_stats(self, what: int, ranks):
        """hica_p1_order_stats: the elements at the given zero-based ascending ranks of D (what = 0), S (1) or
        the resident matrix as stored (2), and the number of NaN elements."""
        k = np.ascontiguousarray(ranks, dtype=np.int64)
        out = np.empty(k.size)
        nan = C.c_int64(0)
        self._check(self._L.hica_p1_order_stats(self._h, int(what), int(k.size), _ptr(k), _ptr(out), C.byref(nan)))
        return out, nan.value

    def order_stats_host(self, x: np.ndarray, ranks):
        """hica_order_stats_host: the same for a float64 array in host memory (flattened)."""
        a = np.ascontiguousarray(x, dtype=np.float64)
        k = np.ascontiguousarray(ranks, dtype=np.int64)
        out = np.empty(k.size)
        nan = C.c_int64(0)
        self._check(self._L.hica_order_stats_host(self._h, _ptr(a), int(a.size), int(k.size), _ptr(k), _ptr(out), C.byref(nan)))
        return out, nan.value

    def p1_rank(self):
        self._check(self._L.hica_p1_rank(self._h))

    def p1_get_rank(self) -> np.ndarray:
        R = np.empty((self.n, self.n), dtype=np.int64)
        self._check(self._L.hica_p1_get_rank(self._h, _ptr(R)))
        return R

    def p1_set_rank(self, R: np.ndarray):
        R = _c(R, np.int64)
        self.n = R.shape[0]
        self._check(self._L.hica_p1_set_rank(self._h, _ptr(R), self.n))

    def p1_window_sig(self, start: int, M: int, psig: float = .05, mode: int = 0, cut: int = 0):
        L = self.n - start
        sig = np.empty(L, dtype=np.uint8)
        x = np.empty(L, dtype=np.int64)
        amb = C.c_int64(0)
        self._check(self._L.hica_p1_window_sig(self._h, start, cut, M, psig, mode, _ptr(sig), _ptr(x), C.byref(amb)))
        return sig, x, amb.value

    def p1_breakpoints(self, min_size: int = 5, min_frac: float = .05):
        cuts = np.empty(max(16, self.n), dtype=np.int64)
        n = C.c_int64(0)
        amb = C.c_int64(0)
        self._check(self._L.hica_p1_breakpoints(self._h, int(min_size), float(min_frac), _ptr(cuts), cuts.size,
                                                C.byref(n), C.byref(amb)))
        return [int(v) for v in cuts[: n.value]], amb.value

    def p1_filter(self, cuts: Sequence[int], psig: float = .05):
        cin = _c(list(cuts), np.int64)
        cout = np.empty(max(1, cin.size), dtype=np.int64)
        n = C.c_int64(0)
        amb = C.c_int64(0)
        self._check(self._L.hica_p1_filter(self._h, _ptr(cin), cin.size, float(psig), _ptr(cout), C.byref(n),
                                           C.byref(amb)))
        return [int(v) for v in cout[: n.value]], amb.value

    def tie_report(self) -> dict:
        """hica_p1_tie_report: significance decisions of the last breakpoint / filter searches that depend on the
        order of exactly tied similarities (numpy.argsort's order is implementation-defined, :1132)."""
        out = np.zeros(4, dtype=np.int64)
        self._check(self._L.hica_p1_tie_report(self._h, _ptr(out)))
        return {"available": bool(out[0]), "breakpoints": int(out[1]), "filter": int(out[2]), "rows_with_ties": int(out[3])}

    def tie_sensitive(self) -> int:
        r = self.tie_report()
        return max(r["breakpoints"], 0) + max(r["filter"], 0) if r["available"] else 0

    def hypergeom_less(self, x, M, n, N, psig: float = .05) -> np.ndarray:
        x, M, n, N = (_c(np.broadcast_to(np.asarray(v), np.broadcast(x, M, n, N).shape).ravel(), np.int64)
                      for v in (x, M, n, N))
        out = np.empty(x.size, dtype=np.uint8)
        self._check(self._L.hica_hypergeom_less(self._h, _ptr(x), _ptr(M), _ptr(n), _ptr(N), x.size, psig, _ptr(out)))
        return out

    # -- 10 kb configuration: fp32 storage, row-block sharding (see hica.h) ----
    def s32_alloc(self, n: int, r0: int, r1: int, with_linkage_matrix: bool):
        self._check(self._L.hica_s32_alloc(self._h, int(n), int(r0), int(r1), 1 if with_linkage_matrix else 0))
        self.s32 = (int(n), int(r0), int(r1), (int(n) + 15) // 16 * 16)

    def s32_buffer(self, which: int) -> Tuple[int, int]:
        ptr, nbytes = C.c_void_p(0), C.c_int64(0)
        self._check(self._L.hica_s32_buffer(self._h, int(which), C.byref(ptr), C.byref(nbytes)))
        return int(ptr.value or 0), int(nbytes.value)

    def s32_row_sums(self):
        self._check(self._L.hica_s32_row_sums(self._h))

    def s32_build_w(self, dst_ptr: int = 0, ld: int = 0, dst_row0: int = 0):
        self._check(self._L.hica_s32_build_w(self._h, C.c_void_p(dst_ptr) if dst_ptr else None, int(ld), int(dst_row0)))

    def s32_ipc_export(self, which: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self._L.hica_s32_ipc_export(self._h, int(which), buf))
        return buf.raw

    def s32_ipc_open(self, handle: bytes) -> int:
        ptr = C.c_void_p(0)
        buf = C.create_string_buffer(handle, 64)
        self._check(self._L.hica_s32_ipc_open(self._h, buf, C.byref(ptr)))
        return int(ptr.value or 0)

    def s32_ipc_close(self, ptr: int):
        self._check(self._L.hica_s32_ipc_close(self._h, C.c_void_p(ptr)))

    def s32_linkage(self, want_Z: bool = True):
        n = self.s32[0]
        Z = np.empty((n - 1, 4)) if want_Z else None
        leaves = np.empty(n, dtype=np.int32)
        self._check(self._L.hica_s32_linkage(self._h, _ptr(Z), _ptr(leaves)))
        return Z, leaves

    # -- Part 2 -----------------------------------------------------------
    def p2_upload(self, G: np.ndarray):
        G = _c(G, np.float64)
        self._check(self._L.hica_p2_upload(self._h, _ptr(G), G.shape[0]))

    def p2_use_p1_matrix(self):
        self._check(self._L.hica_p2_use_p1_matrix(self._h))

    def p2_select(self, idx: Sequence[int]):
        idx = _c(idx, np.int32)
        self._check(self._L.hica_p2_select(self._h, _ptr(idx), idx.size))

    def p2_set_segments(self, seg_ptr: Sequence[int], seg_bins: Sequence[int]):
        sp = _c(seg_ptr, np.int32)
        sb = _c(seg_bins, np.int32)
        self._check(self._L.hica_p2_set_segments(self._h, _ptr(sp), _ptr(sb), sp.size - 1))

    def p2_total(self, cand: Sequence[int]) -> float:
        c = _c(cand, np.int32)
        t = C.c_double(0)
        self._check(self._L.hica_p2_total(self._h, _ptr(c), c.size, C.byref(t)))
        return t.value

    def _score_out(self, B, want_scores):
        return (np.empty(B) if want_scores else None), C.c_double(0), C.c_int64(-1)

    def p2_score_candidates(self, cands: np.ndarray, total: float, carried: float = 0.0, want_scores: bool = False):
        cands = _c(cands, np.int32)
        B, ns = cands.shape
        sc, best, bi = self._score_out(B, want_scores)
        self._check(self._L.hica_p2_score_candidates(self._h, _ptr(cands), B, ns, total, carried, C.byref(best),
                                                     C.byref(bi), _ptr(sc)))
        return best.value, bi.value, sc

    def p2_score_window(self, base: Sequence[int], win_start: int, win_n: int, order_table: np.ndarray,
                        orient_table: np.ndarray, total: float, carried: float = 0.0,
                        cand_lo: int = 0, cand_hi: Optional[int] = None, want_scores: bool = False):
        base = _c(base, np.int32)
        ot = _c(order_table, np.uint8)
        rt = _c(orient_table, np.uint8)
        ncand = ot.shape[0] * rt.shape[0]
        cand_hi = ncand if cand_hi is None else cand_hi
        sc, best, bi = self._score_out(cand_hi - cand_lo, want_scores)
        self._check(self._L.hica_p2_score_window(self._h, _ptr(base), base.size, win_start, win_n, _ptr(ot), ot.shape[0],
                                                 _ptr(rt), rt.shape[0], cand_lo, cand_hi, total, carried,
                                                 C.byref(best), C.byref(bi), _ptr(sc)))
        return best.value, bi.value, sc

    def p2_order_groups(self, grp_ptr, seg_ptr, bin_rows, n_scaffolds: int, scan: int, brute_tables: dict,
                        scan_tables: tuple):
        """brute_tables: {k: (order_table(k), orient_table(k))} for every k in use; scan_tables: (orders, orients)."""
        grp_ptr, seg_ptr, bin_rows = _c(grp_ptr, np.int32), _c(seg_ptr, np.int32), _c(bin_rows, np.int32)
        ngroups = grp_ptr.size - 1
        ob, rb, oo, ro = [], [], [0, 0], [0, 0]
        for k in range(1, 9):
            if k in brute_tables:
                ob.append(_c(brute_tables[k][0], np.uint8).ravel())
                rb.append(_c(brute_tables[k][1], np.uint8).ravel())
            else:
                ob.append(np.zeros(0, np.uint8))
                rb.append(np.zeros(0, np.uint8))
            oo.append(oo[-1] + ob[-1].size)
            ro.append(ro[-1] + rb[-1].size)
        ob, rb = np.concatenate(ob + [np.zeros(1, np.uint8)]), np.concatenate(rb + [np.zeros(1, np.uint8)])
        oo, ro = _c(oo, np.int32), _c(ro, np.int32)
        so, sr = _c(scan_tables[0], np.uint8), _c(scan_tables[1], np.uint8)
        layout = np.empty(seg_ptr.size - 1, dtype=np.int32)
        cost = np.empty(ngroups)
        rounds = np.empty(ngroups, dtype=np.int32)
        stats = np.zeros((ngroups, 4), dtype=np.int64)
        self._check(self._L.hica_p2_order_groups(self._h, ngroups, _ptr(grp_ptr), _ptr(seg_ptr), _ptr(bin_rows),
                                                 int(n_scaffolds), int(scan), _ptr(ob), _ptr(oo), _ptr(rb), _ptr(ro),
                                                 _ptr(so), so.shape[0], _ptr(sr), sr.shape[0], _ptr(layout),
                                                 _ptr(cost), _ptr(rounds), _ptr(stats)))
        return layout, cost, rounds, stats

    def p2_score_orders(self, orders: np.ndarray, total: float, carried: float = 0.0, want_scores: bool = False):
        orders = _c(orders, np.int32)
        B, n = orders.shape
        sc, best, bi = self._score_out(B, want_scores)
        self._check(self._L.hica_p2_score_orders(self._h, _ptr(orders), B, n, total, carried, C.byref(best),
                                                 C.byref(bi), _ptr(sc)))
        return best.value, bi.value, sc


def io_read_loci(bed_file: str, bias_file: str, keep_ids=None):
    """hica_io_read_loci + hica_io_loci_fetch: (ids, starts, stops, biases, chromosome names) of the bins
    initiateLoci keeps (host side, no GPU needed).  Raises ValueError where the reference raises."""
    L = load_library()
    n, cb = C.c_int64(0), C.c_int64(0)
    if keep_ids is None:
        keep, nk = None, -1
    else:
        keep = np.unique(np.asarray(list(keep_ids), dtype=np.int64))
        nk = int(keep.size)
    rc = L.hica_io_read_loci(os.fsencode(bed_file), os.fsencode(bias_file), _ptr(keep), nk, C.byref(n), C.byref(cb))
    if rc:
        msg = L.hica_io_last_error().decode()
        raise (FileNotFoundError if msg.startswith("cannot open") else ValueError)(msg)
    ids, starts, stops = (np.empty(n.value, dtype=np.int64) for _ in range(3))
    bias = np.empty(n.value)
    off = np.empty(n.value + 1, dtype=np.int64)
    buf = np.empty(max(1, cb.value), dtype=np.uint8)
    rc = L.hica_io_loci_fetch(_ptr(ids), _ptr(starts), _ptr(stops), _ptr(bias), _ptr(off), _ptr(buf))
    if rc:
        raise ValueError(L.hica_io_last_error().decode())
    raw = buf.tobytes()
    names, cache = [], {}
    o = off.tolist()
    for i in range(n.value):
        b = raw[o[i]:o[i + 1]]
        s = cache.get(b)
        if s is None:
            s = cache[b] = b.decode()
        names.append(s)
    return ids, starts, stops, bias, names


def io_log_transform(M: np.ndarray, base: float = 10.0, reverse: bool = False, threads: int = 0,
                     offset: float = 1.0) -> np.ndarray:
    """hica_io_log_transform on a float64 array (host side): a new array of the same shape."""
    L = load_library()
    a = np.ascontiguousarray(M, dtype=np.float64)
    out = np.empty_like(a)
    rc = L.hica_io_log_transform(_ptr(a), _ptr(out), a.size, float(base), float(offset), 1 if reverse else 0, int(threads))
    if rc:
        raise ValueError(L.hica_io_last_error().decode())
    return out
