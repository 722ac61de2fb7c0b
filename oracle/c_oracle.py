"""ctypes binding of oracle/_build/libhica_oracle.so (test infrastructure)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhica_oracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "hica_oracle.c")):
            build()
        L = C.CDLL(_SO)
        L.hica_o_pairwise_sum.restype = C.c_double
        L.hica_o_pairwise_sum.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.hica_o_row_sums_pairwise.argtypes = [_dp, C.c_int64, _dp]
        L.hica_o_row_sums_sequential.argtypes = [_dp, C.c_int64, _dp]
        L.hica_o_convert_distance.argtypes = [_dp, C.c_int64, _dp]
        L.hica_o_nn_chain_average.argtypes = [_dp, C.c_int64, _dp]
        L.hica_o_nn_chain_average.restype = C.c_int
        L.hica_o_label.argtypes = [_dp, C.c_int64, _dp]
        L.hica_o_leaf_order.argtypes = [_dp, C.c_int64, _ip]
        L.hica_o_total_weight.restype = C.c_double
        L.hica_o_total_weight.argtypes = [_dp, C.c_int64]
        L.hica_o_ordering_cost.restype = C.c_double
        L.hica_o_ordering_cost.argtypes = [_dp, C.c_int64, _ip, C.c_int64, C.c_double]
        L.hica_o_score_batch.argtypes = [_dp, C.c_int64, _ip, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                         C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.hica_o_window_counts.argtypes = [_ip, C.c_int64, C.c_int64, _lp]
        L.hica_o_filter_counts.argtypes = [_ip, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _lp]
        L.hica_o_rank_rows.argtypes = [_dp, C.c_int64, _ip]
        _lib = L
    return _lib


def pairwise_sum(a: np.ndarray) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().hica_o_pairwise_sum(a.ctypes.data, a.size, 1)


def row_sums_pairwise(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    s = np.empty(A.shape[0])
    lib().hica_o_row_sums_pairwise(A, A.shape[0], s)
    return s


def row_sums_sequential(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    s = np.empty(A.shape[0])
    lib().hica_o_row_sums_sequential(A, A.shape[0], s)
    return s


def convert_distance(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    D = np.empty_like(A)
    lib().hica_o_convert_distance(A, A.shape[0], D)
    return D


def linkage_average(D):
    """(Z, leaves) from the square asymmetric distance matrix D."""
    D = np.ascontiguousarray(D, dtype=np.float64)
    n = D.shape[0]
    raw = np.empty((max(n - 1, 0), 4))
    if lib().hica_o_nn_chain_average(D, n, raw):
        raise MemoryError
    Z = np.empty_like(raw)
    lib().hica_o_label(raw, n, Z)
    leaves = np.empty(n, dtype=np.int32)
    lib().hica_o_leaf_order(Z, n, leaves)
    return Z, leaves, raw


def total_weight(M):
    M = np.ascontiguousarray(M, dtype=np.float64)
    return lib().hica_o_total_weight(M, M.shape[0])


def ordering_cost(M, order, total):
    M = np.ascontiguousarray(M, dtype=np.float64)
    o = np.ascontiguousarray(order, dtype=np.int32)
    return lib().hica_o_ordering_cost(M, M.shape[0], o, o.size, float(total))


def score_batch(M, orders, total, carried=0.0, want_scores=True):
    M = np.ascontiguousarray(M, dtype=np.float64)
    o = np.ascontiguousarray(orders, dtype=np.int32)
    B, n = o.shape
    scores = np.empty(B) if want_scores else None
    best = C.c_double()
    bi = C.c_int64()
    lib().hica_o_score_batch(M, M.shape[0], o, B, n, float(total), float(carried),
                             scores.ctypes.data if want_scores else None, C.byref(best), C.byref(bi))
    return best.value, bi.value, scores


def window_counts(R, start):
    R = np.ascontiguousarray(R, dtype=np.int32)
    n = R.shape[0]
    x = np.zeros(n - start, dtype=np.int64)
    lib().hica_o_window_counts(R, n, start, x)
    return x


def filter_counts(R, start, cut, rows):
    R = np.ascontiguousarray(R, dtype=np.int32)
    x = np.zeros(rows, dtype=np.int64)
    lib().hica_o_filter_counts(R, R.shape[0], start, cut, rows, x)
    return x


def rank_rows(S):
    S = np.ascontiguousarray(S, dtype=np.float64)
    R = np.empty(S.shape, dtype=np.int32)
    lib().hica_o_rank_rows(S, S.shape[0], R)
    return R
