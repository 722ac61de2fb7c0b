"""Device-backed stand-ins for the N x N matrices that travel between the reference's pipeline steps.

scaffoldToChromosomes.runPipeline (:1116-1157) hands the matrix from seam to seam through Python variables:
convertMatrix -> averageClusterNodes -> reorderMatrix -> plot -> convertMatrix -> numpy.argsort ->
pre_process_all_matrix_breakpoints -> filter_noisy_breakpoints -> logTransformMatrix -> convertMatrix -> plot.
Only the seams themselves ever look inside it (the plots do too, when matplotlib is installed).  Returning real
host arrays from every seam costs three 8 N^2-byte read-backs plus an int64 rank matrix (4 x 2 GB at 100 kb); the
objects below are returned instead.  They know their shape, carry a tag that lets the next seam recognise the
device state they stand for, and build the host array - through exactly the code path the eager seams use, so the
values are the same bit for bit - the first time anything else touches them.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


class LazyMatrix:
    """An N x N float64 matrix that lives on the device (or is a deferred host computation on one)."""

    __array_priority__ = 1000.0
    _OWN = ("_shape", "_thunk", "_host", "_hica_tag", "kind")

    def __init__(self, shape: Tuple[int, int], thunk: Callable[[], np.ndarray], tag=None, kind: str = ""):
        object.__setattr__(self, "_shape", tuple(int(v) for v in shape))
        object.__setattr__(self, "_thunk", thunk)
        object.__setattr__(self, "_host", None)
        object.__setattr__(self, "_hica_tag", tag)
        object.__setattr__(self, "kind", kind)

    # ---- what can be answered without the data
    @property
    def shape(self):
        return self._shape

    @property
    def ndim(self):
        return 2

    @property
    def dtype(self):
        return np.dtype(np.float64)

    @property
    def size(self):
        return self._shape[0] * self._shape[1]

    def __len__(self):
        return self._shape[0]

    @property
    def materialised(self) -> bool:
        return self._host is not None

    # ---- everything else builds the host matrix once
    def materialise(self) -> np.matrix:
        if self._host is None:
            object.__setattr__(self, "_host", np.asmatrix(self._thunk()))
            object.__setattr__(self, "_thunk", None)
        return self._host

    def __array__(self, dtype=None, copy=None):
        m = np.asarray(self.materialise())
        return m.astype(dtype) if dtype is not None and np.dtype(dtype) != m.dtype else m

    def __getitem__(self, key):
        return self.materialise()[key]

    def __setitem__(self, key, value):
        self.materialise()[key] = value
        object.__setattr__(self, "_hica_tag", None)   # no longer the device's state

    def __iter__(self):
        return iter(self.materialise())

    def __getattr__(self, name):   # only reached for names not defined above
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return getattr(self.materialise(), name)

    def __repr__(self):
        return "<LazyMatrix %s %dx%d%s>" % (self.kind, self._shape[0], self._shape[1],
                                            " (on host)" if self.materialised else "")


def _binary(name):
    def op(self, other):
        other = other.materialise() if isinstance(other, LazyMatrix) else other
        return getattr(self.materialise(), name)(other)
    op.__name__ = name
    return op


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__",
           "__pow__", "__matmul__", "__rmatmul__", "__eq__", "__ne__", "__lt__", "__le__", "__gt__", "__ge__"):
    setattr(LazyMatrix, _n, _binary(_n))
LazyMatrix.__hash__ = object.__hash__
LazyMatrix.__neg__ = lambda self: -self.materialise()


class LazyRank:
    """The result of ``numpy.argsort(S, axis=1)[:, ::-1]`` (:1132) left on the device as the uint16/int32 rank
    matrix the cut searches read; the int64 host array exists only if something asks for it."""

    def __init__(self, n: int, fetch: Callable[[], np.ndarray], tag=None):
        self._n, self._fetch, self._host, self._hica_tag = int(n), fetch, None, tag

    @property
    def shape(self):
        return (self._n, self._n)

    ndim = 2
    dtype = np.dtype(np.int64)

    def __len__(self):
        return self._n

    def materialise(self) -> np.ndarray:
        if self._host is None:
            self._host = self._fetch()
        return self._host

    def __array__(self, dtype=None, copy=None):
        m = self.materialise()
        return m.astype(dtype) if dtype is not None and np.dtype(dtype) != m.dtype else m

    def __getitem__(self, key):
        # the reference immediately takes [:, ::-1] of numpy.argsort's result; the device rank is already descending
        return self.materialise()[key]

    def __iter__(self):
        return iter(self.materialise())

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return getattr(self.materialise(), name)


def host_array(M) -> Optional[np.ndarray]:
    """float64 C-contiguous ndarray of whatever a seam was handed (lazy object, numpy.matrix, list of lists)."""
    if isinstance(M, LazyMatrix):
        return np.ascontiguousarray(np.asarray(M.materialise(), dtype=np.float64))
    return np.ascontiguousarray(np.asarray(M, dtype=np.float64))
