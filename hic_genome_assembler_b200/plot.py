"""The plot inputs either side of the hot path: the colour-scale clip of plotContactMap.

plotContactMap (plotContactMaps.py:15-31) calls ``numpy.percentile(adjMat, lP)`` and ``numpy.percentile(adjMat, hP)``
on the N x N matrix it draws: two flattened copies and two introselect partitions (seconds each at 16,000 bins).
Here the two order statistics next to each percentile come from the device (``hica_p1_order_stats`` on the resident
matrix - a percentile does not care about the row order -, ``hica_order_stats_host`` for a host array) and the
interpolation below restates numpy's own arithmetic (numpy/lib/_function_base_impl.py: ``percentile`` ->
``_quantile`` with method "linear" -> ``_lerp``), so the result equals numpy's bit for bit.
"""
from __future__ import annotations

from typing import Callable, Sequence, Tuple

import numpy as np


def _lerp(a, b, t):
    """numpy's _lerp: a + (b - a) t, switched to b - (b - a)(1 - t) for t >= 0.5."""
    diff_b_a = np.subtract(b, a)
    out = np.asanyarray(np.add(a, diff_b_a * t))
    np.subtract(b, diff_b_a * (1 - t), out=out, where=t >= 0.5)
    return out


def percentile_from_order_stats(count: int, q, fetch: Callable[[Sequence[int]], Tuple[np.ndarray, int]]):
    """numpy.percentile(x, q) (method "linear", flattened) of ``count`` values, given ``fetch(ranks)`` -> (the
    elements at those zero-based ascending ranks among the non-NaN values, number of NaN values)."""
    qf = np.true_divide(q, np.float64(100))
    if qf.ndim > 1:
        raise ValueError("q must be a scalar or 1d")
    if not (np.all(0 <= qf) and np.all(qf <= 1)):
        raise ValueError("Percentiles must be in the range [0, 100]")
    virt = np.asanyarray((count - 1) * qf)                      # _QuantileMethods['linear']['get_virtual_index']
    prev = np.asanyarray(np.floor(virt))                        # _get_indexes
    nxt = np.asanyarray(prev + 1)
    above = virt >= count - 1
    if above.any():
        prev[above] = -1
        nxt[above] = -1
    below = virt < 0
    if below.any():
        prev[below] = 0
        nxt[below] = 0
    prev = prev.astype(np.intp)
    nxt = nxt.astype(np.intp)
    ranks = sorted({int(v) % count for v in np.concatenate((prev.ravel(), nxt.ravel()))})
    vals, nan = fetch(ranks)
    if nan:                                                     # a NaN sorts last: numpy answers nan everywhere
        res = np.full(virt.shape, np.nan)
        return res[()] if res.ndim == 0 else res
    at = dict(zip(ranks, np.asarray(vals, dtype=np.float64)))
    previous = np.asarray([at[int(v) % count] for v in prev.ravel()]).reshape(prev.shape)
    following = np.asarray([at[int(v) % count] for v in nxt.ravel()]).reshape(nxt.shape)
    gamma = np.asanyarray(virt - prev, dtype=virt.dtype)        # _get_gamma
    res = _lerp(previous, following, gamma)
    return res[()] if res.ndim == 0 else res


def percentile_device(h, x, q, what=None):
    """numpy.percentile(x, q) with the order statistics taken on the device: ``what`` = 0 / 1 / 2 reads the
    handle's resident matrix (D / S / as stored; ``x`` is then only asked for its size), None uploads ``x``."""
    if what is not None:
        count = int(h.n) * int(h.n)
        return percentile_from_order_stats(count, q, lambda r: h.p1_order_stats(what, r))
    a = np.ascontiguousarray(np.asarray(x), dtype=np.float64)
    return percentile_from_order_stats(int(a.size), q, lambda r: h.order_stats_host(a, r))
