"""CPU oracle for the contact-map scoring hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, function by function, what the reference does on the hot
path.  It is imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s CPU-baseline / ``--impl
reference`` legs.  The product package never imports it and has no CPU fallback.

Citations are to ``/root/reference/HIC_ASSEMBLER/<file>:<lines>``.

Third-party arithmetic the reference delegates to (none of it is vendored in
the reference tree, and the reference pins no versions -
packageInstallCommands.txt:9-17).  The oracle calls the same functions at the
same call sites, with the versions of this image:

  * numpy 2.3.5   - ``matrix.sum`` (pairwise tree), ``argsort``, ``ix_``, ``trace``
  * scipy 1.18.1  - ``spatial.distance.squareform``, ``cluster.hierarchy.average``
                    (-> ``_hierarchy.nn_chain``), ``cluster.hierarchy.dendrogram``,
                    ``stats.hypergeom.sf`` (Boost.Math ``_hypergeom_sf``)
  * numba 0.65.0  - ``np.trace`` overload inside ``costFunction_numba``: a strictly
                    sequential sum (numba/np/linalg.py:2642-2670), restated here
                    with ``numpy.add.accumulate`` which is also strictly sequential.

Parity pinning: the reference ships no tests or golden vectors.
The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in
the build container: ``tests/golden/make_golden.py`` imports the unmodified
reference modules (via ``oracle/ref_harness.py``) and stores their results as
fixtures under ``tests/golden/``; ``tests/test_oracle_vs_golden.py`` checks this
file against them, and ``tests/test_oracle_vs_reference.py`` checks it live when
the reference tree is mounted.
"""
from __future__ import annotations

import sys
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------- #
# Part 1 - matrix transforms
# --------------------------------------------------------------------------- #


def pairwise_row_sums(A: np.ndarray) -> np.ndarray:
    """``row.sum()`` of every row, numpy pairwise-tree order.

    scaffoldToChromosomes.py:147 evaluates ``row.sum()`` on a 1xN ``np.matrix``
    row; a contiguous reduction over the last axis uses the same pairwise
    routine on the same n elements.
    """
    A = np.ascontiguousarray(A, dtype=np.float64)
    return np.add.reduce(A, axis=1)


def numpy_pairwise_sum(a: np.ndarray) -> float:
    """Pure-Python restatement of numpy's pairwise summation (small cases only).

    n < 8: sequential from 0.; n <= 128: eight interleaved accumulators combined
    as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then a sequential tail; otherwise
    split at n/2 rounded down to a multiple of 8.  This is the specification the
    CUDA row-sum kernel follows.
    """
    n = len(a)
    if n < 8:
        res = 0.0
        for v in a:
            res += float(v)
        return res
    if n <= 128:
        r = [float(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += float(a[i + j])
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += float(a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:])


def sequential_row_sums(A: np.ndarray) -> np.ndarray:
    """``binList[i].rowSum`` as set by removeRows: builtin ``sum`` over the row,
    i.e. strictly left-to-right (scaffoldToChromosomes.py:134)."""
    A = np.asarray(A, dtype=np.float64)
    if A.shape[1] == 0:
        return np.zeros(A.shape[0])
    return np.add.accumulate(A, axis=1)[:, -1]


def convert_distance(A: np.ndarray) -> np.ndarray:
    """convertMatrix(distance=True): D_ij = (1 - A_ij/s_i) + 1 for ALL j
    (scaffoldToChromosomes.py:145-147).  Returns a new array."""
    A = np.asarray(A, dtype=np.float64)
    s = pairwise_row_sums(A)
    return (1.0 - (A / s[:, None])) + 1.0


def convert_similarity(D: np.ndarray, row_sum: np.ndarray) -> np.ndarray:
    """convertMatrix(similarity=True): S_ij = rowSum_i * (1 - (D_ij - 1))
    (scaffoldToChromosomes.py:148-149)."""
    D = np.asarray(D, dtype=np.float64)
    return np.asarray(row_sum, dtype=np.float64)[:, None] * (1.0 - (D - 1.0))


def reorder(M: np.ndarray, order: Sequence[int]) -> np.ndarray:
    """reorderMatrix: M[:, o][o] (scaffoldToChromosomes.py:157-163)."""
    o = np.asarray(order, dtype=np.int64)
    return np.asarray(M)[:, o][o]


# --------------------------------------------------------------------------- #
# Part 1 - UPGMA and leaf order
# --------------------------------------------------------------------------- #


def condensed_upper(D: np.ndarray) -> np.ndarray:
    """squareform(checks=False) on a square matrix = row-major strict upper
    triangle (scaffoldToChromosomes.py:194)."""
    import scipy.spatial.distance as ssd

    return ssd.squareform(np.asarray(D), checks=False)


def linkage_average(D: np.ndarray) -> np.ndarray:
    """scipy.cluster.hierarchy.average on the condensed upper triangle
    (scaffoldToChromosomes.py:194-197) -> Z (n-1, 4)."""
    import scipy.cluster.hierarchy as sch

    return sch.average(condensed_upper(D))


def leaf_order_scipy(Z: np.ndarray) -> List[int]:
    """dendrogram(no_plot=True, get_leaves=True, count_sort='ascending')['leaves']
    (scaffoldToChromosomes.py:204)."""
    import scipy.cluster.hierarchy as sch

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 100000))  # the reference does the same, :9
    try:
        return list(sch.dendrogram(Z, no_plot=True, get_leaves=True,
                                   count_sort='ascending')['leaves'])
    finally:
        sys.setrecursionlimit(old)


def nn_chain_average(y: np.ndarray, n: int) -> np.ndarray:
    """Pure-Python restatement of scipy's ``_hierarchy.nn_chain(y, n, average)``
    followed by the stable sort on heights and union-find relabel (small cases
    only; ``oracle/hica_oracle.c`` holds the fast C twin).

    Rules (SURVEY section 7 hard part 2): chain restarts at the lowest live index;
    neighbours scanned in ascending index with strict ``<``; the minimum is
    seeded with the previous chain element; merged pair ordered x<y, the union
    lives in slot y; update d(i, y) = (nx*d(i,x) + ny*d(i,y)) / (nx+ny).
    """
    D = {}

    def cidx(i, j):
        if i > j:
            i, j = j, i
        return n * i - (i * (i + 1) // 2) + (j - i - 1)

    y = np.array(y, dtype=np.float64, copy=True)
    size = [1] * n
    chain: List[int] = []
    raw = []
    for _ in range(n - 1):
        if not chain:
            for i in range(n):
                if size[i] > 0:
                    chain.append(i)
                    break
        while True:
            x = chain[-1]
            if len(chain) > 1:
                yy = chain[-2]
                cur = y[cidx(x, yy)]
            else:
                yy = -1
                cur = np.inf
            for i in range(n):
                if size[i] == 0 or i == x:
                    continue
                d = y[cidx(x, i)]
                if d < cur:
                    cur = d
                    yy = i
            if len(chain) > 1 and yy == chain[-2]:
                break
            chain.append(yy)
        chain.pop()
        chain.pop()
        a, b = (x, yy) if x < yy else (yy, x)
        nx, ny = size[a], size[b]
        raw.append((a, b, float(cur), nx + ny))
        size[a] = 0
        size[b] = nx + ny
        for i in range(n):
            ni = size[i]
            if ni == 0 or i == b:
                continue
            y[cidx(i, b)] = (nx * y[cidx(i, a)] + ny * y[cidx(i, b)]) / (nx + ny)
    return label_merges(np.array(raw, dtype=np.float64).reshape(n - 1, 4), n)


def label_merges(raw: np.ndarray, n: int) -> np.ndarray:
    """Stable sort of raw NN-chain merges by height + union-find relabel
    (smaller root id first; new cluster id = n + k)."""
    order = np.argsort(raw[:, 2], kind='mergesort')
    Z = raw[order].copy()
    parent = list(range(2 * n - 1))
    sizes = [1] * n + [0] * (n - 1)
    nxt = n

    def find(x):
        p = x
        while parent[x] != x:
            x = parent[x]
        while parent[p] != x:
            p, parent[p] = parent[p], x
        return x

    for i in range(n - 1):
        a, b = find(int(Z[i, 0])), find(int(Z[i, 1]))
        if a > b:
            a, b = b, a
        Z[i, 0], Z[i, 1] = a, b
        parent[a] = nxt
        parent[b] = nxt
        sizes[nxt] = sizes[a] + sizes[b]
        Z[i, 3] = sizes[nxt]
        nxt += 1
    return Z


def leaf_order(Z: np.ndarray) -> List[int]:
    """Iterative restatement of scipy's dendrogram leaf order with
    count_sort='ascending': at every node the child with FEWER leaves is
    visited first; on equal counts Z[i,0] goes first."""
    Z = np.asarray(Z)
    n = Z.shape[0] + 1
    if n == 1:
        return [0]

    def count(c):
        return 1 if c < n else int(Z[c - n, 3])

    out: List[int] = []
    stack = [2 * n - 2]
    while stack:
        c = stack.pop()
        if c < n:
            out.append(c)
            continue
        a, b = int(Z[c - n, 0]), int(Z[c - n, 1])
        if count(a) > count(b):
            a, b = b, a
        stack.append(b)
        stack.append(a)
    return out


def average_cluster_leaves(D: np.ndarray) -> Tuple[np.ndarray, List[int]]:
    """averageClusterNodes (scaffoldToChromosomes.py:187-208) -> (Z, leaves)."""
    Z = linkage_average(D)
    return Z, leaf_order_scipy(Z)


# --------------------------------------------------------------------------- #
# Part 1 - rank matrix and hypergeometric cut indices
# --------------------------------------------------------------------------- #


def rank_matrix(S: np.ndarray) -> np.ndarray:
    """numpy.argsort(S, axis=1)[:, ::-1] (scaffoldToChromosomes.py:1132).

    NOTE numpy's default sort is unstable; rows with exactly tied values come
    back in an implementation-defined order."""
    return np.asarray(np.argsort(np.asarray(S), axis=1)[:, ::-1])


def rank_matrix_stable(S: np.ndarray) -> np.ndarray:
    """Deterministic variant: stable ascending argsort, reversed (ties come out
    in DESCENDING column order).  Equals ``rank_matrix`` on tie-free rows; the
    CUDA rank kernel implements this rule."""
    return np.asarray(np.argsort(np.asarray(S), axis=1, kind='stable')[:, ::-1])


def hyper_geom(x, M, n, N):
    """hyper_geom (scaffoldToChromosomes.py:352-368): P[X >= x]."""
    import scipy.stats

    return scipy.stats.hypergeom.sf(np.asarray(x) - 1, M, n, N)


def window_counts(R: np.ndarray, start: int) -> np.ndarray:
    """x_i = #{k < i-start : start <= R[i,k] <= i} for i in (start, N)
    (scaffoldToChromosomes.py:452-459).  Returns an array of length N-start with
    entry 0 unused (=0)."""
    N = R.shape[0]
    out = np.zeros(N - start, dtype=np.int64)
    for i in range(start + 1, N):
        c = i - start
        pr = R[i, :c]
        out[c] = np.count_nonzero((pr >= start) & (pr <= i))
    return out


def window_significance(x: np.ndarray, M: int, psig: float = 0.05) -> np.ndarray:
    """dist_sigs of find_matrix_pvalue_breakpoints for a given world size M
    (scaffoldToChromosomes.py:446-469): entry 0 is 0; entry c is 0 when
    ``pval >= psig`` and 1 otherwise (so NaN counts as significant)."""
    L = len(x)
    sig = np.zeros(L, dtype=np.int64)
    if L > 1:
        c = np.arange(1, L)
        with np.errstate(all='ignore'):
            p = hyper_geom(x[1:], M, c, c)
        sig[1:] = np.where(p >= psig, 0, 1)
    return sig


def sliding_break_signal(sig: np.ndarray, w: int) -> Optional[np.ndarray]:
    """get_sliding_window_distance_metrics (scaffoldToChromosomes.py:370-411):
    b[t] = sum(sig[t:t+w]) - sum(sig[t+w:t+2w]) for t in [0, len-w), 0 where
    the right half is truncated.  None when w >= len(sig) ("NA")."""
    L = len(sig)
    if w >= L:
        return None
    cs = np.concatenate([[0], np.cumsum(sig)])
    b = np.zeros(L - w, dtype=np.int64)
    for t in range(L - w):
        if t + 2 * w <= L:
            b[t] = (cs[t + w] - cs[t]) - (cs[t + 2 * w] - cs[t + w])
    return b


def find_breakpoints(R: np.ndarray, start: int, min_size: int, world_size: int,
                     psig: float = 0.05) -> List[int]:
    """find_matrix_pvalue_breakpoints (scaffoldToChromosomes.py:413-511) ->
    relative cut offsets (pre_cut_inds)."""
    M = world_size
    ws = min_size
    break_sig = 0
    loop_count = 0
    x = window_counts(R, start)   # does not depend on M
    while True:
        while True:
            sig = window_significance(x, M, psig)
            loop_count += 1
            if (int(sig.sum()) / len(sig)) >= .9:
                M = int(M - start)
            else:
                break_sig = 1
            if break_sig == 1 or loop_count >= 5:
                break
        b = sliding_break_signal(sig, ws)
        hits = [] if b is None else [int(t) + min_size for t in np.nonzero(b == min_size)[0]]
        if hits:
            return hits
        ws -= 1
        if ws == 0:
            return []


def pre_process_breakpoints(R: np.ndarray, min_size: int = 5, min_frac: float = .05,
                            psig: float = .05) -> List[int]:
    """pre_process_all_matrix_breakpoints (scaffoldToChromosomes.py:513-551).
    The first-stage psig is the literal .05 regardless of the argument (:535)."""
    N = len(R)
    stop_ind = int(N - (N * min_frac))
    ind = 0
    cinds: List[int] = []
    if min_frac == 1:
        return cinds
    while True:
        cuts = find_breakpoints(R, ind, min_size, N - ind, psig=.05)
        if len(cuts) == 0:
            break
        ind += cuts[0]
        cinds.append(int(ind))
        if ind >= stop_ind or (N - ind) <= min_size:
            break
    return cinds


def filter_row_significance(R: np.ndarray, start: int, c: int, psig: float) -> np.ndarray:
    """The per-row 0/1 vector of filter_noisy_breakpoints for one (start, cut)
    (scaffoldToChromosomes.py:617-636); index = ii - start, length N - start.
    NaN p-values count as NOT significant here (``pval < psig`` is False)."""
    N = len(R)
    MD = int(N / 5)
    M = N - start
    local = c - start
    out = np.zeros(N - start, dtype=np.int64)
    hi = min(N, start + MD + 1)
    sub = R[start:hi, :local]
    x = np.count_nonzero((sub >= start) & (sub <= c), axis=1)
    with np.errstate(all='ignore'):
        p = hyper_geom(x, M, local, local)
    out[: hi - start] = np.where(p < psig, 1, 0)
    return out


def filter_noisy_breakpoints(R: np.ndarray, original_inds: Sequence[int],
                             psig: float = .05) -> List[int]:
    """filter_noisy_breakpoints (scaffoldToChromosomes.py:553-727)."""
    if len(original_inds) == 0:
        return []
    N = len(R)
    max_rounds = 10 * len(original_inds)
    altered = list(original_inds)
    prev_filtered: Dict[int, str] = {}
    while True:
        start = 0
        filtered: Dict[int, str] = {}
        round_count = 0
        while True:
            if round_count >= max_rounds:
                break
            M = N - start
            noise_found = 0
            keep_from = 0
            for i, c in enumerate(altered):
                local = c - start
                sig = filter_row_significance(R, start, c, psig)
                cs = np.concatenate([[0], np.cumsum(sig)])
                right_most = None
                right_most_ind = None
                fc_prev = start
                for ai_ind, ai in enumerate(altered):
                    lo, hi_ = fc_prev, ai
                    if ai == fc_prev:
                        continue
                    fc_prev = ai
                    npts = max(0, hi_ - lo)
                    if npts == 0:
                        break
                    xs = int(cs[hi_ - start] - cs[lo - start])
                    with np.errstate(all='ignore'):
                        noise_p = hyper_geom(xs, M, local, npts)
                    if noise_p < psig:
                        right_most = ai
                        right_most_ind = ai_ind
                if right_most is not None:
                    start = right_most
                    filtered[right_most] = ''
                    noise_found = 1
                    keep_from = right_most_ind
                    break
                filtered[c] = ''
                keep_from = i
            round_count += 1
            if noise_found == 0:
                break
            altered = altered[keep_from:]
        if prev_filtered != filtered:
            altered = sorted(filtered)
            prev_filtered = filtered
        else:
            break
    return sorted(filtered.keys())


# ---- tie-order sensitivity of the window counts (hard part 3) ---------------------------------
#
# numpy.argsort(S, axis=1)[:, ::-1] (scaffoldToChromosomes.py:1132) orders exactly tied similarities arbitrarily.
# A count x = #{k < c : lo <= R[i][k] <= hi} (:455-459, :631) depends on that order only when sorted position c
# falls inside a run of tied values.  Over all orders of the run it ranges over [x_lo, x_hi]:


def count_bounds(vals_sorted: np.ndarray, cols_sorted: np.ndarray, c: int, lo: int, hi: int) -> Tuple[int, int, int]:
    """(x, x_lo, x_hi) for one row: ``vals_sorted`` = the row's similarities in rank order (descending),
    ``cols_sorted`` = R[i] (the columns in that order), window = the first ``c`` entries, in range = [lo, hi]."""
    n = len(vals_sorted)
    inr = (cols_sorted >= lo) & (cols_sorted <= hi)
    x = int(np.count_nonzero(inr[:c]))
    if c <= 0 or c >= n or vals_sorted[c] != vals_sorted[c - 1]:
        return x, x, x
    v = vals_sorted[c]
    g0 = c - 1
    while g0 > 0 and vals_sorted[g0 - 1] == v:
        g0 -= 1
    g1 = c + 1
    while g1 < n and vals_sorted[g1] == v:
        g1 += 1
    x_fixed = int(np.count_nonzero(inr[:g0]))
    m_in = int(np.count_nonzero(inr[g0:g1]))
    k_in, g = c - g0, g1 - g0
    return x, x_fixed + max(0, k_in - (g - m_in)), x_fixed + min(k_in, m_in)


def _sig0(x, M, psig=.05):
    """mode-0 decision vector for counts x[1:], n = N = c (NaN counts as significant, :466-469)."""
    sig = np.zeros(len(x), dtype=np.int64)
    if len(x) > 1:
        c = np.arange(1, len(x))
        with np.errstate(all='ignore'):
            p = hyper_geom(np.asarray(x[1:]), M, c, c)
        sig[1:] = np.where(p >= psig, 0, 1)
    return sig


def tie_sensitive_breakpoints(S: np.ndarray, R: np.ndarray, min_size: int = 5, min_frac: float = .05) -> int:
    """How many row decisions of pre_process_all_matrix_breakpoints (:513-551, through
    find_matrix_pvalue_breakpoints :413-511) depend on the order of exactly tied similarities; every evaluation
    of the significances counts (they are re-evaluated when M shrinks and for every narrower window)."""
    N = len(R)
    stop_ind = int(N - (N * min_frac))
    ind, total = 0, 0
    if min_frac == 1:
        return 0
    Ssorted = np.take_along_axis(np.asarray(S), np.asarray(R), 1)
    while True:
        start = ind
        L = N - start
        x = np.zeros(L, dtype=np.int64)
        xlo, xhi = x.copy(), x.copy()
        for i in range(start + 1, N):
            c = i - start
            x[c], xlo[c], xhi[c] = count_bounds(Ssorted[i], R[i], c, start, i)
        M, ws, break_sig, loop_count, cuts = N - start, min_size, 0, 0, []
        while True:
            while True:
                sig = _sig0(x, M)
                total += int(np.count_nonzero(_sig0(xlo, M) != _sig0(xhi, M)))
                loop_count += 1
                if (int(sig.sum()) / len(sig)) >= .9:
                    M = int(M - start)
                else:
                    break_sig = 1
                if break_sig == 1 or loop_count >= 5:
                    break
            b = sliding_break_signal(sig, ws)
            cuts = [] if b is None else [int(t) + min_size for t in np.nonzero(b == min_size)[0]]
            if cuts:
                break
            ws -= 1
            if ws == 0:
                break
        if not cuts:
            break
        ind += cuts[0]
        if ind >= stop_ind or (N - ind) <= min_size:
            break
    return total


def tie_sensitive_filter(S: np.ndarray, R: np.ndarray, original_inds: Sequence[int], psig: float = .05) -> int:
    """The same for the row decisions of filter_noisy_breakpoints (:617-636), over the (start, cut) pairs its loop
    evaluates (NaN p-values count as not significant there)."""
    if len(original_inds) == 0:
        return 0
    N = len(R)
    MD = int(N / 5)
    Ssorted = np.take_along_axis(np.asarray(S), np.asarray(R), 1)
    total = 0
    max_rounds = 10 * len(original_inds)
    altered = list(original_inds)
    prev_filtered: Dict[int, str] = {}
    while True:
        start = 0
        filtered: Dict[int, str] = {}
        round_count = 0
        while True:
            if round_count >= max_rounds:
                break
            M = N - start
            noise_found = 0
            keep_from = 0
            for i, c in enumerate(altered):
                local = c - start
                sig = filter_row_significance(R, start, c, psig)
                hi = min(N, start + MD + 1)
                lo_x = np.zeros(hi - start, dtype=np.int64)
                hi_x = lo_x.copy()
                for ii in range(start, hi):
                    _, lo_x[ii - start], hi_x[ii - start] = count_bounds(Ssorted[ii], R[ii], local, start, c)
                with np.errstate(all='ignore'):
                    total += int(np.count_nonzero((hyper_geom(lo_x, M, local, local) < psig) !=
                                                  (hyper_geom(hi_x, M, local, local) < psig)))
                cs = np.concatenate([[0], np.cumsum(sig)])
                right_most = None
                right_most_ind = None
                fc_prev = start
                for ai_ind, ai in enumerate(altered):
                    lo, hi_ = fc_prev, ai
                    if ai == fc_prev:
                        continue
                    fc_prev = ai
                    npts = max(0, hi_ - lo)
                    if npts == 0:
                        break
                    xs = int(cs[hi_ - start] - cs[lo - start])
                    with np.errstate(all='ignore'):
                        noise_p = hyper_geom(xs, M, local, npts)
                    if noise_p < psig:
                        right_most = ai
                        right_most_ind = ai_ind
                if right_most is not None:
                    start = right_most
                    filtered[right_most] = ''
                    noise_found = 1
                    keep_from = right_most_ind
                    break
                filtered[c] = ''
                keep_from = i
            round_count += 1
            if noise_found == 0:
                break
            altered = altered[keep_from:]
        if prev_filtered != filtered:
            altered = sorted(filtered)
            prev_filtered = filtered
        else:
            break
    return total


def part1_cut_indices(A: np.ndarray, min_size: int = 5, min_frac: float = 0.0,
                      psig: float = .05, stable_rank: bool = False):
    """The Part1 hot path from the dense ICE matrix to (Z, leaves, initial cuts,
    filtered cuts): scaffoldToChromosomes.py:1116-1134 without file I/O."""
    A = np.asarray(A, dtype=np.float64)
    rs = sequential_row_sums(A)
    D = convert_distance(A)
    Z, leaves = average_cluster_leaves(D)
    Dp = reorder(D, leaves)
    S = convert_similarity(Dp, rs[np.asarray(leaves)])
    R = rank_matrix_stable(S) if stable_rank else rank_matrix(S)
    init = pre_process_breakpoints(R, min_size=min_size, min_frac=min_frac, psig=psig)
    cuts = filter_noisy_breakpoints(R, init, psig=psig)
    return dict(Z=Z, leaves=leaves, R=R, initial_cuts=init, cuts=cuts)


# --------------------------------------------------------------------------- #
# Part 2 - ordering score
# --------------------------------------------------------------------------- #


def total_weight(M: np.ndarray):
    """sum([numpy.trace(M, offset=i) for i in range(1, len(M))])
    (orderGenome.py:343, 448, 506): numpy pairwise on each strided diagonal,
    then a left-to-right Python sum starting from int 0."""
    M = np.asarray(M)
    return sum([np.trace(M, offset=i) for i in range(1, len(M))])


def ordering_cost(M: np.ndarray, total: float) -> float:
    """costFunction_numba (orderGenome.py:184-191).  numba's np.trace is a
    sequential sum down the diagonal; so is ``add.accumulate``."""
    M = np.asarray(M, dtype=np.float64)
    L = M.shape[0]
    cum = 0.0
    cost = 0.0
    for i in range(1, L):
        d = np.diagonal(M, i)
        cum += float(np.add.accumulate(d)[-1])
        cost += (cum / total / float(i))
    return cost


_numba_cost = None


def ordering_cost_numba(M: np.ndarray, total: float) -> float:
    """The same loop compiled with numba's nopython JIT, exactly as the reference runs it
    (orderGenome.py:183-191); used for CPU-baseline timing and as a cross-check of ``ordering_cost``."""
    global _numba_cost
    if _numba_cost is None:
        from numba import jit
        from numpy import trace as np_trace

        @jit(nopython=True)
        def _cost(matrix, total):
            cum, cost = 0., 0.
            for i in range(1, len(matrix)):
                cum += np_trace(matrix, offset=i)
                cost += (cum / total / float(i))
            return cost

        _numba_cost = _cost
    return float(_numba_cost(np.ascontiguousarray(M, dtype=np.float64), float(total)))


def ordering_cost_of(M: np.ndarray, order: Sequence[int], total: float) -> float:
    """cost of M[ix_(order, order)] (orderGenome.py:348, 463, 534)."""
    o = np.asarray(order, dtype=np.int64)
    return ordering_cost(np.asarray(M)[np.ix_(o, o)], total)


# ---- enumeration (orderGenome.py:374-430) --------------------------------- #


def _swap_permutations(items: list) -> List[list]:
    """All permutations in the order of the reference's recursive swap
    generator (orderGenome.py:381-394)."""
    out: List[list] = []
    a = list(items)
    n = len(a)

    def rec(k):
        if k == n:
            out.append(list(a))
            return
        for i in range(k, n):
            a[k], a[i] = a[i], a[k]
            rec(k + 1)
            a[k], a[i] = a[i], a[k]

    rec(0)
    return out


def order_candidates(n: int) -> List[Tuple[int, ...]]:
    """Index permutations surviving removeReverseDuplicates, in enumeration
    order (orderGenome.py:396-411): of each {p, reversed p} pair the one seen
    first is kept."""
    seen = set()
    out = []
    for p in _swap_permutations(list(range(n))):
        t = tuple(p)
        r = t[::-1]
        if r in seen:
            seen.discard(r)
            continue
        seen.add(t)
        out.append(t)
    return out


def orient_candidates(n: int) -> List[Tuple[str, ...]]:
    """plusMinusPerms (orderGenome.py:413-430): all-plus first, then for
    i = 0..n-1 the distinct arrangements of i '+' and n-i '-' in swap order."""
    seq = [tuple(['+'] * n)]
    for i in range(n):
        seq.extend(tuple(p) for p in _swap_permutations(['+'] * i + ['-'] * (n - i)))
    seen = {}
    for p in seq:
        seen.setdefault(p, None)
    return list(seen.keys())


class Scaf:
    """Minimal stand-in for orderGenome.Scaffold (:238-254): ``bins`` are the
    bin ids in current 5'->3' reading order; flipping reverses them."""

    __slots__ = ("name", "bins", "orientation")

    def __init__(self, name, bins, orientation="+"):
        self.name = name
        self.bins = list(bins)
        self.orientation = orientation

    def flip(self):
        self.orientation = "-" if self.orientation == "+" else "+"
        self.bins = self.bins[::-1]

    def set_orientation(self, o):
        if self.orientation != o:
            self.flip()


def _sub_matrix(G: np.ndarray, index_of_bin: Dict[int, int], scafs: Sequence[Scaf]):
    """giveNewAdjMat (orderGenome.py:297-308)."""
    bins = [b for s in scafs for b in s.bins]
    local = {b: i for i, b in enumerate(bins)}
    idx = [index_of_bin[b] for b in bins]
    return np.asarray(G)[np.ix_(idx, idx)], local


def brute_force(scafs: List[Scaf], M: np.ndarray, local: Dict[int, int]):
    """bruteForceBestScore (orderGenome.py:432-473) -> (order idx tuple,
    orientation tuple, best cost); first strict maximum wins."""
    n = len(scafs)
    orders = order_candidates(n)
    orients = orient_candidates(n)
    total = total_weight(M)
    if total == 0:
        return orders[0], orients[0], 0.0
    best = (None, None, 0.0)
    for od in orders:
        for orr in orients:
            bins = []
            for si, o in zip(od, orr):
                scafs[si].set_orientation(o)
                bins += scafs[si].bins
            cost = ordering_cost_of(M, [local[b] for b in bins], total)
            if cost > best[2]:
                best = (od, orr, cost)
    return best


def check_all_scores(M: np.ndarray, local: Dict[int, int], ordered: List[Scaf], new: Scaf):
    """checkAllScores (orderGenome.py:332-372): 2(k+1) candidates with the
    reference's alternating orientation state."""
    best_cost = 0.0
    best_pos = 0
    best_orient = "+"
    total = total_weight(M)
    for i in range(len(ordered) + 1):
        ordered.insert(i, new)
        for _ in range(2):
            o = [local[b] for s in ordered for b in s.bins]
            cost = ordering_cost_of(M, o, total)
            if cost > best_cost:
                best_cost, best_pos, best_orient = cost, i, ordered[i].orientation
            if _ == 0:
                ordered[i].flip()
        new = ordered.pop(i)
    new.set_orientation(best_orient)
    ordered.insert(best_pos, new)
    return ordered, best_cost


def scan_ordering(G: np.ndarray, index_of_bin: Dict[int, int], ordered: List[Scaf],
                  best_cost: float, scan: int = 5):
    """scanOrdering (orderGenome.py:495-549)."""
    M, local = _sub_matrix(G, index_of_bin, ordered)
    total = total_weight(M)
    best_state = [(s, s.orientation) for s in ordered]
    rounds = 0
    while True:
        improved = False
        for i in range(0, len(ordered) - scan + 1):
            window = ordered[i:i + scan]
            head, tail = ordered[:i], ordered[i + scan:]
            for od in order_candidates(scan):
                for orr in orient_candidates(scan):
                    mid = []
                    for wi, o in zip(od, orr):
                        window[wi].set_orientation(o)
                        mid.append(window[wi])
                    cand = head + mid + tail
                    o_idx = [local[b] for s in cand for b in s.bins]
                    cost = ordering_cost_of(M, o_idx, total)
                    if cost > best_cost:
                        best_cost = cost
                        best_state = [(s, s.orientation) for s in cand]
                        improved = True
            for s, o in best_state:
                s.set_orientation(o)
            ordered = [s for s, _ in best_state]
            M, local = _sub_matrix(G, index_of_bin, ordered)
        rounds += 1
        if not improved:
            break
    return ordered, best_cost, rounds


def order_chromosome(G: np.ndarray, index_of_bin: Dict[int, int],
                     nodes: Sequence[Tuple[int, str]], n_scaffolds: int = 6, scan: int = 5):
    """orderChromosome (orderGenome.py:551-586) on the genome matrix ``G``;
    ``nodes`` = [(binID, scaffoldName), ...] of one chromosome group.
    Returns [(name, orientation), ...]."""
    if n_scaffolds >= 9:
        n_scaffolds = 8
    if scan > n_scaffolds:
        scan = n_scaffolds
    by_name: Dict[str, Scaf] = {}
    for b, nm in nodes:
        by_name.setdefault(nm, Scaf(nm, []))
    for b, nm in nodes:
        by_name[nm].bins.append(b)
    for s in by_name.values():
        s.bins.sort()
    pool = sorted(by_name.values(), key=lambda s: len(s.bins), reverse=True)
    ordered, pool = pool[:n_scaffolds], pool[n_scaffolds:]
    M, local = _sub_matrix(G, index_of_bin, ordered)
    od, orr, cost = brute_force(ordered, M, local)
    for si, o in zip(od, orr):
        ordered[si].set_orientation(o)
    ordered = [ordered[si] for si in od]
    best_cost = cost
    # orderRemainderScaffolds (orderGenome.py:475-493) is a do-while: even with
    # nothing left to place, the LAST brute-forced scaffold is popped and
    # re-inserted by checkAllScores.
    while True:
        if pool:
            ordered.append(pool.pop(0))
        M, local = _sub_matrix(G, index_of_bin, ordered)
        new = ordered.pop(-1)
        ordered, best_cost = check_all_scores(M, local, ordered, new)
        if not pool:
            break
    if len(ordered) > n_scaffolds:
        ordered, best_cost, _ = scan_ordering(G, index_of_bin, ordered, best_cost, scan)
    return [(s.name, s.orientation) for s in ordered], best_cost
