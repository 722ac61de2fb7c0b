"""Host side of Part1 (scaffoldToChromosomes.py): the device pipeline and the
seam functions with the reference's signatures.

``run_part1`` drives the C ABI from a dense ICE matrix in host memory to
``(leaves, cut indices)`` - the timing boundary of SURVEY.md 8(d).  The
``Seams`` class holds drop-in replacements for the reference's module-level
functions (same names, arguments and return conventions); matrices returned to
Python are ``numpy.matrix`` objects tagged with the device state they mirror,
so that a following seam call can skip the re-upload.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import Handle
from .lazy import LazyMatrix, LazyRank, host_array


def run_part1(h: Handle, A: np.ndarray, min_size: int = 5, min_frac: float = .05, psig: float = .05,
              row_sums: Optional[np.ndarray] = None, want_Z: bool = False) -> dict:
    """scaffoldToChromosomes.runPipeline :1116-1134 without file I/O and plots:
    distance transform -> UPGMA -> leaf order -> similarity rank -> breakpoints -> noise filter."""
    h.p1_upload(A)
    return run_part1_resident(h, min_size, min_frac, psig, row_sums, want_Z)


def run_part1_resident(h: Handle, min_size: int = 5, min_frac: float = .05, psig: float = .05,
                       row_sums: Optional[np.ndarray] = None, want_Z: bool = False) -> dict:
    """Same, for a matrix that is already on the device (p1_upload / p1_upload_device)."""
    h.p1_set_row_sums(row_sums)
    h.p1_distance()
    Z, leaves = h.p1_linkage(want_Z=want_Z)
    h.p1_rank()
    init, amb0 = h.p1_breakpoints(min_size, min_frac)
    cuts, amb1 = h.p1_filter(init, psig)
    tie = h.tie_report()
    return dict(Z=Z, leaves=leaves, initial_cuts=init, cuts=cuts, ambiguous=amb0 + amb1,
                tie_sensitive=(tie["breakpoints"] + tie["filter"]) if tie["available"] else None)


# --------------------------------------------------------------------------- #
# bin -> scaffold -> chromosome group assignment (scaffoldToChromosomes.py:945-1100)
# --------------------------------------------------------------------------- #


def chromosome_groups(cuts: Sequence[int], leaves: Sequence[int], bin_ids: Sequence[int],
                      scaffold_of_bin: Sequence[str], scaffold_size_bp: Dict[str, int],
                      percent_to_assign: float = 51.) -> List[List[Tuple[int, str]]]:
    """In-memory restatement of writeBinGroupingsToFile (:945-962) -> assessChromosomeClustering
    (:1038-1077) -> writeChromosomeGroupingsToFile (:1079-1100): split the dendrogram order at the cut
    indices, assign a scaffold to a group when >= 51 % of its bins fall in it, sort groups by bp size.
    ``bin_ids`` / ``scaffold_of_bin`` are in matrix (pre-reorder) index order."""
    order = list(leaves)
    bounds = [0] + list(cuts) + [len(order)]
    bin_groups = [[order[k] for k in range(bounds[g], bounds[g + 1])] for g in range(len(bounds) - 1)]
    # all bins of every scaffold, sorted by bin id (:1047-1056)
    per_scaf: Dict[str, List[int]] = {}
    for grp in bin_groups:
        for i in grp:
            per_scaf.setdefault(scaffold_of_bin[i], []).append(int(bin_ids[i]))
    for s in per_scaf:
        per_scaf[s].sort()
    final = []
    for grp in bin_groups:
        counts: Dict[str, int] = {}
        for i in grp:
            counts[scaffold_of_bin[i]] = counts.get(scaffold_of_bin[i], 0) + 1
        nodes: List[Tuple[int, str]] = []
        for s, c in counts.items():  # dict order = first appearance, as in the reference
            pct = round((float(c) / float(len(per_scaf[s]))) * 100., 2)
            if pct >= percent_to_assign:
                nodes += [(b, s) for b in per_scaf[s]]
        if nodes:
            final.append(nodes)
    sizes = [sum(scaffold_size_bp[s] for s in dict.fromkeys(nm for _, nm in grp)) for grp in final]
    ranked = sorted(zip(final, sizes), key=lambda p: p[1], reverse=True)
    return [g for g, _ in ranked]


def chromosome_groups_compact(cuts: Sequence[int], leaves: Sequence[int], bin_ids: np.ndarray,
                             scaffold_code: np.ndarray, scaffold_names: Sequence[str],
                             scaffold_size_bp: Sequence[int], percent_to_assign: float = 51.):
    """Array form of :func:`chromosome_groups` for large N (no per-bin Python work): returns, per group,
    ``[(scaffold name, sorted bin-id array), ...]`` in the order the reference lists the scaffolds (first
    appearance inside the group), groups sorted by bp size.  ``scaffold_code[i]`` indexes ``scaffold_names`` /
    ``scaffold_size_bp`` for matrix row i."""
    n = len(leaves)
    leaves = np.asarray(leaves, dtype=np.int64)
    code = np.asarray(scaffold_code, dtype=np.int64)
    bin_ids = np.asarray(bin_ids, dtype=np.int64)
    S = len(scaffold_names)
    gid = np.searchsorted(np.asarray(list(cuts), dtype=np.int64), np.arange(n), side="right")
    code_o = code[leaves]
    key = gid * S + code_o
    n_keys = (len(cuts) + 1) * S
    if n_keys <= (1 << 22):
        # small key space: counting instead of np.unique's sort (1.2 ms at 16k bins)
        counts_full = np.bincount(key, minlength=n_keys)
        uniq = np.nonzero(counts_full)[0]
        counts = counts_full[uniq]
        first_full = np.empty(n_keys, dtype=np.int64)
        first_full[key[::-1]] = np.arange(n - 1, -1, -1)   # repeated index: the last assignment (lowest row) wins
        first = first_full[uniq]
    else:
        uniq, first, counts = np.unique(key, return_index=True, return_counts=True)
    total = np.bincount(code, minlength=S)
    by_scaf = np.lexsort((bin_ids, code))            # bins of every scaffold, ascending bin id
    start = np.concatenate([[0], np.cumsum(total)])
    grp_of, sc = uniq // S, uniq % S
    # round((float(c) / float(total)) * 100., 2) >= percent_to_assign (:1060-1066), decided without a Python call
    # per entry: rounding to two decimals moves x by at most 0.005, so only entries within 0.01 of the threshold
    # need the reference's own round()
    x = counts.astype(np.float64) / total[sc].astype(np.float64) * 100.
    assigned = x >= percent_to_assign + 0.01
    for k in np.nonzero((~assigned) & (x > percent_to_assign - 0.01))[0]:
        assigned[k] = round(float(x[k]), 2) >= percent_to_assign
    order = np.lexsort((first, grp_of))              # (group, first appearance)
    order = order[assigned[order]]
    sorted_bins = bin_ids[by_scaf]
    size_arr = np.asarray(scaffold_size_bp, dtype=np.int64)
    g_of, s_of = grp_of[order].tolist(), sc[order].tolist()
    lo, hi = start[sc[order]].tolist(), start[sc[order] + 1].tolist()
    bp_of = size_arr[sc[order]].tolist()
    final, sizes = [], []
    prev = -1
    for g, s_, a_, b_, bp in zip(g_of, s_of, lo, hi, bp_of):
        if g != prev:
            final.append([])
            sizes.append(0)
            prev = g
        final[-1].append((scaffold_names[s_], sorted_bins[a_:b_]))
        sizes[-1] += bp
    ranked = sorted(zip(final, sizes), key=lambda p: p[1], reverse=True)
    return [g for g, _ in ranked]


# --------------------------------------------------------------------------- #
# seams
# --------------------------------------------------------------------------- #


def _weak(obj):
    """weakref to ``obj`` or None (lists cannot be weakly referenced): identity caches hold no bare addresses."""
    import weakref
    try:
        return weakref.ref(obj)
    except TypeError:
        return None


class Seams:
    """Replacements for the module-level functions of scaffoldToChromosomes.py on the hot path (same names,
    arguments and return conventions).  The N x N matrices they hand back are :class:`lazy.LazyMatrix` /
    :class:`lazy.LazyRank` objects: the device keeps the state, and the host copy is built only if something
    other than the next seam looks at it (runPipeline :1116-1157 never does, unless plotting is installed).
    A seam that is handed a matrix it does not recognise as its own device state uploads it - any call
    sequence works, the resident one is just the fast lane."""

    def __init__(self, h: Handle):
        self.h = h
        self._epoch = 0
        self._live = []          # weak references to lazies that still need the current device state
        self._order = None
        self._rank_obj = None    # LazyRank of the resident rank matrix
        self._rank_host = None   # weakref to a host rank matrix that is known to equal the resident one

    # ---- bookkeeping
    def _fresh(self, kind):
        self._epoch += 1
        return (kind, self._epoch)

    def _is_current(self, M, kind) -> bool:
        tag = getattr(M, "_hica_tag", None)
        return tag is not None and tag == (kind, self._epoch)

    def _lazy(self, n, thunk, tag, kind):
        import weakref
        m = LazyMatrix((n, n), thunk, tag, kind)
        self._live = [w for w in self._live if w() is not None and not w().materialised]
        self._live.append(weakref.ref(m))
        return m

    def _before_overwrite(self, kinds=None):
        """Device state that live lazies still stand for is about to change (a new upload replaces the matrix:
        all kinds; a new order: "Dperm" and "S"; new row sums: "S"): they are brought to the host first - they stay
        valid, nothing is ever silently stale."""
        keep = []
        for w in self._live:
            m = w()
            if m is None or m.materialised:
                continue
            if m._hica_tag is not None and (kinds is None or m.kind in kinds):
                m.materialise()
            else:
                keep.append(w)
        self._live = keep
        if kinds is None:
            self._rank_obj = None
            self._rank_host = None

    @staticmethod
    def _row_sums(binList):
        return np.asarray([b.rowSum for b in binList], dtype=np.float64)

    # convertMatrix(adjacencyMatrix, binList, distance=True, similarity=False) :138-155
    def convertMatrix(self, adjacencyMatrix, binList, distance=True, similarity=False):
        t0 = time.time()
        h = self.h
        if distance == True:  # noqa: E712 - the reference compares with == True
            rs = self._row_sums(binList)
            if isinstance(adjacencyMatrix, LazyMatrix) and not adjacencyMatrix.materialised:
                # a deferred input (the pipeline's closing convertMatrix(logTransformMatrix(.., reverse=True)),
                # :1155, feeds only a plot): stay deferred, do the work if and when somebody looks
                src = adjacencyMatrix
                out = self._lazy(len(src), lambda: np.asarray(self._distance(host_array(src), rs).materialise()),
                                 None, "D(deferred)")
            else:
                out = self._distance(host_array(adjacencyMatrix), rs)
        elif similarity == True:  # noqa: E712
            rs = self._row_sums(binList)
            if self._is_current(adjacencyMatrix, "Dperm"):
                # reordered distance matrix of the resident ICE matrix: row sums follow the reordered binList
                full = np.empty(h.n)
                full[np.asarray(self._order)] = rs
                self._before_overwrite({"S"})
                h.p1_set_row_sums(full)
            else:
                D = host_array(adjacencyMatrix)
                self._before_overwrite()
                h.p1_upload_distance(D)
                h.p1_set_row_sums(rs)
                h.p1_set_order(np.arange(D.shape[0], dtype=np.int32))
                self._order = list(range(D.shape[0]))
            out = self._lazy(h.n, lambda: h.p1_get_reordered(1), self._fresh("S"), "S")
        else:
            out = np.asmatrix(adjacencyMatrix)
        print("Time to convert matrix " + str(time.time() - t0))
        return out

    def _distance(self, A, rs):
        h = self.h
        self._before_overwrite()
        h.p1_upload(A)
        h.p1_set_row_sums(rs)
        h.p1_distance()
        return self._lazy(h.n, lambda: h.p1_get_distance(), self._fresh("D"), "D")

    # averageClusterNodes(adjacencyMatrix, nodeLabels, noPlot=True) :187-208
    def averageClusterNodes(self, adjacencyMatrix, nodeLabels, noPlot=True):
        t0 = time.time()
        h = self.h
        if not self._is_current(adjacencyMatrix, "D"):
            D = host_array(adjacencyMatrix)
            self._before_overwrite()
            h.p1_upload_distance(D)
            self._dhost = (self._fresh("Dhost"), _weak(adjacencyMatrix))
        Z, leaves = h.p1_linkage(want_Z=True)
        leaves = [int(v) for v in leaves]
        self._order = leaves
        # a linkage consumes W; the D tag stays valid for reorderMatrix
        print("Time to cluster " + str(time.time() - t0))
        return {"leaves": leaves, "ivl": [nodeLabels[i] for i in leaves], "Z": Z}

    # reorderMatrix(matrix, binList, newOrder) :157-163
    def reorderMatrix(self, matrix, binList, newOrder):
        h = self.h
        order = [int(v) for v in newOrder]
        tag = getattr(matrix, "_hica_tag", None)
        resident = tag is not None and tag == ("D", self._epoch)
        dh = getattr(self, "_dhost", None)   # the host matrix averageClusterNodes uploaded, if it is this very object
        if not resident and dh is not None and dh[0] == ("Dhost", self._epoch) and dh[1] is not None and dh[1]() is matrix:
            resident = True
        if not resident:
            # a matrix that is not the resident one: upload it as is and gather on the device
            D = host_array(matrix)
            self._before_overwrite()
            h.p1_upload_distance(D)
            self._fresh("Dhost")
        self._before_overwrite({"Dperm", "S"})
        h.p1_set_order(np.asarray(order, dtype=np.int32))
        self._order = order
        out = self._lazy(h.n, lambda: h.p1_get_reordered(0), self._fresh("Dperm"), "Dperm")
        return out, [binList[i] for i in order]

    # numpy.percentile(adjMat, lP) / numpy.percentile(adjMat, hP), plotContactMaps.py:30-31 (bound through the plot
    # module's `numpy` name): the colour-scale clip of the two Part1 plots (:1124, :1156)
    def percentile(self, adjMat, q):
        from . import plot as hplot
        if isinstance(adjMat, LazyMatrix) and adjMat._hica_tag is not None:
            kind, epoch = adjMat._hica_tag
            if epoch == self._epoch and kind in ("D", "Dperm", "S"):
                # the resident matrix in some row order: a percentile does not depend on the order
                return hplot.percentile_device(self.h, adjMat, q, what=1 if kind == "S" else 0)
        return hplot.percentile_device(self.h, host_array(adjMat), q)

    # numpy.argsort(adjMat, axis=1)[:, ::-1] :1132 (bound through the module's `numpy` name)
    def argsort_desc(self, S) -> LazyRank:
        h = self.h
        if not self._is_current(S, "S"):
            Sh = host_array(S)
            self._before_overwrite()
            h.p1_upload_similarity(Sh)
            self._fresh("Shost")
        h.p1_rank()
        self._rank_obj = LazyRank(h.n, h.p1_get_rank, ("R", self._epoch))
        self._rank_host = None
        return self._rank_obj

    def _ensure_rank(self, argsorted_mat):
        if argsorted_mat is self._rank_obj and self._rank_obj is not None and self._rank_obj._hica_tag == ("R", self._epoch):
            return                                     # the device already holds this rank matrix
        if self._rank_host is not None and self._rank_host() is argsorted_mat:
            return                                     # the very array object that was uploaded last (still alive)
        if self._rank_obj is not None and self._rank_obj._host is not None and argsorted_mat is self._rank_obj._host \
                and self._rank_obj._hica_tag == ("R", self._epoch):
            return                                     # the host copy of the resident rank matrix
        R = np.ascontiguousarray(np.asarray(argsorted_mat), dtype=np.int64)
        self._rank_obj = None
        self.h.p1_set_rank(R)
        self._fresh("Rhost")
        self._rank_host = _weak(argsorted_mat)

    def _report(self, amb, tie):
        if amb:
            print("- WARNING - {} hypergeometric decisions fell within 1e-7 of the significance level".format(amb))
        if tie:
            print("- WARNING - {} significance decisions depend on the order numpy.argsort gives exactly tied "
                  "values (implementation-defined, scaffoldToChromosomes.py:1132)".format(tie))

    # pre_process_all_matrix_breakpoints(argsorted_mat, min_size=5, min_frac=.05, psig=.05) :513-551
    def pre_process_all_matrix_breakpoints(self, argsorted_mat, min_size=5, min_frac=.05, psig=.05):
        self._ensure_rank(argsorted_mat)
        cuts, amb = self.h.p1_breakpoints(int(min_size), float(min_frac))
        self._report(amb, self.h.tie_sensitive() if hasattr(self.h, "tie_sensitive") else 0)
        print("- Breakpoints found {}".format(len(cuts)))
        return cuts

    # filter_noisy_breakpoints(argsorted_mat, original_inds, psig=.05) :553-727
    def filter_noisy_breakpoints(self, argsorted_mat, original_inds, psig=.05):
        if len(original_inds) == 0:
            return []
        self._ensure_rank(argsorted_mat)
        cuts, amb = self.h.p1_filter([int(v) for v in original_inds], float(psig))
        self._report(amb, self.h.tie_sensitive() if hasattr(self.h, "tie_sensitive") else 0)
        print("- Original cut indices {}".format(list(original_inds)))
        print("- Filtered cut indices {}".format(cuts))
        return cuts
