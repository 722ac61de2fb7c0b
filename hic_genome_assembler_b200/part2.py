"""Host side of Part2 (orderGenome.py): candidate enumeration and search control.

The arithmetic (gather + ordering score + first-strict-maximum) runs in
libhica's CUDA kernels; this module mirrors the reference's interface for the
path - same function names, argument meaning and accept rules - and turns each
search loop into ONE batched device call:

  reference loop (orderGenome.py)                 device call
  ---------------------------------------------   ---------------------------
  bruteForceBestScore :432-473  (n!/2 * 2^n)      hica_p2_score_window
  checkAllScores      :332-372  (2(k+1))          hica_p2_score_candidates
  scanOrdering window :513-541  (60 * 32 at 5)    hica_p2_score_window

A chromosome's contact matrix is gathered ONCE on the device
(hica_p2_select); every candidate is a signed permutation of scaffold
segments of that matrix, so no permuted matrix is ever materialised.
"""
from __future__ import annotations

import functools
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import Handle

# --------------------------------------------------------------------------- #
# enumeration order of the reference (orderGenome.py:374-430)
# --------------------------------------------------------------------------- #


def _swap_order_permutations(symbols: Sequence) -> List[tuple]:
    """Every arrangement of ``symbols`` in the order the reference's recursive
    swap generator emits them (:381-394); duplicates are kept."""
    work = list(symbols)
    n = len(work)
    found: List[tuple] = []
    # explicit stack instead of recursion: frame = (depth k, next swap index i)
    stack = [[0, 0]]
    while stack:
        k, i = stack[-1]
        if k == n:
            found.append(tuple(work))
            stack.pop()
            if stack:
                kk, ii = stack[-1]
                work[kk], work[ii - 1] = work[ii - 1], work[kk]  # undo the swap that led here
            continue
        if i == n:
            stack.pop()
            if stack:
                kk, ii = stack[-1]
                work[kk], work[ii - 1] = work[ii - 1], work[kk]
            continue
        work[k], work[i] = work[i], work[k]
        stack[-1][1] = i + 1
        stack.append([k + 1, k + 1])
    return found


@functools.lru_cache(maxsize=None)
def order_table(n: int) -> np.ndarray:
    """(n!/2, n) uint8: index permutations that survive removeReverseDuplicates
    (:396-411) - of every {p, reversed(p)} pair the first one enumerated."""
    kept, seen = [], set()
    for p in _swap_order_permutations(range(n)):
        r = p[::-1]
        if r in seen:
            seen.discard(r)
            continue
        seen.add(p)
        kept.append(p)
    return np.asarray(kept, dtype=np.uint8).reshape(len(kept), n)


@functools.lru_cache(maxsize=None)
def orient_table(n: int) -> np.ndarray:
    """(2^n, n) uint8, 0 = '+', 1 = '-': plusMinusPerms (:413-430)."""
    seq = [tuple([0] * n)]
    for i in range(n):
        seq.extend(_swap_order_permutations([0] * i + [1] * (n - i)))
    uniq = list(dict.fromkeys(seq))
    return np.asarray(uniq, dtype=np.uint8).reshape(len(uniq), n)


def calc_possible_perms(n: int) -> float:
    import math
    return math.factorial(n) * (2 ** n) / 2


# --------------------------------------------------------------------------- #
# device-resident chromosome
# --------------------------------------------------------------------------- #


class Scaf:
    """A scaffold of the chromosome being ordered: ``bins`` = bin ids in '+'
    (ascending id) reading order, ``seg`` = segment index on the device."""

    __slots__ = ("name", "bins", "orientation", "seg")

    def __init__(self, name, bins, orientation="+", seg=-1):
        self.name = name
        self.bins = list(bins)
        self.orientation = orientation
        self.seg = seg

    def entry(self, orientation: Optional[str] = None) -> int:
        o = self.orientation if orientation is None else orientation
        return (self.seg << 1) | (1 if o == "-" else 0)


def _flip(o: str) -> str:
    return "-" if o == "+" else "+"


STATS = {"candidates": 0, "bin_pairs": 0, "calls": 0}


def _account(n_cand: int, scafs) -> None:
    L = sum(len(s.bins) for s in scafs)
    STATS["candidates"] += int(n_cand)
    STATS["bin_pairs"] += int(n_cand) * (L * (L - 1) // 2)
    STATS["calls"] += 1


class DeviceChromosome:
    """One chromosome group resident on the GPU: its sub-matrix gathered from
    the genome matrix and its scaffolds registered as segments."""

    def __init__(self, h: Handle, genome_index_of_bin: Dict[int, int], scafs: Sequence[Scaf]):
        self.h = h
        self.scafs = list(scafs)
        idx, ptr, local = [], [0], []
        for s_i, s in enumerate(self.scafs):
            s.seg = s_i
            for b in s.bins:
                local.append(len(idx))
                idx.append(genome_index_of_bin[b])
            ptr.append(len(idx))
        h.p2_select(np.asarray(idx, dtype=np.int32))
        h.p2_set_segments(np.asarray(ptr, dtype=np.int32), np.asarray(local, dtype=np.int32))

    # -- the three searches -------------------------------------------------
    def total(self, layout: Sequence[Scaf]) -> float:
        return self.h.p2_total([s.entry() for s in layout])

    def brute_force(self, layout: List[Scaf]):
        """bruteForceBestScore (:432-473) -> (order idx tuple, orientation tuple of '+'/'-', cost).
        ``layout`` is the pulled scaffold list in its current state."""
        n = len(layout)
        orders, orients = order_table(n), orient_table(n)
        total = self.total(layout)
        if total == 0:
            return tuple(int(v) for v in orders[0]), tuple("-" if v else "+" for v in orients[0]), 0.0
        base = [s.entry() for s in layout]
        best, idx, _ = self.h.p2_score_window(base, 0, n, orders, orients, total, 0.0)
        _account(orders.shape[0] * orients.shape[0], layout)
        if idx < 0:
            raise RuntimeError("no ordering scored above 0 (the reference would fail on 'NA' here)")
        oi, ri = divmod(idx, orients.shape[0])
        return tuple(int(v) for v in orders[oi]), tuple("-" if v else "+" for v in orients[ri]), best

    def insert_best(self, ordered: List[Scaf], new: Scaf) -> Tuple[List[Scaf], float]:
        """checkAllScores (:332-372): 2(k+1) candidates, orientation alternating
        with the insertion index exactly as the reference's in-place flips do."""
        k = len(ordered)
        total = self.total(ordered + [new])
        base = [s.entry() for s in ordered]
        o0 = new.orientation
        cands, tried = [], []
        for i in range(k + 1):
            first = o0 if i % 2 == 0 else _flip(o0)
            for o in (first, _flip(first)):
                cands.append(base[:i] + [new.entry(o)] + base[i:])
                tried.append((i, o))
        best, idx, _ = self.h.p2_score_candidates(np.asarray(cands, dtype=np.int32), total, 0.0)
        _account(len(cands), ordered + [new])
        pos, orient = (0, "+") if idx < 0 else tried[idx]
        new.orientation = orient
        ordered.insert(pos, new)
        return ordered, (best if idx >= 0 else 0.0)

    def scan(self, ordered: List[Scaf], best_cost: float, scan: int) -> Tuple[List[Scaf], float, int]:
        """scanOrdering (:495-549): sliding window, carried best cost, rounds to a fixed point."""
        total = self.total(ordered)
        orders, orients = order_table(scan), orient_table(scan)
        rounds = 0
        while True:
            improved = False
            for i in range(0, len(ordered) - scan + 1):
                base = [s.entry() for s in ordered]
                best, idx, _ = self.h.p2_score_window(base, i, scan, orders, orients, total, best_cost)
                _account(orders.shape[0] * orients.shape[0], ordered)
                if idx >= 0:
                    oi, ri = divmod(idx, orients.shape[0])
                    window = ordered[i:i + scan]
                    arranged = []
                    for j in range(scan):
                        s = window[int(orders[oi][j])]
                        s.orientation = "-" if orients[ri][j] else "+"
                        arranged.append(s)
                    ordered[i:i + scan] = arranged
                    best_cost = best
                    improved = True
            rounds += 1
            if not improved:
                break
        return ordered, best_cost, rounds


def order_chromosome(h: Handle, genome_index_of_bin: Dict[int, int], nodes: Sequence[Tuple[int, str]],
                     n_scaffolds: int = 6, scan: int = 5):
    """orderChromosome (:551-586) for one chromosome group; the genome matrix must
    already be resident (Handle.p2_upload / p2_use_p1_matrix).
    Returns ([(scaffold name, '+'/'-'), ...], best cost, scan rounds)."""
    if n_scaffolds >= 9:
        n_scaffolds = 8
    if scan > n_scaffolds:
        scan = n_scaffolds
    by_name: Dict[str, Scaf] = {}
    for b, nm in nodes:
        if nm not in by_name:
            by_name[nm] = Scaf(nm, [])
    for b, nm in nodes:
        by_name[nm].bins.append(b)
    for s in by_name.values():
        s.bins.sort()
    pool = sorted(by_name.values(), key=lambda s: len(s.bins), reverse=True)
    chrom = DeviceChromosome(h, genome_index_of_bin, pool)
    ordered, pool = pool[:n_scaffolds], pool[n_scaffolds:]
    od, orr, _ = chrom.brute_force(ordered)
    for si, o in zip(od, orr):
        ordered[si].orientation = o
    ordered = [ordered[si] for si in od]
    best_cost = 0.0
    # orderRemainderScaffolds (:475-493) is a do-while: with nothing left to place the last
    # brute-forced scaffold is still popped and re-inserted.
    while True:
        if pool:
            ordered.append(pool.pop(0))
        new = ordered.pop(-1)
        ordered, best_cost = chrom.insert_best(ordered, new)
        if not pool:
            break
    rounds = 0
    if len(ordered) > n_scaffolds:
        ordered, best_cost, rounds = chrom.scan(ordered, best_cost, scan)
    return [(s.name, s.orientation) for s in ordered], best_cost, rounds


def _pool(nodes: Sequence[Tuple[int, str]]) -> List[Scaf]:
    """initiateBinsAndScaffolds (:256-281): bins ascending inside a scaffold, scaffolds by bin count
    descending (stable: ties keep first-appearance order)."""
    by_name: Dict[str, Scaf] = {}
    for b, nm in nodes:
        if nm not in by_name:
            by_name[nm] = Scaf(nm, [])
    for b, nm in nodes:
        by_name[nm].bins.append(b)
    for s in by_name.values():
        s.bins.sort()
    return sorted(by_name.values(), key=lambda s: len(s.bins), reverse=True)


def reference_work(lens: Sequence[int], n_scaffolds: int, scan: int, rounds: int) -> Tuple[int, int]:
    """What the reference evaluates for one chromosome group whose scaffolds (pool order: bin count descending)
    have ``lens`` bins and whose sliding-window search ran ``rounds`` rounds: (candidates, candidates x
    L(L-1)/2 bin pairs each) over bruteForceBestScore (:432-473), the checkAllScores steps (:475-493, a
    do-while) and scanOrdering (:495-549).  The device engine reaches the same decisions by filtering the
    candidates in closed form and re-scoring a short list exactly, so this is REFERENCE-EQUIVALENT work, not
    memory traffic of the GPU path; bench.py and its reference arm both count with this function."""
    if n_scaffolds >= 9:
        n_scaffolds = 8
    scan = min(scan, n_scaffolds)
    lens = list(lens)
    k = min(n_scaffolds, len(lens))
    pairs = lambda n: n * (n - 1) // 2  # noqa: E731
    nb = sum(lens[:k])
    cand = order_table(k).shape[0] * orient_table(k).shape[0] if nb >= 2 else 0
    work = cand * pairs(nb)
    placed, bins, first = k, nb, True
    while first or placed < len(lens):
        first = False
        if placed < len(lens):
            m, bins, placed = placed, bins + lens[placed], placed + 1
        else:
            m = placed - 1
        cand += 2 * (m + 1)
        work += 2 * (m + 1) * pairs(bins)
    if len(lens) > n_scaffolds:
        c = int(rounds) * (len(lens) - scan + 1) * order_table(scan).shape[0] * orient_table(scan).shape[0]
        cand += c
        work += c * pairs(sum(lens))
    return cand, work


def order_groups_device(h: Handle, genome_index_of_bin: Dict[int, int],
                        chrom_groups: Sequence[Sequence[Tuple[int, str]]], n_scaffolds: int = 6, scan: int = 5):
    """orderGenome (:591-628) for all groups in ONE device call (hica_p2_order_groups): every chromosome runs its
    brute force / insertion / scan on its own CUDA stream with device-side decisions.
    Returns ([[(name, orientation), ...] per group], costs, rounds, stats)."""
    if n_scaffolds >= 9:
        n_scaffolds = 8
    if scan > n_scaffolds:
        scan = n_scaffolds
    compact = bool(chrom_groups) and bool(chrom_groups[0]) and isinstance(chrom_groups[0][0][1], np.ndarray)
    if compact:
        # groups from part1.chromosome_groups_compact: [(name, sorted bin-id array), ...]; the pool order of
        # initiateBinsAndScaffolds is a stable sort by bin count, descending
        pools, pool_lens = [], []
        for g in chrom_groups:
            lens = [len(ids) for _, ids in g]
            order = np.argsort(-np.asarray(lens, dtype=np.int64), kind="stable").tolist()
            pools.append([g[i] for i in order])
            pool_lens.append([lens[i] for i in order])
        lookup = genome_index_of_bin if isinstance(genome_index_of_bin, np.ndarray) else None
        all_ids = [ids for pool in pools for _, ids in pool]
        lens_all = np.asarray([l for pl in pool_lens for l in pl], dtype=np.int64)
        seg_ptr = np.concatenate([[0], np.cumsum(lens_all)]).tolist()
        grp_ptr = np.concatenate([[0], np.cumsum([len(p) for p in pools])]).tolist()
        flat = np.concatenate(all_ids) if all_ids else np.zeros(0, dtype=np.int64)
        if lookup is not None:
            rows = lookup[flat]
        else:
            rows = np.fromiter((genome_index_of_bin[int(b)] for b in flat), dtype=np.int64, count=len(flat))
        pool_names = [[nm for nm, _ in pool] for pool in pools]
    else:
        pools = [_pool(g) for g in chrom_groups]
        grp_ptr, seg_ptr, rows, nrows = [0], [0], [], 0
        for pool in pools:
            for s in pool:
                rows.append(np.fromiter((genome_index_of_bin[int(b)] for b in s.bins), dtype=np.int64, count=len(s.bins)))
                nrows += len(s.bins)
                seg_ptr.append(nrows)
            grp_ptr.append(len(seg_ptr) - 1)
        rows = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
        pool_names = [[s.name for s in pool] for pool in pools]
        pool_lens = [[len(s.bins) for s in pool] for pool in pools]
    ks = sorted({min(n_scaffolds, len(p)) for p in pools})
    brute = {k: (order_table(k), orient_table(k)) for k in ks}
    layout, cost, rounds, stats = h.p2_order_groups(grp_ptr, seg_ptr, rows, n_scaffolds, scan, brute,
                                                    (order_table(scan), orient_table(scan)))
    out = []
    ent_all = layout.tolist()
    for gi in range(len(pools)):
        names = pool_names[gi]
        out.append([(names[e >> 1], "-" if (e & 1) else "+") for e in ent_all[grp_ptr[gi]:grp_ptr[gi + 1]]])
    # work accounting (reference-equivalent candidates / bin pairs) is bookkeeping for the bench, not part of the
    # call: it is queued here and computed when somebody asks (flush_stats)
    _PENDING_STATS.append((pool_lens, n_scaffolds, scan, np.array(rounds, copy=True), np.array(stats[:, 1], copy=True)))
    return out, cost, rounds, stats


_PENDING_STATS: list = []


def flush_stats() -> dict:
    """Fold the queued per-call work accounting of order_groups_device into STATS and return it."""
    while _PENDING_STATS:
        pool_lens, n_scaffolds, scan, rounds, exact = _PENDING_STATS.pop(0)
        for gi, lens in enumerate(pool_lens):
            cand, work = reference_work(lens, n_scaffolds, scan, int(rounds[gi]))
            STATS["candidates"] += cand
            STATS["bin_pairs"] += work
            STATS["exact"] = STATS.get("exact", 0) + int(exact[gi])
            Lg = sum(lens)
            STATS["exact_pairs"] = STATS.get("exact_pairs", 0) + int(exact[gi]) * (Lg * (Lg - 1) // 2)
        STATS["calls"] += 1
    return STATS


def order_genome(h: Handle, G: Optional[np.ndarray], bin_ids: Sequence[int],
                 chrom_groups: Sequence[Sequence[Tuple[int, str]]], n_scaffolds: int = 6, scan: int = 5,
                 chrom_subset: Optional[Sequence[int]] = None):
    """orderGenome (:591-628) without plotting.  ``G`` = genome matrix in ``bin_ids`` order (None: already
    resident).  ``chrom_subset`` restricts the work to some groups (multi-GPU sharding: chromosomes are
    independent, :608-611).  Returns {group index: [(name, orientation), ...]}."""
    if G is not None:
        h.p2_upload(G)
    index_of = {int(b): i for i, b in enumerate(bin_ids)}
    out = {}
    for ci, group in enumerate(chrom_groups):
        if chrom_subset is not None and ci not in chrom_subset:
            continue
        out[ci] = order_chromosome(h, index_of, group, n_scaffolds, scan)
    return out


# --------------------------------------------------------------------------- #
# seams with the reference signatures (duck-typed on orderGenome.Scaffold)
# --------------------------------------------------------------------------- #


class _Seams:
    """Callable replacements for the module-level functions of orderGenome.py; the launcher binds them
    into the loaded reference module.  They operate on the reference's own Scaffold objects
    (``name``, ``binList``, ``orientation``, ``flipOrientation()``).

    Two levels are rebound.  ``orderGenome`` / ``orderChromosome`` (:591-628, :551-586) are the fast lane: the
    genome matrix goes to the device once and every chromosome group runs its brute force, insertion steps and
    window scans in ONE device-resident call (``order_groups_device`` - the engine bench.py measures).  The
    inner functions (``bruteForceBestScore``, ``checkAllScores``, ``scanOrdering``, ``costFunction_numba``) stay
    rebound as well for callers that drive them one by one (the notebook does)."""

    def __init__(self, h: Handle, module=None):
        self.h = h
        self.mod = module          # the loaded reference module (Scaffold, initiateBinsAndScaffolds, plotModule)
        self._genome = None        # (weakref to the host matrix object, shape) that is resident as the P2 matrix

    # ---- residency of the genome matrix: identity of the host object, never a bare address
    def _upload_genome(self, matrix):
        import weakref
        cur = self._genome
        if cur is not None and cur[0]() is matrix:
            return
        M = np.ascontiguousarray(np.asarray(matrix, dtype=np.float64))
        self.h.p2_upload(M)
        try:
            self._genome = (weakref.ref(matrix),)
        except TypeError:          # a list of lists: uploaded, not remembered
            self._genome = None

    def _forget_genome(self):
        self._genome = None

    # orderGenome(matrix, chromList, binList, resolution, nScaffolds=6, scanScaffolds=5, plotChrom=True, ...) :591-628
    def orderGenome(self, matrix, chromList, binList, resolution, nScaffolds=6, scanScaffolds=5, plotChrom=True,
                    showPlot=True, savePlotDir=False, plotTitleSuffix=False):
        import time
        t0 = time.time()
        mod = self.mod
        self._upload_genome(matrix)
        ids = np.fromiter((b.ID for b in binList), dtype=np.int64, count=len(binList))
        lut = np.full(int(ids.max()) + 2 if len(ids) else 2, -1, dtype=np.int64)
        lut[ids] = np.arange(len(ids))
        groups = [[(int(b), nm) for b, nm in grp] for grp in chromList]
        live = [i for i, g in enumerate(groups) if g]
        orderings, costs, rounds, _ = order_groups_device(self.h, lut, [groups[i] for i in live], int(nScaffolds),
                                                          int(scanScaffolds))
        by_group = dict(zip(live, zip(orderings, costs, rounds)))
        fullGenomeOrder = []
        noop_plot = mod is None or getattr(mod.plotModule.plotContactMap, "_hica_noop", False)
        for i, chromGroup in enumerate(chromList):
            print("#####################" + "\n" + "#####################")
            print("Working on Chr_" + str(i + 1) + "...")
            scaffoldList, scaffoldDict = mod.initiateBinsAndScaffolds(chromGroup)
            chromOrder = []
            if i in by_group:
                got, cost, rnd = by_group[i]
                for name, o in got:
                    sObj = scaffoldDict[name]
                    if sObj.orientation != o:
                        sObj.flipOrientation()
                    chromOrder.append(sObj)
                print("Best cost at the end of the final step = " + str(float(cost)) + " (" + str(int(rnd)) + " scan rounds)")
            print("Final ordering...")
            for sObj in chromOrder:
                print(sObj.name, sObj.orientation)
            fullGenomeOrder.append(chromOrder)
            if plotChrom != True or noop_plot:  # noqa: E712
                continue
            chrName = "Chr_" + str(i + 1)
            adjMat, orderDict = mod.giveNewAdjMat(np.asmatrix(matrix), chromOrder, binList)
            mod.plotModule.plotContactMap(adjMat, resolution=resolution, tickCount=11, highlightChroms=False,
                                          wInches=24, hInches=24, lP=1, hP=98, reverseColorMap='',
                                          showPlot=showPlot, savePlot=savePlotDir + "/" + chrName + ".png",
                                          title=chrName, titleSuffix=plotTitleSuffix)
        print("RunTime for total genome with plotting and saving .pngs = " + str(time.time() - t0))
        return fullGenomeOrder

    # orderChromosome(chromGroup, matrix, binList, nScaffolds=6, scanScaffolds=5) :551-586
    def orderChromosome(self, chromGroup, matrix, binList, nScaffolds=6, scanScaffolds=5):
        self._upload_genome(matrix)
        lut_ids = np.fromiter((b.ID for b in binList), dtype=np.int64, count=len(binList))
        lut = np.full(int(lut_ids.max()) + 2, -1, dtype=np.int64)
        lut[lut_ids] = np.arange(len(lut_ids))
        got, cost, rounds, _ = order_groups_device(self.h, lut, [[(int(b), nm) for b, nm in chromGroup]],
                                                   int(nScaffolds), int(scanScaffolds))
        scaffoldList, scaffoldDict = self.mod.initiateBinsAndScaffolds(chromGroup)
        out = []
        for name, o in got[0]:
            sObj = scaffoldDict[name]
            if sObj.orientation != o:
                sObj.flipOrientation()
            out.append(sObj)
        print("Best cost at the end of the final step = " + str(float(cost[0])))
        print("Final ordering...")
        for s in out:
            print(s.name, s.orientation)
        return out

    # giveNewAdjMat(matrix, scaffList, binList) :297-308 - the N x N gather is deferred: after orderGenome it only
    # feeds the whole-genome plot (:700-704)
    def giveNewAdjMat(self, matrix, scaffList, binList):
        from .lazy import LazyMatrix
        nList = [n for sObj in scaffList for n in sObj.binList]
        orderDict = {gID: i for i, gID in enumerate(nList)}
        binIndices = {b.ID: i for i, b in enumerate(binList)}
        indices = [binIndices[n] for n in nList]
        src = matrix
        return LazyMatrix((len(indices), len(indices)),
                          lambda: np.asmatrix(np.asarray(src)[np.ix_(indices, indices)]), None, "giveNewAdjMat"), orderDict

    # costFunction_numba(matrix, total) :184-191
    def costFunction_numba(self, matrix, total):
        M = np.ascontiguousarray(np.asarray(matrix, dtype=np.float64))
        n = M.shape[0]
        self.h.p2_upload(M)
        self._forget_genome()
        self.h.p2_select(np.arange(n, dtype=np.int32))
        _, _, sc = self.h.p2_score_orders(np.arange(n, dtype=np.int32)[None, :], float(total), 0.0, want_scores=True)
        return float(sc[0])

    @staticmethod
    def _plus_bins(sObj):
        return list(sObj.binList) if sObj.orientation == "+" else list(sObj.binList)[::-1]

    @staticmethod
    def _set_orientation(sObj, o):
        if sObj.orientation != o:
            sObj.flipOrientation()

    def _resident_matrix(self, matrix, orderDict, sObjs) -> DeviceChromosome:
        M = np.ascontiguousarray(np.asarray(matrix, dtype=np.float64))
        self.h.p2_upload(M)
        self._forget_genome()
        scafs = [Scaf(s.name, self._plus_bins(s), s.orientation) for s in sObjs]
        # every row of `matrix` is addressed through orderDict (binID -> row)
        return DeviceChromosome(self.h, dict(orderDict), scafs), scafs

    # bruteForceBestScore(sObjList, scaffDict, matrix, orderDict) :432-473
    def bruteForceBestScore(self, sObjList, scaffDict, matrix, orderDict):
        chrom, scafs = self._resident_matrix(matrix, orderDict, sObjList)
        names = [s.name for s in sObjList]
        od, orr, cost = chrom.brute_force(scafs)
        n = len(names)
        # leave the shared Scaffold objects in the state of the LAST candidate, as the reference's loop does
        last_o, last_r = order_table(n)[-1], orient_table(n)[-1]
        if cost != 0.0 or chrom.total(scafs) != 0:
            for j in range(n):
                self._set_orientation(scaffDict[names[int(last_o[j])]], "-" if last_r[j] else "+")
        return [names[i] for i in od], list(orr), cost

    # checkAllScores(adjMat, orderDict, orderedScaffs, scaffToCheck) :332-372
    def checkAllScores(self, adjMat, orderDict, orderedScaffs, scaffToCheck):
        chrom, scafs = self._resident_matrix(adjMat, orderDict, list(orderedScaffs) + [scaffToCheck])
        ordered, new = scafs[:-1], scafs[-1]
        ordered, best = chrom.insert_best(ordered, new)
        pos = ordered.index(new)
        self._set_orientation(scaffToCheck, new.orientation)
        orderedScaffs.insert(pos, scaffToCheck)
        return orderedScaffs, best

    # scanOrdering(orderedScaffolds, scaffoldDict, orderDict, matrix, binList, bestCost, scanScaffolds=5) :495-549
    def scanOrdering(self, orderedScaffolds, scaffoldDict, orderDict, matrix, binList, bestCost, scanScaffolds=5):
        self._upload_genome(matrix)
        index_of = {b.ID: i for i, b in enumerate(binList)}
        scafs = [Scaf(s.name, self._plus_bins(s), s.orientation) for s in orderedScaffolds]
        chrom = DeviceChromosome(self.h, index_of, scafs)
        ordered, best, rounds = chrom.scan(scafs, float(bestCost), int(scanScaffolds))
        out = []
        for s in ordered:
            ref = scaffoldDict[s.name]
            self._set_orientation(ref, s.orientation)
            out.append(ref)
        print("Sliding window conversion after " + str(rounds) + " rounds")
        print("Best cost at the end of the final step = " + str(best))
        return out, best
