"""Batch forms of the Part2 oracle for golden vectors at the published sizes.  TEST INFRASTRUCTURE ONLY.

``hica_oracle.order_chromosome`` follows orderGenome.py:551-586 one candidate at a time through Python loops,
which is fine for the small cases of the unit tests and hopeless at 100 kb (1,920 candidates x 36 windows x
rounds x 24 groups at L = 800).  The functions below enumerate exactly the same candidates in exactly the same
order, build their bin orders as int32 rows and hand each batch to the C restatement of the score
(``hica_o_score_batch``: numba's summation order, first strict maximum above the carried cost -
orderGenome.py:184-191, :349, :464, :535).  ``tests/test_oracle_fast.py`` pins them against the one-at-a-time
oracle; ``tests/golden/make_golden_large.py`` uses them.

Only tests/ and the golden generators import this file.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

import c_oracle as CO
import hica_oracle as O


class FScaf:
    """A scaffold as (name, bins in '+' reading order, current orientation)."""

    __slots__ = ("name", "plus", "orientation")

    def __init__(self, name, plus_bins, orientation="+"):
        self.name = name
        self.plus = np.asarray(plus_bins, dtype=np.int64)
        self.orientation = orientation

    def bins(self, o=None):
        o = self.orientation if o is None else o
        return self.plus if o == "+" else self.plus[::-1]


def _flip(o):
    return "-" if o == "+" else "+"


def _sub(G, index_of, scafs):
    """giveNewAdjMat (orderGenome.py:297-308) for the scaffolds in their current state."""
    bins = np.concatenate([s.bins() for s in scafs])
    idx = index_of[bins]
    M = np.ascontiguousarray(G[np.ix_(idx, idx)])
    local = np.full(int(bins.max()) + 1, -1, dtype=np.int64)
    local[bins] = np.arange(len(bins))
    return M, local


def _rows(cands: Sequence[Sequence[Tuple[FScaf, str]]], local) -> np.ndarray:
    return np.asarray([np.concatenate([local[s.bins(o)] for s, o in c]) for c in cands], dtype=np.int32)


def brute_force(scafs: List[FScaf], M, local):
    """bruteForceBestScore (orderGenome.py:432-473)."""
    n = len(scafs)
    orders, orients = O.order_candidates(n), O.orient_candidates(n)
    total = CO.total_weight(M)
    if total == 0:
        return orders[0], orients[0], 0.0
    best, best_idx = 0.0, -1
    # one batch per order keeps the candidate matrix small (2^n rows)
    for oi, od in enumerate(orders):
        rows = _rows([[(scafs[si], o) for si, o in zip(od, orr)] for orr in orients], local)
        b, bi, _ = CO.score_batch(M, rows, total, best, want_scores=False)
        if bi >= 0:
            best, best_idx = b, oi * len(orients) + bi
    if best_idx < 0:
        return None, None, 0.0
    oi, ri = divmod(best_idx, len(orients))
    return orders[oi], orients[ri], best


def check_all_scores(M, local, ordered: List[FScaf], new: FScaf):
    """checkAllScores (orderGenome.py:332-372): (0,o)(0,-o)(1,-o)(1,o)(2,o)... - the scaffold is popped still
    flipped, so the first orientation tried alternates with the insertion index."""
    total = CO.total_weight(M)
    k = len(ordered)
    cands, tried = [], []
    o0 = new.orientation
    for i in range(k + 1):
        first = o0 if i % 2 == 0 else _flip(o0)
        for o in (first, _flip(first)):
            cands.append([(s, s.orientation) for s in ordered[:i]] + [(new, o)] + [(s, s.orientation) for s in ordered[i:]])
            tried.append((i, o))
    best, bi, _ = CO.score_batch(M, _rows(cands, local), total, 0.0, want_scores=False)
    pos, orient = (0, "+") if bi < 0 else tried[bi]
    new.orientation = orient
    ordered.insert(pos, new)
    return ordered, (best if bi >= 0 else 0.0)


def scan_ordering(G, index_of, ordered: List[FScaf], best_cost: float, scan: int = 5):
    """scanOrdering (orderGenome.py:495-549): `total` from the entry ordering, carried best cost, the matrix is
    re-gathered after every window, rounds to a fixed point."""
    M, local = _sub(G, index_of, ordered)
    total = CO.total_weight(M)
    orders, orients = O.order_candidates(scan), O.orient_candidates(scan)
    rounds = 0
    while True:
        improved = False
        for i in range(0, len(ordered) - scan + 1):
            window = ordered[i:i + scan]
            head = [(s, s.orientation) for s in ordered[:i]]
            tail = [(s, s.orientation) for s in ordered[i + scan:]]
            cands = [head + [(window[wi], o) for wi, o in zip(od, orr)] + tail for od in orders for orr in orients]
            b, bi, _ = CO.score_batch(M, _rows(cands, local), total, best_cost, want_scores=False)
            if bi >= 0:
                best_cost = b
                improved = True
                for s, o in cands[bi]:
                    s.orientation = o
                ordered = [s for s, _ in cands[bi]]
            # (the reference re-gathers the matrix in the new order here, :540-541; candidates are expressed in
            #  bins, so the scores do not depend on how M is arranged and the gather is skipped)
        rounds += 1
        if not improved:
            break
    return ordered, best_cost, rounds


def order_chromosome(G: np.ndarray, index_of: np.ndarray, nodes: Sequence[Tuple[int, str]],
                     n_scaffolds: int = 6, scan: int = 5):
    """orderChromosome (orderGenome.py:551-586).  ``index_of[bin id]`` = row of ``G``; ``nodes`` =
    [(binID, scaffoldName), ...] or the compact form [(name, sorted bin-id array), ...].
    Returns ([(name, orientation), ...], best cost, scan rounds)."""
    if n_scaffolds >= 9:
        n_scaffolds = 8
    if scan > n_scaffolds:
        scan = n_scaffolds
    if nodes and isinstance(nodes[0][1], np.ndarray):
        pool = [FScaf(nm, np.sort(ids)) for nm, ids in nodes]
    else:
        by_name: Dict[str, list] = {}
        for b, nm in nodes:
            by_name.setdefault(nm, []).append(int(b))
        pool = [FScaf(nm, sorted(v)) for nm, v in by_name.items()]
    pool.sort(key=lambda s: len(s.plus), reverse=True)     # stable, as initiateBinsAndScaffolds :279
    ordered, pool = pool[:n_scaffolds], pool[n_scaffolds:]
    M, local = _sub(G, index_of, ordered)
    od, orr, cost = brute_force(ordered, M, local)
    if od is None:
        raise RuntimeError("no ordering scored above 0 (the reference fails on 'NA' here)")
    for si, o in zip(od, orr):
        ordered[si].orientation = o
    ordered = [ordered[si] for si in od]
    best_cost = cost
    while True:                                            # do-while, orderGenome.py:475-493
        if pool:
            ordered.append(pool.pop(0))
        M, local = _sub(G, index_of, ordered)
        new = ordered.pop(-1)
        ordered, best_cost = check_all_scores(M, local, ordered, new)
        if not pool:
            break
    rounds = 0
    if len(ordered) > n_scaffolds:
        ordered, best_cost, rounds = scan_ordering(G, index_of, ordered, best_cost, scan)
    return [(s.name, s.orientation) for s in ordered], best_cost, rounds


# --------------------------------------------------------------------------- #
# Part1 at large N: the same numpy/scipy calls as hica_oracle, with the per-row Python loops of the window
# counts replaced by the C restatement (bitwise pinned against hica_oracle.window_counts in the tests)
# --------------------------------------------------------------------------- #


def find_breakpoints(R32: np.ndarray, start: int, min_size: int, world_size: int, psig: float = .05) -> List[int]:
    """hica_oracle.find_breakpoints with C window counts."""
    M = world_size
    ws = min_size
    break_sig = 0
    loop_count = 0
    x = CO.window_counts(R32, start)
    while True:
        while True:
            sig = O.window_significance(x, M, psig)
            loop_count += 1
            if (int(sig.sum()) / len(sig)) >= .9:
                M = int(M - start)
            else:
                break_sig = 1
            if break_sig == 1 or loop_count >= 5:
                break
        b = O.sliding_break_signal(sig, ws)
        hits = [] if b is None else [int(t) + min_size for t in np.nonzero(b == min_size)[0]]
        if hits:
            return hits
        ws -= 1
        if ws == 0:
            return []


def pre_process_breakpoints(R32: np.ndarray, min_size: int = 5, min_frac: float = .05) -> List[int]:
    N = len(R32)
    stop_ind = int(N - (N * min_frac))
    ind = 0
    cinds: List[int] = []
    if min_frac == 1:
        return cinds
    while True:
        cuts = find_breakpoints(R32, ind, min_size, N - ind, psig=.05)
        if len(cuts) == 0:
            break
        ind += cuts[0]
        cinds.append(int(ind))
        if ind >= stop_ind or (N - ind) <= min_size:
            break
    return cinds
