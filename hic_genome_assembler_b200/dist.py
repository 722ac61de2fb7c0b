"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

What shards
  * Part2 - chromosome groups are independent (orderGenome.py:608-611): groups are dealt to ranks largest
    first (``shard_groups``); there is NO data-path collective, only one small ``all_gather_object`` of the
    finished orderings.
  * A single very large window search (bruteForceBestScore with nScaffolds = 8: 5.16 M candidates) can be
    split by candidate range; the winner is the first strict maximum over all ranks, i.e. an arg-max
    all-reduce on (score, -candidate index) (``reduce_first_max``).
  * Part1's linkage is one sequential chain and runs on one GPU; at <= 50 kb the whole of Part1 is a fraction
    of a second, so every rank runs it redundantly (deterministic, bit-identical) instead of exchanging the
    working matrix.  Row-block sharding of the distance/rank/count kernels for the 10 kb configuration is not
    built yet.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def shard_groups(group_sizes: Sequence[int], world: int, rank: int) -> List[int]:
    """Indices of the groups owned by ``rank``: greedy longest-processing-time assignment (largest group to
    the least loaded rank, ties to the lowest rank), so every rank computes the same partition."""
    order = sorted(range(len(group_sizes)), key=lambda i: (-int(group_sizes[i]), i))
    load = [0] * world
    owner = [0] * len(group_sizes)
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += int(group_sizes[i]) ** 2   # the scan cost of a group grows with L^2
    return [i for i in range(len(group_sizes)) if owner[i] == rank]


def merge_orderings(local: Dict[int, object], n_groups: int, group=None) -> List[object]:
    """all_gather_object of {group index: result}; returns the list in group order on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        parts = [local]
    else:
        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, local, group=group)
    merged: Dict[int, object] = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise RuntimeError("group %d was ordered by two ranks" % k)
            merged[k] = v
    missing = [i for i in range(n_groups) if i not in merged]
    if missing:
        raise RuntimeError("groups %s were not ordered by any rank" % missing)
    return [merged[i] for i in range(n_groups)]


def reduce_first_max(best: float, idx: int, carried: float, device=None, group=None) -> Tuple[float, int]:
    """First strict maximum above ``carried`` over all ranks: the highest score wins, ties go to the lowest
    candidate index (the sequential accept rule of orderGenome.py:464).  ``idx`` = -1 when the rank found
    nothing above ``carried``."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return (best, idx) if idx >= 0 else (carried, -1)
    world = dist.get_world_size(group)
    mine = torch.tensor([best if idx >= 0 else float("-inf"), float(idx)], dtype=torch.float64, device=device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    win_b, win_i = carried, -1
    for p in parts:
        b, i = float(p[0]), int(p[1])
        if i < 0:
            continue
        if b > win_b or (b == win_b and win_i >= 0 and i < win_i):
            win_b, win_i = b, i
    return win_b, win_i


def order_genome_distributed(h, genome_index_of_bin: Dict[int, int], chrom_groups, n_scaffolds: int = 6,
                             scan: int = 5, group=None):
    """orderGenome over all ranks: every rank orders its shard of the chromosome groups on its own GPU
    (``part2.order_groups_device``), then the orderings are exchanged.  Returns the list of
    ``([(name, orientation), ...], cost, rounds)`` in group order on every rank."""
    import torch.distributed as dist
    from . import part2
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    mine = shard_groups([len(g) for g in chrom_groups], world, rank)
    local: Dict[int, object] = {}
    if mine:
        got, cost, rounds, _ = part2.order_groups_device(h, genome_index_of_bin, [chrom_groups[i] for i in mine],
                                                         n_scaffolds, scan)
        for k, gi in enumerate(mine):
            local[gi] = (got[k], float(cost[k]), int(rounds[k]))
    return merge_orderings(local, len(chrom_groups), group)


def brute_force_sharded(h, chrom, layout, carried: float = 0.0, device=None, group=None):
    """bruteForceBestScore (:432-473) with the candidate range split over the ranks (every rank holds the same
    chromosome): returns (order idx tuple, orientation tuple, cost) on every rank."""
    import torch.distributed as dist
    from . import part2
    n = len(layout)
    orders, orients = part2.order_table(n), part2.orient_table(n)
    B = orders.shape[0] * orients.shape[0]
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    per = (B + world - 1) // world
    lo, hi = min(B, rank * per), min(B, (rank + 1) * per)
    total = chrom.total(layout)
    base = [s.entry() for s in layout]
    best, idx = carried, -1
    if hi > lo:
        best, idx, _ = h.p2_score_window(base, 0, n, orders, orients, total, carried, cand_lo=lo, cand_hi=hi)
    best, idx = reduce_first_max(best, idx, carried, device=device, group=group)
    if idx < 0:
        return None, None, carried
    oi, ri = divmod(idx, orients.shape[0])
    return tuple(int(v) for v in orders[oi]), tuple("-" if v else "+" for v in orients[ri]), best
