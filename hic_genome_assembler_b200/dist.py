"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

What shards
  * Part2 - chromosome groups are independent (orderGenome.py:608-611): groups are dealt to ranks largest
    first (``shard_groups``); there is NO data-path collective, only one small ``all_gather_object`` of the
    finished orderings.
  * A single very large window search (bruteForceBestScore with nScaffolds = 8: 5.16 M candidates) can be
    split by candidate range; the winner is the first strict maximum over all ranks, i.e. an arg-max
    all-reduce on (score, -candidate index) (``reduce_first_max``).
  * Part1 (``run_part1_sharded``) - the contact matrix is cut into contiguous row blocks: every rank uploads
    its own block over its own PCIe link, the full matrix is assembled on every GPU with an in-place
    all-gather over NVLink (the one data-path collective of "assembling the distance matrix"), the sequential
    linkage chain runs on ONE GPU (rank 0) and the leaf order is broadcast (N int32), the similarity rank
    matrix is computed by row block and all-gathered, and the cut searches (launch-latency bound, 10 ms) run
    replicated on every rank - deterministic, so every rank ends with the same cut indices without a
    collective per break search.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def row_blocks(n: int, world: int, balance_upper_triangle: bool = False, align: int = 1) -> List[Tuple[int, int]]:
    """Contiguous row blocks [r0, r1) of an n-row matrix, one per rank (possibly empty for world > n).
    ``balance_upper_triangle``: the blocks hold equal shares of the UPPER TRIANGLE (what a symmetric upload
    sends over PCIe: row i contributes n - i entries) instead of equal row counts; boundaries are rounded to a
    multiple of ``align`` rows."""
    bounds = [0]
    for k in range(1, world):
        if balance_upper_triangle:
            # rows [0, b) hold b*n - b(b-1)/2 ~ n^2/2 * (1 - (1 - b/n)^2) entries of the upper triangle
            b = n * (1.0 - (1.0 - k / world) ** 0.5)
        else:
            b = n * k / world
        b = int(round(b / align)) * align
        bounds.append(min(n, max(bounds[-1], b)))
    bounds.append(n)
    return [(bounds[k], bounds[k + 1]) for k in range(world)]


class _DevBuf:
    """A device buffer of libhica as a zero-copy torch tensor (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(h, which: int, device):
    """The handle's matrix (which = 0, fp64 n x n) or rank matrix (which = 1) as a torch tensor that aliases
    the library's device memory - the buffer the collectives of ``run_part1_sharded`` run on."""
    import torch
    ptr, nbytes, eb = h.p1_device_buffer(which)
    n = h.n
    if which == 0:
        shape, typestr, dtype = (n, n), "<f8", torch.float64
    else:
        # uint16 / int32 ranks as raw bytes: the collective only moves them (NCCL has no 16-bit integer type)
        shape, typestr, dtype = (n, n * eb), "|u1", torch.uint8
    t = torch.as_tensor(_DevBuf(ptr, shape, typestr), device=device)
    assert t.dtype == dtype and t.data_ptr() == ptr
    return t


def gather_row_blocks(full, blocks: Sequence[Tuple[int, int]], group=None):
    """In-place all-gather of the row blocks of ``full`` (a 2-D tensor whose rows ``blocks[r]`` are valid on
    rank r): afterwards every rank holds every row.  Equal blocks go through one ``all_gather_into_tensor``
    (NCCL in-place all-gather over NVLink); ragged blocks through one broadcast per rank.  Works on CPU tensors
    with the gloo backend (tests)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return full
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    assert len(blocks) == world
    sizes = {r1 - r0 for r0, r1 in blocks}
    r0, r1 = blocks[rank]
    if len(sizes) == 1 and blocks[0][0] == 0 and blocks[-1][1] == full.shape[0] and full.is_cuda:
        dist.all_gather_into_tensor(full, full[r0:r1], group=group)
    else:
        works = []
        for src, (b0, b1) in enumerate(blocks):
            if b1 > b0:
                works.append(dist.broadcast(full[b0:b1], src=dist.get_global_rank(group, src) if group is not None else src,
                                            group=group, async_op=True))
        for w in works:
            w.wait()
    return full


def _dist_state(group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def upload_sharded(h, A_host: np.ndarray, device, symmetric: bool = True, group=None) -> dict:
    """Row-block sharded upload: this rank copies only its row block of ``A_host`` (the full n x n matrix in
    the pinned host memory of every rank) over its own PCIe link; the matrix is then assembled on every GPU by
    an in-place gather over NVLink, and the lower triangle mirrored on the device for a symmetric upload (then
    only the part right of the diagonal crossed PCIe, the blocks being balanced by upper-triangle area)."""
    import time
    import torch
    world, rank = _dist_state(group)
    if world == 1:
        (h.p1_upload_symmetric if symmetric else h.p1_upload)(A_host)
        return {"row_block": (0, A_host.shape[0])}
    n = A_host.shape[0]
    t0 = time.perf_counter()
    blocks = row_blocks(n, world, balance_upper_triangle=symmetric, align=512 if symmetric and n >= 512 * world else 1)
    r0, r1 = blocks[rank]
    h.p1_alloc(n)
    h.p1_upload_rows(A_host[r0:r1], r0, r1, symmetric)
    t1 = time.perf_counter()
    A = device_tensor(h, 0, device)
    gather_row_blocks(A, blocks, group)
    torch.cuda.current_stream(device).synchronize()
    if symmetric:
        h.p1_mirror_lower()
    t2 = time.perf_counter()
    return {"row_block": (r0, r1), "upload_rows_ms": (t1 - t0) * 1e3, "gather_matrix_ms": (t2 - t1) * 1e3,
            "h2d_bytes": int(sum(8 * min(512, r1 - b0) * (n - b0) for b0 in range(r0, r1, 512))) if symmetric
            else 8 * (r1 - r0) * n}


def run_part1_resident_sharded(h, device, min_size: int = 5, min_frac: float = .05, psig: float = .05,
                               linkage_rank: int = 0, group=None) -> dict:
    """Part1 on a matrix that is resident on every GPU of the group (``upload_sharded`` / ``p1_upload_device``):
    row sums + distance transform everywhere (1 ms; the rank step needs the row sums on every GPU), the
    sequential linkage chain on ONE GPU with its leaf order broadcast (N int32), the similarity rank matrix by
    row block of the reordered matrix + in-place all-gather, the cut searches replicated (deterministic).
    Returns the dict of ``part1.run_part1_resident`` on every rank."""
    import time
    import torch
    import torch.distributed as dist
    from . import part1
    world, rank = _dist_state(group)
    if world == 1:
        return part1.run_part1_resident(h, min_size, min_frac, psig)
    n = h.n
    t2 = time.perf_counter()
    h.p1_set_row_sums(None)
    h.p1_distance()
    leaves_t = torch.empty(n, dtype=torch.int32, device=device)
    leaves = None
    # n <= 65535: every GPU sorts all rows by value in matrix order while the linkage chain runs (the linkage GPU
    # on its 132 free SMs - hica_p1_linkage starts that itself - the others while they wait for the broadcast);
    # the rank step is then a local column relabel and needs no gather.  (If the library declines the pre-sort,
    # p1_rank simply sorts all rows in one step on every GPU: still correct, no gather either.)
    presort = n <= 65535 and os.environ.get("HICA_RANK_OVERLAP", "1") != "0" and (
        n <= 16384 or os.environ.get("HICA_RANK_OVERLAP_CHUNKED", "1") != "0")
    if presort and rank != linkage_rank:
        h.p1_rank_presort()
    if rank == linkage_rank:
        _, leaves = h.p1_linkage(want_Z=False)
        leaves_t.copy_(torch.from_numpy(np.ascontiguousarray(leaves, dtype=np.int32)))
    src = dist.get_global_rank(group, linkage_rank) if group is not None else linkage_rank
    dist.broadcast(leaves_t, src=src, group=group)
    if rank != linkage_rank:
        leaves = leaves_t.cpu().numpy()
        h.p1_set_order(leaves)
    t3 = time.perf_counter()
    if presort:
        h.p1_rank()
    else:
        rb = row_blocks(n, world)
        h.p1_rank_rows(*rb[rank])
        R = device_tensor(h, 1, device)
        gather_row_blocks(R, rb, group)
        torch.cuda.current_stream(device).synchronize()
        h.p1_rank_commit()
    t4 = time.perf_counter()
    init, amb0 = h.p1_breakpoints(min_size, min_frac)
    cuts, amb1 = h.p1_filter(init, psig)
    t5 = time.perf_counter()
    tie = h.tie_report()
    return dict(Z=None, leaves=np.asarray(leaves), initial_cuts=init, cuts=cuts, ambiguous=amb0 + amb1,
                tie_sensitive=(tie["breakpoints"] + tie["filter"]) if tie["available"] else None,
                stage_ms={"distance_linkage_bcast": (t3 - t2) * 1e3, "rank_rows_gather": (t4 - t3) * 1e3,
                          "cuts": (t5 - t4) * 1e3})


def run_part1_sharded(h, A_host: np.ndarray, device, min_size: int = 5, min_frac: float = .05, psig: float = .05,
                      symmetric: bool = True, linkage_rank: int = 0, group=None) -> dict:
    """Part1 over all ranks of the process group: ``upload_sharded`` + ``run_part1_resident_sharded``."""
    up = upload_sharded(h, A_host, device, symmetric, group)
    res = run_part1_resident_sharded(h, device, min_size, min_frac, psig, linkage_rank, group)
    res["upload"] = up
    return res


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
