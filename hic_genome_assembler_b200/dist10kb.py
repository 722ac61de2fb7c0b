"""The 10 kb configuration (BASELINE.json configs[4]): 160,000 bins, fp32 matrix (102 GB) sharded by row block over the
GPUs of one box, linkage on one GPU.  One process per GPU, torch.distributed (NCCL) for the plumbing.

    stage                         where                        exchange
    ---------------------------   --------------------------   ---------------------------------------------------
    row sums (fp64 accumulation)  every GPU, own rows          all-gather of the N sums (1.3 MB)
    distance transform -> rows    every GPU, own rows          written straight into the linkage GPU's working
      of the mirrored matrix W                                 matrix over NVLink (peer mapping, CUDA IPC): one
                                                               kernel does the transform AND the gather
    UPGMA chain + leaf order      the linkage GPU              broadcast of the leaf order (N int32)

replacing convertMatrix(distance=True) (scaffoldToChromosomes.py:138-147), squareform (:194), hierarchy.average
(:197) and dendrogram (:204) for a matrix the reference cannot hold.  When peer mappings are not
available the row blocks of W are built locally and sent with NCCL instead (``transport`` in the result says which).

The synthetic matrix is generated IN PLACE by row block with the portable generator of :mod:`.synth` (fp32): the
values do not depend on how many GPUs share the work, so runs at 2 / 4 / 8 GPUs can be compared by digest.
"""
from __future__ import annotations

import hashlib
import json
import os
import time
from typing import List, Optional, Tuple

import numpy as np


class _DevBuf:
    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


def row_blocks(n: int, world: int, align: int = 4) -> List[Tuple[int, int]]:
    bounds = [min(n, int(round(n * k / world / align)) * align) for k in range(world)] + [n]
    return [(bounds[k], bounds[k + 1]) for k in range(world)]


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_world_size(), dist.get_rank()
    return None, 1, 0


def gather_vector(vec, blocks, dist_mod):
    """In-place all-gather of a vector whose block ``blocks[r]`` is valid on rank r (ragged blocks: one broadcast
    per rank; works with gloo on CPU tensors too)."""
    if dist_mod is None:
        return vec
    works = [dist_mod.broadcast(vec[b0:b1], src=r, async_op=True) for r, (b0, b1) in enumerate(blocks) if b1 > b0]
    for w in works:
        w.wait()
    return vec


class ShardLayout:
    """Ground truth of the synthetic genome (chromosomes, scaffolds in file order) - the same on every rank."""

    def __init__(self, n_bins: int, n_chrom: int, seed: int, scaffolds_per_chrom: int = 40):
        from .synth import _scaffold_layout
        rng = np.random.RandomState(seed)
        base = n_bins // n_chrom
        chrom_len = np.full(n_chrom, base, dtype=np.int64)
        chrom_len[: n_bins - base * n_chrom] += 1
        self.chrom_of = np.repeat(np.arange(n_chrom), chrom_len)
        self.pos_in_chrom = np.concatenate([np.arange(l) for l in chrom_len])
        res = int(round(1.6e9 / n_bins))
        self.perm, self.scaffold_of_bin, _, _, self.names = _scaffold_layout(rng, chrom_len, scaffolds_per_chrom, res)
        self.true_chrom_of_bin = self.chrom_of[self.perm]
        self.n, self.seed = n_bins, seed


def generate_shard(M, layout: ShardLayout, r0: int, r1: int, blocks, *, noise_sigma: float = 0.15, ice_sweeps: int = 50,
                   row_sum_target: float = 1000.0, dist_mod=None, distance_scale: Optional[float] = None):
    """Fill ``M`` (a torch tensor, (r1-r0) x n, fp32 or fp64, any device) with rows [r0, r1) of the portable
    synthetic ICE matrix.  Every operation is row-local except the row sums, which are all-gathered once per
    ICE sweep; the result is bit-identical for every number of ranks."""
    import torch
    from .synth import portable_tile, tree_row_sums, _host_mean
    n = layout.n
    dev = M.device
    fdtype = M.dtype
    if distance_scale is None:
        # contacts decay with GENOMIC distance: one bin of a finer map is a shorter stretch of DNA.  SURVEY 8d's
        # recipe counts in bins and was tuned at 100 kb (16,000 bins); finer maps scale the bin distance back to
        # 100 kb units, otherwise the two halves of a 8,000-bin chromosome touch each other less than the 0.3
        # inter-chromosome background and no clustering can keep the chromosome together.
        distance_scale = min(1.0, 16000.0 / n)
    conv = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64), device=dev)  # noqa: E731
    t_all, c_of, p_of = conv(layout.perm), conv(layout.chrom_of), conv(layout.pos_in_chrom)
    rows = r1 - r0
    blk = max(1, min(max(rows, 1), (1 << 24) // n))
    for b0 in range(0, rows, blk):
        b1 = min(rows, b0 + blk)
        M[b0:b1] = portable_tile(t_all[r0 + b0:r0 + b1], t_all, c_of, p_of, layout.seed, noise_sigma, fdtype, torch,
                                 dscale=distance_scale)
    s = torch.zeros(n, dtype=torch.float64, device=dev)

    def sums():
        # the halving tree in slabs of rows (the tree is per row, so slabs change nothing but the peak memory)
        slab = max(1, (1 << 28) // n)
        for b0 in range(0, rows, slab):
            b1 = min(rows, b0 + slab)
            s[r0 + b0:r0 + b1] = tree_row_sums(M[b0:b1], torch).to(torch.float64)
        gather_vector(s, blocks, dist_mod)
        return s.cpu().numpy()

    for _ in range(ice_sweeps):
        sh = sums()
        q = np.sqrt(sh / _host_mean(sh))
        qd = torch.as_tensor(q, device=dev).to(fdtype)
        for b0 in range(0, rows, blk):
            b1 = min(rows, b0 + blk)
            M[b0:b1] /= qd[r0 + b0:r0 + b1, None] * qd[None, :]
    sh = sums()
    M *= float(np.asarray(row_sum_target / _host_mean(sh)).astype(np.float32 if fdtype == torch.float32 else np.float64))
    return M


class Shard32:
    """One rank's part of the sharded fp32 Part1 front end (see the module docstring)."""

    def __init__(self, h, n: int, device, linkage_rank: int = 0):
        import torch
        self.h, self.n, self.dev, self.linkage_rank = h, int(n), device, linkage_rank
        self.dist, self.world, self.rank = _dist()
        self.blocks = row_blocks(self.n, self.world)
        self.r0, self.r1 = self.blocks[self.rank]
        self.is_linkage = self.rank == linkage_rank
        h.s32_alloc(self.n, self.r0, self.r1, self.is_linkage)
        self.ld = h.s32[3]
        ptr, _ = h.s32_buffer(0)
        self.A = torch.as_tensor(_DevBuf(ptr, (self.r1 - self.r0, self.n), "<f4"), device=device)
        ptr, _ = h.s32_buffer(2)
        self.sums = torch.as_tensor(_DevBuf(ptr, (self.n,), "<f8"), device=device)
        self.W = None
        if self.is_linkage:
            ptr, _ = h.s32_buffer(1)
            self.W = torch.as_tensor(_DevBuf(ptr, (self.n, self.ld), "<f4"), device=device)
        self.peer_W = 0          # the linkage GPU's working matrix mapped into this process (0: not mapped)
        self.transport = "local" if self.world == 1 else None
        self._stage = None
        if self.world > 1:
            self._connect()

    def _connect(self):
        """Map the linkage GPU's working matrix into every other rank (CUDA IPC); fall back to NCCL sends."""
        import torch
        handle = [self.h.s32_ipc_export(1) if self.is_linkage else None]
        self.dist.broadcast_object_list(handle, src=self.linkage_rank)
        ok = 1
        if not self.is_linkage and os.environ.get("HICA_S32_TRANSPORT", "ipc") == "ipc":
            try:
                self.peer_W = self.h.s32_ipc_open(handle[0])
            except Exception as e:  # noqa: BLE001 - any failure means "no peer mapping here"
                self.peer_W, ok = 0, 0
                self._ipc_error = str(e)
        elif not self.is_linkage:
            ok = 0
        flag = torch.tensor([ok], device=self.dev)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            self.transport = "peer stores over NVLink (CUDA IPC mapping of the linkage GPU's working matrix)"
        else:
            if self.peer_W:
                self.h.s32_ipc_close(self.peer_W)
                self.peer_W = 0
            self.transport = "NCCL send/recv of locally built row blocks"
            if not self.is_linkage:
                self._stage = torch.empty((self.r1 - self.r0, self.ld), dtype=torch.float32, device=self.dev)

    def distance_and_gather(self) -> dict:
        """Row sums -> all-gather -> rows of W into the linkage GPU's working matrix.  Returns stage times (ms)."""
        import torch
        t0 = time.perf_counter()
        self.h.s32_row_sums()
        gather_vector(self.sums, self.blocks, self.dist)
        torch.cuda.current_stream(self.dev).synchronize()
        t1 = time.perf_counter()
        if self.is_linkage or self.world == 1:
            self.h.s32_build_w()
        elif self.peer_W:
            self.h.s32_build_w(self.peer_W, self.ld)
        else:
            self.h.s32_build_w(self._stage.data_ptr(), self.ld, self.r0)
        k1_ms = self.h.timers()["distance"]
        if self.world > 1 and not self.peer_W and self.transport.startswith("NCCL"):
            if self.is_linkage:
                works = [self.dist.irecv(self.W[b0:b1], src=r) for r, (b0, b1) in enumerate(self.blocks)
                         if r != self.rank and b1 > b0]
                for w in works:
                    w.wait()
            elif self.r1 > self.r0:
                self.dist.send(self._stage, dst=self.linkage_rank)
        torch.cuda.synchronize(self.dev)
        if self.dist is not None:
            self.dist.barrier()       # every row block has landed in W
        t2 = time.perf_counter()
        return {"row_sums_gather_ms": (t1 - t0) * 1e3, "distance_and_gather_ms": (t2 - t1) * 1e3, "k1_kernel_ms": k1_ms}

    def linkage(self, want_Z: bool = False):
        """hierarchy.average + leaf order on the linkage GPU, leaf order broadcast.  Returns (Z or None, leaves)."""
        import torch
        Z, leaves = None, None
        if self.is_linkage:
            Z, leaves = self.h.s32_linkage(want_Z=want_Z)
        if self.dist is not None:
            lt = torch.empty(self.n, dtype=torch.int32, device=self.dev)
            if self.is_linkage:
                lt.copy_(torch.from_numpy(np.ascontiguousarray(leaves, dtype=np.int32)))
            self.dist.broadcast(lt, src=self.linkage_rank)
            leaves = lt.cpu().numpy()
        return Z, leaves

    def close(self):
        if self.peer_W:
            self.h.s32_ipc_close(self.peer_W)
            self.peer_W = 0


def chromosome_breaks(leaves, true_chrom_of_bin) -> int:
    """Places where the leaf order passes from one true chromosome to another (19 = every chromosome contiguous)."""
    ch = np.asarray(true_chrom_of_bin)[np.asarray(leaves)]
    return int(np.count_nonzero(np.diff(ch) != 0))


# --------------------------------------------------------------------------- bench.py --config 10kb
def bench(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import bench as B
    from ._lib import Handle
    n = args.bins if args.bins != 16000 else 160000
    dev = torch.device("cuda", local_rank)
    h = Handle(local_rank)
    layout = ShardLayout(n, args.chroms, args.seed)
    t0 = time.perf_counter()
    sh = Shard32(h, n, dev)
    generate_shard(sh.A, layout, sh.r0, sh.r1, sh.blocks, dist_mod=sh.dist,
                   ice_sweeps=int(os.environ.get("HICA_10KB_ICE_SWEEPS", "50")))
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        st = sh.distance_and_gather()
        t1 = time.perf_counter()
        _, leaves = sh.linkage()
        st["linkage_leaforder_bcast_ms"] = (time.perf_counter() - t1) * 1e3
        st["linkage_kernel_ms"] = h.timers()["linkage"] if sh.is_linkage else None
        st["leaforder_ms"] = h.timers()["leaforder"] if sh.is_linkage else None
        return st, leaves

    sampler = B.ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = h.launches
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = [step() for _ in range(args.steps)]
    e1.record()
    barrier()
    secs = max(time.perf_counter() - w0, e0.elapsed_time(e1) / 1e3)
    t = torch.tensor([secs], dtype=torch.float64, device=dev)
    k1 = torch.tensor([res[-1][0]["k1_kernel_ms"]], dtype=torch.float64, device=dev)
    k1_all = [k1.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_gather(k1_all, k1)
    clocks = sampler.stop([(w0, time.perf_counter())])
    launches = h.launches - l0
    sh.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    secs = float(t.item())
    st, leaves = res[-1]
    pairs = n * (n - 1) // 2
    peak, peak_src = B.hbm_peak()
    rows = sh.r1 - sh.r0
    k1_per_rank = [float(v.item()) for v in k1_all]
    k1_ms = k1_per_rank[0]             # the linkage rank's own row block: local loads, local stores -> HBM-bound
    k1_bytes = 8.0 * rows * n          # 4 B read + 4 B written per entry of the row block
    inbound = 4.0 * (n - rows) * sh.ld  # bytes the linkage GPU receives over NVLink
    link_ms = st["linkage_kernel_ms"]
    line = {
        "metric": "Part1 distance + linkage bin-pairs/s (and secs) at %dkb" % int(round(1.6e9 / n / 1000)),
        "value": pairs / (secs / args.steps), "unit": "bin-pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 storage, f64 arithmetic", "data": "synthetic",
        "config": {"workload": "synthetic ICE genome 1.6Gb %d chr @%dkb (%d bins): distance matrix sharded by row block, "
                               "linkage on one GPU (BASELINE configs[4])" % (args.chroms, int(round(1.6e9 / n / 1000)), n),
                   "bins": n, "chromosomes": args.chroms, "seed": args.seed, "generator": "synth portable, by row block, fp32",
                   "matrix_GB": 4.0 * n * n / 1e9, "shard_GB_per_gpu": 4.0 * rows * n / 1e9,
                   "l2": "inputs (%.1f GB per GPU) larger than L2" % (4.0 * rows * n / 1e9)},
        "sharding": "rows %s; row sums all-gathered; W rows -> linkage GPU by %s; linkage on rank 0; leaf order broadcast"
                    % (sh.blocks if world <= 8 else "...", sh.transport),
        "stage_ms": st, "per_step_stage_ms": [r[0] for r in res], "generate_seconds": t_gen,
        "digests": {"leaves_sha": hashlib.sha256(np.ascontiguousarray(leaves, dtype=np.int32).tobytes()).hexdigest(),
                    "chromosome_breaks_in_leaf_order": chromosome_breaks(leaves, layout.true_chrom_of_bin),
                    "expected_breaks": args.chroms - 1},
        "e2e": None, "e2e_note": "the 102 GB matrix is generated in place on the GPUs (no host copy exists to upload)",
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_build_w32 (K1 fp32: distance transform of a row block, stores into the linkage GPU's W)",
                     "achieved": k1_bytes / (k1_ms / 1e3) / 1e9 if k1_ms else None, "peak": peak, "unit": "GB/s",
                     "frac": (k1_bytes / (k1_ms / 1e3) / 1e9 / peak) if k1_ms else None, "peak_source": peak_src,
                     "alg_bytes_per_launch": k1_bytes, "kernel_ms": k1_ms, "kernel_ms_per_rank": k1_per_rank,
                     "note": "achieved/frac: the linkage rank's launch (local loads and stores, HBM-bound). On the other ranks the "
                             "same kernel stores across NVLink and is bound by the linkage GPU's inbound link: see nvlink",
                     "nvlink": {"inbound_bytes_linkage_gpu": inbound, "peak_GBps_per_direction": 900.0,
                                "GBps": inbound / (st["distance_and_gather_ms"] / 1e3) / 1e9 if world > 1 else None},
                     "linkage": {"kernel_ms": link_ms, "merges_per_s": (n - 1) / (link_ms / 1e3) if link_ms else None,
                                 "alg_bytes": 12.0 * n * n, "GBps": 12.0 * n * n / (link_ms / 1e3) / 1e9 if link_ms else None,
                                 "stats": h.linkage_stats()}},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0
