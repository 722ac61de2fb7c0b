"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (no kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hic_genome_assembler_b200 import dist as hd


def test_shard_groups_is_a_partition_and_balanced():
    sizes = [800, 790, 640, 900, 10, 12, 400, 800, 805, 33, 1]
    for world in (1, 2, 3, 4, 8):
        shards = [hd.shard_groups(sizes, world, r) for r in range(world)]
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(sizes)))
        loads = [sum(sizes[i] ** 2 for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(sizes) ** 2
    assert hd.shard_groups(sizes, 2, 0) == hd.shard_groups(list(sizes), 2, 0)   # deterministic


def test_row_blocks_cover_all_rows_and_balance():
    for n in (1, 5, 320, 3200, 16000, 16001):
        for world in (1, 2, 3, 4, 8):
            for bal in (False, True):
                for align in (1, 512):
                    b = hd.row_blocks(n, world, balance_upper_triangle=bal, align=align)
                    assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
                    assert all(b[k][1] == b[k + 1][0] for k in range(world - 1))
                    assert all(r0 <= r1 for r0, r1 in b)
    # equal rows
    assert hd.row_blocks(16000, 8) == [(2000 * k, 2000 * (k + 1)) for k in range(8)]
    # equal upper-triangle area: the PCIe bytes of a symmetric upload are balanced to within one aligned step
    n, world = 16000, 4
    b = hd.row_blocks(n, world, balance_upper_triangle=True, align=512)
    area = [sum(n - i for i in range(r0, r1)) for r0, r1 in b]
    assert max(area) - min(area) <= 512 * n
    assert b[0][1] - b[0][0] < b[-1][1] - b[-1][0]       # the first block has the longest rows, so the fewest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = [50, 40, 30, 20, 10, 5, 5]
        mine = hd.shard_groups(sizes, world, rank)
        local = {i: ("ordering-of-%d" % i, float(sizes[i])) for i in mine}
        merged = hd.merge_orderings(local, len(sizes))
        assert merged == [("ordering-of-%d" % i, float(sizes[i])) for i in range(len(sizes))]
        # arg-max over candidate shards: rank 1 holds the global maximum, rank 0 an equal score at a lower index
        cases = [
            # (per-rank (best, idx), carried, expected)
            ([(3.0, 7), (3.0, 100)], 0.0, (3.0, 7)),          # tie -> lowest candidate index
            ([(2.0, 5), (4.5, 900)], 0.0, (4.5, 900)),        # plain maximum
            ([(0.0, -1), (1.5, 64)], 1.0, (1.5, 64)),         # one rank found nothing above the carried cost
            ([(0.0, -1), (0.0, -1)], 2.5, (2.5, -1)),         # nobody did
        ]
        for per_rank, carried, want in cases:
            b, i = per_rank[rank]
            assert hd.reduce_first_max(b, i, carried) == want
        # in-place gather of ragged and of equal row blocks (the collective of run_part1_sharded, CPU tensors)
        for n_rows, blocks in ((7, [(0, 3), (3, 7)]), (6, [(0, 3), (3, 6)]), (4, [(0, 4), (4, 4)])):
            want = torch.arange(n_rows * 5, dtype=torch.float64).reshape(n_rows, 5)
            full = torch.full((n_rows, 5), -1.0, dtype=torch.float64)
            r0, r1 = blocks[rank]
            full[r0:r1] = want[r0:r1]
            hd.gather_row_blocks(full, blocks)
            assert torch.equal(full, want)
        with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_merge_and_argmax(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def _gen_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from hic_genome_assembler_b200 import dist10kb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 203
        layout = dist10kb.ShardLayout(n, 4, seed=3, scaffolds_per_chrom=5)
        blocks = dist10kb.row_blocks(n, world)
        r0, r1 = blocks[rank]
        M = torch.empty((r1 - r0, n), dtype=torch.float32)
        dist10kb.generate_shard(M, layout, r0, r1, blocks, ice_sweeps=7, dist_mod=dist)
        np.save(os.path.join(out_dir, "shard_%d.npy" % rank), M.numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_generator_does_not_depend_on_the_number_of_ranks(tmp_path):
    """dist10kb.generate_shard (the 10 kb workload, generated in place by row block): 2 gloo ranks give the very
    rows one rank gives - row sums are the only exchange - and the matrix is exactly symmetric."""
    import torch
    import torch.multiprocessing as mp
    from hic_genome_assembler_b200 import dist10kb
    mp.spawn(_gen_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    n = 203
    layout = dist10kb.ShardLayout(n, 4, seed=3, scaffolds_per_chrom=5)
    M = torch.empty((n, n), dtype=torch.float32)
    dist10kb.generate_shard(M, layout, 0, n, [(0, n)], ice_sweeps=7)
    full = M.numpy()
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "shard_%d.npy" % r)) for r in range(2)])
    assert np.array_equal(got, full)
    assert np.array_equal(full, full.T)
    assert np.allclose(full.sum(1), 1000.0, rtol=1e-3)
