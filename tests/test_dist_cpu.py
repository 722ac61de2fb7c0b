"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (no kernels)."""
import os
import socket

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
        with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_merge_and_argmax(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
