"""GPU, >= 2 devices: chromosome-sharded Part2 and candidate-sharded brute force give the single-GPU results."""
import os
import pickle
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from hic_genome_assembler_b200 import dist as hd, part2
    from hic_genome_assembler_b200._lib import Handle
    from hic_genome_assembler_b200.synth import make_genome
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n, nc = 600, 6
        g = make_genome(n, n_chrom=nc, seed=3, scaffolds_per_chrom=9)
        A = g.A.numpy()
        names = [g.scaffold_names[s] for s in g.scaffold_of_bin]
        groups = [[(i + 1, names[i]) for i in range(n) if g.true_chrom_of_bin[i] == ch] for ch in range(nc)]
        index_of = {i + 1: i for i in range(n)}
        h = Handle(rank)
        h.p2_upload(A)
        merged = hd.order_genome_distributed(h, index_of, groups, 5, 4)
        # candidate-range sharded brute force on the first group
        pool = part2._pool(groups[0])
        chrom = part2.DeviceChromosome(h, index_of, pool)
        sharded = hd.brute_force_sharded(h, chrom, pool[:6], device=torch.device("cuda", rank))
        if rank == 0:
            single, cost, rounds, _ = part2.order_groups_device(h, index_of, groups, 5, 4)
            want = [(single[i], float(cost[i]), int(rounds[i])) for i in range(nc)]
            bf = chrom.brute_force(pool[:6])
            with open(os.path.join(out_dir, "res.pkl"), "wb") as f:
                pickle.dump(dict(merged=merged, want=want, sharded=sharded, bf=bf), f)
        h.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharding_matches_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    res = pickle.load(open(os.path.join(str(tmp_path), "res.pkl"), "rb"))
    assert res["merged"] == res["want"]
    assert res["sharded"] == res["bf"]
