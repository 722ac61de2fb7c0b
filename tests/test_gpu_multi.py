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


def _worker_p1(rank, world, port, out_dir):
    import torch.distributed as dist
    from hic_genome_assembler_b200 import dist as hd, part1
    from hic_genome_assembler_b200._lib import Handle
    from hic_genome_assembler_b200.synth import make_genome
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        out = {}
        for n, nc, sym in ((1200, 6, True), (1201, 5, False), (2100, 7, True)):  # all below 16,384 bins: pre-sorted rank
            g = make_genome(n, n_chrom=nc, seed=5, scaffolds_per_chrom=8)
            A = np.ascontiguousarray(g.A.numpy())
            h = Handle(rank)
            res = hd.run_part1_sharded(h, A, dev, min_size=5, min_frac=0.0, symmetric=sym)
            R = h.p1_get_rank()
            if rank == 0:
                h1 = Handle(rank)
                h1.p1_upload(A)
                want = part1.run_part1_resident(h1, min_size=5, min_frac=0.0)
                R1 = h1.p1_get_rank()
                h1.close()
                out[n] = dict(leaves=(res["leaves"].tolist(), list(want["leaves"])), cuts=(list(res["cuts"]), list(want["cuts"])),
                              init=(list(res["initial_cuts"]), list(want["initial_cuts"])), rank_equal=bool((R == R1).all()))
            # every rank ends with the same cuts
            mine = torch.tensor(list(res["cuts"]) + [-1] * (64 - len(res["cuts"])), dtype=torch.int64, device=dev)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            assert all(torch.equal(p, mine) for p in parts)
            h.close()
        if rank == 0:
            with open(os.path.join(out_dir, "p1.pkl"), "wb") as f:
                pickle.dump(out, f)
    finally:
        dist.destroy_process_group()


def test_two_gpu_row_block_part1_matches_single_gpu(tmp_path):
    """Row-block sharded upload + NVLink all-gather, linkage on rank 0 + leaf-order broadcast, row-block rank +
    all-gather: leaves, rank matrix and cut indices equal the single-GPU run bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker_p1, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    res = pickle.load(open(os.path.join(str(tmp_path), "p1.pkl"), "rb"))
    assert len(res) == 3
    for n, r in res.items():
        assert r["leaves"][0] == r["leaves"][1], n
        assert r["rank_equal"], n
        assert r["init"][0] == r["init"][1], n
        assert r["cuts"][0] == r["cuts"][1], n
