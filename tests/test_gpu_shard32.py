"""GPU: the fp32 / row-block sharded Part1 front end of the 10 kb configuration (hica_s32_*, dist10kb.Shard32) against
this library's own fp64 path at a size both can run - the numerical contract SURVEY 7.7 states for 10 kb, where no
reference exists: merge heights within 5e-6 relative, identical chromosome-level clusters, every chromosome
contiguous in the leaf order."""
import os
import pickle
import socket

import numpy as np
import pytest
import torch

from hic_genome_assembler_b200 import dist10kb

pytestmark = pytest.mark.gpu
HEIGHT_RTOL = 5e-6


def _clusters(Z, k):
    import scipy.cluster.hierarchy as sch
    lab = sch.fcluster(Z, k, criterion="maxclust")
    # canonical form: label of a cluster = its smallest member
    first = {}
    for i, l in enumerate(lab):
        first.setdefault(l, i)
    return np.asarray([first[l] for l in lab])


@pytest.mark.parametrize("n,nc", [(1200, 6), (3000, 20)])
def test_fp32_sharded_front_end_matches_fp64_path(hica, n, nc):
    dev = torch.device("cuda", 0)
    layout = dist10kb.ShardLayout(n, nc, seed=5, scaffolds_per_chrom=8)
    sh = dist10kb.Shard32(hica, n, dev)
    dist10kb.generate_shard(sh.A, layout, 0, n, sh.blocks, ice_sweeps=20)
    A32 = sh.A.cpu().numpy()
    assert np.array_equal(A32, A32.T)                       # symmetric by construction, no transpose pass
    st = sh.distance_and_gather()
    # row sums: fp64 accumulation of the fp32 entries
    assert np.allclose(sh.sums.cpu().numpy(), A32.astype(np.float64).sum(1), rtol=1e-12)
    # K1: W[i][j] = (1 - A_ij / s_min(i,j)) + 1, one rounding to fp32
    s = sh.sums.cpu().numpy()
    W = sh.W.cpu().numpy()[:, :n]
    smin = np.minimum.outer(np.arange(n), np.arange(n))
    want = ((1.0 - A32.astype(np.float64) / s[smin]) + 1.0).astype(np.float32)
    off = ~np.eye(n, dtype=bool)
    assert np.array_equal(W[off], want[off])
    Z32, leaves32 = sh.linkage(want_Z=True)
    # the same matrix through the fp64 path
    hica.p1_upload(A32.astype(np.float64))
    hica.p1_set_row_sums(None)
    hica.p1_distance()
    Z64, leaves64 = hica.p1_linkage(want_Z=True)
    assert np.allclose(Z32[:, 2], Z64[:, 2], rtol=HEIGHT_RTOL, atol=0)
    # (merges whose heights agree to within the tolerance may swap places in the sorted list - the inter-chromosome
    #  merges all sit within 1e-4 of 2.0 - so the merge LIST is not compared entry by entry; the clusters are)
    assert np.array_equal(_clusters(Z32, nc), _clusters(Z64, nc))      # chromosome-level clusters identical
    assert dist10kb.chromosome_breaks(leaves32, layout.true_chrom_of_bin) == nc - 1
    assert sorted(leaves32.tolist()) == list(range(n))
    sh.close()


@pytest.mark.parametrize("n,nc,cluster", [(36, 1, ""), (72, 2, ""), (1200, 6, ""), (3000, 20, ""), (1500, 6, "2"), (900, 4, "1")])
def test_fp32_cached_chain_equals_the_full_search_chain_bitwise(hica, n, nc, cluster, monkeypatch):
    """k_nn_chain3g<float> (nearest-neighbour cache, tables in global memory; the default of hica_s32_linkage) against
    nn2::k_nn_chain2<float> (every search in full): same fp32 working matrix, same merge order, same heights - bit for
    bit, ties included (the fp32 rounding creates plenty)."""
    if cluster:
        monkeypatch.setenv("HICA_NN_CLUSTER", cluster)
    dev = torch.device("cuda", 0)
    out = {}
    for which in ("2", "3"):
        monkeypatch.setenv("HICA_S32_NN", which)
        layout = dist10kb.ShardLayout(n, nc, seed=9, scaffolds_per_chrom=3 if n < 100 else 8)
        sh = dist10kb.Shard32(hica, n, dev)
        dist10kb.generate_shard(sh.A, layout, 0, n, sh.blocks, ice_sweeps=20)
        sh.distance_and_gather()
        out[which] = sh.linkage(want_Z=True)
        st = hica.linkage_stats()
        if which == "3" and n >= 72:
            assert st["cached_searches"] > 0
        sh.close()
    assert np.array_equal(out["2"][0], out["3"][0])
    assert np.array_equal(out["2"][1], out["3"][1])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, transport):
    import torch.distributed as dist
    from hic_genome_assembler_b200._lib import Handle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["HICA_S32_TRANSPORT"] = transport
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n, nc = 2000, 8
        dev = torch.device("cuda", rank)
        h = Handle(rank)
        layout = dist10kb.ShardLayout(n, nc, seed=9, scaffolds_per_chrom=8)
        sh = dist10kb.Shard32(h, n, dev)
        dist10kb.generate_shard(sh.A, layout, sh.r0, sh.r1, sh.blocks, ice_sweeps=10, dist_mod=sh.dist)
        sh.distance_and_gather()
        Z, leaves = sh.linkage(want_Z=True)
        if rank == 0:
            with open(os.path.join(out_dir, "res_%s.pkl" % transport), "wb") as f:
                pickle.dump(dict(Z=Z, leaves=leaves, transport=sh.transport, A0=sh.A.cpu().numpy(), r=(sh.r0, sh.r1)), f)
        sh.close()
        h.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["ipc", "nccl"])
def test_two_gpu_row_blocks_equal_single_gpu(hica, tmp_path, transport):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), transport), nprocs=2, join=True)
    res = pickle.load(open(os.path.join(str(tmp_path), "res_%s.pkl" % transport), "rb"))
    n, nc = 2000, 8
    dev = torch.device("cuda", 0)
    layout = dist10kb.ShardLayout(n, nc, seed=9, scaffolds_per_chrom=8)
    sh = dist10kb.Shard32(hica, n, dev)
    dist10kb.generate_shard(sh.A, layout, 0, n, sh.blocks, ice_sweeps=10)
    r0, r1 = res["r"]
    assert np.array_equal(sh.A.cpu().numpy()[r0:r1], res["A0"])        # the generator does not depend on the sharding
    sh.distance_and_gather()
    Z, leaves = sh.linkage(want_Z=True)
    assert np.array_equal(Z, res["Z"]) and np.array_equal(leaves, res["leaves"])   # bitwise: same fp32 working matrix
    if transport == "ipc":
        assert "IPC" in res["transport"], res["transport"]
    sh.close()
