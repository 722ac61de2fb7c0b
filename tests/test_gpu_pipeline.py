"""GPU: whole -part1 -part2 hot path through the host mirrors vs the oracle."""
import numpy as np
import pytest

import hica_oracle as O
from hic_genome_assembler_b200 import part1, part2
from hic_genome_assembler_b200.synth import make_genome

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,nc,spc,seed,nscaf,scan", [(400, 8, 8, 7, 4, 3), (300, 3, 12, 5, 5, 4), (240, 4, 9, 3, 6, 5)])
def test_part1_part2_pipeline(hica, n, nc, spc, seed, nscaf, scan):
    g = make_genome(n, n_chrom=nc, seed=seed, scaffolds_per_chrom=spc)
    A = g.A.numpy()
    ref = O.part1_cut_indices(A, min_frac=0.0)
    res = part1.run_part1(hica, A, 5, 0.0, .05, want_Z=True)
    assert np.array_equal(res["Z"], ref["Z"])
    assert list(res["leaves"]) == list(ref["leaves"])
    assert res["cuts"] == ref["cuts"]
    names = [g.scaffold_names[s] for s in g.scaffold_of_bin]
    sizes = {nm: int(sz) for nm, sz in zip(g.scaffold_names, g.scaffold_sizes_bp)}
    bin_ids = np.arange(1, n + 1)
    groups = part1.chromosome_groups(res["cuts"], res["leaves"], bin_ids, names, sizes)
    assert len(groups) >= 1
    hica.p2_use_p1_matrix()
    index_of = {int(b): i for i, b in enumerate(bin_ids)}
    for grp in groups:
        got, cost, rounds = part2.order_chromosome(hica, index_of, grp, nscaf, scan)
        want, wcost = O.order_chromosome(A, index_of, grp, nscaf, scan)
        assert got == want
        assert cost == wcost


def test_single_bin_scaffolds_orientation_parity(hica):
    """1-bin scaffolds score identically in both orientations; the reference's alternating flips decide the sign."""
    rng = np.random.RandomState(4)
    n = 40
    pos = np.arange(n)
    A = 100.0 / (1.0 + np.abs(pos[:, None] - pos[None, :])) * np.exp(0.1 * rng.randn(n, n))
    A = (A + A.T) / 2
    # scaffolds: sizes 6,5,5,4,4,3 then ten 1-bin scaffolds, three 1-bin ... shuffled ids
    sizes = [6, 5, 5, 4, 4, 3] + [1] * 13
    nodes, b = [], 0
    for si, sz in enumerate(sizes):
        for _ in range(sz):
            nodes.append((b + 1, "s%d" % si))
            b += 1
    index_of = {i + 1: i for i in range(n)}
    hica.p2_upload(A)
    got, cost, _ = part2.order_chromosome(hica, index_of, nodes, 6, 5)
    want, wcost = O.order_chromosome(A, index_of, nodes, 6, 5)
    assert got == want and cost == wcost


@pytest.mark.parametrize("n,nc,spc,seed,nscaf,scan", [(400, 8, 8, 7, 4, 3), (300, 3, 12, 5, 5, 4), (240, 4, 9, 3, 6, 5),
                                                     (200, 5, 3, 9, 6, 5), (120, 6, 1, 2, 6, 5)])
def test_device_engine_matches_oracle(hica, n, nc, spc, seed, nscaf, scan):
    """hica_p2_order_groups (all groups in one call, device-side decisions, closed-form filter + exact re-score)
    must give the orderings and costs of the reference's exhaustive loops."""
    g = make_genome(n, n_chrom=nc, seed=seed, scaffolds_per_chrom=spc)
    A = g.A.numpy()
    names = [g.scaffold_names[s] for s in g.scaffold_of_bin]
    groups = []
    for ch in range(nc):
        groups.append([(i + 1, names[i]) for i in range(n) if g.true_chrom_of_bin[i] == ch])
    index_of = {i + 1: i for i in range(n)}
    hica.p2_upload(A)
    got, cost, rounds, stats = part2.order_groups_device(hica, index_of, groups, nscaf, scan)
    for gi, grp in enumerate(groups):
        want, wcost = O.order_chromosome(A, index_of, grp, nscaf, scan)
        assert got[gi] == want, gi
        assert cost[gi] == wcost, gi
    assert stats[:, 2].sum() == 0 and stats[:, 3].sum() == 0


def test_device_engine_single_bin_scaffolds_and_asymmetric_matrix(hica):
    rng = np.random.RandomState(4)
    n = 40
    pos = np.arange(n)
    A = 100.0 / (1.0 + np.abs(pos[:, None] - pos[None, :])) * np.exp(0.1 * rng.randn(n, n))
    Asym = (A + A.T) / 2
    sizes = [6, 5, 5, 4, 4, 3] + [1] * 13
    nodes, b = [], 0
    for si, sz in enumerate(sizes):
        for _ in range(sz):
            nodes.append((b + 1, "s%d" % si))
            b += 1
    index_of = {i + 1: i for i in range(n)}
    for M, full in ((Asym, 0), (A, 1)):   # asymmetric input falls back to exact scoring of every candidate
        hica.p2_upload(M)
        got, cost, rounds, stats = part2.order_groups_device(hica, index_of, [nodes], 6, 5)
        want, wcost = O.order_chromosome(M, index_of, nodes, 6, 5)
        assert got[0] == want and cost[0] == wcost
        assert stats[0, 2] == full
