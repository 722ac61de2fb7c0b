"""The batch forms in oracle/fast_oracle.py (used for the golden vectors at the published sizes) against the
one-candidate-at-a-time oracle that is pinned to the reference."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import c_oracle as CO  # noqa: E402
import fast_oracle as FO  # noqa: E402
import hica_oracle as O  # noqa: E402
from hic_genome_assembler_b200.synth import make_genome, make_genome_portable  # noqa: E402


@pytest.mark.parametrize("n,spc,nscaf,scan,seed", [(70, 9, 4, 3, 1), (90, 12, 5, 4, 2), (48, 16, 4, 3, 3), (64, 3, 6, 5, 4)])
def test_order_chromosome_batch_equals_sequential(n, spc, nscaf, scan, seed):
    g = make_genome(n, n_chrom=1, seed=seed, scaffolds_per_chrom=spc)
    A = g.A.numpy()
    nodes = [(i + 1, g.scaffold_names[g.scaffold_of_bin[i]]) for i in range(n)]
    index_of = {i + 1: i for i in range(n)}
    want, wcost = O.order_chromosome(A, index_of, nodes, nscaf, scan)
    lut = np.full(n + 2, -1, dtype=np.int64)
    lut[np.arange(1, n + 1)] = np.arange(n)
    got, cost, _ = FO.order_chromosome(A, lut, nodes, nscaf, scan)
    assert got == want and cost == wcost


def test_single_bin_scaffolds_and_asymmetric_matrix():
    rng = np.random.RandomState(5)
    n = 40
    A = rng.rand(n, n)          # asymmetric on purpose: only the upper triangle in candidate order is read
    names = ["s%d" % (i // 2 if i < 20 else i) for i in range(n)]   # ten 2-bin scaffolds and twenty 1-bin ones
    nodes = [(i + 1, names[i]) for i in range(n)]
    want, wcost = O.order_chromosome(A, {i + 1: i for i in range(n)}, nodes, 4, 3)
    lut = np.concatenate([[-1], np.arange(n)])
    got, cost, _ = FO.order_chromosome(A, lut, nodes, 4, 3)
    assert got == want and cost == wcost


def test_breakpoints_with_c_window_counts():
    g = make_genome(260, n_chrom=5, seed=3, scaffolds_per_chrom=5)
    A = g.A.numpy()
    ref = O.part1_cut_indices(A, 5, 0.0, .05)
    R32 = np.ascontiguousarray(ref["R"], dtype=np.int32)
    assert FO.pre_process_breakpoints(R32, 5, 0.0) == ref["initial_cuts"]
    for start in (0, 17, 200):
        assert np.array_equal(CO.window_counts(R32, start), O.window_counts(ref["R"], start))


def test_portable_generator_numpy_equals_torch():
    """make_genome_portable must give the same bytes through numpy and through torch (the golden vectors at
    16,000 bins are made with numpy in the build container and checked against a matrix generated with torch on
    the GPU)."""
    for n, nc in ((97, 4), (256, 8)):
        a = make_genome_portable(n, n_chrom=nc, seed=5)
        b = make_genome_portable(n, n_chrom=nc, seed=5, device="cpu")
        assert np.array_equal(a.A.numpy(), b.A.numpy())
        assert np.array_equal(a.A.numpy(), a.A.numpy().T)
        assert np.array_equal(a.scaffold_of_bin, b.scaffold_of_bin)
