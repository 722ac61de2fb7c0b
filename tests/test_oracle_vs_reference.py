"""Build container only: the oracle against the UNMODIFIED reference modules, live."""
import contextlib
import io

import numpy as np
import pytest

import hica_oracle as O
import ref_harness
from hic_genome_assembler_b200.synth import make_genome

pytestmark = pytest.mark.reference


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


@pytest.mark.parametrize("n,nc,spc,seed", [(90, 3, 5, 21), (260, 5, 7, 22)])
def test_live_part1_and_part2(n, nc, spc, seed):
    p1, p2 = ref_harness.load_reference()
    g = make_genome(n, n_chrom=nc, seed=seed, scaffolds_per_chrom=spc)
    A = g.A.numpy()
    bins = [p1.Bin(i + 1, g.scaffold_names[g.scaffold_of_bin[i]], 0, 1, 1.0, 0.) for i in range(n)]
    mat, bins = quiet(p1.removeRows, [list(map(float, r)) for r in A], bins)
    rs = np.array([b.rowSum for b in bins])
    assert np.array_equal(O.sequential_row_sums(A), rs)
    D_ref = np.asarray(quiet(p1.convertMatrix, mat, bins, True, False))
    D = O.convert_distance(A)
    assert np.array_equal(D, D_ref)
    leaves_ref = quiet(p1.averageClusterNodes, D_ref, [str(i) for i in range(n)])["leaves"]
    Z, leaves = O.average_cluster_leaves(D)
    assert leaves == list(leaves_ref)
    Dp_ref, bins2 = p1.reorderMatrix(np.asmatrix(D_ref), bins, leaves_ref)
    S_ref = np.asarray(quiet(p1.convertMatrix, Dp_ref.copy(), bins2, False, True))
    S = O.convert_similarity(O.reorder(D, leaves), rs[np.asarray(leaves)])
    assert np.array_equal(S, S_ref)
    R = O.rank_matrix(S)
    init_ref = [int(v) for v in quiet(p1.pre_process_all_matrix_breakpoints, R, 5, 0.0, .05)]
    assert O.pre_process_breakpoints(R, 5, 0.0, .05) == init_ref
    assert O.filter_noisy_breakpoints(R, init_ref, .05) == [int(v) for v in quiet(p1.filter_noisy_breakpoints, R, init_ref, .05)]
    G = np.asmatrix(A)
    bins_p2 = [p2.Bin(i + 1, g.scaffold_names[g.scaffold_of_bin[i]], 0, 0, 1., 0.) for i in range(n)]
    idx_of = {i + 1: i for i in range(n)}
    for ch in range(nc):
        nodes = [[i + 1, g.scaffold_names[g.scaffold_of_bin[i]]] for i in range(n) if g.true_chrom_of_bin[i] == ch]
        ref = [(s.name, s.orientation) for s in quiet(p2.orderChromosome, [list(x) for x in nodes], G, bins_p2, 4, 3)]
        got, _ = O.order_chromosome(A, idx_of, [(b, s) for b, s in nodes], 4, 3)
        assert got == ref
