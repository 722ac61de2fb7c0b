"""GPU: tie-order sensitivity report of the cut searches (hica_p1_tie_report) on maps with many exactly tied
similarities - a zero-heavy sparse map and a map rounded the way HiC-Pro's text format rounds (%.6f) - against the
oracle's brute-force-checked bounds (tests/test_oracle_ties.py), through both sources of the tie masks (the rank
pre-sort's own masks, and the masks recomputed from the rank matrix)."""
import numpy as np
import pytest

import hica_oracle as O
from hic_genome_assembler_b200 import part1
from hic_genome_assembler_b200.synth import make_genome

pytestmark = pytest.mark.gpu


def _oracle(A):
    rs = O.sequential_row_sums(A)
    D = O.convert_distance(A)
    Z, leaves = O.average_cluster_leaves(D)
    S = O.convert_similarity(O.reorder(D, leaves), rs[np.asarray(leaves)])
    R = O.rank_matrix_stable(S)
    init = O.pre_process_breakpoints(R, 5, 0.0)
    cuts = O.filter_noisy_breakpoints(R, init, .05)
    return dict(S=S, R=R, leaves=list(leaves), init=init, cuts=cuts,
                tie_b=O.tie_sensitive_breakpoints(S, R, 5, 0.0), tie_f=O.tie_sensitive_filter(S, R, init, .05))


def _sparse_map(n, nc, seed, zero_frac):
    g = make_genome(n, n_chrom=nc, seed=seed, scaffolds_per_chrom=4)
    A = g.A.numpy().copy()
    thr = np.quantile(A[np.triu_indices(n, 1)], zero_frac)
    A[A < thr] = 0.0
    return A


def _rounded_map(n, nc, seed, decimals):
    g = make_genome(n, n_chrom=nc, seed=seed, scaffolds_per_chrom=4)
    return np.round(g.A.numpy(), decimals)


CASES = [("sparse60", lambda: _sparse_map(240, 4, 3, 0.60)), ("sparse85", lambda: _sparse_map(200, 4, 5, 0.85)),
         ("rounded1", lambda: _rounded_map(240, 4, 7, 1)), ("rounded6", lambda: _rounded_map(200, 4, 9, 6))]


@pytest.mark.parametrize("overlap", ["1", "0"])
@pytest.mark.parametrize("name,make", CASES)
def test_tie_report_matches_oracle(hica, monkeypatch, name, make, overlap):
    monkeypatch.setenv("HICA_RANK_OVERLAP", overlap)
    A = make()
    assert np.array_equal(A, A.T) and np.all(A.sum(1) > 0)
    ref = _oracle(A)
    res = part1.run_part1(hica, A, min_size=5, min_frac=0.0, psig=.05)
    assert list(res["leaves"]) == ref["leaves"]
    assert np.array_equal(hica.p1_get_rank(), ref["R"])
    assert res["initial_cuts"] == ref["init"] and res["cuts"] == ref["cuts"]
    rep = hica.tie_report()
    assert rep["available"]
    assert rep["breakpoints"] == ref["tie_b"], (name, rep, ref["tie_b"], ref["tie_f"])
    assert rep["filter"] == ref["tie_f"], (name, rep, ref["tie_b"], ref["tie_f"])
    if name.startswith("sparse"):
        assert rep["breakpoints"] + rep["filter"] > 0    # the boundary does fall inside the run of zeros
    # the `ambiguous` channel carries them (plus p-values within 1e-7 of the level, none here)
    assert res["ambiguous"] >= rep["breakpoints"] + rep["filter"]


def test_no_tie_information_for_a_rank_matrix_from_the_host(hica):
    A = _sparse_map(120, 3, 2, 0.5)
    ref = _oracle(A)
    hica.p1_upload(A)          # (a rank matrix must have the size of the resident matrix)
    hica.p1_set_rank(ref["R"].astype(np.int64))
    init, amb = hica.p1_breakpoints(5, 0.0)
    assert init == ref["init"]
    assert hica.tie_report()["available"] is False


def test_tie_free_map_reports_zero(hica):
    g = make_genome(300, n_chrom=5, seed=3, scaffolds_per_chrom=5)
    res = part1.run_part1(hica, g.A.numpy(), min_size=5, min_frac=0.0, psig=.05)
    rep = hica.tie_report()
    assert rep["available"] and rep["breakpoints"] == 0 and rep["filter"] == 0
