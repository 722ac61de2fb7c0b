"""CPU: the oracle (Python restatement and C twin) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

import c_oracle as CO
import hica_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module", params=[60, 200, 320])
def case(request):
    z = np.load(os.path.join(G, "part1_%d.npz" % request.param), allow_pickle=False)
    return {k: z[k] for k in z.files}


def test_row_sums(case):
    A = case["A"]
    assert np.array_equal(O.sequential_row_sums(A), case["row_sum"])
    assert np.array_equal(CO.row_sums_sequential(A), case["row_sum"])
    assert np.array_equal(O.pairwise_row_sums(A), case["pair_sum"])
    assert np.array_equal(CO.row_sums_pairwise(A), case["pair_sum"])
    assert all(O.numpy_pairwise_sum(r) == s for r, s in zip(A[:7], case["pair_sum"][:7]))


def test_distance_linkage_leaves(case):
    A = case["A"]
    D = O.convert_distance(A)
    assert sha(D) == str(case["D_sha"])
    assert sha(CO.convert_distance(A)) == str(case["D_sha"])
    Z, leaves = O.average_cluster_leaves(D)
    assert np.array_equal(Z, case["Z"]) and leaves == list(case["leaves"])
    assert O.leaf_order(Z) == list(case["leaves"])
    Zc, lc, _ = CO.linkage_average(D)
    assert np.array_equal(Zc, case["Z"]) and list(lc) == list(case["leaves"])
    if len(A) <= 200:
        assert np.array_equal(O.nn_chain_average(O.condensed_upper(D), len(A)), case["Z"])


def test_reorder_similarity_rank(case):
    A = case["A"]
    leaves = case["leaves"]
    D = O.convert_distance(A)
    Dp = O.reorder(D, leaves)
    assert sha(Dp) == str(case["Dp_sha"])
    S = O.convert_similarity(Dp, case["row_sum"][leaves])
    assert sha(S) == str(case["S_sha"])
    assert np.array_equal(O.rank_matrix(S), case["R"])
    assert np.array_equal(O.rank_matrix_stable(S), case["R"])     # these inputs are tie-free
    assert np.array_equal(CO.rank_rows(S), case["R"])


def test_cut_indices(case):
    R = case["R"].astype(np.int64)
    init = O.pre_process_breakpoints(R, 5, 0.0, .05)
    assert init == case["init"].tolist()
    assert O.filter_noisy_breakpoints(R, init, .05) == case["cuts"].tolist()
    assert O.pre_process_breakpoints(R, 5, 0.05, .05) == case["init_minfrac05"].tolist()
    for start in (0, len(R) // 3):
        assert np.array_equal(CO.window_counts(R, start), O.window_counts(R, start))


@pytest.mark.parametrize("n", [60, 200, 320])
def test_part2_orderings(n):
    z = np.load(os.path.join(G, "part1_%d.npz" % n))
    A = z["A"]
    spec = json.load(open(os.path.join(G, "part2_%d.json" % n)))
    index_of = {i + 1: i for i in range(n)}
    for grp in spec["groups"]:
        nodes = [(int(b), s) for b, s in grp["nodes"]]
        idx = np.array([b - 1 for b, _ in nodes])
        sub = np.ascontiguousarray(A[np.ix_(idx, idx)])
        total = O.total_weight(sub)
        assert float(total) == grp["total_identity"]
        assert CO.total_weight(sub) == grp["total_identity"]
        assert O.ordering_cost(sub, total) == grp["cost_identity"]
        assert O.ordering_cost_numba(sub, total) == grp["cost_identity"]
        assert CO.ordering_cost(sub, np.arange(len(sub)), total) == grp["cost_identity"]
        got, cost = O.order_chromosome(A, index_of, nodes, spec["nScaffolds"], spec["scanScaffolds"])
        assert [list(x) for x in got] == grp["ordering"]
        want_cost = grp["cost_after_scan"] if grp["cost_after_scan"] is not None else grp["cost_after_insertion"]
        assert float(cost) == want_cost   # printed with str(float) = shortest round-trip repr, so exact
