"""GPU: the CUDA path through the C ABI against fixtures produced by the UNMODIFIED reference."""
import hashlib
import json
import os

import numpy as np
import pytest

from hic_genome_assembler_b200 import part1, part2

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("n", [60, 200, 320])
def test_part1_seams_match_reference(hica, n):
    z = np.load(os.path.join(G, "part1_%d.npz" % n))
    A = z["A"]
    hica.p1_upload(A)
    hica.p1_set_row_sums(None)
    assert np.array_equal(hica.p1_get_row_sums(), z["row_sum"])          # removeRows :134
    hica.p1_distance()
    assert np.array_equal(hica.p1_get_pairwise_row_sums(), z["pair_sum"])  # row.sum() :147
    assert sha(hica.p1_get_distance()) == str(z["D_sha"])                # convertMatrix(distance)
    Z, leaves = hica.p1_linkage()
    assert np.array_equal(Z, z["Z"]) and list(leaves) == list(z["leaves"])  # averageClusterNodes
    assert sha(hica.p1_get_reordered(0)) == str(z["Dp_sha"])             # reorderMatrix
    assert sha(hica.p1_get_reordered(1)) == str(z["S_sha"])              # convertMatrix(similarity)
    hica.p1_rank()
    assert np.array_equal(hica.p1_get_rank(), z["R"])                    # argsort(...)[:, ::-1]
    init, _ = hica.p1_breakpoints(5, 0.0)
    assert init == z["init"].tolist()                                    # pre_process_all_matrix_breakpoints
    cuts, _ = hica.p1_filter(init, .05)
    assert cuts == z["cuts"].tolist()                                    # filter_noisy_breakpoints
    init05, _ = hica.p1_breakpoints(5, 0.05)
    assert init05 == z["init_minfrac05"].tolist()


@pytest.mark.parametrize("n", [60, 200, 320])
def test_part2_orderings_match_reference(hica, n):
    z = np.load(os.path.join(G, "part1_%d.npz" % n))
    A = z["A"]
    spec = json.load(open(os.path.join(G, "part2_%d.json" % n)))
    hica.p2_upload(A)
    index_of = {i + 1: i for i in range(n)}
    for grp in spec["groups"]:
        nodes = [(int(b), s) for b, s in grp["nodes"]]
        got, cost, _ = part2.order_chromosome(hica, index_of, nodes, spec["nScaffolds"], spec["scanScaffolds"])
        assert [list(x) for x in got] == grp["ordering"]
        want = grp["cost_after_scan"] if grp["cost_after_scan"] is not None else grp["cost_after_insertion"]
        assert cost == want
        # costFunction_numba / total on the identity layout
        idx = np.array([b - 1 for b, _ in nodes], dtype=np.int32)
        hica.p2_select(idx)
        hica.p2_set_segments([0, len(idx)], np.arange(len(idx)))
        assert hica.p2_total([0]) == grp["total_identity"]
        _, _, sc = hica.p2_score_orders(np.arange(len(idx), dtype=np.int32)[None], grp["total_identity"], 0.0, True)
        assert sc[0] == grp["cost_identity"]
