"""GPU: Part1 + Part2 at the PUBLISHED sizes (BASELINE.json configs[1] = 16,000 bins; configs[3] = 32,000) against
golden vectors made in the build container by the oracle at the reference's call sites
(tests/golden/make_golden_large.py: scaffoldToChromosomes.py:1116-1134 through numpy/scipy, orderGenome.py:551-586
through the C restatement of numba's score).  The workload is regenerated on the GPU with torch
(synth.make_genome_portable is bit-identical on every platform; the digest of the matrix is checked first)."""
import hashlib
import json
import os

import numpy as np
import pytest

from hic_genome_assembler_b200 import part1, part2
from hic_genome_assembler_b200.synth import make_genome_portable

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _golden(n):
    p = os.path.join(G, "large_%d.npz" % n)
    if not os.path.exists(p):
        pytest.skip("no golden vectors for %d bins" % n)
    z = np.load(p)
    return z, json.loads(str(z["meta"]))


@pytest.mark.parametrize("n", [16000, 32000])
def test_headline_size_against_golden(hica, n):
    import torch
    z, meta = _golden(n)
    g = make_genome_portable(n, n_chrom=meta["chroms"], seed=meta["seed"], device="cuda:0")
    A_host = g.A.cpu().numpy()
    assert sha(A_host) == meta["A_sha"], "the workload regenerated on the GPU differs from the build container's"
    hica.p1_upload_device(g.A.data_ptr(), n)
    torch.cuda.synchronize()
    del g.A
    hica.p1_set_row_sums(None)
    assert sha(hica.p1_get_row_sums()) == meta["row_sum_sha"]            # removeRows :134 (sequential builtin sum)
    hica.p1_distance()
    Z, leaves = hica.p1_linkage(want_Z=True)
    assert np.array_equal(leaves, z["leaves"])                           # averageClusterNodes :187-208
    assert np.array_equal(Z[:, 2], z["heights"])                         # merge heights, bitwise
    assert sha(Z) == meta["Z_sha"]
    hica.p1_rank()
    R = hica.p1_get_rank()
    key = "R_u16_sha" if n <= 65535 else "R_i32_sha"
    assert sha(R.astype(np.uint16 if n <= 65535 else np.int32)) == meta[key]   # argsort(...)[:, ::-1] :1132
    del R
    init, _ = hica.p1_breakpoints(meta["min_size"], meta["min_frac"])
    assert init == meta["initial_cuts"]                                  # pre_process_all_matrix_breakpoints :513-551
    cuts, _ = hica.p1_filter(init, meta["psig"])
    assert cuts == meta["cuts"]                                          # filter_noisy_breakpoints :553-727
    if not meta["orderings"]:
        return
    bin_ids = np.arange(1, n + 1)
    groups = part1.chromosome_groups_compact(cuts, leaves, bin_ids, np.asarray(g.scaffold_of_bin, dtype=np.int64),
                                             list(g.scaffold_names), [int(v) for v in g.scaffold_sizes_bp])
    assert [[[nm, len(ids)] for nm, ids in grp] for grp in groups] == meta["groups"]
    lut = np.full(n + 2, -1, dtype=np.int64)
    lut[bin_ids] = np.arange(n)
    hica.p2_use_p1_matrix()
    got, cost, rounds, _ = part2.order_groups_device(hica, lut, groups, meta["nScaffolds"], meta["scanScaffolds"])
    for gi, grp in enumerate(got):                                       # orderChromosome :551-586, every group
        assert [list(x) for x in grp] == meta["orderings"][gi], "group %d" % gi
        assert float(cost[gi]) == meta["costs"][gi], "group %d" % gi
        assert int(rounds[gi]) == meta["rounds"][gi], "group %d" % gi
