#!/usr/bin/env python
"""One-off parity check above the sizes the test suite can afford (run on a GPU box; needs ~20 GB of host RAM):
linkage (k_nn_chain3 with the chain in global memory and two column pairs per thread) and the multi-chunk rank
sort at N > 16,384 against scipy / numpy.   python profiles/check_large_n.py [N]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import hica_oracle as O  # noqa: E402
from hic_genome_assembler_b200._lib import Handle  # noqa: E402
from hic_genome_assembler_b200.synth import make_genome  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
g = make_genome(n, n_chrom=20, seed=11, device="cuda")
A = g.A.cpu().numpy()
h = Handle(0)
h.p1_upload(A)
h.p1_set_row_sums(None)
h.p1_distance()
t0 = time.time()
Z, leaves = h.p1_linkage()
print("gpu linkage %.3f s" % (time.time() - t0), h.linkage_stats(), flush=True)
D = O.convert_distance(A)
t0 = time.time()
Z_ref, leaves_ref = O.average_cluster_leaves(D)
print("scipy linkage + leaf order %.1f s" % (time.time() - t0), flush=True)
print("Z bitwise equal:", np.array_equal(Z, Z_ref), " leaves equal:", list(leaves) == list(leaves_ref), flush=True)
h.p1_rank()
R = h.p1_get_rank()
rs = O.sequential_row_sums(A)
rows = np.random.RandomState(0).choice(n, 40, replace=False)
lv = np.asarray(leaves_ref)
ok = True
for r in rows:
    Srow = O.convert_similarity(D[lv[r]][lv][None, :], rs[lv[r:r + 1]])[0]
    want = np.argsort(Srow, kind="stable")[::-1]
    ok &= np.array_equal(R[r], want)
print("rank rows equal (40 sampled):", bool(ok))
