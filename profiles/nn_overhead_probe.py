"""Fixed per-search cost of the NN-chain kernel: tiny matrices make the row loads negligible."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hic_genome_assembler_b200._lib import Handle
h = Handle(0)
rng = np.random.RandomState(0)
for n in (32, 64, 256, 1024, 2048):
    A = rng.rand(n, n) + 0.1
    A = A + A.T
    for rep in range(2):
        h.p1_upload(A); h.p1_distance(); h.p1_linkage(want_Z=False)
    s = h.linkage_stats(); k = s["searches"]
    print("n=%5d linkage %.3f ms  searches %d  per search: load %d red %d xchg %d  merge %d" % (
        n, h.timers()["linkage"], k, s["load_cycles"] // k, s["reduce_cycles"] // k, s["exchange_cycles"] // k,
        s["merge_cycles"] // (n - 1)))
