import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hica_oracle as O
from hic_genome_assembler_b200._lib import Handle
def sym(n, seed):
    rng = np.random.RandomState(seed)
    A = rng.rand(n, n) + 0.1; A = (A + A.T) / 2
    return A
for env in ({}, {"HICA_NN_FLAGS": "1"}, {"HICA_NN_CLUSTER": "1"}, {"HICA_NN_CLUSTER": "2"}, {"HICA_NN_KERNEL": "3"}):
    for k in ("HICA_NN_FLAGS", "HICA_NN_CLUSTER", "HICA_NN_KERNEL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for n in (4, 5, 6, 9, 33, 70, 200, 1200):
        nbad = 0; first = None
        for seed in range(5):
            A = sym(n, 100 * n + seed)
            D = O.convert_distance(A)
            Zr, lr = O.average_cluster_leaves(D)
            with Handle(0) as h:
                h.p1_upload(A); h.p1_distance()
                Z, lv = h.p1_linkage()
            if not np.array_equal(Z, Zr):
                nbad += 1
                if first is None:
                    i = int(np.flatnonzero((Z != Zr).any(axis=1))[0])
                    first = (seed, i, Z[i].tolist(), Zr[i].tolist())
        print(env, "n", n, "bad", nbad, first, flush=True)
