#!/usr/bin/env python
"""Golden vectors at the PUBLISHED sizes (BASELINE.json configs[1] = 16,000 bins, configs[3] = 32,000 bins).

Run in the build container (CPU only; ~15 min at 16,000 bins on 8 cores, ~1 h at 32,000):

    python tests/golden/make_golden_large.py --bins 16000
    python tests/golden/make_golden_large.py --bins 32000

The workload is ``synth.make_genome_portable`` (bit-identical through numpy here and through torch on the GPU
box; the digest of A is stored so that the GPU test can prove it regenerated the same matrix).  Part1 is the
oracle at the reference's call sites - the very numpy/scipy calls scaffoldToChromosomes.py:1116-1134 makes
(``hica_oracle``), with the per-row Python loop of the window counts replaced by its C twin (``fast_oracle``,
pinned in tests/test_oracle_fast.py).  Part2 is ``fast_oracle.order_chromosome`` (orderGenome.py:551-586, the C
restatement of numba's score) on every chromosome group, one process per core.

Stored in tests/golden/large_<bins>.npz: leaves, Z (raw bytes digest + the heights), initial and filtered cuts,
SHA-256 of A / row sums / the uint16 rank matrix, per-group orderings, costs and scan rounds (JSON).
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

import fast_oracle as FO  # noqa: E402
import hica_oracle as O  # noqa: E402
from hic_genome_assembler_b200 import part1  # noqa: E402
from hic_genome_assembler_b200.synth import make_genome_portable  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def orderings_digest(orderings):
    """SHA-256 of the chosen orderings/orientations of all groups (the same function bench.py uses)."""
    txt = json.dumps([[[str(nm), str(o)] for nm, o in grp] for grp in orderings], separators=(",", ":"))
    return hashlib.sha256(txt.encode()).hexdigest()


_G = {}


def _one_group(gi):
    t0 = time.time()
    got, cost, rounds = FO.order_chromosome(_G["A"], _G["lut"], _G["groups"][gi], _G["nscaf"], _G["scan"])
    return gi, got, cost, rounds, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", type=int, default=16000)
    ap.add_argument("--chroms", type=int, default=20)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--no-part2", action="store_true")
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    args = ap.parse_args()
    n = args.bins
    T0 = time.time()

    def log(msg):
        print("[%7.1fs] %s" % (time.time() - T0, msg), flush=True)

    g = make_genome_portable(n, n_chrom=args.chroms, seed=args.seed)
    A = g.A.numpy()
    log("matrix generated")
    out = {"A_sha": sha(A)}
    rs = O.sequential_row_sums(A)
    out["row_sum_sha"] = sha(rs)
    D = O.convert_distance(A)
    log("distance")
    Z, leaves = O.average_cluster_leaves(D)
    leaves = np.asarray(leaves, dtype=np.int32)
    log("linkage + leaf order")
    Dp = O.reorder(D, leaves)
    del D
    S = O.convert_similarity(Dp, rs[leaves])
    del Dp
    R = O.rank_matrix(S)
    # numpy's argsort is unstable on exact ties; the CUDA rank kernel implements "stable ascending,
    # reversed".  Count the tied neighbours, and when there are any take the cut lists under BOTH orders.
    ties = int((np.take_along_axis(S, R[:, 1:], 1) == np.take_along_axis(S, R[:, :-1], 1)).sum())
    R32 = np.ascontiguousarray(R, dtype=np.int32)
    del R
    stable_equal = True
    if ties:
        Rs = np.ascontiguousarray(O.rank_matrix_stable(S), dtype=np.int32)
        stable_equal = bool(np.array_equal(Rs, R32))
    del S
    log("rank matrix (tied neighbours in sorted rows: %d; numpy order == stable order: %s)" % (ties, stable_equal))
    init = FO.pre_process_breakpoints(R32, 5, 0.0)
    log("initial cuts %s" % init)
    cuts = [int(c) for c in O.filter_noisy_breakpoints(R32, init, .05)]
    log("filtered cuts %s" % cuts)
    if not stable_equal:
        init_s = FO.pre_process_breakpoints(Rs, 5, 0.0)
        cuts_s = [int(c) for c in O.filter_noisy_breakpoints(Rs, init_s, .05)]
        out["cuts_equal_under_stable_tie_order"] = bool(init_s == init and cuts_s == cuts)
        log("cut lists under the stable tie order equal: %s" % out["cuts_equal_under_stable_tie_order"])
        R32 = Rs   # the digest below is that of the stable rule (what the GPU implements; ties documented)
    out["rank_tie_rule"] = "numpy == stable" if stable_equal else "stable (numpy's own order differs on %d tied neighbours)" % ties
    out["R_u16_sha" if n <= 65535 else "R_i32_sha"] = sha(R32.astype(np.uint16) if n <= 65535 else R32)
    del R32

    bin_ids = np.arange(1, n + 1)
    groups = part1.chromosome_groups_compact(cuts, leaves, bin_ids, np.asarray(g.scaffold_of_bin, dtype=np.int64),
                                             list(g.scaffold_names), [int(v) for v in g.scaffold_sizes_bp])
    orderings, costs, rounds = [], [], []
    if not args.no_part2:
        lut = np.full(n + 2, -1, dtype=np.int64)
        lut[bin_ids] = np.arange(n)
        _G.update(A=A, lut=lut, groups=groups, nscaf=6, scan=5)
        res = {}
        with mp.get_context("fork").Pool(args.procs) as pool:
            # largest groups first
            order = sorted(range(len(groups)), key=lambda i: -sum(len(ids) for _, ids in groups[i]))
            for gi, got, cost, rnd, dt in pool.imap_unordered(_one_group, order):
                res[gi] = (got, cost, rnd)
                log("group %d: %d scaffolds, cost %.17g, %d scan rounds (%.0fs)" % (gi, len(got), cost, rnd, dt))
        for gi in range(len(groups)):
            orderings.append([[nm, o] for nm, o in res[gi][0]])
            costs.append(res[gi][1])
            rounds.append(res[gi][2])
    meta = dict(bins=n, chroms=args.chroms, seed=args.seed, min_size=5, min_frac=0.0, psig=.05, nScaffolds=6,
                scanScaffolds=5, tied_neighbours=ties, generator="synth.make_genome_portable",
                versions=dict(numpy=np.__version__, scipy=__import__("scipy").__version__),
                seconds=round(time.time() - T0, 1), **out,
                leaves_sha=sha(leaves), Z_sha=sha(Z), initial_cuts=init, cuts=cuts,
                groups=[[[nm, len(ids)] for nm, ids in grp] for grp in groups],
                orderings=orderings, costs=costs, rounds=rounds,
                orderings_sha=orderings_digest(orderings) if orderings else None)
    np.savez_compressed(os.path.join(HERE, "large_%d.npz" % n), leaves=leaves, heights=Z[:, 2].copy(),
                        meta=json.dumps(meta))
    log("written large_%d.npz" % n)


if __name__ == "__main__":
    main()
