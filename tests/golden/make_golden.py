#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What is stored
  part1_<n>.npz    input matrix A (seeded synthetic ICE map) and what the reference's own functions
                   return for it at every seam of Part1: removeRows row sums, convertMatrix(distance),
                   averageClusterNodes (scipy Z + leaves), reorderMatrix, convertMatrix(similarity),
                   numpy.argsort(...)[:, ::-1], pre_process_all_matrix_breakpoints, filter_noisy_breakpoints.
                   Large matrices are stored as SHA-256 digests of their raw float64 bytes.
  part2_<n>.json   for every chromosome group: the reference's orderChromosome result and a few
                   costFunction_numba / total values.
  pipeline_320/    HiC-Pro style input files + config and the five output files that
                   `run_hicAssembler.py -part1 -part2 -config ...` (unchanged entry point) wrote.
"""
import contextlib
import hashlib
import io
import json
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

import ref_harness  # noqa: E402
from hic_genome_assembler_b200.synth import make_genome, write_hicpro_files  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def part1_case(p1, n, nc, spc, seed):
    g = make_genome(n, n_chrom=nc, seed=seed, scaffolds_per_chrom=spc)
    A = g.A.numpy()
    bins = [p1.Bin(i + 1, g.scaffold_names[g.scaffold_of_bin[i]], int(g.bin_start[i]),
                   int(g.bin_start[i]) + g.resolution, float(g.bias[i]), 0.) for i in range(n)]
    mat = [list(map(float, r)) for r in A]
    mat, bins = quiet(p1.removeRows, mat, bins)
    assert len(bins) == n
    row_sum = np.array([b.rowSum for b in bins])
    D = np.asarray(quiet(p1.convertMatrix, mat, bins, True, False)).copy()
    pair_sum = np.array([np.asmatrix(A)[i].sum() for i in range(n)])
    dend = quiet(p1.averageClusterNodes, D, [str(i) for i in range(n)])
    import scipy.cluster.hierarchy as sch
    import scipy.spatial.distance as ssd
    Z = sch.average(ssd.squareform(D, checks=False))
    leaves = np.asarray(dend["leaves"], dtype=np.int32)
    Dp, bins2 = p1.reorderMatrix(np.asmatrix(D), bins, list(leaves))
    S = np.asarray(quiet(p1.convertMatrix, Dp.copy(), bins2, False, True))
    R = np.asarray(np.argsort(S, axis=1)[:, ::-1])
    init = [int(v) for v in quiet(p1.pre_process_all_matrix_breakpoints, R, 5, 0.0, .05)]
    cuts = [int(v) for v in quiet(p1.filter_noisy_breakpoints, R, init, .05)]
    init05 = [int(v) for v in quiet(p1.pre_process_all_matrix_breakpoints, R, 5, 0.05, .05)]
    np.savez_compressed(os.path.join(HERE, "part1_%d.npz" % n), A=A, row_sum=row_sum, pair_sum=pair_sum,
                        D_sha=sha(D), Z=Z, leaves=leaves, Dp_sha=sha(np.asarray(Dp)), S_sha=sha(S),
                        R=R.astype(np.int16), init=np.asarray(init), cuts=np.asarray(cuts),
                        init_minfrac05=np.asarray(init05),
                        scaffold_of_bin=g.scaffold_of_bin, true_chrom=g.true_chrom_of_bin,
                        meta=json.dumps(dict(n=n, n_chrom=nc, scaffolds_per_chrom=spc, seed=seed)))
    print("part1", n, "init", init, "cuts", cuts)
    return g, A


def part2_case(p2, g, A, n, nscaf, scan):
    G = np.asmatrix(A)
    bins = [p2.Bin(i + 1, g.scaffold_names[g.scaffold_of_bin[i]], 0, 0, 1., 0.) for i in range(n)]
    out = {"n": n, "nScaffolds": nscaf, "scanScaffolds": scan, "groups": []}
    for ch in sorted(set(g.true_chrom_of_bin.tolist())):
        nodes = [[i + 1, g.scaffold_names[g.scaffold_of_bin[i]]] for i in range(n) if g.true_chrom_of_bin[i] == ch]
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res = p2.orderChromosome([list(x) for x in nodes], G, bins, nscaf, scan)
        log = buf.getvalue()
        best = [ln for ln in log.splitlines() if ln.startswith("BestCost at the end of first two steps")]
        final = [ln for ln in log.splitlines() if ln.startswith("Best cost at the end of the final step")]
        idx = [i for i, _ in nodes]
        sub = np.ascontiguousarray(A[np.ix_(np.array(idx) - 1, np.array(idx) - 1)])
        total = sum([np.trace(np.asmatrix(sub), offset=i) for i in range(1, len(sub))])
        out["groups"].append({
            "nodes": nodes,
            "ordering": [[s.name, s.orientation] for s in res],
            "cost_after_insertion": float(best[0].split()[-1]) if best else None,
            "cost_after_scan": float(final[0].split(" = ")[-1]) if final else None,
            "total_identity": float(total),
            "cost_identity": float(p2.costFunction_numba(np.asmatrix(sub), total)),
        })
    with open(os.path.join(HERE, "part2_%d.json" % n), "w") as f:
        json.dump(out, f)
    print("part2", n, [len(gp["ordering"]) for gp in out["groups"]])


def pipeline_case(n=320, nc=8, spc=8, seed=7):
    d = os.path.join(HERE, "pipeline_%d" % n)
    os.makedirs(os.path.join(d, "in"), exist_ok=True)
    os.makedirs(os.path.join(d, "out"), exist_ok=True)
    g = make_genome(n, n_chrom=nc, seed=seed, scaffolds_per_chrom=spc)
    paths = write_hicpro_files(g, os.path.join(d, "in"), prefix="synth")
    cfg = os.path.join(d, "config.txt")
    rel = lambda p: p  # absolute paths are rewritten by the tests  # noqa: E731
    with open(cfg, "w") as f:
        f.write("\n".join([
            "resolution = %d" % g.resolution,
            "saveFilesDirectory = %s" % os.path.join(d, "out"),
            "savePlotsDirectory = %s" % os.path.join(d, "out"),
            "hicProBedFile = %s" % rel(paths["bed"]),
            "hicProBiasFile = %s" % rel(paths["biases"]),
            "hicProMatrixFile = %s" % rel(paths["matrix"]),
            "hicProScaffSizeFile = %s" % rel(paths["sizes"]),
            "dendrogramOrderFile = dendrogramOrder.txt",
            "avgClusterPlot = avgCluster.png",
            "avgClusterPlot_outlined = avgCluster_outlined.png",
            "binGroupFile = binGroups.txt",
            "assessmentFile = assessment.txt",
            "hyperGeom = True",
            "hmm = False",
            "minSize = 5",
            "modularity = 0",
            "psig = .05",
            "convergenceRounds = 5",
            "lookAhead = .2",
            "louvainRounds = 20",
            "chromosomeGroupFile = chromosomeGroups.txt",
            "chromosomeOrderFile = chromosomeOrder.txt",
            "chromosomePlotSuffix = synth",
            "fullGenomePlot = fullGenome.png",
            "fullGenomePlotTitle = synth",
            "plotOrderFile = plotOrder.txt",
            "nScaffolds = 4",
            "scanScaffolds = 3",
            "lengthCutoff = 500000",
            "restrictionSiteFile = NA",
            "validPairFile = NA",
            "finalOrderingsFile = finalOrderings.txt",
            "originalFastaFile = NA",
            "assembledFastaFile = NA",
        ]) + "\n")
    ref_harness.load_reference()
    script = os.path.join(ref_harness.REF_DIR, "run_hicAssembler.py")
    old = sys.argv
    sys.argv = [script, "-part1", "-part2", "-config", cfg]
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = old
    # make the config relocatable: the tests substitute @DIR@
    txt = open(cfg).read().replace(d, "@DIR@")
    open(cfg, "w").write(txt)
    print("pipeline", n, sorted(os.listdir(os.path.join(d, "out"))))


def main():
    p1, p2 = ref_harness.load_reference()
    for (n, nc, spc, seed, nscaf, scan) in [(60, 3, 4, 1, 4, 3), (200, 5, 6, 2, 4, 3), (320, 8, 8, 7, 5, 4)]:
        g, A = part1_case(p1, n, nc, spc, seed)
        part2_case(p2, g, A, n, nscaf, scan)
    pipeline_case()


if __name__ == "__main__":
    main()
