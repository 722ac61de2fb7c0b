#!/usr/bin/env python
"""tests/golden/pipeline_<n>/: what the UNMODIFIED reference writes for `run_hicAssembler.py -part1 -part2` on a
synthetic genome at a published size (BASELINE.json configs[0]: 500 kb, 3,200 bins - the reference's own
CPU-runnable case).  Build container only (needs /root/reference; ~6 min at 3,200 bins).

    python tests/golden/make_golden_pipeline.py --bins 3200

The HiC-Pro input files are NOT stored (120 MB of text): they are regenerated on the test machine with
`write_inputs` below - synth.make_genome_portable through numpy is bit-identical everywhere and the values are
formatted with HiC-Pro's %.6f - and their SHA-256 digests are checked against the ones stored here.
"""
import argparse
import contextlib
import hashlib
import io
import json
import os
import runpy
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CONFIG = """resolution = {res}
saveFilesDirectory = {out}
savePlotsDirectory = {out}
hicProBedFile = {inp}/synth_abs.bed
hicProBiasFile = {inp}/synth_iced.matrix.biases
hicProMatrixFile = {inp}/synth_iced.matrix
hicProScaffSizeFile = {inp}/synth.sizes
dendrogramOrderFile = dendrogramOrder.txt
avgClusterPlot = avgCluster.png
avgClusterPlot_outlined = avgCluster_outlined.png
binGroupFile = binGroups.txt
assessmentFile = assessment.txt
hyperGeom = True
hmm = False
minSize = 5
modularity = 0
psig = .05
convergenceRounds = 5
lookAhead = .2
louvainRounds = 20
chromosomeGroupFile = chromosomeGroups.txt
chromosomeOrderFile = chromosomeOrder.txt
chromosomePlotSuffix = synth
fullGenomePlot = fullGenome.png
fullGenomePlotTitle = synth
plotOrderFile = plotOrder.txt
nScaffolds = 6
scanScaffolds = 5
lengthCutoff = 500000
restrictionSiteFile = NA
validPairFile = NA
finalOrderingsFile = finalOrderings.txt
originalFastaFile = NA
assembledFastaFile = NA
"""
OUT_FILES = ("dendrogramOrder.txt", "binGroups.txt", "assessment.txt", "chromosomeGroups.txt", "chromosomeOrder.txt",
             "plotOrder.txt")


def file_sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def write_inputs(n, n_chrom, seed, inp_dir, min_value=0.0):
    """The HiC-Pro style input files of the synthetic genome (numpy path of the portable generator).  Entries
    below ``min_value`` are left out of the sparse matrix file (they read back as 0)."""
    import numpy as np
    from hic_genome_assembler_b200.synth import make_genome_portable
    os.makedirs(inp_dir, exist_ok=True)
    g = make_genome_portable(n, n_chrom=n_chrom, seed=seed)
    A = g.A.numpy()
    res = g.resolution
    with open(os.path.join(inp_dir, "synth_abs.bed"), "w") as fb, open(os.path.join(inp_dir, "synth_iced.matrix.biases"), "w") as fbi:
        for i in range(n):
            name = g.scaffold_names[g.scaffold_of_bin[i]]
            st = int(g.bin_start[i])
            fb.write("%s\t%d\t%d\t%d\n" % (name, st, st + res, i + 1))
            fbi.write("%r\n" % float(g.bias[i]))
    with open(os.path.join(inp_dir, "synth_iced.matrix"), "w") as fm:
        for i in range(n):
            row = A[i, i:]
            js = np.nonzero(row >= min_value)[0] if min_value > 0 else np.nonzero(row != 0.0)[0]
            fm.write("".join("%d\t%d\t%.6f\n" % (i + 1, i + 1 + j, row[j]) for j in js.tolist()))
    with open(os.path.join(inp_dir, "synth.sizes"), "w") as fs:
        for nm, sz in zip(g.scaffold_names, g.scaffold_sizes_bp):
            fs.write("%s\t%d\n" % (nm, int(sz)))
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", type=int, default=3200)
    ap.add_argument("--chroms", type=int, default=20)
    ap.add_argument("--seed", type=int, default=7)
    args = ap.parse_args()
    import ref_harness
    d = os.path.join(HERE, "pipeline_%d" % args.bins)
    tmp = "/tmp/hica_pipeline_%d" % args.bins
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(os.path.join(d, "out"), exist_ok=True)
    t0 = time.time()
    g = write_inputs(args.bins, args.chroms, args.seed, os.path.join(tmp, "in"))
    print("inputs written in %.0fs" % (time.time() - t0), flush=True)
    os.makedirs(os.path.join(tmp, "out"), exist_ok=True)
    cfg = os.path.join(tmp, "config.txt")
    open(cfg, "w").write(CONFIG.format(res=g.resolution, out=os.path.join(tmp, "out"), inp=os.path.join(tmp, "in")))
    ref_harness.load_reference()
    script = os.path.join(ref_harness.REF_DIR, "run_hicAssembler.py")
    old = sys.argv
    sys.argv = [script, "-part1", "-part2", "-config", cfg]
    t0 = time.time()
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = old
    secs = time.time() - t0
    print("reference -part1 -part2: %.0fs" % secs, flush=True)
    for name in OUT_FILES:
        shutil.copy(os.path.join(tmp, "out", name), os.path.join(d, "out", name))
    meta = dict(bins=args.bins, chroms=args.chroms, seed=args.seed, resolution=g.resolution,
                reference_seconds_build_container=round(secs, 1),
                inputs_sha={f: file_sha(os.path.join(tmp, "in", f)) for f in sorted(os.listdir(os.path.join(tmp, "in")))},
                outputs_sha={f: file_sha(os.path.join(d, "out", f)) for f in OUT_FILES})
    json.dump(meta, open(os.path.join(d, "meta.json"), "w"), indent=1)
    print(json.dumps(meta["outputs_sha"], indent=1))


if __name__ == "__main__":
    main()
