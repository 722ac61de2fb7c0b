#!/usr/bin/env python
"""The unmodified `run_hicAssembler.py -part1 -part2` through the launcher at the headline size (100 kb, 16,000 bins)
on a GPU box: wall times by stage and a byte-compare of the six output files with this repository's own file-level
path (io.run_files).  The reference itself cannot produce files at this size in a reasonable time (hours), so the
comparison partner at 100 kb is run_files, which is checked against the reference's files at 320 and 3,200 bins.

    python profiles/launcher_100kb.py [out.json]        # needs baseline/_ref/HIC_ASSEMBLER and a GPU

Inputs: synth.make_genome_portable (torch on the GPU), every entry of the upper triangle written as a HiC-Pro
triplet with six decimals (128 M lines, ~3.3 GB; pyarrow's CSV writer - shortest representation of the rounded
value, which parses to the same double as HiC-Pro's %.6f)."""
import filecmp
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def main():
    import pyarrow as pa
    import pyarrow.csv as pacsv
    from hic_genome_assembler_b200.synth import make_genome_portable
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_golden_pipeline.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    ref = os.path.join(ROOT, "baseline", "_ref", "HIC_ASSEMBLER")
    n = int(os.environ.get("HICA_LAUNCHER_BINS", "16000"))
    tmp = tempfile.mkdtemp(prefix="hica_launcher_")
    inp = os.path.join(tmp, "in")
    os.makedirs(inp)
    rec = {"bins": n}
    t0 = time.time()
    g = make_genome_portable(n, n_chrom=20, seed=7, device="cuda:0")
    A = np.round(g.A.cpu().numpy(), 6)
    with open(os.path.join(inp, "synth_abs.bed"), "w") as fb, open(os.path.join(inp, "synth_iced.matrix.biases"), "w") as fbi:
        for i in range(n):
            st = int(g.bin_start[i])
            fb.write("%s\t%d\t%d\t%d\n" % (g.scaffold_names[g.scaffold_of_bin[i]], st, st + g.resolution, i + 1))
            fbi.write("%r\n" % float(g.bias[i]))
    with open(os.path.join(inp, "synth.sizes"), "w") as fs:
        for nm, sz in zip(g.scaffold_names, g.scaffold_sizes_bp):
            fs.write("%s\t%d\n" % (nm, int(sz)))
    iu = np.triu_indices(n)
    tbl = pa.table({"i": (iu[0] + 1).astype(np.int32), "j": (iu[1] + 1).astype(np.int32), "v": A[iu]})
    pacsv.write_csv(tbl, os.path.join(inp, "synth_iced.matrix"),
                    write_options=pacsv.WriteOptions(include_header=False, delimiter="\t", quoting_style="none"))
    del tbl, iu
    rec["write_inputs_seconds"] = time.time() - t0
    rec["matrix_file_GB"] = os.path.getsize(os.path.join(inp, "synth_iced.matrix")) / 1e9
    outs = {}
    for tag in ("launcher", "run_files"):
        outs[tag] = os.path.join(tmp, "out_" + tag)
        os.makedirs(outs[tag])
    cfg = os.path.join(tmp, "config.txt")
    open(cfg, "w").write(mk.CONFIG.format(res=g.resolution, out=outs["launcher"], inp=inp))
    t0 = time.time()
    r = subprocess.run([sys.executable, "-m", "hic_genome_assembler_b200.launcher", "--reference-dir", ref, "--no-plots",
                        "-part1", "-part2", "-config", cfg], cwd=ROOT, capture_output=True, text=True)
    rec["launcher_wall_seconds"] = time.time() - t0
    rec["launcher_rc"] = r.returncode
    keep = ("Total run-time", "Time to", "RunTime", "kernels launched", "WARNING", "CutIndices")
    rec["launcher_stdout_times"] = [ln for ln in r.stdout.splitlines() if any(k in ln for k in keep)][:40]
    if r.returncode:
        rec["launcher_stderr_tail"] = r.stderr[-2000:]
    # the same files through io.run_files (no reference code at all)
    from hic_genome_assembler_b200 import io as hio
    from hic_genome_assembler_b200._lib import Handle
    c = {}
    for line in open(cfg):
        if " = " in line:
            k, v = line.rstrip("\n").split(" = ")
            c[k] = v
    for k in ("dendrogramOrderFile", "binGroupFile", "assessmentFile", "chromosomeGroupFile", "chromosomeOrderFile", "plotOrderFile"):
        c[k] = os.path.join(outs["run_files"], c[k])
    t0 = time.time()
    with Handle(0) as h:
        res = hio.run_files(h, c)
        rec["tie_report"] = h.tie_report()
    rec["run_files_wall_seconds"] = time.time() - t0
    rec["cuts"] = res["cuts"]
    rec["files_identical_to_run_files"] = {f: filecmp.cmp(os.path.join(outs["launcher"], f), os.path.join(outs["run_files"], f), shallow=False)
                                           for f in mk.OUT_FILES if os.path.exists(os.path.join(outs["launcher"], f))}
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_launcher_100kb.json")
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1)[:3000])


if __name__ == "__main__":
    main()
