"""GPU: the launcher rebinding every seam of a (miniature, test-written) reference tree and running its entry
point; results are compared with the oracle.  See tests/mock_reference/README.txt."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import hica_oracle as O
from hic_genome_assembler_b200.synth import make_genome

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock_reference")


def test_launcher_rebinds_and_runs_both_parts(tmp_path):
    n, nc = 320, 4
    g = make_genome(n, n_chrom=nc, seed=12, scaffolds_per_chrom=9)
    A = g.A.numpy()
    np.save(tmp_path / "A.npy", A)
    names = [g.scaffold_names[s] for s in g.scaffold_of_bin]
    nodes = [[i + 1, names[i]] for i in range(n) if g.true_chrom_of_bin[i] == 1]
    json.dump(nodes, open(tmp_path / "nodes.json", "w"))
    cfg = dict(matrix=str(tmp_path / "A.npy"), nodes=str(tmp_path / "nodes.json"), out1=str(tmp_path / "o1.json"),
               out2=str(tmp_path / "o2.json"), minSize=5, modularity=0.0, psig=.05, nScaffolds=5, scanScaffolds=4)
    json.dump(cfg, open(tmp_path / "cfg.json", "w"))
    r = subprocess.run([sys.executable, "-m", "hic_genome_assembler_b200.launcher", "--reference-dir", MOCK,
                        "-part1", "-part2", "-config", str(tmp_path / "cfg.json")],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "kernels launched" in r.stdout
    ref = O.part1_cut_indices(A, min_size=5, min_frac=0.0, psig=.05)
    o1 = json.load(open(tmp_path / "o1.json"))
    assert o1["leaves"] == [int(v) for v in ref["leaves"]]
    assert o1["initial"] == ref["initial_cuts"] and o1["cuts"] == ref["cuts"]
    assert o1["R_row0"] == [int(v) for v in ref["R"][0][:10]]
    want, wcost = O.order_chromosome(A, {i + 1: i for i in range(n)}, [(b, s) for b, s in nodes], 5, 4)
    o2 = json.load(open(tmp_path / "o2.json"))
    assert [tuple(x) for x in o2["ordering"]] == want
    assert o2["bestCost"] == wcost
