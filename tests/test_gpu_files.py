"""GPU: the whole `-part1 -part2` flow from HiC-Pro text files, byte-for-byte against the files the UNMODIFIED
reference wrote for the same inputs (tests/golden/pipeline_320, produced by tests/golden/make_golden.py)."""
import filecmp
import os

import pytest

from hic_genome_assembler_b200 import io as hio

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden", "pipeline_320")


def _config(tmp):
    cfg = {}
    for line in open(os.path.join(G, "config.txt")):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        k, v = line.split(" = ")
        cfg[k] = v.replace("@DIR@", G)
    out = str(tmp)
    for k in ("dendrogramOrderFile", "binGroupFile", "assessmentFile", "chromosomeGroupFile", "chromosomeOrderFile",
              "plotOrderFile"):
        cfg[k] = os.path.join(out, cfg[k])
    return cfg


def test_output_files_match_reference(hica, tmp_path):
    cfg = _config(tmp_path)
    res = hio.run_files(hica, cfg)
    assert len(res["cuts"]) >= 1
    for name in ("dendrogramOrder.txt", "binGroups.txt", "assessment.txt", "chromosomeGroups.txt",
                 "chromosomeOrder.txt", "plotOrder.txt"):
        got, want = os.path.join(str(tmp_path), name), os.path.join(G, "out", name)
        assert filecmp.cmp(got, want, shallow=False), name
