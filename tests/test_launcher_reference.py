"""The launcher on the REAL reference tree: the unmodified ``run_hicAssembler.py -part1 -part2`` with its hot-path
functions rebound (launcher.install), output files compared byte for byte with the files the unmodified reference
wrote for the same inputs (tests/golden/pipeline_320).

Two variants:
  * CPU (marker ``reference``; build container only): the seams run over a stand-in handle that answers every
    C-ABI call through the oracle.  This pins the host side of the drop-in - the device-backed lazy matrices, the
    numpy shim, the identity caches, the rebound orderGenome - against the reference's own pipeline code.
  * GPU: the same with the real libhica handle.  The reference tree is not part of this repository; it is picked
    up from ``baseline/_ref/HIC_ASSEMBLER`` (git-ignored, travels with gpurun) or /root/reference and the test is
    skipped when neither exists.
"""
import filecmp
import os
import runpy
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden", "pipeline_320")
OUT_FILES = ("dendrogramOrder.txt", "binGroups.txt", "assessment.txt", "chromosomeGroups.txt",
             "chromosomeOrder.txt", "plotOrder.txt")


def reference_dir():
    for d in (os.path.join(ROOT, "baseline", "_ref", "HIC_ASSEMBLER"), "/root/reference/HIC_ASSEMBLER"):
        if os.path.isfile(os.path.join(d, "run_hicAssembler.py")):
            return d
    return None


def write_config(tmp):
    lines = []
    for line in open(os.path.join(G, "config.txt")):
        line = line.rstrip("\n").replace("@DIR@/out", str(tmp)).replace("@DIR@", G)
        lines.append(line)
    path = os.path.join(str(tmp), "config.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    return path


def run_launcher(handle, ref, cfg, extra=()):
    """launcher.main() in-process with an injected handle; the reference modules are removed from sys.modules
    afterwards so that the next test imports fresh ones."""
    from hic_genome_assembler_b200 import launcher
    before = set(sys.modules)
    old_argv, old_path = sys.argv, list(sys.path)
    try:
        launcher.install(ref, no_plots=True, handle=handle, **dict(extra))
        script = os.path.join(ref, "run_hicAssembler.py")
        sys.argv = [script, "-part1", "-part2", "-config", cfg]
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = old_argv
        sys.path[:] = old_path
        for m in set(sys.modules) - before:
            if m in ("scaffoldToChromosomes", "orderGenome", "plotContactMaps", "orientSmallScaffolds", "writeAssembledFasta"):
                del sys.modules[m]


def compare_outputs(tmp):
    for name in OUT_FILES:
        assert filecmp.cmp(os.path.join(str(tmp), name), os.path.join(G, "out", name), shallow=False), name


class OracleHandle:
    """Answers the C-ABI calls the seams make through the oracle (CPU stand-in for tests of the HOST logic)."""

    def __init__(self):
        import hica_oracle as O
        self.O = O
        self.n = 0
        self.launches = 0
        self.calls = []

    def _log(self, name):
        self.calls.append(name)

    def p1_upload(self, A):
        self._log("p1_upload")
        self.A, self.kind, self.n = np.array(A, dtype=np.float64), 0, A.shape[0]

    def p1_upload_distance(self, D):
        self._log("p1_upload_distance")
        self.A, self.kind, self.n = np.array(D, dtype=np.float64), 1, D.shape[0]

    def p1_upload_similarity(self, S):
        self._log("p1_upload_similarity")
        self.A, self.kind, self.n = np.array(S, dtype=np.float64), 2, S.shape[0]
        self.order = np.arange(self.n)

    def p1_set_row_sums(self, rs):
        self.rs = self.O.sequential_row_sums(self.A) if rs is None else np.array(rs, dtype=np.float64)

    def p1_distance(self):
        self.D = self.O.convert_distance(self.A)

    def _D(self):
        return self.D if self.kind == 0 else self.A

    def p1_get_distance(self):
        self._log("p1_get_distance")
        return self._D().copy()

    def p1_linkage(self, want_Z=True):
        Z, leaves = self.O.average_cluster_leaves(self._D())
        self.order = np.asarray(leaves)
        return Z, np.asarray(leaves, dtype=np.int32)

    def p1_set_order(self, order):
        self.order = np.asarray(order)

    def p1_get_reordered(self, what):
        self._log("p1_get_reordered(%d)" % what)
        if self.kind == 2:
            return self.A.copy()
        Dp = self.O.reorder(self._D(), self.order)
        return Dp if what == 0 else self.O.convert_similarity(Dp, self.rs[self.order])

    def p1_rank(self):
        S = self.A if self.kind == 2 else self.O.convert_similarity(self.O.reorder(self._D(), self.order), self.rs[self.order])
        self.R = self.O.rank_matrix_stable(S)

    def p1_get_rank(self):
        self._log("p1_get_rank")
        return self.R.astype(np.int64)

    def p1_set_rank(self, R):
        self._log("p1_set_rank")
        self.R, self.n = np.array(R), R.shape[0]

    def p1_breakpoints(self, min_size, min_frac):
        return self.O.pre_process_breakpoints(self.R, min_size, min_frac), 0

    def p1_filter(self, cuts, psig):
        return [int(c) for c in self.O.filter_noisy_breakpoints(self.R, list(cuts), psig)], 0

    def p2_upload(self, G):
        self._log("p2_upload")
        self.G = np.array(G, dtype=np.float64)

    def p2_order_groups(self, grp_ptr, seg_ptr, bin_rows, n_scaffolds, scan, brute_tables, scan_tables):
        """Decode the packed groups (rows of G per scaffold segment, pool order) and order each with the oracle."""
        import fast_oracle as FO
        self._log("p2_order_groups")
        ng = len(grp_ptr) - 1
        layout = np.zeros(len(seg_ptr) - 1, dtype=np.int32)
        cost, rounds = np.zeros(ng), np.zeros(ng, dtype=np.int32)
        lut = np.arange(self.G.shape[0] + 1)          # "bin id" = row of G
        for g in range(ng):
            segs = list(range(grp_ptr[g], grp_ptr[g + 1]))
            # scaffold "names" are the local segment numbers; bins must stay in '+' order as given
            nodes = [(str(k), np.asarray(bin_rows[seg_ptr[s]:seg_ptr[s + 1]], dtype=np.int64)) for k, s in enumerate(segs)]
            pool = [FO.FScaf(nm, ids) for nm, ids in nodes]
            got, c, r = _order_pool(FO, self.G, lut, pool, n_scaffolds, scan)
            for slot, (nm, o) in enumerate(got):
                layout[grp_ptr[g] + slot] = (int(nm) << 1) | (1 if o == "-" else 0)
            cost[g], rounds[g] = c, r
        return layout, cost, rounds, np.zeros((ng, 4), dtype=np.int64)

    def close(self):
        pass


def _order_pool(FO, G, lut, pool, n_scaffolds, scan):
    """fast_oracle.order_chromosome for a pool that is already in initiateBinsAndScaffolds order."""
    n_scaffolds = 8 if n_scaffolds >= 9 else n_scaffolds
    scan = min(scan, n_scaffolds)
    ordered, pool = pool[:n_scaffolds], pool[n_scaffolds:]
    M, local = FO._sub(G, lut, ordered)
    od, orr, cost = FO.brute_force(ordered, M, local)
    for si, o in zip(od, orr):
        ordered[si].orientation = o
    ordered = [ordered[si] for si in od]
    best = cost
    while True:
        if pool:
            ordered.append(pool.pop(0))
        M, local = FO._sub(G, lut, ordered)
        new = ordered.pop(-1)
        ordered, best = FO.check_all_scores(M, local, ordered, new)
        if not pool:
            break
    rounds = 0
    if len(ordered) > n_scaffolds:
        ordered, best, rounds = FO.scan_ordering(G, lut, ordered, best, scan)
    return [(s.name, s.orientation) for s in ordered], best, rounds


@pytest.mark.reference
def test_reference_pipeline_over_oracle_backed_handle(tmp_path, capsys):
    ref = reference_dir()
    if ref is None:
        pytest.skip("reference tree not available")
    h = OracleHandle()
    run_launcher(h, ref, write_config(tmp_path))
    compare_outputs(tmp_path)
    # the fast lane: no N x N matrix and no rank matrix came back to the host, Part2 was ONE engine call
    assert not [c for c in h.calls if c.startswith("p1_get") or c == "p1_set_rank"], h.calls
    assert h.calls.count("p1_upload") == 1 and h.calls.count("p2_upload") == 1 and h.calls.count("p2_order_groups") == 1


@pytest.mark.gpu
@pytest.mark.parametrize("per_call", [False, True])
def test_reference_pipeline_on_gpu(hica, tmp_path, per_call):
    ref = reference_dir()
    if ref is None:
        pytest.skip("reference tree not available (baseline/_ref/HIC_ASSEMBLER or /root/reference)")
    l0 = hica.launches
    run_launcher(hica, ref, write_config(tmp_path), extra={"per_call_part2": per_call}.items())
    assert hica.launches > l0
    compare_outputs(tmp_path)


def _load_pipeline_maker():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_pipeline", os.path.join(ROOT, "tests", "golden", "make_golden_pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.gpu
def test_reference_pipeline_500kb_on_gpu(hica, tmp_path):
    """BASELINE.json configs[0] - the reference's own CPU-runnable case: 1.6 Gb genome at 500 kb (3,200 bins),
    `run_hicAssembler.py -part1 -part2` through the launcher on the GPU, six output files byte-identical to what the
    unmodified reference wrote in the build container (274 s there; tests/golden/pipeline_3200).  The HiC-Pro input
    files (120 MB of text, %.6f values - so exact ties exist) are regenerated here and checked by digest."""
    import json
    import time
    ref = reference_dir()
    meta_path = os.path.join(ROOT, "tests", "golden", "pipeline_3200", "meta.json")
    if ref is None or not os.path.exists(meta_path):
        pytest.skip("reference tree or pipeline_3200 golden not available")
    meta = json.load(open(meta_path))
    mk = _load_pipeline_maker()
    inp, out = os.path.join(str(tmp_path), "in"), os.path.join(str(tmp_path), "out")
    g = mk.write_inputs(meta["bins"], meta["chroms"], meta["seed"], inp)
    for f, want in meta["inputs_sha"].items():
        assert mk.file_sha(os.path.join(inp, f)) == want, "regenerated input differs: " + f
    os.makedirs(out)
    cfg = os.path.join(str(tmp_path), "config.txt")
    open(cfg, "w").write(mk.CONFIG.format(res=g.resolution, out=out, inp=inp))
    t0 = time.time()
    run_launcher(hica, ref, cfg)
    secs = time.time() - t0
    for name, want in meta["outputs_sha"].items():
        assert mk.file_sha(os.path.join(out, name)) == want, name
    rep = hica.tie_report()
    print("launcher -part1 -part2 at 500 kb: %.1f s (reference in the build container: %.0f s); tie report %s"
          % (secs, meta["reference_seconds_build_container"], rep))
    rec = os.environ.get("HICA_LAUNCHER_RECORD")
    if rec:
        json.dump({"bins": meta["bins"], "launcher_seconds": secs, "reference_seconds_build_container":
                   meta["reference_seconds_build_container"], "files_identical": True, "tie_report": rep}, open(rec, "w"))


@pytest.mark.reference
def test_reference_pipeline_500kb_over_oracle_backed_handle(tmp_path):
    """The same at the reference's own CPU-runnable size (BASELINE configs[0], 3,200 bins, %.6f inputs with exact
    ties) with the oracle answering the C-ABI calls: pins the host side of the fast lane against the six files the
    unmodified reference wrote (320 s in the build container; about 1 min here)."""
    import json
    ref = reference_dir()
    meta_path = os.path.join(ROOT, "tests", "golden", "pipeline_3200", "meta.json")
    if ref is None or not os.path.exists(meta_path):
        pytest.skip("reference tree or pipeline_3200 golden not available")
    meta = json.load(open(meta_path))
    mk = _load_pipeline_maker()
    inp, out = os.path.join(str(tmp_path), "in"), os.path.join(str(tmp_path), "out")
    g = mk.write_inputs(meta["bins"], meta["chroms"], meta["seed"], inp)
    for f, want in meta["inputs_sha"].items():
        assert mk.file_sha(os.path.join(inp, f)) == want, "regenerated input differs: " + f
    os.makedirs(out)
    cfg = os.path.join(str(tmp_path), "config.txt")
    open(cfg, "w").write(mk.CONFIG.format(res=g.resolution, out=out, inp=inp))
    h = OracleHandle()
    run_launcher(h, ref, cfg)
    for name, want in meta["outputs_sha"].items():
        assert mk.file_sha(os.path.join(out, name)) == want, name
    assert h.calls == ["p1_upload", "p2_upload", "p2_order_groups"], h.calls
