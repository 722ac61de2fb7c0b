#!/usr/bin/env python
"""bench.py - Part1+Part2 hot path of HiC_Genome_Assembler on B200 ("Measurement").

A step = one pass of the hot path over one synthetic ICE-normalised genome (SURVEY.md 8d recipe through
synth.make_genome_portable, seed 7 - bit-identical on every platform, so the result digests in every line can be
compared with the committed golden vectors tests/golden/large_<bins>.npz):
  Part1: distance transform -> UPGMA linkage -> leaf order -> similarity rank -> hypergeometric cut indices
  Part2: per chromosome group brute-force + greedy insertion + sliding-window ordering search
Default workload = BASELINE.json configs[1]/metric: 1.6 Gb genome at 100 kb (16,000 bins, 20 chromosomes,
40 scaffolds per chromosome).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--bins 16000] [--impl reference]
    python bench.py --config candidates [--gpus N]     BASELINE configs[2]: 10^5 exactly scored candidates
    python bench.py --config 10kb --gpus 8             BASELINE configs[4]: 160,000 bins, fp32, row-block sharded

metric  : bin-pairs/s over Part1+Part2 (Part1 pairs = N(N-1)/2; Part2 pairs = sum over the candidates THE
          REFERENCE evaluates of L(L-1)/2 - reference-equivalent work, see `work.pairs_touched` for what the GPU
          path actually reads), plus the absolute seconds in `seconds`.
value   : input matrix resident in HBM when the timed region starts.
e2e     : the same step through the public host API with the matrix in pinned HOST memory (H2D inside the
          timed region, results read back to the host); `e2e_pageable` the same from ordinary numpy memory.
digests : SHA-256 of leaves / orderings and the cut lists of the LAST timed step, checked against the golden file.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


# DRAM bytes of one linkage-kernel launch measured with `ncu --set full` (bins -> bytes):
# profiles/r02_ncu_full_summary.txt, nn4::k_nn_chain4 (1.588 GB read + 0.659 GB written; k_nn_chain3 of round 1: 4.95 GB)
NCU_DRAM_BYTES = {16000: 2.247e9}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bins", type=int, default=16000)
    ap.add_argument("--chroms", type=int, default=20)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--impl", default="hica", choices=["hica", "reference"])
    ap.add_argument("--config", default="step", choices=["step", "candidates", "10kb"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--part1-only", action="store_true")
    ap.add_argument("--replicate-part1", action="store_true", help="N > 1: every rank runs the whole of Part1 (no row-block sharding)")
    ap.add_argument("--p2-host-loop", action="store_true", help="Part2 through per-call host loops (exact scoring of every candidate)")
    ap.add_argument("--no-golden-check", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- workload
def build_workload(n_bins, n_chrom, seed, device):
    from hic_genome_assembler_b200.synth import make_genome_portable
    g = make_genome_portable(n_bins, n_chrom=n_chrom, seed=seed, device=device)
    names = [g.scaffold_names[s] for s in g.scaffold_of_bin]
    sizes = {nm: int(sz) for nm, sz in zip(g.scaffold_names, g.scaffold_sizes_bp)}
    bin_ids = np.arange(1, n_bins + 1)
    row_of = np.full(n_bins + 2, -1, dtype=np.int64)
    row_of[bin_ids] = np.arange(n_bins)
    return dict(gen=g, A=g.A, bin_ids=bin_ids, scaffold_of_bin=names, sizes=sizes, resolution=g.resolution,
                scaffold_code=np.asarray(g.scaffold_of_bin, dtype=np.int64), scaffold_names=list(g.scaffold_names),
                scaffold_sizes=[int(v) for v in g.scaffold_sizes_bp], row_of_bin_id=row_of)


def workload_name(args):
    kb = int(round(1.6e9 / args.bins / 1000))
    return "synthetic ICE genome 1.6Gb %d chr @%dkb (%d bins), -part1 -part2" % (args.chroms, kb, args.bins)


def config_dict(args):
    """The workload description both arms print (same keys and values in `--impl reference`)."""
    n = args.bins
    return {"workload": workload_name(args), "bins": n, "chromosomes": args.chroms, "seed": args.seed,
            "generator": "synth.make_genome_portable",
            "part2": "skipped" if args.part1_only else "nScaffolds=6 scanScaffolds=5",
            "l2": "inputs (%.2f GB matrix) larger than L2; no explicit flush" % (8.0 * n * n / 1e9)}


# ----------------------------------------------------------------------------- digests / golden vectors
def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def orderings_digest(orderings) -> str:
    txt = json.dumps([[[str(nm), str(o)] for nm, o in grp] for grp in orderings], separators=(",", ":"))
    return hashlib.sha256(txt.encode()).hexdigest()


def load_golden(args):
    p = os.path.join(ROOT, "tests", "golden", "large_%d.npz" % args.bins)
    if not os.path.exists(p):
        return None
    meta = json.loads(str(np.load(p)["meta"]))
    if meta["chroms"] != args.chroms or meta["seed"] != args.seed:
        return None
    meta["file"] = os.path.relpath(p, ROOT)
    return meta


def digests_of(step_result, part1_only):
    d = {"leaves_sha": sha(np.asarray(step_result["leaves"], dtype=np.int32)),
         "initial_cuts": [int(c) for c in step_result["initial_cuts"]], "cuts": [int(c) for c in step_result["cuts"]]}
    if not part1_only:
        od = step_result["orderings"]
        d["orderings_sha"] = orderings_digest([od[i][0] for i in range(len(od))])
        d["costs_sha"] = sha(np.asarray([od[i][1] for i in range(len(od))], dtype=np.float64))
    return d


def golden_check(dig, golden, part1_only):
    """Compare the digests of this run with the committed golden vectors (oracle at the reference's call sites,
    made in the build container by tests/golden/make_golden_large.py)."""
    if golden is None:
        return {"file": None, "equal": None, "note": "no golden vectors committed for this size"}
    bad = []
    if dig["leaves_sha"] != golden["leaves_sha"]:
        bad.append("leaves")
    if dig["initial_cuts"] != golden["initial_cuts"]:
        bad.append("initial_cuts")
    if dig["cuts"] != golden["cuts"]:
        bad.append("cuts")
    if not part1_only and golden.get("orderings_sha"):
        if dig["orderings_sha"] != golden["orderings_sha"]:
            bad.append("orderings")
        if dig["costs_sha"] != sha(np.asarray(golden["costs"], dtype=np.float64)):
            bad.append("costs")
    return {"file": golden["file"], "equal": not bad, "mismatch": bad}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []   # (arrival time, csv line)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, windows=None):
        """windows: [(t0, t1), ...] in time.perf_counter() seconds - only samples taken inside a timed region count
        (the sampler is started before the warm-up so that nvidia-smi's start-up time does not eat the region)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for ts, s in self.samples:
            if windows and not any(a <= ts <= b for a, b in windows):
                continue
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampled": "inside the timed regions (value and e2e), every 25 ms"}


# ----------------------------------------------------------------------------- one GPU step
def gpu_step(h, wl, args, source, rank, world, dist):
    """source = ('device', ptr) or ('host', ndarray - pinned or pageable).  Returns a dict of results and timings."""
    from hic_genome_assembler_b200 import part1, part2
    n = args.bins
    t0 = time.perf_counter()
    from hic_genome_assembler_b200 import dist as hdist
    import torch
    dev = torch.device("cuda", h.device)
    if source[0] == "device":
        h.p1_upload_device(source[1], n)
        up = None
    else:
        # the contact map is symmetric by construction (the reference writes [i][j] and [j][i] together):
        # only its upper triangle crosses PCIe, the library mirrors it on the device.  With several GPUs every
        # rank uploads its own row block and the matrix is assembled over NVLink (dist.upload_sharded).
        up = hdist.upload_sharded(h, source[1], dev, symmetric=True)
    if world > 1 and not args.replicate_part1:
        res = hdist.run_part1_resident_sharded(h, dev, min_size=5, min_frac=0.0, psig=.05)
    else:
        res = part1.run_part1_resident(h, min_size=5, min_frac=0.0, psig=.05)
    res["upload"] = up
    t1 = time.perf_counter()
    tm1 = h.timers()
    out = {"leaves": res["leaves"], "cuts": res["cuts"], "initial_cuts": res["initial_cuts"],
           "ambiguous": res["ambiguous"], "tie_sensitive": res.get("tie_sensitive"), "t_part1": t1 - t0, "timers_p1": tm1,
           "linkage_stats": h.linkage_stats(), "upload": res.get("upload"), "stage_ms_sharded": res.get("stage_ms")}
    p2_pairs, p2_cands, p2_calls = 0, 0, 0
    orderings = {}
    if not args.part1_only:
        if args.p2_host_loop:
            groups = part1.chromosome_groups(res["cuts"], res["leaves"], wl["bin_ids"], wl["scaffold_of_bin"], wl["sizes"])
            index_of = {int(b): i for i, b in enumerate(wl["bin_ids"])}
        else:
            groups = part1.chromosome_groups_compact(res["cuts"], res["leaves"], wl["bin_ids"], wl["scaffold_code"],
                                                     wl["scaffold_names"], wl["scaffold_sizes"])
            index_of = wl["row_of_bin_id"]
        h.p2_use_p1_matrix()
        part2.flush_stats()
        part2.STATS.update(candidates=0, bin_pairs=0, calls=0, exact=0, exact_pairs=0)
        nbins = [len(g) if args.p2_host_loop else sum(len(ids) for _, ids in g) for g in groups]
        mine = hdist.shard_groups(nbins, world, rank)
        if args.p2_host_loop:
            for ci in mine:
                orderings[ci] = part2.order_chromosome(h, index_of, groups[ci], 6, 5)
        elif mine:
            got, cost, rounds, stats = part2.order_groups_device(h, index_of, [groups[ci] for ci in mine], 6, 5)
            for k, ci in enumerate(mine):
                orderings[ci] = (got[k], float(cost[k]), int(rounds[k]))
            out["p2_host_ms"] = h.debug_stats()
        if world > 1:  # every rank ends the step with all orderings (one small all_gather_object)
            orderings = dict(enumerate(hdist.merge_orderings(orderings, len(groups))))
        out["n_groups"] = len(groups)
    t2 = time.perf_counter()
    if not args.part1_only:  # work accounting of the step (reference-equivalent pairs): outside the timed region
        st = part2.flush_stats()
        out["p2_exact"] = st.get("exact", 0)
        out["p2_exact_pairs"] = st.get("exact_pairs", 0)
        p2_pairs, p2_cands, p2_calls = st["bin_pairs"], st["candidates"], st["calls"]
    out.update(orderings=orderings, t_part2=t2 - t1, t_total=t2 - t0, p2_pairs=p2_pairs, p2_cands=p2_cands,
               p2_calls=p2_calls)
    return out


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
class CpuReference:
    """The reference's CPU path through the oracle (numpy/scipy/numba calls at the reference's call sites, one
    thread - the reference is single-threaded) on the SAME workload.

    full()    the matrix stages P1a-P1g (distance, squareform + linkage, leaf order, reorder, similarity,
              argsort) timed ONCE at the full bin count - nothing extrapolated.
    sample()  a bounded sample of the two stages that do not fit a few minutes of CPU: break searches of the
              hypergeometric stage (P1h/P1i) at the full size - scaled to the whole search sequence by the
              known cost model sum_k (N - start_k)^2 over the golden/GPU cut list - and Part2 scan candidates at
              the chromosome size, scaled to the candidate pairs the reference evaluates (part2.reference_work).
              The noise filter (P1j) is not timed: its share is reported as 0, which favours the CPU.
    Extrapolated terms are labelled in `sample`."""

    def __init__(self, A_host, args, starts, p2_pairs_full, p2_source):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import hica_oracle as O
        self.O, self.A, self.args = O, A_host, args
        self.starts = [0] + [int(c) for c in starts]
        self.p2_pairs_full, self.p2_source = int(p2_pairs_full), p2_source
        self.stage, self.R = None, None

    def full(self):
        O, A = self.O, self.A
        st = {}
        t = time.perf_counter()
        rs = O.sequential_row_sums(A)
        st["row_sums(removeRows)"] = time.perf_counter() - t
        t = time.perf_counter()
        D = O.convert_distance(A)
        st["P1a convertMatrix(distance)"] = time.perf_counter() - t
        t = time.perf_counter()
        y = O.condensed_upper(D)
        st["P1b squareform"] = time.perf_counter() - t
        t = time.perf_counter()
        import scipy.cluster.hierarchy as sch
        Z = sch.average(y)
        del y
        st["P1c hierarchy.average"] = time.perf_counter() - t
        t = time.perf_counter()
        leaves = O.leaf_order_scipy(Z)
        st["P1d dendrogram"] = time.perf_counter() - t
        t = time.perf_counter()
        Dp = O.reorder(D, leaves)
        del D
        st["P1e reorderMatrix"] = time.perf_counter() - t
        t = time.perf_counter()
        S = O.convert_similarity(Dp, rs[np.asarray(leaves)])
        del Dp
        st["P1f convertMatrix(similarity)"] = time.perf_counter() - t
        t = time.perf_counter()
        self.R = O.rank_matrix(S)
        del S
        st["P1g argsort"] = time.perf_counter() - t
        self.stage = st
        self.leaves_sha = sha(np.asarray(leaves, dtype=np.int32))
        return st

    def sample(self, n_searches=1, n_cands=150):
        O, A, args = self.O, self.A, self.args
        n = A.shape[0]
        if self.stage is None:
            self.full()
        t_matrix = sum(self.stage.values())
        # ---- hypergeometric break searches at the full size (the first n_searches starts of the sequence)
        t0 = time.perf_counter()
        done = []
        for s in self.starts[:n_searches]:
            O.find_breakpoints(self.R, s, 5, n - s, .05)
            done.append(s)
        t_cuts = time.perf_counter() - t0
        model = lambda ss: float(sum((n - s) ** 2 for s in ss))  # noqa: E731  window counts: sum_i (i - start) rank reads
        t_cuts_full = t_cuts * model(self.starts) / model(done)
        # ---- Part2: scan candidates of one chromosome-sized block of the same matrix
        t_p2_full, rate_p2, L = 0.0, None, max(8, n // args.chroms)
        if self.p2_pairs_full:
            M = np.ascontiguousarray(A[:L, :L])
            total = O.total_weight(M)
            rng = np.random.RandomState(0)
            O.ordering_cost_numba(M[:16, :16], 1.0)  # JIT warm-up outside the timed region
            t0 = time.perf_counter()
            for _ in range(n_cands):
                o = np.arange(L)
                a = rng.randint(0, L - 100) if L > 120 else 0
                o[a:a + min(100, L)] = o[a:a + min(100, L)][::-1]
                O.ordering_cost_numba(M[np.ix_(o, o)], total)
            t_p2 = time.perf_counter() - t0
            rate_p2 = n_cands * (L * (L - 1) // 2) / t_p2
            t_p2_full = self.p2_pairs_full / rate_p2
        est = t_matrix + t_cuts_full + t_p2_full
        full_p1 = n * (n - 1) // 2
        return {"value": (full_p1 + self.p2_pairs_full) / est, "unit": "bin-pairs/s", "cores": 1, "kind": "port",
                "sample": "oracle = numpy/scipy/numba at the reference's call sites, 1 thread (the reference is "
                          "single-threaded). MEASURED in full at %d bins, once: matrix stages P1a-P1g %.1f s. SAMPLED per "
                          "step: %d of %d break searches at full size (%.2f s, scaled by sum (N-start)^2 to %.1f s; noise "
                          "filter not counted); %d Part2 scan candidates at L=%d (scaled to the %.3g candidate bin-pairs "
                          "the reference evaluates, counted by part2.reference_work from %s)"
                          % (n, t_matrix, len(done), len(self.starts), t_cuts, t_cuts_full, n_cands, L,
                             float(self.p2_pairs_full), self.p2_source),
                "stage_seconds_full_size": {k: round(v, 3) for k, v in self.stage.items()},
                "cuts_seconds_extrapolated": t_cuts_full, "part2_seconds_extrapolated": t_p2_full,
                "part2_pairs_per_s": rate_p2, "est_full_step_seconds": est, "host_cpu_count": os.cpu_count(),
                "leaves_sha": self.leaves_sha}


def reference_p2_pairs(golden, args):
    """Candidate bin pairs the reference evaluates in Part2, from the golden vectors' groups and scan rounds."""
    from hic_genome_assembler_b200 import part2
    if args.part1_only:
        return 0, "part1 only"
    if golden and golden.get("rounds"):
        total = 0
        for grp, r in zip(golden["groups"], golden["rounds"]):
            lens = sorted((int(l) for _, l in grp), reverse=True)
            total += part2.reference_work(lens, 6, 5, int(r))[1]
        return total, golden["file"]
    # no golden vectors for this size: 20 equal chromosomes of 40 equal scaffolds, 1.5 scan rounds (labelled)
    L = args.bins // args.chroms
    lens = [max(1, L // 40)] * 40
    return args.chroms * part2.reference_work(lens, 6, 5, 2)[1], "a model of %d equal groups (no golden vectors for this size)" % args.chroms


def score_kernel_probe(h, wl, last):
    """The exact ordering-score kernel on its own (the device-resident search only re-scores a short list with it):
    one scan window of the largest chromosome group, every candidate scored exactly (hica_p2_score_window),
    3 timed calls after one warm-up.  Algorithmic bytes = candidates x L(L-1)/2 contacts x 8 B; the 5 MB chromosome
    matrix is L2-resident, so the figure is an L2 rate quoted against the HBM peak (profiles/ holds the kernel's
    dram__bytes and lts__t_bytes)."""
    from hic_genome_assembler_b200 import part1, part2
    groups = part1.chromosome_groups_compact(last["cuts"], last["leaves"], wl["bin_ids"], wl["scaffold_code"],
                                             wl["scaffold_names"], wl["scaffold_sizes"])
    groups = [g for g in groups if len(g) >= 5]
    if not groups:
        return None
    g = max(groups, key=lambda g: sum(len(ids) for _, ids in g))
    pool = sorted(g, key=lambda t: len(t[1]), reverse=True)
    rows = np.concatenate([wl["row_of_bin_id"][ids] for _, ids in pool])
    ptr = np.cumsum([0] + [len(ids) for _, ids in pool])
    L = int(ptr[-1])
    h.p2_use_p1_matrix()
    h.p2_select(rows.astype(np.int32))
    h.p2_set_segments(ptr.astype(np.int32), np.arange(L, dtype=np.int32))
    base = [sg << 1 for sg in range(len(pool))]
    total = h.p2_total(base)
    orders, orients = part2.order_table(5), part2.orient_table(5)
    ms = []
    for it in range(4):
        h.p2_score_window(base, 0, 5, orders, orients, total, 0.0)
        if it:
            ms.append(h.timers()["p2_score"])
    ncand = orders.shape[0] * orients.shape[0]
    return {"ms": float(np.median(ms)), "alg_bytes": 8.0 * ncand * (L * (L - 1) // 2), "candidates": int(ncand), "bins": L}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- main
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libhica has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.config == "candidates":
        import bench_extra
        return bench_extra.candidates_config(args, rank, world, local_rank)
    if args.config == "10kb":
        import bench_extra
        return bench_extra.config_10kb(args, rank, world, local_rank)
    from hic_genome_assembler_b200._lib import Handle

    dev = torch.device("cuda", local_rank)
    wl = build_workload(args.bins, args.chroms, args.seed, dev)
    A_dev = wl["A"]
    A_pinned = torch.empty(A_dev.shape, dtype=torch.float64, pin_memory=True)
    A_pinned.copy_(A_dev)
    torch.cuda.synchronize()
    wl["A_host"] = A_pinned.numpy()
    h = Handle(local_rank)
    golden = None if args.no_golden_check else load_golden(args)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(source, steps):
        results = []
        barrier()
        launches0 = h.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            results.append(gpu_step(h, wl, args, source, rank, world, dist))
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        t = torch.tensor([max(wall, dev_ms / 1e3)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return results, float(t.item()), h.launches - launches0

    sampler = ClockSampler(local_rank)
    sampler.start()
    # warm-up (also JIT-free: everything is precompiled)
    for _ in range(max(args.warmup, 3)):
        gpu_step(h, wl, args, ("device", A_dev.data_ptr()), rank, world, dist)

    w0 = time.perf_counter()
    res_dev, secs_dev, launches = timed(("device", A_dev.data_ptr()), args.steps)
    windows = [(w0, time.perf_counter())]
    # the host-source path has its own first-use costs (NCCL communicators of the row-block gather, pinned
    # staging): warm it up as well, untimed
    for _ in range(2):
        gpu_step(h, wl, args, ("host", wl["A_host"]), rank, world, dist)
    w0 = time.perf_counter()
    res_e2e, secs_e2e, _ = timed(("host", wl["A_host"]), args.steps)
    windows.append((w0, time.perf_counter()))
    clocks = sampler.stop(windows)
    # the same from ordinary (pageable) numpy memory, as a caller that never pinned anything would have it
    A_pageable = np.array(wl["A_host"], copy=True)
    gpu_step(h, wl, args, ("host", A_pageable), rank, world, dist)
    res_pg, secs_pg, _ = timed(("host", A_pageable), min(args.steps, 3))
    del A_pageable

    # ---- aggregate work over ranks
    n = args.bins
    p1_pairs = n * (n - 1) // 2
    last = res_dev[-1]
    p2 = torch.tensor([last["p2_pairs"], last["p2_cands"], last["p2_calls"], last.get("p2_exact", 0) or 0,
                       last.get("p2_exact_pairs", 0) or 0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(p2, op=dist.ReduceOp.SUM)
    p2_pairs, p2_cands, p2_calls, p2_exact, p2_exact_pairs = (int(v) for v in p2.tolist())
    total_pairs = p1_pairs + p2_pairs
    ms_per_step = secs_dev / args.steps * 1e3
    value = total_pairs / (secs_dev / args.steps)
    e2e_value = total_pairs / (secs_e2e / args.steps)

    # ---- result digests: every rank holds the full result; all ranks and both paths must agree
    dig = digests_of(last, args.part1_only)
    same = [digests_of(r, args.part1_only) == dig for r in res_dev + res_e2e + res_pg]
    agree = torch.tensor([int(all(same)), int(dig["leaves_sha"][:12], 16) % (1 << 40),
                          int(dig.get("orderings_sha", "0")[:12], 16) % (1 << 40)], dtype=torch.float64, device=dev)
    if world > 1:
        parts = [torch.empty_like(agree) for _ in range(world)]
        dist.all_gather(parts, agree)
        ranks_agree = all(bool(p[0] == 1) and bool((p[1:] == agree[1:]).all()) for p in parts)
    else:
        ranks_agree = bool(all(same))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    tm = last["timers_p1"]
    peak, peak_src = hbm_peak()
    link_bytes = 24.0 * n * n
    link_ms = tm["linkage"]
    achieved = link_bytes / (link_ms / 1e3) / 1e9 if link_ms > 0 else 0.0
    kernels = {
        "distance(K1)": {"ms": tm["distance"], "alg_bytes": 16.0 * n * n},
        "linkage(K2)": {"ms": tm["linkage"], "alg_bytes": link_bytes},
        # the value sort runs on a second stream next to the linkage chain (rank_presort), the part after the
        # leaf order is known (rank) is a column relabel: the kernel's own rate uses the sum of both
        "rank(K5)": {"ms": tm["rank"] + tm.get("rank_presort", 0.0), "alg_bytes": 12.0 * n * n,
                     "overlapped_with_linkage_ms": tm.get("rank_presort", 0.0), "after_linkage_ms": tm["rank"]},
    }
    if not args.part1_only:
        probe = score_kernel_probe(h, wl, last)
        if probe:
            kernels["score(K8: exact ordering score, one candidate per CTA)"] = probe
    for k in kernels.values():
        k["GBps"] = k["alg_bytes"] / (k["ms"] / 1e3) / 1e9 if k["ms"] > 0 else 0.0
        k["frac_of_hbm_peak"] = k["GBps"] / peak
    gcheck = golden_check(dig, golden, args.part1_only)
    cfg = config_dict(args)
    sharding = ("Part1 replicated per rank" if args.replicate_part1 else
                       "row-block sharded upload + NVLink all-gather of the matrix, linkage on rank 0 + leaf-order "
                       "broadcast, " + ("rank: value sort on every GPU while the chain runs, local column relabel"
                                        if n <= 65535 and os.environ.get("HICA_RANK_OVERLAP", "1") != "0"
                                        else "rank matrix by row block + all-gather") +
                       ", cut searches replicated") + "; Part2 chromosome groups dealt to ranks largest first" if world > 1 else "1 GPU"
    h2d = int(sum(8 * min(512, n - r0) * (n - r0) for r0 in range(0, n, 512)))
    line = {
        "metric": "Part1+Part2 bin-pairs/s (and secs) at %dkb" % int(round(1.6e9 / n / 1000)),
        "value": value, "unit": "bin-pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "value_note": "reference-equivalent bin pairs per second: Part2 counts every candidate the reference evaluates "
                      "(part2.reference_work); the GPU path filters them in closed form and reads `work.pairs_touched`. "
                      "`seconds` is the plain figure.",
        "config": cfg, "sharding": sharding,
        "digests": dict(dig, all_steps_and_paths_equal=bool(all(same)), all_ranks_equal=ranks_agree),
        "golden": gcheck,
        "seconds": {"part1": last["t_part1"], "part2": last["t_part2"], "total": last["t_total"],
                    "e2e_total": secs_e2e / args.steps, "e2e_part1": res_e2e[-1]["t_part1"],
                    "e2e_part2": res_e2e[-1]["t_part2"], "e2e_upload_ms": res_e2e[-1]["timers_p1"]["upload"],
                    "e2e_pageable_total": secs_pg / len(res_pg), "e2e_pageable_upload_ms": res_pg[-1]["timers_p1"]["upload"],
                    "e2e_upload_rank0": res_e2e[-1].get("upload"), "sharded_stage_ms_rank0": last.get("stage_ms_sharded")},
        "stage_ms_part1": tm, "linkage_stats": last["linkage_stats"],
        "per_step_seconds": [[round(r["t_part1"], 4), round(r["t_part2"], 4)] for r in res_dev],
        "per_step_seconds_e2e": [[round(r["t_part1"], 4), round(r["t_part2"], 4)] for r in res_e2e],
        "part2_host_phase_ms": [r.get("p2_host_ms") for r in res_dev + res_e2e],
        "work": {"part1_bin_pairs": p1_pairs, "part2_bin_pairs": p2_pairs, "part2_candidates": p2_cands,
                 "part2_device_calls": p2_calls, "part2_exact_rescored": p2_exact,
                 "pairs_touched": {"part1": p1_pairs, "part2_exact_scores": p2_exact_pairs,
                                   "note": "Part2: bin pairs read by the bit-exact re-scores (candidates on the near list x "
                                           "L(L-1)/2); the closed-form filter adds table passes over the window rows"},
                 "cuts": len(last["cuts"]), "groups": last.get("n_groups"),
                 "ambiguous_decisions": last["ambiguous"], "tie_sensitive_rows": last.get("tie_sensitive")},
        "e2e": {"value": e2e_value, "unit": "bin-pairs/s",
                # hica_p1_upload_symmetric: row blocks of 512 rows, columns from the block's first row onwards
                "h2d_bytes_per_step": h2d,
                "h2d_note": "all ranks together; symmetric upload: upper triangle only (%.2f GB of the %.2f GB matrix); "
                            "timed from PINNED host memory" % (h2d / 1e9, 8.0 * n * n / 1e9),
                "d2h_bytes_per_step": int(32 * (n - 1) + 4 * n + 8 * len(last["cuts"]) + 24 * max(1, p2_calls))},
        "e2e_pageable": {"value": total_pairs / (secs_pg / len(res_pg)), "unit": "bin-pairs/s", "steps": len(res_pg),
                         "note": "same step from ordinary (pageable) numpy memory through hica_p1_upload_symmetric's staged copies"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "nn4::k_nn_chain4 (linkage: one thread-block cluster, latency-bound chain)" if n <= 21000 else
                               "nn3::k_nn_chain3 (linkage: one thread-block cluster, latency-bound chain)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_BYTES.get(n), "traffic_source": "profiles/ (ncu --set full, "
                     "dram__bytes_read.sum + dram__bytes_write.sum per launch)" if n in NCU_DRAM_BYTES else None,
                     "alg_bytes_per_launch": link_bytes, "peak_source": peak_src, "per_kernel": kernels},
    }
    if not args.no_cpu_baseline:
        p2_full = p2_pairs
        cpu = CpuReference(wl["A_host"], args, last["initial_cuts"], p2_full, "this run's orderings")
        cpu.full()
        cb = cpu.sample(n_searches=1, n_cands=150)
        cb["leaves_equal_gpu"] = cb.pop("leaves_sha") == dig["leaves_sha"]
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if gcheck["equal"] is False or not ranks_agree or not all(same):
        sys.stderr.write("bench.py: RESULT MISMATCH %s (ranks agree: %s, steps agree: %s)\n" % (gcheck, ranks_agree, all(same)))
        return 3
    return 0


def reference_arm(args):
    """--impl reference: the reference's CPU path on the SAME workload and config (the reference itself is Python
    and cannot travel to the GPU box - no /root/reference there -; the oracle makes the same numpy/scipy/numba
    calls at the reference's call sites).  The matrix stages P1a-P1g are timed once at the full size; every step
    then times a bounded sample of the hypergeometric stage and of Part2 (see CpuReference)."""
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    wl = build_workload(args.bins, args.chroms, args.seed, dev)   # input generation only; nothing of libhica runs
    A = wl["A"].cpu().numpy()
    golden = load_golden(args)
    p2_full, p2_src = reference_p2_pairs(golden, args)
    starts = golden["initial_cuts"] if golden else [args.bins * k // args.chroms for k in range(1, args.chroms)]
    cpu = CpuReference(A, args, starts, p2_full, p2_src)
    t0 = time.perf_counter()
    cpu.full()
    t_full = time.perf_counter() - t0
    vals, secs, cb = [], [], None
    for i in range(max(args.warmup, 0) + args.steps):
        cb = cpu.sample(n_searches=1, n_cands=60)
        if i >= args.warmup:
            vals.append(cb["value"])
            secs.append(cb["est_full_step_seconds"])
    v = statistics.mean(vals)
    cb["value"] = v
    cb["full_size_pass_seconds"] = t_full
    cfg = config_dict(args)
    line = {"impl": "reference", "metric": "Part1+Part2 bin-pairs/s (and secs) at %dkb" % int(round(1.6e9 / args.bins / 1000)),
            "value": v, "unit": "bin-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": statistics.mean(secs) * 1e3, "ms_per_step_note": "estimated full-step time of the CPU path "
            "(matrix stages measured in full once, hypergeometric stage and Part2 extrapolated from the step's sample)",
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "bin-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
