#!/usr/bin/env python
"""bench.py - Part1+Part2 hot path of HiC_Genome_Assembler on B200 ("Measurement").

A step = one pass of the hot path over one synthetic ICE-normalised genome (SURVEY.md 8d recipe, seed 7):
  Part1: distance transform -> UPGMA linkage -> leaf order -> similarity rank -> hypergeometric cut indices
  Part2: per chromosome group brute-force + greedy insertion + sliding-window ordering search
Default workload = BASELINE.json configs[1]/metric: 1.6 Gb genome at 100 kb (16,000 bins, 20 chromosomes,
40 scaffolds per chromosome).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--bins 16000] [--impl reference]

metric  : bin-pairs/s over Part1+Part2 (Part1 pairs = N(N-1)/2; Part2 pairs = sum over scored candidates of
          L(L-1)/2), plus the absolute seconds in `seconds`.
value   : input matrix resident in HBM when the timed region starts.
e2e     : the same step through the public host API with the matrix in pinned HOST memory (H2D inside the
          timed region, results read back to the host).
"""
from __future__ import annotations

import argparse
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


# DRAM bytes of one linkage-kernel launch measured with `ncu --set full` (bins -> bytes); see profiles/
NCU_DRAM_BYTES = {16000: 4.95e9}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bins", type=int, default=16000)
    ap.add_argument("--chroms", type=int, default=20)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--impl", default="hica", choices=["hica", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--part1-only", action="store_true")
    ap.add_argument("--replicate-part1", action="store_true", help="N > 1: every rank runs the whole of Part1 (no row-block sharding)")
    ap.add_argument("--p2-host-loop", action="store_true", help="Part2 through per-call host loops (exact scoring of every candidate)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- workload
def build_workload(n_bins, n_chrom, seed, device):
    from hic_genome_assembler_b200.synth import make_genome
    g = make_genome(n_bins, n_chrom=n_chrom, seed=seed, device=device)
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
    """source = ('device', ptr) or ('host', pinned ndarray).  Returns a dict of results and timings."""
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
           "ambiguous": res["ambiguous"], "t_part1": t1 - t0, "timers_p1": tm1, "linkage_stats": h.linkage_stats(),
           "upload": res.get("upload"), "stage_ms_sharded": res.get("stage_ms")}
    p2_pairs, p2_cands, p2_calls, kernel_ms_p2 = 0, 0, 0, 0.0
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
        part2.STATS.update(candidates=0, bin_pairs=0, calls=0, exact=0)
        nbins = [len(g) if args.p2_host_loop else sum(len(ids) for _, ids in g) for g in groups]
        mine = hdist.shard_groups(nbins, world, rank)
        if args.p2_host_loop:
            for ci in mine:
                orderings[ci] = part2.order_chromosome(h, index_of, groups[ci], 6, 5)
        elif mine:
            got, cost, rounds, stats = part2.order_groups_device(h, index_of, [groups[ci] for ci in mine], 6, 5)
            for k, ci in enumerate(mine):
                orderings[ci] = (got[k], float(cost[k]), int(rounds[k]))
            out["p2_exact"] = part2.STATS.get("exact", 0)
            out["p2_host_ms"] = h.debug_stats()
        if world > 1:  # every rank ends the step with all orderings (one small all_gather_object)
            orderings = dict(enumerate(hdist.merge_orderings(orderings, len(groups))))
        p2_pairs, p2_cands, p2_calls = part2.STATS["bin_pairs"], part2.STATS["candidates"], part2.STATS["calls"]
        out["n_groups"] = len(groups)
    t2 = time.perf_counter()
    out.update(orderings=orderings, t_part2=t2 - t1, t_total=t2 - t0, p2_pairs=p2_pairs, p2_cands=p2_cands,
               p2_calls=p2_calls)
    return out


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline(wl, args, quick=False):
    """The oracle (numpy/scipy/numba calls at the reference's call sites) on a bounded sample of the workload:
    Part1 matrix stages + two break searches on the leading sub-block, Part2 on a sample of scan candidates.
    Rates are extrapolated to the full workload's pair counts (labelled as such)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hica_oracle as O
    A = wl["A_host"]
    n = A.shape[0]
    ns = min(n, 2400 if quick else 6000)
    sub = np.ascontiguousarray(A[:ns, :ns])
    t0 = time.perf_counter()
    rs = O.sequential_row_sums(sub)
    D = O.convert_distance(sub)
    Z, leaves = O.average_cluster_leaves(D)
    Dp = O.reorder(D, leaves)
    S = O.convert_similarity(Dp, rs[np.asarray(leaves)])
    R = O.rank_matrix(S)
    t_matrix = time.perf_counter() - t0
    t0 = time.perf_counter()
    starts_done, ind = 0, 0
    for _ in range(2):
        cuts = O.find_breakpoints(R, ind, 5, ns - ind, .05)
        starts_done += 1
        if not cuts:
            break
        ind += cuts[0]
    t_cuts = time.perf_counter() - t0
    p1_pairs = ns * (ns - 1) // 2
    # a full run at n bins needs ~ (#chromosomes + spurious) break searches and the noise filter; scale the
    # measured per-search cost by the number of searches the GPU run performed
    n_searches = max(1, wl.get("n_searches", args.chroms + 5))
    t_p1_sample = t_matrix + t_cuts * (n_searches / max(1, starts_done)) * ((n / ns) ** 0)
    rate_p1 = p1_pairs / t_p1_sample
    # ---- Part2 sample: one chromosome-sized matrix, real scan-window candidates
    L = max(8, n // args.chroms)
    M = np.ascontiguousarray(A[:L, :L])
    total = O.total_weight(M)
    rng = np.random.RandomState(0)
    O.ordering_cost_numba(M[:16, :16], 1.0)  # JIT warm-up outside the timed region
    n_c = 200 if quick else 5000
    t0 = time.perf_counter()
    for _ in range(n_c):
        o = np.arange(L)
        a = rng.randint(0, L - 100) if L > 120 else 0
        o[a:a + min(100, L)] = o[a:a + min(100, L)][::-1]
        O.ordering_cost_numba(M[np.ix_(o, o)], total)
    t_p2 = time.perf_counter() - t0
    p2_pairs = n_c * (L * (L - 1) // 2)
    rate_p2 = p2_pairs / t_p2
    full_p1, full_p2 = n * (n - 1) // 2, wl.get("p2_pairs_full", 0)
    est_seconds = full_p1 / rate_p1 + (full_p2 / rate_p2 if full_p2 else 0.0)
    value = (full_p1 + full_p2) / est_seconds
    return {"value": value, "unit": "bin-pairs/s", "cores": 1, "kind": "port",
            "sample": "oracle (numpy/scipy/numba at the reference's call sites, 1 thread): Part1 stages on the "
                      "leading %d-bin sub-block (%.1fs matrix stages + %d break searches %.1fs, scaled to %d searches)"
                      "; Part2 %d scan candidates at L=%d (%.1fs); rates extrapolated to the full step's pair counts"
                      % (ns, t_matrix, starts_done, t_cuts, n_searches, n_c, L, t_p2),
            "part1_pairs_per_s": rate_p1, "part2_pairs_per_s": rate_p2, "est_full_step_seconds": est_seconds,
            "host_cpu_count": os.cpu_count()}


def score_kernel_probe(h, wl, last):
    """The exact ordering-score kernel on its own (the device-resident search only re-scores a short list with it):
    one scan window of the largest chromosome group, every candidate scored exactly (hica_p2_score_window),
    3 timed calls after one warm-up.  Algorithmic bytes = candidates x L(L-1)/2 contacts x 8 B; the 5 MB chromosome
    matrix is L2-resident, so the figure is an L2 rate quoted against the HBM peak."""
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
    from hic_genome_assembler_b200._lib import Handle

    dev = torch.device("cuda", local_rank)
    wl = build_workload(args.bins, args.chroms, args.seed, dev)
    A_dev = wl["A"]
    A_pinned = torch.empty(A_dev.shape, dtype=torch.float64, pin_memory=True)
    A_pinned.copy_(A_dev)
    torch.cuda.synchronize()
    wl["A_host"] = A_pinned.numpy()
    h = Handle(local_rank)

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

    # ---- aggregate work over ranks
    n = args.bins
    p1_pairs = n * (n - 1) // 2
    p2_pairs_local = res_dev[-1]["p2_pairs"]
    p2 = torch.tensor([p2_pairs_local, res_dev[-1]["p2_cands"], res_dev[-1]["p2_calls"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(p2, op=dist.ReduceOp.SUM)
    p2_pairs, p2_cands, p2_calls = (int(v) for v in p2.tolist())
    total_pairs = p1_pairs + p2_pairs
    ms_per_step = secs_dev / args.steps * 1e3
    value = total_pairs / (secs_dev / args.steps)
    e2e_value = total_pairs / (secs_e2e / args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    last = res_dev[-1]
    tm = last["timers_p1"]
    # roofline of the dominant kernel (the NN-chain linkage, HBM-class but latency-bound: 24 N^2 algorithmic bytes)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
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
    line = {
        "metric": "Part1+Part2 bin-pairs/s (and secs) at %dkb" % int(round(1.6e9 / n / 1000)),
        "value": value, "unit": "bin-pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "bins": n, "chromosomes": args.chroms,
                   "part2": "skipped" if args.part1_only else "nScaffolds=6 scanScaffolds=5",
                   "l2": "inputs (%.2f GB matrix) larger than L2; no explicit flush" % (8.0 * n * n / 1e9),
                   "sharding": ("Part1 replicated per rank" if args.replicate_part1 else
                                "row-block sharded upload + NVLink all-gather of the matrix, linkage on rank 0 + leaf-order "
                                "broadcast, " + ("rank: value sort on every GPU while the chain runs, local column relabel"
                                                 if n <= 65535 and os.environ.get("HICA_RANK_OVERLAP", "1") != "0"
                                                 else "rank matrix by row block + all-gather") +
                                ", cut searches replicated") +
                               "; Part2 chromosome groups dealt to ranks largest first" if world > 1 else "1 GPU"},
        "seconds": {"part1": last["t_part1"], "part2": last["t_part2"], "total": last["t_total"],
                    "e2e_total": secs_e2e / args.steps, "e2e_part1": res_e2e[-1]["t_part1"],
                    "e2e_part2": res_e2e[-1]["t_part2"], "e2e_upload_ms": res_e2e[-1]["timers_p1"]["upload"],
                    "e2e_upload_rank0": res_e2e[-1].get("upload"), "sharded_stage_ms_rank0": last.get("stage_ms_sharded")},
        "stage_ms_part1": tm, "linkage_stats": last["linkage_stats"],
        "per_step_seconds": [[round(r["t_part1"], 4), round(r["t_part2"], 4)] for r in res_dev],
        "per_step_seconds_e2e": [[round(r["t_part1"], 4), round(r["t_part2"], 4)] for r in res_e2e],
        "part2_host_phase_ms": [r.get("p2_host_ms") for r in res_dev + res_e2e],
        "work": {"part1_bin_pairs": p1_pairs, "part2_bin_pairs": p2_pairs, "part2_candidates": p2_cands,
                 "part2_device_calls": p2_calls, "part2_exact_rescored": last.get("p2_exact"), "cuts": len(last["cuts"]), "groups": last.get("n_groups"),
                 "ambiguous_decisions": last["ambiguous"]},
        "e2e": {"value": e2e_value, "unit": "bin-pairs/s",
                # hica_p1_upload_symmetric: row blocks of 512 rows, columns from the block's first row onwards
                "h2d_bytes_per_step": int(sum(8 * min(512, n - r0) * (n - r0) for r0 in range(0, n, 512))),
                "h2d_note": "all ranks together; symmetric upload: upper triangle only (%.2f GB of the %.2f GB matrix)" % (
                    sum(8 * min(512, n - r0) * (n - r0) for r0 in range(0, n, 512)) / 1e9, 8.0 * n * n / 1e9),
                "d2h_bytes_per_step": int(32 * (n - 1) + 4 * n + 8 * len(last["cuts"]) + 24 * max(1, p2_calls))},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "nn3::k_nn_chain3 (linkage: one 16-CTA cluster, latency-bound chain)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_BYTES.get(n), "traffic_source": "profiles/r01_ncu_full_final2_summary.txt (ncu --set full, "
                     "dram__bytes_read.sum + dram__bytes_write.sum per launch)" if n in NCU_DRAM_BYTES else None,
                     "alg_bytes_per_launch": link_bytes, "peak_source": peak_src, "per_kernel": kernels},
    }
    if not args.no_cpu_baseline:
        wl["p2_pairs_full"] = p2_pairs
        line["cpu_baseline"] = cpu_baseline(wl, args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def reference_arm(args):
    """--impl reference: the reference's CPU path (oracle port: the same numpy/scipy/numba calls) on a bounded
    sample per step.  The reference is Python and cannot travel to the GPU box (no /root/reference there)."""
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    wl = build_workload(args.bins, args.chroms, args.seed, dev)
    wl["A_host"] = wl["A"].cpu().numpy()
    wl["p2_pairs_full"] = int(2.8e6 * (args.bins // args.chroms) ** 2 / 2) if not args.part1_only else 0
    vals, secs = [], []
    for i in range(max(args.warmup, 0) + args.steps):
        t0 = time.perf_counter()
        cb = cpu_baseline(wl, args, quick=True)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            vals.append(cb["value"])
            secs.append(dt)
    v = statistics.mean(vals)
    cb["value"] = v
    line = {"impl": "reference", "metric": "Part1+Part2 bin-pairs/s (and secs) at %dkb" % int(round(1.6e9 / args.bins / 1000)),
            "value": v, "unit": "bin-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": statistics.mean(secs) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "bins": args.bins, "chromosomes": args.chroms},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "bin-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
