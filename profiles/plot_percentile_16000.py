"""Round-2 measurement of the plot-input path: numpy.percentile(adjMat, 1) and (adjMat, 98) of the
16,000-bin distance matrix (plotContactMaps.py:30-31) - device radix select on the resident matrix against numpy on
a host copy, results compared bit for bit.  Writes gpurun_out/r2_plot_percentile.json.

    python profiles/plot_percentile_16000.py [bins] [--device-only]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from hic_genome_assembler_b200 import plot as hplot
    from hic_genome_assembler_b200._lib import Handle
    from hic_genome_assembler_b200.synth import make_genome_portable
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    device_only = "--device-only" in sys.argv          # for runs under ncu: no numpy comparison, nothing written
    n = int(args[0]) if args else 16000
    g = make_genome_portable(n, n_chrom=20, seed=7, device="cuda:0")
    out = {"bins": n, "peak_hbm_gbs": None}
    try:
        out["peak_hbm_gbs"] = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        pass
    with Handle(0) as h:
        h.p1_upload_device(g.A.data_ptr(), n)
        h.p1_set_row_sums(None)
        h.p1_distance()
        ms = {}
        for what, name in ((0, "D"), (2, "A")):
            t = []
            for _ in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                got = hplot.percentile_device(h, None, [1, 98], what=what)
                t.append((time.perf_counter() - t0) * 1e3)
            ms[name] = {"device_ms_per_call_two_percentiles": float(np.median(t)), "all_ms": t, "values": [float(v) for v in got]}
            # 6 counting passes over 8 n^2 bytes
            ms[name]["algorithmic_gb"] = 6 * 8.0 * n * n / 1e9
            ms[name]["gb_per_s"] = ms[name]["algorithmic_gb"] / (ms[name]["device_ms_per_call_two_percentiles"] / 1e3)
        if device_only:
            print(json.dumps(ms))
            return
        D = h.p1_get_distance()
        t0 = time.perf_counter()
        want = [np.percentile(D, 1), np.percentile(D, 98)]          # the reference's two calls
        ms["numpy_s_two_calls"] = time.perf_counter() - t0
        got = hplot.percentile_device(h, None, [1, 98], what=0)
        ms["equal_to_numpy"] = bool(np.asarray(got).tobytes() == np.asarray(want).tobytes())
        t0 = time.perf_counter()
        got_h = hplot.percentile_device(h, D, [1, 98])               # host array: upload + select
        ms["host_array_ms"] = (time.perf_counter() - t0) * 1e3
        ms["host_array_equal"] = bool(np.asarray(got_h).tobytes() == np.asarray(want).tobytes())
        out.update(ms)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_plot_percentile.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
