"""Turn an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list of `bench.py --steps 1 --warmup 1`
into the compact per-step list + per-kernel summary kept under profiles/.

    python profiles/summarise_launches.py gpurun_out/final_launches.csv profiles/r02_launches_100kb_final

The process first launches torch kernels for the synthetic workload, then warm-up and timed steps of this
repository's kernels (every step starts with k_rowsum_sequential or k_rowsum_pairwise).  The LAST complete step that
starts from resident data is written out.
"""
import csv
import sys
from collections import OrderedDict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else (v * 1e6 if u in ("s", "second") else v))
        name = r[ik].split("(")[0].strip()
        rows.append((name, us))
    ours = [i for i, (n, _) in enumerate(rows) if "at::" not in n and "cutlass" not in n and "elementwise" not in n]
    starts = [i for i in ours if "k_rowsum_pairwise" in rows[i][0]]
    if not starts:
        raise SystemExit("no step found")
    # steps are delimited by the distance kernels; take the last step whose successor exists, else the last one
    s0 = starts[-2] if len(starts) >= 2 else starts[-1]
    s1 = starts[-1] if len(starts) >= 2 else len(rows)
    if s0 > 0 and "k_rowsum_sequential" in rows[s0 - 1][0]:
        s0 -= 1
        if len(starts) >= 2 and "k_rowsum_sequential" in rows[s1 - 1][0]:
            s1 -= 1
    step = rows[s0:s1]
    with open(dst + ".csv", "w") as f:
        f.write("index,kernel,gpu__time_duration_us\n")
        for i, (n, us) in enumerate(step):
            f.write('%d,"%s",%.3f\n' % (i, n, us))
    agg = OrderedDict()
    for n, us in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(us for _, us in step)
    with open(dst + "_summary.txt", "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 12000, python bench.py --steps 1 --warmup 1 --no-cpu-baseline\n")
        f.write("%d launches in the process, launches %d..%d are one step (%d launches)\n" % (len(rows), s0, s1 - 1, len(step)))
        f.write("Per-launch times under ncu are cold-cache and serialised: compare the shares, not the absolutes.\n")
        f.write("%d launches, %.2f ms in total\n\n" % (len(step), tot / 1e3))
        for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-62s n=%4d %11.1f us %6.1f %%\n" % (n[:62], c, us, 100 * us / tot))
    print(open(dst + "_summary.txt").read())


if __name__ == "__main__":
    main()
