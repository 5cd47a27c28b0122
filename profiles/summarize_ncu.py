"""Turns `ncu` outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.md
    python profiles/summarize_ncu.py full gpurun_out/prof_r1_agg.ncu-rep  > profiles/r1_agg_full.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        t = float(r[mv].replace(",", ""))
        # same kernel is launched on whole Jacobians (timed region) and on 4M-column chunks (e2e leg):
        # bucket by duration decade so the two do not average together
        bucket = "<10us" if t < 1e4 else ("10-100us" if t < 1e5 else ">=100us")
        key = (r[kn].split("(")[0][-70:], r[gs] + " " + bucket)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print("| launches | grid, duration bucket | total us | avg us | share | kernel |\n|---|---|---|---|---|---|")
    for (n, g), (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
        print(f"| {c} | {g} | {t / 1e3:.1f} | {t / 1e3 / c:.2f} | {100 * t / tot:.1f}% | `{n}` |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    kn = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join(m for m, _ in cols) + " |")
    print("|---|" + "---|" * len(cols))
    print("| (unit) | " + " | ".join(units[i] for _, i in cols) + " |")
    for r in rows[2:]:
        print(f"| `{r[kn][:48]}` | " + " | ".join(r[i] for _, i in cols) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
