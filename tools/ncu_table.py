#!/usr/bin/env python
"""One row per kernel NAME of an `ncu --set full` capture with many launches (the wavefront variant launches
three kernels per wave): number of launches captured and the mean of duration, issue-slot utilisation, active
lanes, pipe utilisation, L2 and HBM bytes / bandwidth and the dominant stall.  Markdown on stdout.

    python tools/ncu_table.py report.ncu-rep [title]      (needs ncu on PATH; no GPU)
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "µs", 1.0), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
     ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes", 1.0),
     ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA %", 1.0),
     ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU %", 1.0),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1.0),
     ("lts__t_sectors.sum", "L2 MB", 32.0), ("dram__bytes_read.sum", "HBM read MB", 1.0), ("dram__bytes_write.sum", "HBM write MB", 1.0),
     ("dram__bytes.sum.per_second", "HBM GB/s", 1.0), ("lts__t_sectors.sum.per_second", "L2 GB/s", 32.0)]   # a sector is 32 bytes


def to_base(v, unit):
    v = float(v.replace(",", "")) if v else 0.0
    u = unit.lower()
    for k, f in (("tbyte", 1e12), ("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if u.startswith(k):
            return v * f
    for k, f in (("msecond", 1e-3), ("usecond", 1e-6), ("nsecond", 1e-9), ("second", 1.0), ("ms", 1e-3), ("us", 1e-6), ("ns", 1e-9)):
        if u == k:
            return v * f
    return v


agg = collections.defaultdict(lambda: collections.defaultdict(list))
stall = collections.defaultdict(lambda: collections.Counter())
for vals in rows[2:]:
    name = vals[col["Kernel Name"]].split("(")[0].replace("void ", "")
    for key, _, mult in M:
        if key in col:
            agg[name][key].append(mult * to_base(vals[col[key]], units[col[key]]))
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and vals[i]:
            stall[name][h.split("issue_stalled_")[1].split("_per_issue")[0]] += float(vals[i])
print(f"# {title}\n")
print("| kernel | launches | " + " | ".join(lbl for _, lbl, _ in M) + " | top stalls (cycles / issued instruction) |")
print("|---|---|" + "---|" * (len(M) + 1))
for name, d in agg.items():
    n = len(d[M[0][0]])
    cells = []
    for key, lbl, _ in M:
        v = sum(d[key]) / max(len(d[key]), 1)
        if lbl == "µs":
            v *= 1e6
        elif "MB" in lbl:
            v /= 1e6
        elif "GB/s" in lbl:
            v /= 1e9
        cells.append(f"{v:.1f}")
    top = ", ".join(f"{k} {v / n:.1f}" for k, v in stall[name].most_common(3))
    print(f"| `{name}` | {n} | " + " | ".join(cells) + f" | {top} |")
