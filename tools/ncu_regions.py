#!/usr/bin/env python
"""Per-region issue share and lane occupancy of one kernel from an `ncu --set full --import-source on`
capture: consecutive SASS instructions with the same execution count are one region (a loop body, a divergent
block); for each, its share of all executed warp instructions and the average number of active lanes.  This is
what showed that the primary-ray block of the Cornell kernel ran in every iteration at 5 of 32 lanes, and that
the leaf tests of the Random scene's BVH walk are 25 % of its issue slots at 5 lanes.

    python tools/ncu_regions.py report.ncu-rep [min_share]     (needs ncu on PATH; no GPU)
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ia, isrc = hdr.index("Address"), hdr.index("Source")
ie, it, iss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = int(data[0][ia], 16)
tot_e = sum(int(r[ie]) for r in data)
tot_t = sum(int(r[it]) for r in data)
print(f"{rows[0][1]}\ntotal warp instructions {tot_e:.3e}, active lanes per instruction {tot_t / tot_e:.2f}")
regions, cur = [], None
for r in data:
    a, e, t, s, sm = int(r[ia], 16) - base, int(r[ie]), int(r[it]), r[isrc].strip(), int(r[iss])
    if cur and cur["e0"] > 0 and abs(e - cur["e0"]) <= 0.03 * cur["e0"]:
        cur["n"] += 1; cur["E"] += e; cur["T"] += t; cur["S"] += sm; cur["end"] = a
    else:
        cur = {"start": a, "end": a, "e0": e, "n": 1, "E": e, "T": t, "S": sm, "first": s}
        regions.append(cur)
print("   range        instr  share  lanes  stall-samples  executions  first instruction")
for c in regions:
    if c["E"] / tot_e > min_share:
        print(f"{c['start']:05x}-{c['end']:05x}  {c['n']:5d} {100 * c['E'] / tot_e:5.1f}% {c['T'] / max(c['E'], 1):6.1f} {c['S']:14d} {c['e0'] / 1e6:9.1f}M  {c['first'][:48]}")
