#!/usr/bin/env python
"""Adds the counters of one `ncu --set full` capture of the dominant kernel to profiles/r02_ncu_bench_kernel.json,
keyed by bench workload: bench.py copies them into `roofline.traffic` / `roofline.ncu` of its JSON line, so that the
HBM traffic and the FP32-issue utilisation printed next to the roofline fraction are the measured ones of the same
kernel at the same configuration (never numbers taken while bench.py itself runs under a profiler).

    python tools/ncu_bench_json.py report.ncu-rep cornell_box_1080p_1024spp "v23, bench.py --steps 1 --warmup 0"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def num(key, scale_units=True):
    v = float(vals[col[key]].replace(",", ""))
    u = units[col[key]].lower()
    if scale_units:
        for k, f in (("tbyte", 1e12), ("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3)):
            if u.startswith(k):
                return v * f
        for k, f in (("msecond", 1e-3), ("usecond", 1e-6), ("nsecond", 1e-9), ("ms", 1e-3), ("us", 1e-6), ("ns", 1e-9)):
            if u == k:
                return v * f
    return v


entry = {
    "kernel": vals[col["Kernel Name"]].split("(")[0],
    "source": f"ncu --set full --clock-control none --import-source on, one launch ({os.path.basename(rep)}; {note}); "
              "summary in profiles/",
    "duration_ms": num("gpu__time_duration.sum") * 1e3,
    "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
    "l2_bytes_per_launch": 32.0 * num("lts__t_sectors.sum"),
    "registers_per_thread": num("launch__registers_per_thread"),
    "achieved_occupancy_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "issue_slots_busy_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fma_pipe_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "alu_pipe_pct": num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    "xu_pipe_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    "active_lanes_per_instruction": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "branch_efficiency_pct": num("smsp__sass_average_branch_targets_threads_uniform.pct"),
    "warp_instructions": num("smsp__inst_executed.sum"),
}
# executed FP32 arithmetic from the SASS page: thread-instructions of FFMA / FMUL / FADD (packed forms carry two
# lanes' worth), and the flops they amount to (FMA = 2) — the ISSUED counterpart of the algorithmic flops the
# roofline fraction credits
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2 and "Thread Instructions Executed" in srows[1]:
    h = srows[1]
    isrc, it = h.index("Source"), h.index("Thread Instructions Executed")
    lane_ops = flops = 0.0
    for r in srows[2:]:
        ins = r[isrc].strip()
        if ins.startswith("@"):
            ins = ins.split(None, 1)[1]
        op = ins.split()[0].split(".")[0]
        n = float(r[it])
        w = {"FFMA": (1, 2), "FMUL": (1, 1), "FADD": (1, 1), "FFMA2": (2, 4), "FMUL2": (2, 2), "FADD2": (2, 2)}.get(op)
        if w:
            lane_ops += w[0] * n
            flops += w[1] * n
    entry["fp32_lane_instructions_executed"] = lane_ops
    entry["fp32_flops_executed"] = flops
    entry["fp32_tflops_executed"] = flops / (entry["duration_ms"] * 1e-3) / 1e12
out = os.path.join(ROOT, "profiles", "r02_ncu_bench_kernel.json")
data = json.load(open(out)) if os.path.exists(out) else {}
data[workload] = entry
json.dump(data, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(entry, indent=1))
