#!/usr/bin/env python
"""Summarises one `ncu --set full` capture (.ncu-rep) as markdown: duration, registers,
occupancy, issue utilisation, pipe utilisation, branch efficiency, lane efficiency,
L2/HBM bytes, stall reasons and the SASS opcode mix.  Usage: ncu_summary.py REP [TITLE]"""
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
print(f"# {title}\n")
for k, vals in enumerate(rows[2:]):
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    name = m.get("Kernel Name", ("?", ""))[0]
    print(f"## launch {k}: `{name}`  grid {m.get('Grid Size', ('', ''))[0]} block {m.get('Block Size', ('', ''))[0]}\n")
    want = [
        ("gpu__time_duration.sum", "duration"),
        ("launch__registers_per_thread", "registers / thread"),
        ("launch__occupancy_limit_registers", "CTAs/SM allowed by registers"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps)"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (FP32-issue utilisation proxy)"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per warp instruction (of 32)"),
        ("smsp__sass_average_branch_targets_threads_uniform.pct", "branch efficiency (uniform branch targets)"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe"),
        ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread-instructions"),
        ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread-instructions"),
        ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread-instructions"),
        ("smsp__sass_thread_inst_executed_op_integer_pred_on.sum", "integer thread-instructions"),
        ("dram__bytes_read.sum", "HBM read"),
        ("dram__bytes_write.sum", "HBM written"),
        ("dram__bytes.sum.per_second", "HBM bandwidth"),
        ("lts__t_sectors.sum", "L2 sectors (32 B each)"),
        ("lts__t_sectors.sum.per_second", "L2 sector rate"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
        ("l1tex__t_bytes.sum", "L1 bytes"),
        ("idc__request_cycles_active.avg.pct_of_peak_sustained_elapsed", "indexed-constant cache busy"),
    ]
    print("| metric | value |\n|---|---|")
    for key, label in want:
        if key in m:
            print(f"| {label} (`{key}`) | {m[key][0]} {m[key][1]} |")
    stalls = [(h.split("issue_stalled_")[1].split("_per_issue")[0], float(v[0])) for h, v in m.items()
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v[0]]
    stalls.sort(key=lambda x: -x[1])
    print("\nstall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for n, v in stalls[:8]) + "\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2 and "Instructions Executed" in rows[1]:
    h = rows[1]
    ci, ti = h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    tot, by = 0, collections.Counter()
    for r in rows[2:]:
        ins = r[1].strip()
        if ins.startswith("@"):
            ins = ins.split(None, 1)[1]
        op = ins.split()[0].split(".")[0]
        n = int(r[ci])
        tot += n
        by[op] += n
    print(f"SASS mix of launch 0 ({len(rows) - 2} static instructions, {tot} executed warp instructions): "
          + ", ".join(f"{op} {100 * n / tot:.1f}%" for op, n in by.most_common(16)))
