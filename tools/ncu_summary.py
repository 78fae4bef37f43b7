#!/usr/bin/env python
"""Summarise an ncu report exported as CSV (raw page + source page).  Usage:
   ncu -i X.ncu-rep --page raw --csv > raw.csv; ncu -i X.ncu-rep --page source --csv > src.csv
   python tools/ncu_summary.py raw.csv src.csv"""
import collections
import csv
import re
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_issued.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    if k in M:
        print(f"{k:80s} {M[k][0]:>16s} {M[k][1]}")
print("-- stall reasons (warp-cycles per issued instruction)")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = float(M[h][0])
        if v >= 0.05:
            print(f"   {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {v:6.2f}")
rows = list(csv.reader(open(src)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ops, samples = collections.Counter(), collections.Counter()
tot = totS = 0
for r in rows[2:]:
    sass = r[ci["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    op = ".".join((m.group(2) if m else sass).split(".")[:2])
    n, s = int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])
    ops[op] += n
    samples[op] += s
    tot += n
    totS += s
print(f"-- opcode mix: {tot} warp instructions, {totS} samples")
for op, n in ops.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 24):
    print(f"   {op:22s} {n:12d} {100*n/tot:5.1f}%   samples {100*samples[op]/max(totS,1):5.1f}%")
