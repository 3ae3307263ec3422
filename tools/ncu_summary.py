#!/usr/bin/env python
"""Markdown summary of one `ncu --set full` launch from its raw-page CSV (+ optional per-line file).
usage: python tools/ncu_summary.py <name.raw.csv> [name.lines.txt] > profiles/xxx.md"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
r = list(csv.reader(open(sys.argv[1])))
h, u, v = r[0], r[1], r[2]
kernel = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
print(f"Kernel: `{kernel}`\n")
print("| metric | unit | value |\n|---|---|---|")
for name in WANT:
    if name in h:
        i = h.index(name)
        print(f"| {name} | {u[i]} | {v[i]} |")
print("\nWarp stall reasons (cycles per issued instruction, > 0.2):\n")
print("| stall | ratio |\n|---|---|")
for i, n in enumerate(h):
    if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
        try:
            x = float(v[i])
        except ValueError:
            continue
        if x > 0.2:
            print(f"| {n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | {x:.2f} |")
if len(sys.argv) > 2:
    print("\nHottest source lines (`-lineinfo`; share of executed warp-instructions / of stall samples):\n\n```")
    print(open(sys.argv[2]).read().strip())
    print("```")
