#!/usr/bin/env python
"""Groups per-line ncu counters of k_encode_features by feature group (the `// ----` markers in csrc/mjx_obs.cuh).
usage: ncu -i rep --page source --csv --print-source cuda,sass | python tools/ncu_sections.py"""
import bisect
import collections
import csv
import os
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "mortal_b200/csrc/mjx_obs.cuh")).read().split("\n")
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if l.strip().startswith("// ----") or l.startswith("MJX_D") or l.startswith("template") or l.startswith("MJX_DN")]
starts = [m[0] for m in marks]
inst, smp = collections.Counter(), collections.Counter()
fp = hdr = None
for row in csv.reader(sys.stdin):
    if not row:
        continue
    if row[0] == "File Path":
        fp = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        continue
    if hdr is None or len(row) < len(hdr) or row[2] != "-":
        continue
    try:
        n, s, ln = int(row[hdr.index("Instructions Executed")]), int(row[hdr.index("# Samples")]), int(row[0])
    except ValueError:
        continue
    if fp == "mjx_obs.cuh":
        k = bisect.bisect_right(starts, ln) - 1
        key = "obs: " + (marks[k][1][:60] if k >= 0 else "head")
    else:
        key = fp
    inst[key] += n
    smp[key] += s
ti, ts = sum(inst.values()) or 1, sum(smp.values()) or 1
for k, v in smp.most_common(30):
    print(f"samples {100 * v / ts:5.1f}%  inst {100 * inst[k] / ti:5.1f}%  {k}")
