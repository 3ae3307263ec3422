#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: ncu -i rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [top_n]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows, fpath, hdr = [], None, None
for row in csv.reader(sys.stdin):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        continue
    if hdr is None or len(row) < len(hdr) or row[2] != "-":
        continue  # keep only the per-source-line summary rows (Address == "-")
    try:
        n = int(row[hdr.index("Instructions Executed")])
        smp = int(row[hdr.index("# Samples")])
        thr = float(row[hdr.index("Avg. Threads Executed")]) if n else 0
    except ValueError:
        continue
    rows.append((n, smp, thr, fpath, row[0], row[1].strip()[:100]))
tot = sum(r[0] for r in rows) or 1
tsm = sum(r[1] for r in rows) or 1
print(f"total warp-instructions {tot}, samples {tsm}")
print("-- by instructions")
for n, smp, thr, f, ln, src in sorted(rows, reverse=True)[:top]:
    print(f"{100 * n / tot:5.1f}%  smp {100 * smp / tsm:5.1f}%  {f}:{ln}  {src}")
print("-- by stall samples")
for n, smp, thr, f, ln, src in sorted(rows, key=lambda r: -r[1])[:top]:
    print(f"smp {100 * smp / tsm:5.1f}%  inst {100 * n / tot:5.1f}%  {f}:{ln}  {src}")
