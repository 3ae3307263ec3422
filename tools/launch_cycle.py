#!/usr/bin/env python
"""Per-kernel times of the last full env cycle in an ncu launch list (gpu__time_duration.sum CSV)."""
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
L = []
for x in csv.DictReader(lines):
    try:
        L.append((x["Kernel Name"], float(x["Metric Value"].replace(",", ""))))
    except Exception:
        pass
idx = [i for i, (k, _) in enumerate(L) if k.startswith("k_begin_step")]
enc = [i for i in idx if i + 2 < len(L) and "k_encode" in L[i + 2][0]]
s, e = enc[-2], enc[-1]
tot = 0
for k, v in L[s:e]:
    print(f"{k.split('(')[0][:34]:34s} {v / 1000:9.1f} us")
    tot += v
print("total", tot / 1e6, "ms")
