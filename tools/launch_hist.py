#!/usr/bin/env python
"""Per-kernel-name totals of the LAST `n` launches of an ncu launch list (gpu__time_duration.sum CSV)."""
import csv
import sys
from collections import defaultdict

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = []
for x in csv.DictReader(lines):
    try:
        L.append((x["Kernel Name"], float(x["Metric Value"].replace(",", ""))))
    except Exception:
        pass
if n:
    L = L[-n:]
tot, cnt = defaultdict(float), defaultdict(int)
for k, v in L:
    k = k.split("(")[0][:70]
    tot[k] += v
    cnt[k] += 1
all_ = sum(tot.values())
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{tot[k] / 1000:10.1f} us {cnt[k]:5d}x {100 * tot[k] / all_:5.1f}%  {k}")
print(f"{all_ / 1000:10.1f} us total, {len(L)} launches")
