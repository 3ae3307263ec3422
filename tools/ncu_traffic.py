#!/usr/bin/env python
"""profiles/ncu_traffic.json from an ncu metrics CSV of the env-only loop:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/traffic.csv python tools/profile_env.py --cycles 5
usage: python tools/ncu_traffic.py gpurun_out/traffic.csv > profiles/ncu_traffic.json
Per kernel name: DRAM bytes (read + write) and time of the LAST full cycle; `encode_full` = every kernel of mjx_env_encode_obs
(k_encode_features + k_encode_store + k_sp_*), which is what bench.py's `roofline.traffic` quotes."""
import csv
import json
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
launch = {}
order = []
for r in rows:
    i = int(r["ID"])
    if i not in launch:
        launch[i] = {"name": r["Kernel Name"].split("(")[0].replace("void ", ""), "bytes": 0.0, "ns": 0.0}
        order.append(i)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    if r["Metric Name"].startswith("dram__bytes"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        launch[i]["bytes"] += v * mult
    elif r["Metric Name"] == "gpu__time_duration.sum":
        mult = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        launch[i]["ns"] = v * mult
L = [launch[i] for i in order]
starts = [k for k, x in enumerate(L) if x["name"].startswith("k_begin_step")]
enc = [k for k in starts if k + 2 < len(L) and "k_encode" in L[k + 2]["name"]]
s, e = enc[-2], enc[-1]
out = {}
for x in L[s:e]:
    d = out.setdefault(x["name"], {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
    d["launches"] += 1
    d["dram_bytes"] += x["bytes"]
    d["us"] += x["ns"] / 1e3
full = [x for x in L[s:e] if x["name"].startswith(("k_encode", "k_sp_"))]
res = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (one steady-state env step at 4096 tables; cold-cache, serialised launches)",
       "encode_full": sum(x["bytes"] for x in full), "k_encode_store": out.get("k_encode_store", {}).get("dram_bytes"),
       "per_kernel": out}
print(json.dumps(res, indent=1))
