#!/usr/bin/env python3
"""Per-launch table of an ncu --csv --metrics capture (several metrics per launch): the DP launches of the LAST call.
  python tools/launch_metrics.py LAUNCHES.csv [REGEX]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else "poa_dp")
hdr, data = None, {}
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        e = data.setdefault(int(d["ID"]), {"k": d["Kernel Name"], "g": d["Grid Size"]})
        e[d["Metric Name"]] = (d["Metric Value"], d["Metric Unit"])
ids = sorted(data)
last = max(i for i in ids if data[i]["k"].startswith("init_call_kernel"))
metrics = sorted({m for i in ids for m in data[i] if m not in ("k", "g")})
print("id,kernel,grid," + ",".join(metrics))
for i in ids:
    if i >= last and pat.search(data[i]["k"]):
        d = data[i]
        name = re.sub(r"\(.*", "", d["k"]).replace("void ", "")
        print("%d,%s,%s,%s" % (i, name, d["g"].replace(",", ""), ",".join("%s %s" % d.get(m, ("", "")) for m in metrics)))
