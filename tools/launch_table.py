#!/usr/bin/env python3
"""prints the kernel launches of an ncu `--metrics gpu__time_duration.sum --csv` log: id, kernel, grid, microseconds
  python tools/launch_table.py LOG.csv [first_id]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if int(d["ID"]) >= first:
            name = d["Kernel Name"].split("(")[0].replace("void ", "")
            print("%4s %-46s %-14s %10.1f us" % (d["ID"], name[:46], d["Grid Size"], float(d["Metric Value"].replace(",", "")) / 1e3))
