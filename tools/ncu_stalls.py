#!/usr/bin/env python3
"""Top warp-stall reasons (stalled warps per issue) of the large launches of an `ncu --set full` report.
  python tools/ncu_stalls.py REPORT.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    grid = int(r[ix["Grid Size"]].strip("()").split(",")[0])
    if grid < 1000:
        continue
    vals = sorted(((float(r[ix[h]].replace(",", "")), h) for h in stall), reverse=True)[:6]
    print(r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), r[ix["Grid Size"]])
    for v, h in vals:
        print("    %6.2f %s" % (v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
