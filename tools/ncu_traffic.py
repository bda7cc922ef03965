#!/usr/bin/env python3
"""DRAM traffic of the phase-2 launch set from an `ncu --set full` report -> profiles/traffic.json.  With K the report holds
the poa_dp2 launches of several calls and the LAST K of them are the set of one steady-state call (the first call of a process
grows the scratch pools and runs some segments twice).
  python tools/ncu_traffic.py REPORT.ncu-rep OUT.json [K]"""
import csv
import json
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
out = {"report": sys.argv[1], "launches": []}
tot = 0.0
units = rows[1]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


body = rows[2:]
if len(sys.argv) > 3 and int(sys.argv[3]) > 0:
    body = body[-int(sys.argv[3]):]
for r in body:
    name = r[ix["Kernel Name"]]
    rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    out["launches"].append({"kernel": name.split("(")[0].replace("void ", ""), "grid": r[ix["Grid Size"]], "dram_read_bytes": rd, "dram_write_bytes": wr,
                            "us": float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[units[ix["gpu__time_duration.sum"]]]})
    if "poa_dp2" in name:
        tot += rd + wr
out["poa_dp2_kernel_bytes_per_launch_set"] = tot
json.dump(out, open(sys.argv[2], "w"), indent=1)
print("phase-2 launch set: %.1f MB of DRAM traffic over %d launches" % (tot / 1e6, sum("poa_dp2" in l["kernel"] for l in out["launches"])))
