#!/usr/bin/env python3
"""Issued-instruction view of the phase-2 launch set, from an `ncu --csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum,
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_active` launch list of
`pipe_driver PREFIX 2` (the launches of the LAST call) -> profiles/issue.json, which bench.py attaches to `roofline.issue`.
  python tools/ncu_issue.py LAUNCHES.csv OUT.json"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, data = None, {}
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        e = data.setdefault(int(d["ID"]), {"k": d["Kernel Name"], "g": d["Grid Size"]})
        e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
ids = sorted(data)
last = max(i for i in ids if data[i]["k"].startswith("init_call_kernel"))
out = {"source": sys.argv[1], "launches": []}
inst = ns = ipc_w = alu_w = 0.0
for i in ids:
    d = data[i]
    if i < last or "poa_dp2" not in d["k"]:
        continue
    name = re.sub(r"\(.*", "", d["k"]).replace("void ", "")
    t, n = d["gpu__time_duration.sum"], d["smsp__inst_executed.sum"]
    out["launches"].append({"kernel": name, "grid": d["g"], "us": t / 1e3, "warp_instructions": n,
                            "ipc_active": d["sm__inst_executed.avg.per_cycle_active"],
                            "alu_pipe_pct": d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]})
    inst += n
    ns += t
    ipc_w += d["sm__inst_executed.avg.per_cycle_active"] * t
    alu_w += d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"] * t
out["phase2_warp_instructions_per_launch_set"] = inst
out["phase2_serialised_us"] = ns / 1e3
out["ipc_active_time_weighted"] = ipc_w / ns          # warp instructions per cycle and SM, of 4
out["issue_slot_utilisation"] = ipc_w / ns / 4.0
out["alu_pipe_pct_time_weighted"] = alu_w / ns
json.dump(out, open(sys.argv[2], "w"), indent=1)
print("phase 2: %.0f M warp instructions, IPC %.2f of 4 (time weighted), ALU pipe %.0f %%" % (inst / 1e6, ipc_w / ns, alu_w / ns))
