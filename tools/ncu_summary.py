#!/usr/bin/env python3
"""Compact per-launch summary of an `ncu --set full` report: time, registers, occupancy, pipe and issue utilisation,
active threads per instruction, DRAM bytes, cache hit rates.
  python tools/ncu_summary.py REPORT.ncu-rep > summary.csv"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [c for c in WANT if c in ix]
w = csv.writer(sys.stdout)
w.writerow(["kernel", "grid"] + ["%s [%s]" % (c, units[ix[c]]) for c in cols])
for r in rows[2:]:
    w.writerow([r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), r[ix["Grid Size"]]] + [r[ix[c]] for c in cols])
