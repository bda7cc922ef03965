#!/usr/bin/env python3
"""Writes a synthetic workload as raw CSR arrays for tools/pipe_driver (C, no Python in the profiled process).
  python tools/dump_csr.py READS CONFIG PREFIX  ->  PREFIX.{ref,cor,unc}.bin (uint8), PREFIX.{ref,cor,unc}_off.bin, PREFIX.read_first.bin (int64)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads  # noqa: E402

reads, cfg, pre = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
wl = workloads.make_windows(cfg, reads)
for k in ("ref", "cor", "unc", "ref_off", "cor_off", "unc_off", "read_first"):
    wl[k].tofile("%s.%s.bin" % (pre, k))
print("%d windows, %d reads -> %s.*.bin" % (len(wl["ref_off"]) - 1, len(wl["read_first"]) - 1, pre))
