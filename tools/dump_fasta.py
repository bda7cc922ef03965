#!/usr/bin/env python3
"""Writes the windows of a synthetic workload as the three FASTA files the `poa` CLI reads.
  python tools/dump_fasta.py READS CONFIG PREFIX   ->  PREFIX.ref.fa / .cor.fa / .unc.fa"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads  # noqa: E402

reads, cfg, pre = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
wl = workloads.make_windows(cfg, reads)
n = len(wl["ref_off"]) - 1
for key in ("ref", "cor", "unc"):
    off, seq = wl[key + "_off"], wl[key].tobytes()
    with open("%s.%s.fa" % (pre, key), "wb") as f:
        f.write(b"".join(b">w%d\n%s\n" % (w, seq[off[w]:off[w + 1]]) for w in range(n)))
print("%d windows -> %s.{ref,cor,unc}.fa" % (n, pre))
