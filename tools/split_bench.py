#!/usr/bin/env python3
"""Reads in, counters out (elector_reads_run: window cutting + alignment + merge + tally, windows never leave the device) on
synthetic reads of one config, next to the reference masterSplitter on the same host for a sample of the reads.
  python tools/split_bench.py CONFIG N_READS [REPS]"""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import elector_b200  # noqa: E402
import workloads  # noqa: E402

cfg, n_reads = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
work = tempfile.mkdtemp(prefix="splitbench_")
pre = os.path.join(work, "r")
subprocess.check_call([workloads.ensure_gen(), str(cfg), str(n_reads), "0", pre])
hr, ref, ro = workloads.parse_two_line_fasta(pre + ".ref.fa")
_, unc, uo = workloads.parse_two_line_fasta(pre + ".unc.fa")
_, cor, co = workloads.parse_two_line_fasta(pre + ".cor.fa")
hl = np.asarray([len(h) for h in hr], np.int32)
n = len(hr)
out = {"config": cfg, "triplets": n, "letters_ref": int(ro[-1])}
with elector_b200.PoaContext(0) as ctx:
    for rep in range(reps + 1):
        t0 = time.perf_counter()
        got = ctx.reads_run(ref, ro, unc, uo, cor, co, hl, 0.1)
        wall = time.perf_counter() - t0
        ms = ctx.last_reads_ms()
        if rep:   # the first call grows the buffers
            out.setdefault("calls", []).append({"wall_ms": round(wall * 1e3, 2), "split_ms": round(ms[0], 3), "poa_ms": round(ms[1], 3), "merge_tally_ms": round(ms[2], 3)})
    out["windows"] = got["n_windows"]
    out["status_counts"] = [int((got["status"] == k).sum()) for k in range(3)]
    out["k_used"] = {int(k): int((got["k_used"][got["status"] == 0] == k).sum()) for k in (15, 13, 11, 9)}
best = min(c["wall_ms"] for c in out["calls"])
out["triplets_per_s_wall"] = round(n / best * 1e3)
ref_exe = workloads.SPLITTER
if os.path.exists(ref_exe):
    sample = min(n_reads, 400 if cfg != 3 else 40)
    spre = os.path.join(work, "s")
    subprocess.check_call([workloads.ensure_gen(), str(cfg), str(sample), "0", spre])
    o = os.path.join(work, "o"); os.makedirs(o)
    t0 = time.perf_counter()
    subprocess.call([ref_exe, spre + ".ref.fa", spre + ".unc.fa", spre + ".cor.fa", o + "/out1", o + "/out2", o + "/out3", "7", "200", "10000", "0.1", o], stdout=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    ns = len(open(spre + ".ref.fa").read().split("\n")) // 2
    out["reference_masterSplitter"] = {"triplets": ns, "seconds": round(dt, 2), "ms_per_triplet": round(dt / ns * 1e3, 2), "threads": 1}
print(json.dumps(out))
