#!/usr/bin/env python3
"""Device timeline of the segment launches of one call (ELECTOR_TRACE) for a cached workload.
  python tools/seg_trace.py [reads] [config] [chunks]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elector_b200  # noqa: E402
import workloads  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 1
os.environ["ELECTOR_PIPELINE_CHUNKS"] = sys.argv[3] if len(sys.argv) > 3 else "1"
wl = workloads.make_windows(cfg, reads)
with elector_b200.PoaContext(0) as ctx:
    for k in range(3):
        if k == 2:
            os.environ["ELECTOR_TRACE"] = os.environ.get("DIAG_TRACE", "1")
        res, counters, sums = ctx.pipeline_csr(wl["ref"], wl["ref_off"], wl["cor"], wl["cor_off"], wl["unc"], wl["unc_off"], wl["read_first"])
        ms, n = ctx.last_kernel_ms()
        print("call %d: %d windows, %d launches, %.3f ms kernels" % (k, len(res.nring), n, ms), flush=True)
