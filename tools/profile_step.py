#!/usr/bin/env python3
"""Profiling driver (run under ncu): N plain steps of the hot path through the host C-ABI on a
cached workload; no torch, no subprocesses, so the profiler only sees our kernels.
  python tools/profile_step.py [reads] [steps] [config] [poa|pipeline]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elector_b200  # noqa: E402
import workloads  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mode = sys.argv[4] if len(sys.argv) > 4 else "poa"   # "poa": alignment only; "pipeline": + merge + tally (elector_pipeline_run)
wl = workloads.make_windows(cfg, reads)
with elector_b200.PoaContext(0) as ctx:
    for _ in range(steps):
        if mode == "pipeline":
            res, counters, sums = ctx.pipeline_csr(wl["ref"], wl["ref_off"], wl["cor"], wl["cor_off"], wl["unc"], wl["unc_off"], wl["read_first"])
        else:
            res = ctx.run_csr(wl["ref"], wl["ref_off"], wl["cor"], wl["cor_off"], wl["unc"], wl["unc_off"])
        ms, k = ctx.last_kernel_ms()
        a, b = ctypes.c_float(0), ctypes.c_float(0)
        ctx._lib.elector_last_phase_ms(ctx._ctx, ctypes.byref(a), ctypes.byref(b))
        print("step: %d windows, %d launches, %.3f ms kernels (phase 1 %.3f, phase 2 %.3f of %.3f), %.1f GCUPS"
              % (len(res.nring), k, ms, a.value, b.value - a.value, b.value, res.cells.sum() / ms / 1e6))
