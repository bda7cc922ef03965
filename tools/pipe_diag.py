#!/usr/bin/env python3
"""Where does the time of the host-buffer call (elector_pipeline_run) go?  PCIe bandwidth of this box with
pinned buffers, then the call at several chunk counts with the per-chunk device timeline (ELECTOR_TRACE).
  python tools/pipe_diag.py [reads] [config]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elector_b200  # noqa: E402
import workloads  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
nbytes = 350 << 20
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


print("H2D pinned %.1f GB/s" % (nbytes / timed(lambda: d.copy_(h, non_blocking=True)) / 1e9))
print("D2H pinned %.1f GB/s" % (nbytes / timed(lambda: h2.copy_(d2, non_blocking=True)) / 1e9))
print("H2D + D2H together: %.1f GB/s each way" % (nbytes / timed(both) / 1e9))

wl = workloads.make_windows(cfg, reads)
keys = ("ref", "ref_off", "cor", "cor_off", "unc", "unc_off", "read_first")
hp = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).pin_memory() for k in keys}
n, n_reads = len(wl["ref_off"]) - 1, len(wl["read_first"]) - 1
K = len(elector_b200.TALLY_FIELDS)
with elector_b200.PoaContext(0) as ctx:
    lib = ctx._lib
    ptr = {k: hp[k].numpy().ctypes.data for k in keys}
    bound = int(lib.elector_poa_rows_bound(n, ptr["ref_off"], ptr["cor_off"], ptr["unc_off"]))
    rows = torch.empty(bound, dtype=torch.uint8).pin_memory()
    ro = torch.empty(n, dtype=torch.int64).pin_memory()
    st = torch.empty(n, dtype=torch.int32).pin_memory()
    nr = torch.empty(n, dtype=torch.int32).pin_memory()
    cnt = torch.empty(n_reads * K, dtype=torch.int64).pin_memory()
    sums = torch.empty(K, dtype=torch.int64).pin_memory()

    def call():
        ctx._check(lib.elector_pipeline_run(ctx._ctx, n, ptr["ref"], ptr["ref_off"], ptr["cor"], ptr["cor_off"], ptr["unc"], ptr["unc_off"],
                                            n_reads, ptr["read_first"], rows.data_ptr(), bound, ro.data_ptr(), st.data_ptr(), nr.data_ptr(),
                                            None, None, None, cnt.data_ptr(), sums.data_ptr()))

    for chunks in [int(x) for x in os.environ.get("DIAG_CHUNKS", "1,2,3,4,6,8").split(",")]:
        os.environ["ELECTOR_PIPELINE_CHUNKS"] = str(chunks)
        os.environ.pop("ELECTOR_TRACE", None)
        for _ in range(3):
            call()
        t = timed(call, reps=5)
        ms, k = ctx.last_kernel_ms()
        print("chunks %d: %.2f ms per call (%.0f triplets/s), kernels %.2f ms, %d launches" % (chunks, t * 1e3, n_reads / t, ms, k), flush=True)
        os.environ["ELECTOR_TRACE"] = os.environ.get("DIAG_TRACE", "1")
        call()
        sys.stderr.flush()
