#!/usr/bin/env python3
"""A BASELINE.json config at its STATED size, streamed in chunks of reads through elector_reads_run (reads in, per-read counters out:
window cutting + alignment + merge + tally on the device, windows never leave it).  One JSON line: triplets/s on the device clock
(CUDA events of the three stages) and on the host clock of the calls (pageable host buffers in, counters out), windows, letters.
  python tools/config_bench.py CONFIG [N_READS] [CHUNK]
Parity of these shapes is held by tests/test_gpu_split.py and tests/test_split_emul.py (configs 1-4 against the compiled reference)."""
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import elector_b200  # noqa: E402
import workloads  # noqa: E402

cfg = int(sys.argv[1])
total = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else workloads.CONFIG_READS[cfg]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else {1: 10000, 2: 10000, 3: 2000, 4: 20000}[cfg]
work = tempfile.mkdtemp(prefix="cfgbench_")
gen = workloads.ensure_gen()


def load(first):
    n = min(chunk, total - first)
    pre = os.path.join(work, "r%d" % first)
    subprocess.check_call([gen, str(cfg), str(n), str(first), pre])
    hr, ref, ro = workloads.parse_two_line_fasta(pre + ".ref.fa")
    _, unc, uo = workloads.parse_two_line_fasta(pre + ".unc.fa")
    _, cor, co = workloads.parse_two_line_fasta(pre + ".cor.fa")
    for k in ("ref", "unc", "cor"):
        os.remove(pre + "." + k + ".fa")
    return ref, ro, unc, uo, cor, co, np.asarray([len(h) for h in hr], np.int32)


K = elector_b200.TALLY_FIELDS
acc = dict(triplets=0, windows=0, letters=0, split_ms=0.0, poa_ms=0.0, merge_tally_ms=0.0, wall_ms=0.0, calls=0)
sums = np.zeros(len(K), np.int64)
status = np.zeros(3, np.int64)
t_start = time.perf_counter()
with elector_b200.PoaContext(0) as ctx, ThreadPoolExecutor(1) as pool:
    nxt = pool.submit(load, 0)
    first = 0
    warm = True
    while first < total:
        ref, ro, unc, uo, cor, co, hl = nxt.result()
        n = len(hl)
        if first + n < total:
            nxt = pool.submit(load, first + n)
        if warm:      # the first call creates the scratch pools: run it twice, count the second
            ctx.reads_run(ref, ro, unc, uo, cor, co, hl, 0.1)
            warm = False
        t0 = time.perf_counter()
        got = ctx.reads_run(ref, ro, unc, uo, cor, co, hl, 0.1)
        acc["wall_ms"] += (time.perf_counter() - t0) * 1e3
        ms = ctx.last_reads_ms()
        acc["split_ms"] += ms[0]; acc["poa_ms"] += ms[1]; acc["merge_tally_ms"] += ms[2]
        acc["triplets"] += n; acc["windows"] += got["n_windows"]; acc["letters"] += int(ro[-1] + uo[-1] + co[-1]); acc["calls"] += 1
        sums += got["sums"]
        for k in range(3):
            status[k] += int((got["status"] == k).sum())
        first += n
dev_ms = acc["split_ms"] + acc["poa_ms"] + acc["merge_tally_ms"]
out = {"config": cfg, "workload": workloads.CONFIG_NAMES[cfg], "stated_size": workloads.CONFIG_READS[cfg], "triplets": acc["triplets"], "chunk_triplets": chunk,
       "calls": acc["calls"], "windows": acc["windows"], "letters": acc["letters"],
       "device_ms": {k: round(acc[k], 2) for k in ("split_ms", "poa_ms", "merge_tally_ms")},
       "triplets_per_s_device": round(acc["triplets"] / dev_ms * 1e3), "windows_per_s_device": round(acc["windows"] / dev_ms * 1e3),
       "triplets_per_s_host_clock": round(acc["triplets"] / acc["wall_ms"] * 1e3), "host_clock_ms": round(acc["wall_ms"], 1),
       "status_counts": {"cut": int(status[0]), "small_reads": int(status[1]), "wrongly_cor_reads": int(status[2])},
       "counters": {k: int(sums[K.index(k)]) for k in ("TP", "FP", "FN", "insC", "delC", "subsC", "insU", "delU", "subsU", "assessed")},
       "total_wall_s_with_generation": round(time.perf_counter() - t_start, 1)}
print(json.dumps(out))
