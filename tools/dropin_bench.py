#!/usr/bin/env python3
"""ELECTOR's alignment + report stage three ways on the same prepared read files, wall clock of each:
  reference   the reference's alignment.getPOA + computeStats.outputRecallPrecision, its own binaries, `threads` processes
  swap        the same Python with bin/poa and bin/masterSplitter swapped for the CUDA executables (INTEGRATION.md level 1)
  inprocess   elector_b200.alignment.getPOA + elector_b200.computeStats.outputRecallPrecision (level 3, one persistent context)
and whether msa.fa / per_read_metrics.txt are the same bytes.
  python tools/dropin_bench.py example|CONFIG:N_READS [THREADS]"""
import hashlib
import io
import json
import os
import contextlib
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
TREE = os.path.join(ROOT, "oracle", "_ref", "elector_tree")

what = sys.argv[1] if len(sys.argv) > 1 else "example"
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
tmp = tempfile.mkdtemp(prefix="dropin_")
work = os.path.join(tmp, "work")
os.makedirs(work)
if what == "example":
    from oracle import example_prep as ep
    src = os.path.join(tmp, "src"); os.makedirs(src)
    ep.unpack(src); ep.sort_and_duplicate(src, work)
else:
    import workloads
    cfg, n = what.split(":")
    pre = os.path.join(work, "r")
    subprocess.check_call([workloads.ensure_gen(), cfg, n, "0", pre])
    for k in ("ref", "unc", "cor"):
        os.rename(pre + "." + k + ".fa", os.path.join(work, k + ".fa"))


def md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest() if os.path.exists(p) else None


def run_tree(tree, out):
    code = ("import time, sys\nimport elector.alignment as a, elector.computeStats as c\nt0 = time.perf_counter()\n"
            "r = a.getPOA(%r, %r, %r, %d, %r, 0.1)\nt1 = time.perf_counter()\n"
            "c.outputRecallPrecision(%r, %r, open(%r, 'w'), r[0], r[1], 5, 0.1, 'read_size_distribution.txt', {}, 0, 0, None)\nt2 = time.perf_counter()\n"
            "print('TIMES', t1 - t0, t2 - t1)\n" % (work + "/cor.fa", work + "/ref.fa", work + "/unc.fa", threads, out, work + "/cor.fa", out, out + "/log"))
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, "-c", code], cwd=tree, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        return {"error": p.stderr[-500:]}
    t = [l for l in p.stdout.split("\n") if l.startswith("TIMES")][0].split()
    return {"getPOA_s": round(float(t[1]), 3), "outputRecallPrecision_s": round(float(t[2]), 3), "process_wall_s": round(wall, 3),
            "msa_md5": md5(out + "/msa.fa"), "per_read_metrics_md5": md5(out + "/per_read_metrics.txt")}


res = {"input": what, "threads": threads, "triplets": open(work + "/ref.fa").read().count(">")}
out = os.path.join(tmp, "out_ref"); os.makedirs(out)
res["reference"] = run_tree(TREE, out)
from test_gpu_dropin import swapped_tree  # noqa: E402
tree = swapped_tree(os.path.join(tmp, "tree"))
for rep in range(2):    # the second run has the binaries and the CUDA driver state in the page cache
    out = os.path.join(tmp, "out_swap%d" % rep); os.makedirs(out)
    res["swap" if rep else "swap_first"] = run_tree(tree, out)
from elector_b200 import alignment, computeStats  # noqa: E402
for rep in range(2):    # the first call creates the context and grows the buffers
    out = os.path.join(tmp, "out_in%d" % rep); os.makedirs(out)
    t0 = time.perf_counter()
    r = alignment.getPOA(work + "/cor.fa", work + "/ref.fa", work + "/unc.fa", threads, out, 0.1)
    t1 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        computeStats.outputRecallPrecision(work + "/cor.fa", out, open(out + "/log", "w"), r[0], r[1], 5, 0.1, "read_size_distribution.txt", {}, 0, 0, None)
    t2 = time.perf_counter()
    res["inprocess" if rep else "inprocess_first"] = {"getPOA_s": round(t1 - t0, 3), "outputRecallPrecision_s": round(t2 - t1, 3), "msa_md5": md5(out + "/msa.fa"),
                                                      "per_read_metrics_md5": md5(out + "/per_read_metrics.txt")}
res["same_bytes"] = len({res[k].get("msa_md5") for k in ("reference", "swap", "inprocess")}) == 1 and \
    len({res[k].get("per_read_metrics_md5") for k in ("reference", "swap", "inprocess")}) == 1
shutil.rmtree(tmp, ignore_errors=True)
print(json.dumps(res))
