#!/usr/bin/env python3
"""Wall time of single `poa` / `masterSplitter` calls through the persistent service (csrc/service.h) and without it, on the
shards of a few hundred config-1 reads:  python tools/service_probe.py [N_READS]"""
import hashlib
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import elector_b200  # noqa: E402
import workloads  # noqa: E402

n = sys.argv[1] if len(sys.argv) > 1 else "400"
d = tempfile.mkdtemp(prefix="svc_")
elector_b200.write_default_matrix(d + "/blosum80.mat")
subprocess.check_call([workloads.ensure_gen(), "1", n, "0", d + "/r"])
B = os.path.join(ROOT, "elector_b200", "bin")
for mode in ("1", "0"):
    env = dict(os.environ, ELECTOR_SERVICE=mode)
    for i in range(3):
        o = "%s/o%s_%d" % (d, mode, i)
        os.makedirs(o)
        t0 = time.perf_counter()
        rc = subprocess.call([B + "/masterSplitter", d + "/r.ref.fa", d + "/r.unc.fa", d + "/r.cor.fa", o + "/out1", o + "/out2", o + "/out3", "7", "200", "10000", "0.1", o],
                             stdout=subprocess.DEVNULL, env=env)
        print("service=%s masterSplitter call %d rc=%d %.3f s" % (mode, i, rc, time.perf_counter() - t0), flush=True)
    for i in range(6):
        sh = i % 3
        t0 = time.perf_counter()
        rc = subprocess.call([B + "/poa", "-pir", o + "/smsa%d" % sh, "-preserve_seqorder", "-corrected_reads_fasta", o + "/out3%d" % sh, "-reference_reads_fasta", o + "/out1%d" % sh,
                              "-uncorrected_reads_fasta", o + "/out2%d" % sh, "-preserve_seqorder", "-threads", "1", "-pathMatrix", d + "/blosum80.mat"], stdout=subprocess.DEVNULL, env=env)
        print("service=%s poa shard %d call %d rc=%d %.3f s md5 %s" % (mode, sh, i, rc, time.perf_counter() - t0, hashlib.md5(open(o + "/smsa%d" % sh, "rb").read()).hexdigest()), flush=True)
