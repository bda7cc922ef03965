#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- the WHOLE README example (README.md:132-160) as a golden fixture.

Runs here (the authoring container has /root/reference); writes
  tests/golden/example_reads.tar.xz   the three example read files (perfect / uncorrected / corrected
                                      `*_elector.fa`, 13.6 MB of letters -> 1.6 MB): the INPUT of the chain
  tests/golden/example_full.json.gz   what the UNMODIFIED reference makes of them, stage by stage:
      splitter   md5 + record count of every non-empty out1<i>/out2<i>/out3<i> (oracle/_ref/masterSplitter,
                 alignment.py:98-104 command line), small_reads / wrongly_cor_reads
      poa        md5 of every smsa<i> (oracle/_ref/poa, alignment.py:60 command line)
      merge      md5 of msa.fa (oracle/_ref/Donatello over smsa0..199, alignment.py:121-127)
      tally      per merged record the integer counters of the reference's own computeStats.py functions
                 (all records, not a selection)
      report     per_read_metrics.txt, read_size_distribution.txt, the summary block printed and logged by
                 computeStats.outputRecallPrecision (imported from /root/reference, Bio stubbed)
The GPU box has no /root/reference: tests unpack the reads, run the chain under test and compare with this file.
"""
import contextlib
import gzip
import hashlib
import io
import json
import os
import subprocess
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")
EXAMPLE = "/root/reference/example"
sys.path.insert(0, ROOT)
from oracle import make_golden as mg  # noqa: E402
from oracle import make_golden_tally as mt  # noqa: E402  (imports the reference's computeStats as mt.cs)

FILES = ["perfect_reads_elector.fa", "uncorrected_reads_elector.fa", "corrected_reads_elector.fa"]


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def main():
    work = tempfile.mkdtemp(prefix="golden_example_")
    # ---- the inputs, as they are ----
    tar_path = os.path.join(GOLD, "example_reads.tar.xz")
    with tarfile.open(tar_path, "w:xz", preset=9) as tf:
        for f in FILES:
            ti = tf.gettarinfo(os.path.join(EXAMPLE, f), arcname=f)
            ti.mtime = 0; ti.uid = ti.gid = 0; ti.uname = ti.gname = ""
            with open(os.path.join(EXAMPLE, f), "rb") as fh:
                tf.addfile(ti, fh)
    info = {"files": {f: md5(os.path.join(EXAMPLE, f)) for f in FILES}}
    # ---- readAndSortFiles (emulated, Biopython is not here: oracle/make_golden.py prep_example) ----
    mg.prep_example(work)
    info["sorted"] = {k: md5(os.path.join(work, k + ".fa")) for k in ("ref", "unc", "cor")}
    # ---- splitter, alignment.py:98-104 ----
    out = os.path.join(work, "out")
    os.makedirs(out)
    subprocess.check_call([os.path.join(REF, "masterSplitter"), os.path.join(work, "ref.fa"), os.path.join(work, "unc.fa"),
                           os.path.join(work, "cor.fa"), out + "/out1", out + "/out2", out + "/out3", "7", "200", "10000", "0.1", out],
                          stdout=subprocess.DEVNULL)
    shards = [i for i in range(200) if os.path.getsize("%s/out3%d" % (out, i)) > 0]
    info["shards"] = shards
    info["splitter"] = {str(i): {k: md5("%s/%s%d" % (out, k, i)) for k in ("out1", "out2", "out3")} for i in shards}
    info["splitter_records"] = {str(i): open("%s/out1%d" % (out, i)).read().count(">") for i in shards}
    info["small_reads"] = int(open(out + "/small_reads.txt").read())
    info["wrongly_cor_reads"] = int(open(out + "/wrongly_cor_reads.txt").read())
    # ---- poa per shard, alignment.py:59-63 ----
    procs = [subprocess.Popen([os.path.join(REF, "poa"), "-pir", "%s/smsa%d" % (out, i), "-preserve_seqorder", "-corrected_reads_fasta",
                               "%s/out3%d" % (out, i), "-reference_reads_fasta", "%s/out1%d" % (out, i), "-uncorrected_reads_fasta",
                               "%s/out2%d" % (out, i), "-preserve_seqorder", "-threads", "1", "-pathMatrix", os.path.join(REF, "blosum80.mat")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for i in shards]
    for p in procs:
        assert p.wait() == 0
    info["smsa_md5"] = {str(i): md5("%s/smsa%d" % (out, i)) for i in shards}
    # ---- Donatello, alignment.py:121-127 ----
    msa = os.path.join(work, "msa.fa")
    for i in range(200):
        if os.path.exists("%s/smsa%d" % (out, i)):
            subprocess.call([os.path.join(REF, "Donatello"), "%s/smsa%d" % (out, i), msa])
    info["msa_md5"] = md5(msa)
    assert info["msa_md5"] == "c91df333ac7b41a819fa1bd5d40677cd", info["msa_md5"]
    # ---- tally: every merged record through the reference's functions ----
    lines = open(msa).read().split("\n")
    recs = [(lines[i], lines[i + 1], lines[i + 3], lines[i + 5]) for i in range(0, len(lines) - 5, 6)]
    info["records"] = [{"header": h, "expect": mt.reference_counters(R, C, U)} for h, R, C, U in recs]
    # ---- report: the reference's outputRecallPrecision on its own msa.fa ----
    rep = os.path.join(work, "report")
    os.makedirs(rep)
    os.symlink(msa, os.path.join(rep, "msa.fa"))
    log = io.StringIO()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        mt.cs.outputRecallPrecision(os.path.join(work, "cor.fa"), rep, log, info["small_reads"], info["wrongly_cor_reads"], 5, 0.1,
                                    "read_size_distribution.txt", {}, 0, 0, None)
    info["report"] = {"per_read_metrics": open(rep + "/per_read_metrics.txt").read(),
                      "per_read_metrics_md5": md5(rep + "/per_read_metrics.txt"),
                      "read_size_distribution_md5": md5(rep + "/read_size_distribution.txt"),
                      "read_size_distribution_lines": open(rep + "/read_size_distribution.txt").read().count("\n"),
                      "stdout": buf.getvalue(), "log": log.getvalue()}
    assert info["report"]["per_read_metrics_md5"] == "5507a6528f193ac9a87b420d6856964d", info["report"]["per_read_metrics_md5"]
    with gzip.GzipFile(os.path.join(GOLD, "example_full.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(info, sort_keys=True).encode())
    print("example: %d shards, %d merged records (%d assessed); reads %d bytes, json %d bytes; work dir %s"
          % (len(shards), len(recs), sum(r["expect"]["assessed"] for r in info["records"]), os.path.getsize(tar_path),
             os.path.getsize(os.path.join(GOLD, "example_full.json.gz")), work))
    print(buf.getvalue())


if __name__ == "__main__":
    main()
