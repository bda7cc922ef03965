#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/ from the UNMODIFIED reference
binaries in oracle/_ref (built by oracle/build_ref.sh from /root/reference).

Runs here (the authoring container has /root/reference); the fixtures it writes are
committed so that the GPU box, where the reference tree does not exist, can check
parity against the reference's own outputs.

Fixture sets (each: <set>.ref.fa / .cor.fa / .unc.fa inputs in the splitter's shard
format, <set>.pir = output of the reference `poa`, <set>.dump = DP scores and
alignment maps from oracle/_ref/ref_dump; all gzip'ed):
  example_head   first 500 windows of shard 0 of the README example (README.md:132)
  example_tail   the 150 largest example windows + 150 `N`-placeholder windows
  hard           3000 adversarial synthetic windows (oracle/synth.py, seed 11)
  edge           hand-written edge cases (tiny, placeholders, IUPAC, titles, mixed case)
Also writes example_md5.json: md5 of every example shard's PIR and of the merged msa.fa
(cross-check of SURVEY.md 8c) and the window/cell totals.
"""
import gzip
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")
EXAMPLE = "/root/reference/example"
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402


def read_fa(path):
    recs, h, s = [], None, []
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            if h is not None:
                recs.append((h, "".join(s)))
            h, s = line[1:], []
        else:
            s.append(line.strip())
    if h is not None:
        recs.append((h, "".join(s)))
    return recs


def prep_example(work):
    """Biopython-free emulation of elector/readAndSortFiles.py for the example (oracle/example_prep.py)."""
    from oracle import example_prep
    example_prep.sort_and_duplicate(EXAMPLE, work)


def run_poa(prefix_in, out_pir):
    cmd = [os.path.join(REF, "poa"), "-pir", out_pir, "-preserve_seqorder", "-corrected_reads_fasta", prefix_in[2],
           "-reference_reads_fasta", prefix_in[0], "-uncorrected_reads_fasta", prefix_in[1], "-preserve_seqorder",
           "-threads", "1", "-pathMatrix", os.path.join(REF, "blosum80.mat")]
    return subprocess.call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def gz(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)


def emit_set(name, wins, work, titles=False):
    """wins: list of (header_without_gt, ref, cor, unc); runs the reference on them and stores the fixture"""
    pre = os.path.join(work, name)
    with open(pre + ".ref.fa", "w") as fr, open(pre + ".cor.fa", "w") as fc, open(pre + ".unc.fa", "w") as fu:
        for h, r, c, u in wins:
            fr.write(">%s\n%s\n" % (h, r)); fc.write(">%s\n%s\n" % (h, c)); fu.write(">%s\n%s\n" % (h, u))
    rc = run_poa((pre + ".ref.fa", pre + ".unc.fa", pre + ".cor.fa"), pre + ".pir")
    assert rc == 0, (name, rc)
    subprocess.check_call([os.path.join(REF, "ref_dump"), os.path.join(REF, "blosum80.mat"), pre + ".ref.fa",
                           pre + ".cor.fa", pre + ".unc.fa", pre + ".hpir", pre + ".dump"], stdout=subprocess.DEVNULL)
    assert open(pre + ".hpir", "rb").read() == open(pre + ".pir", "rb").read(), "harness PIR != poa PIR"
    for ext in (".ref.fa", ".cor.fa", ".unc.fa", ".pir", ".dump"):
        gz(pre + ext, os.path.join(GOLD, name + ext + ".gz"))
    print("  %-14s %6d windows" % (name, len(wins)))


EDGE = [
    ("e_single", "A", "A", "A"), ("e_single_mm", "A", "C", "G"), ("e_two", "AC", "CA", "AC"),
    ("e_placeholder", "AAA", "AAA", "AAA"), ("e_N", "ACGTACGTACGTACGTACGT", "N", "ACGTACTACGTACGGTACGT"),
    ("e_N_short", "ACG", "N", "AG"), ("e_lower", "acgtacgtac", "acgtacgtac", "acgtaccgtac"),
    ("e_mixed", "AcGtAcGtAc", "aCgTaCgTaC", "ACGTACGTAC"), ("e_iupac", "ACGTRYKMACGT", "ACGTRYKMACGT", "ACGTNNNNACGT"),
    ("e_u_vs_t", "ACGUACGU", "ACGTACGT", "ACGUACGT"), ("e_bracket", "AC]GT?AC", "AC]GT?AC", "ACGTAC"),
    ("e_digits", "AC1GT2AC", "ACGTAC", "AC-GT.AC"), ("e_title some title here", "ACGTAC", "ACGTAC", "ACGAC"),
    ("e_homopolymer", "AAAAAAAAAAAAAAAA", "AAAAAAAAAAAAAA", "AAAAAAAAAAAAAAAAAAA"),
    ("e_all_diff", "AAAAAAAAAA", "CCCCCCCCCC", "GGGGGGGGGG"), ("e_cor_long", "ACGT", "ACGTACGTACGTACGTACGTACGT", "ACGT"),
    ("e_unc_long", "ACGT", "ACGT", "TTTTACGTACGTACGTACGTACGTTTTT"), ("e_ref_long", "ACGTACGTACGTACGTACGTACGTAAAA", "ACGT", "CGTA"),
    ("e_trim_left", "ACGTTGCAACGTTGCAGGCC", "GCAGGCC", "ACGTTGCACGTTGCAGGCC"),
    ("e_trim_right", "ACGTTGCAACGTTGCAGGCC", "ACGTTGC", "ACGTTGCAACGTGCAGGCC"),
    ("e_tie_ins", "ACAC", "ACACAC", "ACAC"), ("e_tie_del", "ACACAC", "ACAC", "ACACAC"),
    ("e_sub_run", "ACGTACGTACGT", "ACGTTTTTACGT", "ACGTAAAAACGT"), ("e_n_inside", "ACGTNACGT", "ACGTnACGT", "ACGTACGT"),
]


def main():
    if not os.path.exists(os.path.join(REF, "poa")):
        sys.exit("oracle/_ref/poa missing: run oracle/build_ref.sh (needs /root/reference)")
    os.makedirs(GOLD, exist_ok=True)
    work = tempfile.mkdtemp(prefix="golden_")
    info = {}
    # ---- example set through the unmodified splitter / poa / Donatello ----
    prep_example(work)
    out = os.path.join(work, "out")
    os.makedirs(out)
    subprocess.check_call([os.path.join(REF, "masterSplitter"), os.path.join(work, "ref.fa"), os.path.join(work, "unc.fa"),
                           os.path.join(work, "cor.fa"), out + "/out1", out + "/out2", out + "/out3", "7", "200", "10000", "0.1",
                           out], stdout=subprocess.DEVNULL)
    shards = [i for i in range(200) if os.path.getsize("%s/out3%d" % (out, i)) > 0]
    procs = []
    for i in shards:
        procs.append(subprocess.Popen([os.path.join(REF, "poa"), "-pir", "%s/smsa%d" % (out, i), "-corrected_reads_fasta",
                                       "%s/out3%d" % (out, i), "-reference_reads_fasta", "%s/out1%d" % (out, i),
                                       "-uncorrected_reads_fasta", "%s/out2%d" % (out, i), "-pathMatrix",
                                       os.path.join(REF, "blosum80.mat")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in procs:
        assert p.wait() == 0
    msa = os.path.join(work, "msa.fa")
    for i in range(200):
        subprocess.call([os.path.join(REF, "Donatello"), "%s/smsa%d" % (out, i), msa])
    info["shards"] = shards
    info["smsa_md5"] = {str(i): md5("%s/smsa%d" % (out, i)) for i in shards}
    info["msa_md5"] = md5(msa)
    info["survey_msa_md5"] = "c91df333ac7b41a819fa1bd5d40677cd"
    info["survey_smsa0_md5"] = "4d20cfdcf8b7494608878977f941f9dd"
    assert info["msa_md5"] == info["survey_msa_md5"] and info["smsa_md5"]["0"] == info["survey_smsa0_md5"]
    allw = []
    for i in shards:
        r = read_fa("%s/out1%d" % (out, i)); u = read_fa("%s/out2%d" % (out, i)); c = read_fa("%s/out3%d" % (out, i))
        assert len(r) == len(u) == len(c)
        allw += [(a[0], a[1], cc[1], b[1]) for a, b, cc in zip(r, u, c)]
    info["example_windows"] = len(allw)
    info["example_triplets"] = 484
    print("example: %d windows in %d shards" % (len(allw), len(shards)))
    emit_set("example_head", allw[:500], work)
    big = sorted(range(len(allw)), key=lambda k: -(len(allw[k][1]) + len(allw[k][2]) + len(allw[k][3])))[:150]
    nwin = [k for k in range(len(allw)) if allw[k][2] == "N"][:150]
    emit_set("example_tail", [allw[k] for k in sorted(set(big + nwin))], work)
    emit_set("hard", [(("%s some title %d" % (w[0], k)) if k % 3 == 0 else w[0], w[1], w[2], w[3])
                      for k, w in enumerate(synth.hard_windows(3000, seed=11))], work)
    emit_set("edge", EDGE, work)
    # merged msa of the first reads for the Donatello / tally fixtures
    shutil.copy(msa, os.path.join(work, "msa_full.fa"))
    with open(os.path.join(GOLD, "example_md5.json"), "w") as f:
        json.dump(info, f, indent=1, sort_keys=True)
    print("work dir kept at", work)


if __name__ == "__main__":
    main()
