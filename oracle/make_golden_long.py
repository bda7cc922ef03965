#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- golden set `long` (SURVEY.md section 4 edge list): windows far beyond the ~50-letter bulk,
run through the UNMODIFIED reference (oracle/_ref/poa and oracle/_ref/ref_dump, built by oracle/build_ref.sh).

  w0  3 300-letter window (cor 3 %, unc 12 % errors): scores leave the 16-bit range of the packed kernels -> INT32 tier
  w1  33 500-letter reference and a 33 400-letter uncorrected sequence, each on ONE FASTA line longer than the
      reference's 32 KiB line buffer (fasta_format.c:20: the line arrives in several fgets chunks), corrected = `N`
      placeholder: a window longer than the former 32 000-letter cap of the CUDA path
  w2  a 1 000-letter window whose FASTA lines are wrapped at 70 columns, with blank lines and a trailing blank
  w3  a 700-letter window (warp-cooperative tier) after the long ones: the reader is back in step
Writes tests/golden/long.{ref,cor,unc}.fa.gz (the FASTA files as the reader must take them), long.pir.gz and
long.dump.gz (scores and len(P1) only: the alignment maps of a 33 000-letter window would be 0.5 MB of text).
"""
import gzip
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402


def wrap(s, n=70):
    return "\n".join(s[i:i + n] for i in range(0, len(s), n))


def main():
    rng = synth.SplitMix64(4242)
    work = tempfile.mkdtemp(prefix="golden_long_")

    def rand(L):
        return "".join("ACGT"[rng.below(4)] for _ in range(L))
    wins = []
    r = rand(3300); wins.append(("long_int32", r, synth.mutate(rng, r, 0.03, "ACGT"), synth.mutate(rng, r, 0.12, "ACGT")))
    r = rand(33500); u = synth.mutate(rng, r, 0.10, "ACGT")[:33400]; wins.append(("long_line title of the long one", r, "N", u))
    r = rand(1000); wins.append(("long_wrapped", r, synth.mutate(rng, r, 0.01, "ACGT"), synth.mutate(rng, r, 0.10, "ACGT")))
    r = rand(700); wins.append(("long_after", r, r, synth.mutate(rng, r, 0.10, "ACGT")))
    assert len(wins[1][1]) > 32768 and len(wins[1][3]) > 32768
    pre = os.path.join(work, "long")
    for k, ext in ((1, "ref"), (2, "cor"), (3, "unc")):
        with open("%s.%s.fa" % (pre, ext), "w") as f:
            for i, w in enumerate(wins):
                if i == 2:
                    f.write(">%s\n\n%s\n\n" % (w[0], wrap(w[k])))
                else:
                    f.write(">%s\n%s\n" % (w[0], w[k]))
    mat = os.path.join(REF, "blosum80.mat")
    rc = subprocess.call([os.path.join(REF, "poa"), "-pir", pre + ".pir", "-preserve_seqorder", "-corrected_reads_fasta", pre + ".cor.fa",
                          "-reference_reads_fasta", pre + ".ref.fa", "-uncorrected_reads_fasta", pre + ".unc.fa", "-preserve_seqorder",
                          "-threads", "1", "-pathMatrix", mat], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert rc == 0, rc
    subprocess.check_call([os.path.join(REF, "ref_dump"), mat, pre + ".ref.fa", pre + ".cor.fa", pre + ".unc.fa", pre + ".hpir", pre + ".fulldump"],
                          stdout=subprocess.DEVNULL)
    assert open(pre + ".hpir", "rb").read() == open(pre + ".pir", "rb").read(), "harness PIR != poa PIR"
    with open(pre + ".dump", "w") as f:
        for line in open(pre + ".fulldump"):
            if line.split()[0] in ("W", "S1", "S2", "L"):
                f.write(line)
    for ext in (".ref.fa", ".cor.fa", ".unc.fa", ".pir", ".dump"):
        with open(pre + ext, "rb") as f, gzip.GzipFile(os.path.join(GOLD, "long" + ext + ".gz"), "wb", mtime=0) as g:
            g.write(f.read())
    print(open(pre + ".dump").read())
    print("sizes:", {e: os.path.getsize(os.path.join(GOLD, "long" + e + ".gz")) for e in (".ref.fa", ".cor.fa", ".unc.fa", ".pir", ".dump")})


if __name__ == "__main__":
    main()
