"""TEST INFRASTRUCTURE ONLY -- turns the README example reads (tests/golden/example_reads.tar.xz, or the reference's own
example/ directory) into what alignment.py hands to the splitter and `poa`.

  unpack(dst)              the three `*_elector.fa` files
  sort_and_duplicate(...)  Biopython-free restatement of elector/readAndSortFiles.py for files whose headers are
                           already in ELECTOR's format: formatHeader (:212, the `_<n>` suffix of corrected headers
                           dropped), readAndSortFasta (:150-166, records sorted by header), duplicateRefReads
                           (:171-191, one `_<k>` copy of the reference / uncorrected read per corrected fragment)
  reference_split(...)     oracle/_ref/masterSplitter with alignment.py:104's arguments -> non-empty shard ids
"""
import os
import re
import subprocess
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
TAR = os.path.join(ROOT, "tests", "golden", "example_reads.tar.xz")
FILES = ("perfect_reads_elector.fa", "uncorrected_reads_elector.fa", "corrected_reads_elector.fa")


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


def unpack(dst):
    with tarfile.open(TAR, "r:xz") as tf:
        for m in tf.getmembers():
            assert m.name in FILES and m.isfile()
            with open(os.path.join(dst, m.name), "wb") as f:
                f.write(tf.extractfile(m).read())
    return dst


def sort_and_duplicate(src_dir, work):
    """-> work/ref.fa, work/unc.fa, work/cor.fa (one sequence per line)"""
    ref = read_fa(os.path.join(src_dir, FILES[0]))
    unc = read_fa(os.path.join(src_dir, FILES[1]))
    cor = [(re.sub(r"_[0-9]*$", "", h), s) for h, s in read_fa(os.path.join(src_dir, FILES[2]))]
    ref.sort(key=lambda x: x[0]); unc.sort(key=lambda x: x[0]); cor.sort(key=lambda x: x[0])
    occ = {}
    for h, _ in cor:
        occ[h] = occ.get(h, 0) + 1
    with open(os.path.join(work, "cor.fa"), "w") as f:
        for h, s in cor:
            f.write(">%s\n%s\n" % (h, s))
    with open(os.path.join(work, "ref.fa"), "w") as fr, open(os.path.join(work, "unc.fa"), "w") as fu:
        for (hr, sr), (_, su) in zip(ref, unc):
            for t in range(occ.get(hr, 0)):
                fr.write(">%s_%d\n%s\n" % (hr, t, sr))
                fu.write(">%s_%d\n%s\n" % (hr, t, su))


def reference_split(work, out, amount=10000, threshold="0.1", nfiles=200, exe=None):
    """alignment.py:104: masterSplitter ref unc cor out1 out2 out3 7 200 <amount> <threshold> <outDir>"""
    os.makedirs(out, exist_ok=True)
    rc = subprocess.call([exe or os.path.join(REF, "masterSplitter"), os.path.join(work, "ref.fa"), os.path.join(work, "unc.fa"),
                          os.path.join(work, "cor.fa"), out + "/out1", out + "/out2", out + "/out3", "7", str(nfiles), str(amount),
                          str(threshold), out], stdout=subprocess.DEVNULL)
    return rc, [i for i in range(nfiles) if os.path.getsize("%s/out3%d" % (out, i)) > 0]
