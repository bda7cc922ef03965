#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- golden vectors for the merge + tally half of the path, produced
by the REFERENCE itself: elector/computeStats.py imported from /root/reference (with an empty
stub for the unused Bio import) and the compiled oracle/_ref/Donatello.  Writes
  tests/golden/tally_example.json.gz   rows of selected example reads + the reference's counters
  tests/golden/tally_random.json.gz    reference counters for oracle.synth.random_msa_rows(3000, 21)
  tests/golden/donatello.pir.gz / donatello.msa.gz   poa output of the first reads of example
                                        shard 0 and Donatello's merge of it
Run after oracle/make_golden.py (needs its work dir: pass it as argv[1]).
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
stub = tempfile.mkdtemp()
os.makedirs(os.path.join(stub, "Bio"))
open(os.path.join(stub, "Bio", "__init__.py"), "w").close()
open(os.path.join(stub, "Bio", "SeqIO.py"), "w").close()
sys.path.insert(0, stub)
sys.path.insert(0, "/root/reference")
import elector.computeStats as cs  # noqa: E402  (the reference, unmodified)
from oracle import synth  # noqa: E402


def reference_counters(R, C, U):
    L = len(R)
    out = {"ncols": L, "assessed": int(L > 10)}
    if L <= 10:
        return out
    ext = []
    gp, isExt, ext, missing, stretches, isTrim, tot = cs.gapsAndExtensions(R, C, U, [], False, False, ext, 0)
    existing, _ = cs.getCorrectedPositions(stretches, C, 0, R, {}, "h", gp)
    FP, TP, FN, cor, uncor, ucc, ucu, gcr, gcc, insU, delU, subsU, insC, delC, subsC, _ = cs.getTPFNFP(R, C, U, existing, 5, [], gp)
    out.update(dict(TP=TP, FP=FP, FN=FN, cor=cor, uncor=uncor, uncorCor=ucc, uncorUncor=ucu, insC=insC, delC=delC, subsC=subsC,
                    insU=insU, delU=delU, subsU=subsU, GCrateRef=gcr, GCrateCor=gcc, lenRef=cs.getLen(R), lenCor=cs.getLen(C),
                    lenUnc=cs.getLen(U), gapsLeft=min(cs.nbLeftGaps(R), cs.nbLeftGaps(U)), gapsRight=min(cs.nbRightGaps(R), cs.nbRightGaps(U)),
                    missing=missing, extended=(sum(ext) if isExt else -1), stretches=sorted([int(k), int(v)] for k, v in stretches.items()),
                    isTrimmed=bool(isTrim), mask_true=int(sum(existing))))
    return out


def main():
    work = sys.argv[1]
    lines = open(os.path.join(work, "msa.fa")).read().split("\n")
    recs = [(lines[i], lines[i + 1], lines[i + 3], lines[i + 5]) for i in range(0, len(lines) - 5, 6)]
    print(len(recs), "merged example records")
    chosen = []
    for k, (h, R, C, U) in enumerate(recs):
        if len(R) <= 10:
            if sum(1 for c in chosen if c[4]) < 3:
                chosen.append((h, R, C, U, True))
            continue
        gl = min(cs.nbLeftGaps(R), cs.nbLeftGaps(U)); gr = min(cs.nbRightGaps(R), cs.nbRightGaps(U))
        st = cs.findGapStretches(C, R, [])
        interesting = gl >= 5 or gr >= 5 or len(st) > 0
        if (interesting and sum(1 for c in chosen if not c[4] and c[5]) < 22) or (not interesting and k % 40 == 0 and sum(1 for c in chosen if not c[4] and not c[5]) < 8):
            chosen.append((h, R, C, U, False, interesting))
    ex = []
    for c in chosen:
        h, R, C, U = c[:4]
        ex.append({"header": h, "R": R, "C": C, "U": U, "expect": reference_counters(R, C, U)})
    with gzip.GzipFile(os.path.join(GOLD, "tally_example.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(ex).encode())
    print("tally_example:", len(ex), "records,", sum(1 for e in ex if e["expect"].get("stretches")), "with stretches")
    rnd = [reference_counters(R, C, U) for R, C, U in synth.random_msa_rows(3000, 21)]
    with gzip.GzipFile(os.path.join(GOLD, "tally_random.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(rnd).encode())
    print("tally_random:", len(rnd), "rows,", sum(1 for e in rnd if e.get("stretches")), "with stretches,", sum(1 for e in rnd if e.get("extended", -1) >= 0), "extended")
    # Donatello fixture: the first 6 reads of shard 0 plus one 'AAA' placeholder read
    pir = open(os.path.join(work, "out", "smsa0")).read().split("\n")
    take, heads = [], []
    for i in range(0, len(pir) - 5, 6):
        h = pir[i + 4]
        if h not in heads:
            if len(heads) == 6:
                break
            heads.append(h)
        take += pir[i:i + 6]
    sub = os.path.join(work, "donatello.pir")
    open(sub, "w").write("\n".join(take) + "\n")
    msa = os.path.join(work, "donatello.msa")
    if os.path.exists(msa):
        os.remove(msa)
    subprocess.check_call([os.path.join(HERE, "_ref", "Donatello"), sub, msa])
    for name in ("donatello.pir", "donatello.msa"):
        with open(os.path.join(work, name), "rb") as f, gzip.GzipFile(os.path.join(GOLD, name + ".gz"), "wb", mtime=0) as g:
            g.write(f.read())
    print("donatello:", len(take) // 6, "windows ->", open(msa).read().count(">") // 3, "reads")


if __name__ == "__main__":
    main()
