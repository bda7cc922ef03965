"""CPU: the report half of computeStats.py through the C-ABI (elector_report_write is host code: no device needed)
against what the UNMODIFIED computeStats.outputRecallPrecision wrote for the README example (tests/golden/example_full.json.gz:
per_read_metrics.txt -- md5 5507a652..., SURVEY.md 8c --, read_size_distribution.txt, the log block and the printed block).
Inputs: the per-record counters and gap stretches the reference's own functions gave (the golden), rows from the oracle chain."""
import hashlib
import os

import numpy as np
import pytest

from conftest import md5_file


def counters_from_golden(records, rows):
    import elector_b200
    F = elector_b200.TALLY_FIELDS
    n = len(records)
    cnt = np.zeros((n, len(F)), np.int64)
    st = np.zeros((n, 17), np.int32)
    for i, (rec, (h, R, C, U)) in enumerate(zip(records, rows)):
        e = rec["expect"]
        cnt[i, F.index("ncols")] = e["ncols"]
        cnt[i, F.index("extended")] = -1
        if not e["assessed"]:
            continue
        for k in F:
            if k in e:
                cnt[i, F.index(k)] = e[k]
        cnt[i, F.index("GCref")] = sum(R.count(x) for x in "gcGC")
        cnt[i, F.index("GCcor")] = sum(C.count(x) for x in "gcGC")
        st[i, 0] = len(e["stretches"])
        for k, (a, b) in enumerate(e["stretches"]):
            st[i, 1 + 2 * k], st[i, 2 + 2 * k] = a, b
    return cnt, st


README_BLOCK = """*********** SUMMARY ***********
Assessed reads:  459
Throughput (uncorrected) 4367089
Throughput (corrected):  4454164
Recall: 0.995006
Precision: 0.9938972
Average correct bases rate (uncorrected):  0.8970857918784844
Error rate (uncorrected): 0.10291420812151564
Average correct bases rate (corrected):  0.9938413
Error rate (corrected): 0.006158699999999961
Number of trimmed/split reads: 37
Mean missing size in trimmed/split reads: 2406.6
Number of over-corrected reads by extention:  4
Mean extension size in over-corrected reads:  40.0
%GC in reference reads:  51.1
%GC in corrected reads:  51.1
Number of corrected reads which length is < 10.0 % of the original read: 25
Number of very low quality corrected reads:  0
Number of insertions in uncorrected:  159839
Number of insertions in corrected:  8119
Number of deletions in uncorrected:  160207
Number of deletions in corrected:  8161
Number of substitutions in uncorrected:  154842
Number of substitutions in corrected:  12148
Ratio of homopolymer sizes in corrected vs reference: 0.9925
"""   # the reference's own pin of this path: README.md:137-161 (written with a Python whose sum() was not yet compensated)


def check_against_golden(res, out_dir, g):
    rep = g["report"]
    assert open(os.path.join(out_dir, "per_read_metrics.txt")).read() == rep["per_read_metrics"]
    assert md5_file(os.path.join(out_dir, "per_read_metrics.txt")) == rep["per_read_metrics_md5"] == "5507a6528f193ac9a87b420d6856964d"
    assert md5_file(os.path.join(out_dir, "read_size_distribution.txt")) == rep["read_size_distribution_md5"]
    assert res["log"] == rep["log"]
    assert res["stdout"] == rep["stdout"]
    assert res["assessed_reads"] == 459 and res["trimmed_or_split"] == 37 and res["extended_reads"] == 4
    assert res["size_distribution_complete"] == 1


def test_report_from_reference_counters_equals_reference_report(example_chain, example_oracle_msa, tmp_path):
    import elector_b200
    g = example_chain["gold"]
    rows = example_oracle_msa["records"]
    assert [h for h, _, _, _ in rows] == [r["header"] for r in g["records"]]
    cnt, st = counters_from_golden(g["records"], rows)
    res = elector_b200.report_write([h[1:] for h, _, _, _ in rows], cnt, st, [r[1] for r in rows], [r[2] for r in rows],
                                    small_reads=g["small_reads"], wrongly_cor_reads=g["wrongly_cor_reads"], size_threshold=0.1, homopolymer_threshold=5,
                                    corrected_fasta=os.path.join(example_chain["work"], "cor.fa"), out_dir=str(tmp_path), compensated_sum=True)
    check_against_golden(res, str(tmp_path), g)   # the golden was written by the reference under Python 3.12
    res = elector_b200.report_write([h[1:] for h, _, _, _ in rows], cnt, st, [r[1] for r in rows], [r[2] for r in rows],
                                    small_reads=g["small_reads"], wrongly_cor_reads=g["wrongly_cor_reads"], size_threshold=0.1, homopolymer_threshold=5,
                                    corrected_fasta=os.path.join(example_chain["work"], "cor.fa"), out_dir=str(tmp_path), compensated_sum=False)
    assert res["stdout"] == "None\n" + README_BLOCK


def test_python_float_text():
    """str(float) / round() as the report prints them, on values where a printf("%g") would differ"""
    import elector_b200
    # a one-record report whose ratios are known: TP/(TP+FN) etc.
    F = elector_b200.TALLY_FIELDS
    cases = [(1, 2), (1, 3), (2, 3), (99999, 100000), (1, 1), (123456, 1234567), (1, 30000), (7, 8)]
    for tp, tot in cases:
        cnt = np.zeros((1, len(F)), np.int64)
        st = np.zeros((1, 17), np.int32)
        for k, v in dict(TP=tp, FN=tot - tp, FP=tot - tp, cor=tp, uncor=tot - tp, uncorCor=tp, uncorUncor=tot - tp, GCref=5, GCcor=5, lenRef=20, lenCor=20, lenUnc=20,
                         ncols=20, assessed=1, extended=-1).items():
            cnt[0, F.index(k)] = v
        res = elector_b200.report_write(["r "], cnt, st, ["a" * 20], ["a" * 20])
        assert ("Average correct bases rate (uncorrected):%s\n" % repr(tp / tot)) in res["log"]
        assert ("Recall (computed only on corrected bases):%s\n" % repr(round(tp / tot, 7))) in res["log"]
        assert ("Error rate (corrected): %s\n" % repr(1 - round(tp / tot, 7))) in res["log"]


def test_report_rejects_records_without_assessed_reads():
    import elector_b200
    F = elector_b200.TALLY_FIELDS
    cnt = np.zeros((1, len(F)), np.int64)
    cnt[0, F.index("ncols")] = 3
    with pytest.raises(elector_b200.ElectorError):
        elector_b200.report_write(["r "], cnt, np.zeros((1, 17), np.int32))
