"""Replacement of elector/computeStats.py:outputRecallPrecision with the reference's signature and return value.

The reference reads `<outDir>/msa.fa` back and walks it column by column in Python (14.9 s for the 459 reads of its example).  Here the
integer part runs on the device (tally_read_kernel) and the report is written from its counters (elector_report_write); when the
msa.fa is the one elector_b200.alignment.getPOA has just written, its counters are still there and the file is only read for the rows
of split reads and of the last read."""
import os
import sys

from . import alignment
from .report import report_run, report_write


def read_msa(path):
    """msa.fa -> (headers without '>', rows_ref, rows_cor, rows_unc) (six lines per record, Donatello.cpp:72-74)"""
    lines = open(path, "rb").read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    n = len(lines) // 6

    def dec(b):
        return b.decode("latin-1")
    return ([dec(lines[6 * i])[1:] for i in range(n)], [dec(lines[6 * i + 1]) for i in range(n)], [dec(lines[6 * i + 3]) for i in range(n)],
            [dec(lines[6 * i + 5]) for i in range(n)])


def outputRecallPrecision(correctedFileName, outDir, logFile, smallReadNumber, wronglyCorrectedReadsNumber, reportedHomopolThreshold,
                          SIZE_CORRECTED_READ_THRESHOLD, fileSizeName, clipsNb, beg=0, end=0, soft=None, compensated_sum=None):
    """computeStats.py:196-264.  clipsNb must be empty (`-simulator real` clipping is not part of this path).
    compensated_sum: None = like the running Python's sum() (compensated from 3.12 on)."""
    if clipsNb:
        raise NotImplementedError("clipped reads (-simulator real) are outside this path")
    if compensated_sum is None:
        compensated_sum = sys.version_info >= (3, 12)
    msa = outDir + ("/msa_" + soft + ".fa" if soft is not None else "/msa.fa")
    kw = dict(small_reads=smallReadNumber, wrongly_cor_reads=wronglyCorrectedReadsNumber, size_threshold=SIZE_CORRECTED_READ_THRESHOLD,
              homopolymer_threshold=reportedHomopolThreshold, corrected_fasta=correctedFileName, out_dir=outDir, soft=soft, size_file_name=fileSizeName,
              compensated_sum=compensated_sum)
    heads, R, C, U = read_msa(msa)
    last = alignment.LAST
    if last is not None and last["msa_path"] == os.path.abspath(msa) and last["headers"] == heads:
        res = report_write(heads, last["counters"], last["stretches"], R, C, **kw)
    else:
        res = report_run(alignment.context(), heads, R, C, U, **kw)
    print(res["stdout"], end="")
    logFile.write(res["log"])
    return (res["assessed_reads"], res["throughput_corrected"], res["precision"], res["recall"], res["correct_rate_corrected"], 1 - res["correct_rate_corrected"],
            smallReadNumber, wronglyCorrectedReadsNumber, res["gc_ref"], res["gc_cor"], str(res["trimmed_or_split"]), res["mean_missing"], str(res["extended_reads"]),
            res["mean_extension"], SIZE_CORRECTED_READ_THRESHOLD, [res["ins_u"], res["del_u"], res["subs_u"]], [res["ins_c"], res["del_c"], res["subs_c"]],
            res["trimmed_or_split"], res["homopolymer_ratio"])
