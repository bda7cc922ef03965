"""Host-side mirror of the report boundary (include/elector_poa.h, "report"): what a caller of
computeStats.outputRecallPrecision (computeStats.py:196-264) gets -- per_read_metrics.txt, the read size distribution, the
summary text and numbers -- from the per-record counters of the device tally instead of a Python pass over msa.fa."""
import ctypes

import numpy as np

from .lib import load_library
from .poa import ElectorError, _p, windows_to_csr

STRETCH_K = 17


class ReportSummaryC(ctypes.Structure):   # struct elector_report_summary
    _fields_ = [("assessed_reads", ctypes.c_int64), ("throughput_uncorrected", ctypes.c_int64), ("throughput_corrected", ctypes.c_int64),
                ("recall", ctypes.c_double), ("precision", ctypes.c_double), ("correct_rate_uncorrected", ctypes.c_double),
                ("correct_rate_corrected", ctypes.c_double), ("error_rate", ctypes.c_double),
                ("trimmed_or_split", ctypes.c_int64), ("split_reads", ctypes.c_int64), ("trimmed_reads", ctypes.c_int64),
                ("mean_missing", ctypes.c_double), ("extended_reads", ctypes.c_int64), ("mean_extension", ctypes.c_double),
                ("gc_ref", ctypes.c_double), ("gc_cor", ctypes.c_double), ("small_reads", ctypes.c_int64), ("wrongly_cor_reads", ctypes.c_int64),
                ("ins_u", ctypes.c_int64), ("del_u", ctypes.c_int64), ("subs_u", ctypes.c_int64),
                ("ins_c", ctypes.c_int64), ("del_c", ctypes.c_int64), ("subs_c", ctypes.c_int64),
                ("homopolymer_ratio", ctypes.c_double), ("size_distribution_complete", ctypes.c_int32)]


def _headers(headers):
    blob = "".join(headers).encode("latin-1")
    off = np.zeros(len(headers) + 1, np.int64)
    off[1:] = np.cumsum([len(h) for h in headers])
    return np.frombuffer(blob, np.uint8).copy() if blob else np.zeros(1, np.uint8), off


def _tail(small_reads, wrongly_cor_reads, size_threshold, homopolymer_threshold, compensated_sum, corrected_fasta, out_dir, soft, size_file_name):
    enc = lambda s: None if s is None else s.encode()
    summary = ReportSummaryC()
    log, out = ctypes.create_string_buffer(8192), ctypes.create_string_buffer(8192)
    args = [int(small_reads), int(wrongly_cor_reads), float(size_threshold), int(homopolymer_threshold), int(bool(compensated_sum)), enc(corrected_fasta), enc(out_dir), enc(soft),
            enc(size_file_name), ctypes.addressof(summary), ctypes.addressof(log), 8192, ctypes.addressof(out), 8192]
    return summary, log, out, args


def _result(summary, log, out):
    d = {k: getattr(summary, k) for k, _ in ReportSummaryC._fields_}
    d["log"] = log.value.decode()
    d["stdout"] = out.value.decode()
    return d


def report_write(headers, counters, stretches, rows_ref=None, rows_cor=None, small_reads=0, wrongly_cor_reads=0, size_threshold=0.1,
                 homopolymer_threshold=5, corrected_fasta=None, out_dir=None, soft=None, size_file_name="read_size_distribution.txt", compensated_sum=False):
    """elector_report_write: the report from counters (host only).  headers without '>'; rows: lists of strings or None."""
    lib = load_library()
    n = len(headers)
    hb, hoff = _headers(headers)
    counters = np.ascontiguousarray(counters, np.int64)
    stretches = np.ascontiguousarray(stretches, np.int32)
    assert counters.shape[0] == n and stretches.shape == (n, STRETCH_K)
    r = c = off = None
    if rows_ref is not None:
        r, off = windows_to_csr(rows_ref)
    if rows_cor is not None:
        c, off = windows_to_csr(rows_cor)
    summary, log, out, tail = _tail(small_reads, wrongly_cor_reads, size_threshold, homopolymer_threshold, compensated_sum, corrected_fasta, out_dir, soft, size_file_name)
    rc = lib.elector_report_write(n, _p(hb), _p(hoff), _p(counters), _p(stretches), None if r is None else _p(r), None if c is None else _p(c),
                                  None if off is None else _p(off), *tail)
    if rc:
        raise ElectorError(rc, lib.elector_last_error(None).decode())
    return _result(summary, log, out)


def report_run(ctx, headers, rows_ref, rows_cor, rows_unc, small_reads=0, wrongly_cor_reads=0, size_threshold=0.1, homopolymer_threshold=5,
               corrected_fasta=None, out_dir=None, soft=None, size_file_name="read_size_distribution.txt", compensated_sum=False):
    """elector_report_run: merged rows in (the records of msa.fa), tally on the device, report out."""
    lib = load_library()
    n = len(headers)
    hb, hoff = _headers(headers)
    r, off = windows_to_csr(rows_ref)
    c, _ = windows_to_csr(rows_cor)
    u, _ = windows_to_csr(rows_unc)
    summary, log, out, tail = _tail(small_reads, wrongly_cor_reads, size_threshold, homopolymer_threshold, compensated_sum, corrected_fasta, out_dir, soft, size_file_name)
    ctx._check(lib.elector_report_run(ctx._ctx, n, _p(hb), _p(hoff), _p(r), _p(c), _p(u), _p(off), *tail))
    return _result(summary, log, out)
