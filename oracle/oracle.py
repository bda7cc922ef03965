"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU oracle (oracle/libpoa_oracle.so).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (elector_b200/) never imports this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OraMatrix(ctypes.Structure):
    _fields_ = [("nsymbol", ctypes.c_int), ("symbol", ctypes.c_char * 129),
                ("score", (ctypes.c_int * 128) * 128), ("gap_set", (ctypes.c_int * 3) * 2),
                ("trunc_gap_length", ctypes.c_int), ("decay_gap_length", ctypes.c_int),
                ("max_gap_length", ctypes.c_int), ("gap_penalty_x", ctypes.c_int * 64),
                ("gap_penalty_y", ctypes.c_int * 64)]


def build(force=False):
    so = os.path.join(_HERE, "libpoa_oracle.so")
    src = os.path.join(_HERE, "poa_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libpoa_oracle.so", "poa_oracle_cli"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp = ctypes.c_void_p
        L.ora_read_matrix.argtypes = [ctypes.c_char_p, ctypes.POINTER(OraMatrix)]
        L.ora_read_matrix.restype = ctypes.c_int
        L.ora_default_matrix.argtypes = [ctypes.POINTER(OraMatrix)]
        L.ora_batch.argtypes = [ctypes.POINTER(OraMatrix), ctypes.c_int] + [vp] * 12 + [ctypes.c_int]
        L.ora_batch.restype = ctypes.c_int
        L.ora_poa_files.argtypes = [ctypes.c_char_p] * 5 + [ctypes.c_int]
        L.ora_poa_files.restype = ctypes.c_int
        L.ora_dump_files.argtypes = [ctypes.c_char_p] * 5
        L.ora_dump_files.restype = ctypes.c_int
        _LIB = L
    return _LIB


def matrix(path=None):
    m = OraMatrix()
    if path is None:
        lib().ora_default_matrix(ctypes.byref(m))
    elif lib().ora_read_matrix(path.encode(), ctypes.byref(m)) <= 0:
        raise RuntimeError("oracle: cannot read matrix %s" % path)
    return m


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def batch(ref, ref_off, cor, cor_off, unc, unc_off, m=None, nthreads=1):
    """Runs the oracle on CSR windows.  Returns dict(rows, row_off, nring, score1, score2, cells);
    window w's three rows are rows[row_off[w] + s*nring[w] : ... + nring[w]], s = 0, 1, 2."""
    m = m or matrix()
    n = len(ref_off) - 1
    ref_off, cor_off, unc_off = (np.ascontiguousarray(o, dtype=np.int64) for o in (ref_off, cor_off, unc_off))
    ref, cor, unc = (np.ascontiguousarray(s, dtype=np.uint8) for s in (ref, cor, unc))
    cap = 3 * int(ref_off[-1] + cor_off[-1] + unc_off[-1]) + 16
    out = dict(rows=np.zeros(cap, np.uint8), row_off=np.zeros(n, np.int64), nring=np.zeros(n, np.int32),
               score1=np.zeros(n, np.int32), score2=np.zeros(n, np.int32), cells=np.zeros(n, np.int64))
    lib().ora_batch(ctypes.byref(m), n, _p(ref), _p(ref_off), _p(cor), _p(cor_off), _p(unc), _p(unc_off),
                    _p(out["rows"]), _p(out["row_off"]), _p(out["nring"]), _p(out["score1"]), _p(out["score2"]),
                    _p(out["cells"]), int(nthreads))
    return out


def window_rows(out, w):
    o, k = int(out["row_off"][w]), int(out["nring"][w])
    return tuple(out["rows"][o + s * k:o + (s + 1) * k].tobytes().decode("latin-1") for s in range(3))


def poa_files(matrix_path, ref_fa, cor_fa, unc_fa, pir_out):
    return lib().ora_poa_files(matrix_path.encode(), ref_fa.encode(), cor_fa.encode(), unc_fa.encode(),
                               pir_out.encode(), 0)


def dump_files(matrix_path, ref_fa, cor_fa, unc_fa, dump_out):
    return lib().ora_dump_files(matrix_path.encode(), ref_fa.encode(), cor_fa.encode(), unc_fa.encode(),
                                dump_out.encode())
