"""Host-side mirror of the file preparation in front of the splitter (elector/readAndSortFiles.py:150-191) over the C-ABI."""
import ctypes

from .lib import load_library
from .poa import ElectorError


def sort_fasta(in_path, out_path):
    """readAndSortFasta (:150-166): records sorted by header line; returns the number of records."""
    lib = load_library()
    n = ctypes.c_int64()
    rc = lib.elector_sort_fasta(in_path.encode(), out_path.encode(), ctypes.cast(ctypes.byref(n), ctypes.c_void_p))
    if rc:
        raise ElectorError(rc, lib.elector_last_error(None).decode())
    return n.value


def duplicate_reads(sorted_ref, sorted_unc, sorted_cor, new_ref, new_unc):
    """duplicateRefReads (:171-191): one `_<k>` copy of the reference / uncorrected read per corrected read; returns the triplets."""
    lib = load_library()
    n = ctypes.c_int64()
    rc = lib.elector_duplicate_reads(sorted_ref.encode(), sorted_unc.encode(), sorted_cor.encode(), new_ref.encode(), new_unc.encode(),
                                     ctypes.cast(ctypes.byref(n), ctypes.c_void_p))
    if rc:
        raise ElectorError(rc, lib.elector_last_error(None).decode())
    return n.value
