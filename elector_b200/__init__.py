"""elector_b200 -- B200-native (sm_100a) implementation of ELECTOR's POA hot path.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C-ABI of
include/elector_poa.h + the `poa` drop-in executable) and this thin host-side mirror
(`poa.PoaContext`, ctypes over the C-ABI).  There is no CPU implementation: importing
works anywhere, but creating a context without the built library or without a CUDA
device raises.
"""
from .lib import LibraryNotBuilt, load_library, library_path  # noqa: F401
from .poa import ElectorError, PackedLetters, PoaContext, PoaResult, TALLY_FIELDS, pack_letters, windows_to_csr  # noqa: F401
from .matrix import write_default_matrix  # noqa: F401
from .report import report_run, report_write  # noqa: F401

__all__ = ["PoaContext", "PoaResult", "PackedLetters", "pack_letters", "ElectorError", "LibraryNotBuilt", "load_library", "library_path",
           "windows_to_csr", "write_default_matrix", "TALLY_FIELDS", "report_run", "report_write"]
