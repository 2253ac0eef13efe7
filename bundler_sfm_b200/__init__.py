"""bundler_sfm_b200 -- B200-native (sm_100a) drop-in for the two data-parallel hot paths of
snavely/bundler_sfm: the all-pairs SIFT match (KeyMatchFull / MatchKeys) and the sparse
Levenberg-Marquardt bundle adjustment (run_sfm / sba_motstr_levmar_x).

The product is the C-ABI shared library ``libbsfm_b200.so`` (include/bsfm_b200.h); this package is
the thin host-side mirror of the reference interfaces on top of it (ctypes, numpy).  There is no
CPU fallback: importing works without a GPU, computing does not.
"""
from ._lib import load_library, library_path, LibraryMissing  # noqa: F401

__all__ = ["load_library", "library_path", "LibraryMissing"]
__version__ = "0.1"
