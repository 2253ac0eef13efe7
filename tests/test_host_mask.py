"""Host-side compression of the dense visibility mask (csrc/common.cpp: host_scan_vmask; sba_levmar.c:652-663 builds the same CRS):
pure host code behind a dev hook of the C-ABI library, so it runs without a GPU."""
import ctypes
import os

import numpy as np
import pytest

from bundler_sfm_b200 import _lib


def _scan(vmask, threads):
    lib = _lib.load_library()
    fn = lib.bsfm_debug_scan_vmask
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    fn.restype = ctypes.c_int
    n, m = vmask.shape
    rowptr = np.zeros(n + 1, dtype=np.int32)
    cap = int(np.count_nonzero(vmask)) + 8
    obs_cam = np.full(cap, -1, dtype=np.int32)
    os.environ["BSFM_BA_MASK_THREADS"] = str(threads)
    nvis = fn(vmask.ctypes.data, n, m, rowptr.ctypes.data, obs_cam.ctypes.data, cap)
    return nvis, rowptr, obs_cam[:max(nvis, 0)]


@pytest.mark.parametrize("n,m,density,threads", [(1, 1, 1.0, 1), (7, 5, 0.5, 3), (1000, 63, 0.05, 4), (513, 200, 0.01, 16),
                                                   (64, 1000, 0.006, 2), (3, 129, 0.0, 2), (40, 64, 1.0, 5)])
def test_scan_equals_numpy_nonzero(n, m, density, threads):
    rng = np.random.default_rng(n * 1000 + m)
    vmask = (rng.random((n, m)) < density).astype(np.int8)
    if density > 0:
        vmask[rng.integers(0, n), rng.integers(0, m)] = 7          # any non-zero byte counts (the reference tests `!= 0`)
    nvis, rowptr, obs_cam = _scan(vmask, threads)
    pts, cams = np.nonzero(vmask)                                  # row-major: ascending point, ascending camera within a point
    assert nvis == len(cams)
    assert np.array_equal(obs_cam, cams.astype(np.int32))
    assert np.array_equal(rowptr, np.concatenate([[0], np.cumsum(np.count_nonzero(vmask, axis=1))]).astype(np.int32))


def test_scan_reports_short_output_buffer():
    lib = _lib.load_library()
    fn = lib.bsfm_debug_scan_vmask
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    fn.restype = ctypes.c_int
    vmask = np.ones((4, 4), dtype=np.int8)
    rowptr = np.zeros(5, dtype=np.int32)
    obs = np.zeros(3, dtype=np.int32)
    assert fn(vmask.ctypes.data, 4, 4, rowptr.ctypes.data, obs.ctypes.data, 3) == -1
