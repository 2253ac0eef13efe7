"""oracle/loader.py -- TEST INFRASTRUCTURE ONLY.

ctypes bindings of (a) oracle/liboracle.so, our CPU restatement of the two hot paths, and
(b) oracle/_ref/*.so, the UNMODIFIED reference compiled from /root/reference by oracle/Makefile.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product (bundler_sfm_b200/, libbsfm_b200.so) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OB = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"


def build(verbose=False):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    r = subprocess.run(["make", "-C", _HERE, "-j8", "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


def _load(path):
    if not os.path.exists(path):
        return None
    return ctypes.CDLL(path)


_port = None
_ref_match = None
_ref_sba = None


def port():
    global _port
    if _port is None:
        p = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(p):
            build()
        _port = ctypes.CDLL(p)
        c = ctypes
        _port.oracle_match_pair.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_int, c.c_double, c.c_void_p, c.c_int]
        _port.oracle_match_pair.restype = c.c_int
        _port.oracle_match_pair_test.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_int, c.c_double, c.c_int, c.c_void_p, c.c_int]
        _port.oracle_match_pair_test.restype = c.c_int
        _port.oracle_top2_all.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p]
        _port.oracle_top2_all.restype = None
        _port.oracle_match_all_pairs.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_double, c.c_int,
                                                 c.c_void_p, c.c_long, c.c_void_p]
        _port.oracle_match_all_pairs.restype = c.c_long
    return _port


def ref_match():
    """reference ANN+keys2a library, or None when oracle/_ref was not built"""
    global _ref_match
    if _ref_match is None:
        lib = _load(os.path.join(_HERE, "_ref", "libref_match.so"))
        if lib is None:
            return None
        c = ctypes
        lib.ref_match_pair.argtypes = [c.c_int, c.c_void_p, c.c_int, c.c_void_p, c.c_double, c.c_int, c.c_void_p, c.c_int]
        lib.ref_match_pair.restype = c.c_int
        lib.ref_create_tree.argtypes = [c.c_int, c.c_void_p]
        lib.ref_create_tree.restype = c.c_void_p
        lib.ref_delete_tree.argtypes = [c.c_void_p]
        lib.ref_match_tree.argtypes = [c.c_int, c.c_void_p, c.c_void_p, c.c_double, c.c_int, c.c_void_p, c.c_int]
        lib.ref_match_tree.restype = c.c_int
        _ref_match = lib
    return _ref_match


_ref_keys = None


def ref_keys():
    """reference in-bundler matcher (src/keys.cpp + ann_1.1_char), or None when oracle/_ref was not built"""
    global _ref_keys
    if _ref_keys is None:
        lib = _load(os.path.join(_HERE, "_ref", "libref_keys.so"))
        if lib is None:
            return None
        c = ctypes
        lib.ref_keys_match.argtypes = [c.c_int, c.c_void_p, c.c_int, c.c_void_p, c.c_void_p, c.c_int, c.c_double, c.c_int, c.c_void_p, c.c_int]
        lib.ref_keys_match.restype = c.c_int
        _ref_keys = lib
    return _ref_keys


def keys_match_ref(k1, k2, extra2=None, registered=False, ratio=0.6, exhaustive=True, fn=None):
    """MatchKeysExhaustive / MatchKeys of src/keys.cpp through the doorway (or the same-shaped `fn` of the shim)"""
    if fn is None:
        lib = ref_keys()
        assert lib is not None, "oracle/_ref/libref_keys.so not built"
        fn = lib.ref_keys_match
    k1, k2 = _u8(k1), _u8(k2)
    ex = None if extra2 is None else np.ascontiguousarray(extra2, dtype=np.int32)
    out = np.empty((max(k1.shape[0], 1), 2), np.int32)
    import sys
    sys.stdout.flush()
    saved = os.dup(1); devnull = os.open(os.devnull, os.O_WRONLY); os.dup2(devnull, 1); os.close(devnull)   # "[MatchKeys] Found ..." chatter
    try:
        n = fn(k1.shape[0], k1.ctypes.data, k2.shape[0], k2.ctypes.data, None if ex is None else ex.ctypes.data, int(bool(registered)),
               float(ratio), int(bool(exhaustive)), out.ctypes.data, out.shape[0])
    finally:
        os.dup2(saved, 1); os.close(saved)
    return out[:n].copy()


def match_pair_port_test(k1, k2, ratio, mode, extra2=None, registered=False):
    """the C restatement with the selectable acceptance test (mode 1 = keys.cpp) and the `registered` subset"""
    k1, k2 = _u8(k1), _u8(k2)
    sel = np.arange(k2.shape[0])
    if registered:
        sel = np.nonzero(np.asarray(extra2) >= 0)[0]
    db = np.ascontiguousarray(k2[sel])
    out = np.empty((max(k1.shape[0], 1), 2), np.int32)
    n = port().oracle_match_pair_test(k1.ctypes.data, k1.shape[0], db.ctypes.data, db.shape[0], float(ratio), int(mode), out.ctypes.data, out.shape[0])
    res = out[:n].copy()
    res[:, 1] = sel[res[:, 1]] if n else res[:, 1]
    return res


def ref_sba():
    """reference sba-1.5 + sfm-driver library (needs OpenBLAS from the venv), or None"""
    global _ref_sba
    if _ref_sba is None:
        os.environ.setdefault("OPENBLAS_CORETYPE", "HASWELL")
        if not os.path.exists(os.path.join(_HERE, "_ref", "libref_sba.so")):
            return None
        # OpenBLAS 0.3.15 from the venv's opencv wheel needs its sibling libquadmath/libgfortran
        for dep in ("libquadmath-2284e583.so.0.0.0", "libgfortran-83c28eba.so.5.0.0", "libopenblasp-r0-59ffcd50.3.15.so"):
            dp = os.path.join(_OB, dep)
            if os.path.exists(dp):
                ctypes.CDLL(dp, mode=ctypes.RTLD_GLOBAL)
        lib = _load(os.path.join(_HERE, "_ref", "libref_sba.so"))
        if lib is None:
            return None
        _ref_sba = lib
    return _ref_sba


# ---- MATCH helpers --------------------------------------------------------------------------
def _u8(k):
    k = np.ascontiguousarray(k, dtype=np.uint8)
    assert k.ndim == 2 and k.shape[1] == 128
    return k


def match_pair_port(k1, k2, ratio=0.6):
    k1, k2 = _u8(k1), _u8(k2)
    out = np.empty((max(k1.shape[0], 1), 2), np.int32)
    n = port().oracle_match_pair(k1.ctypes.data, k1.shape[0], k2.ctypes.data, k2.shape[0], float(ratio), out.ctypes.data, out.shape[0])
    return out[:n].copy()


def top2_port(k1, k2):
    k1, k2 = _u8(k1), _u8(k2)
    d0 = np.empty(k1.shape[0], np.int32); d1 = np.empty_like(d0); i0 = np.empty_like(d0)
    port().oracle_top2_all(k1.ctypes.data, k1.shape[0], k2.ctypes.data, k2.shape[0], d0.ctypes.data, d1.ctypes.data, i0.ctypes.data)
    return d0, d1, i0


def match_pair_ref(k1, k2, ratio=0.6, max_pts_visit=0):
    lib = ref_match()
    assert lib is not None, "oracle/_ref/libref_match.so not built"
    k1, k2 = _u8(k1), _u8(k2)
    out = np.empty((max(k1.shape[0], 1), 2), np.int32)
    n = lib.ref_match_pair(k1.shape[0], k1.ctypes.data, k2.shape[0], k2.ctypes.data, float(ratio), int(max_pts_visit), out.ctypes.data, out.shape[0])
    return out[:n].copy()


def match_all_pairs_port(keys_list, window_radius=-1, ratio=0.6, min_matches=16):
    """-> (table_text, pair_counts[N,N])"""
    ns = [k.shape[0] for k in keys_list]
    N = len(ns)
    key_off = np.zeros(N + 1, np.int64); np.cumsum(ns, out=key_off[1:])
    keys = np.ascontiguousarray(np.concatenate([_u8(k) for k in keys_list], 0)) if key_off[-1] else np.zeros((0, 128), np.uint8)
    counts = np.zeros((N, N), np.int32)
    need = port().oracle_match_all_pairs(keys.ctypes.data, key_off.ctypes.data, N, window_radius, float(ratio), min_matches, None, 0, counts.ctypes.data)
    buf = ctypes.create_string_buffer(max(need, 1))
    port().oracle_match_all_pairs(keys.ctypes.data, key_off.ctypes.data, N, window_radius, float(ratio), min_matches, buf, need, None)
    return buf.raw[:need].decode(), counts


# ---- BA helpers -----------------------------------------------------------------------------
def run_sfm_ref(scene, quiet=True, **kw):
    """run the UNMODIFIED reference run_sfm (oracle/_ref/libref_sba.so) on a scene dict"""
    from bundler_sfm_b200 import bundle
    lib = ref_sba()
    assert lib is not None, "oracle/_ref/libref_sba.so not built"
    fn = lib.run_sfm
    bundle._bind_run_sfm(fn)
    fn.restype = None
    if quiet:
        os.environ["REF_SBA_VERBOSE"] = "0"
    saved = None
    if quiet:
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        os.close(devnull)
    try:
        out = bundle.call_run_sfm(fn, scene, **kw)
    finally:
        if saved is not None:
            os.dup2(saved, 1)
            os.close(saved)
    info = np.zeros(10)
    ret = ctypes.c_int()
    lib.ref_last_info.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.ref_last_info(info.ctypes.data, ctypes.byref(ret))
    out["info"] = info
    out["rc"] = ret.value
    return out


def run_sfm_oracle(scene, **kw):
    """reference build when present, else the C restatement (oracle/ba_oracle.c)"""
    if ref_sba() is not None:
        return run_sfm_ref(scene, **kw)
    return run_sfm_port(scene, **kw)


def run_sfm_port(scene, quiet=True, **kw):
    """C restatement oracle/ba_oracle.c (oracle_run_sfm has run_sfm's signature + info[10])"""
    from bundler_sfm_b200 import bundle
    fn = port().oracle_run_sfm
    bundle._bind_run_sfm(fn)
    fn.argtypes = fn.argtypes + [ctypes.c_void_p]
    fn.restype = ctypes.c_int
    return bundle.call_run_sfm(fn, scene, extra_info=True, **kw)
