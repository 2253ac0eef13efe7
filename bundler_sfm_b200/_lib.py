"""ctypes binding of libbsfm_b200.so (the C ABI declared in include/bsfm_b200.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LibraryMissing(RuntimeError):
    pass


def library_path():
    # BSFM_LIB_PATH: development override to A/B-test a differently configured build of the same library
    return os.environ.get("BSFM_LIB_PATH") or os.path.join(_HERE, "libbsfm_b200.so")


def load_library():
    """Load libbsfm_b200.so; fail loudly if it was not built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise LibraryMissing(
            f"{path} not found: build it with `make -C bundler_sfm_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "bundler_sfm_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    c = ctypes
    u8p, i32p, i64p = c.POINTER(c.c_uint8), c.POINTER(c.c_int32), c.POINTER(c.c_int64)
    lib.bsfm_last_error.restype = c.c_char_p
    lib.bsfm_version.restype = c.c_char_p
    lib.bsfm_kernel_launches.restype = c.c_int64
    lib.bsfm_set_device.argtypes = [c.c_int]
    lib.bsfm_match_pair.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_int, c.c_double, c.c_void_p, c.c_int]
    lib.bsfm_match_pair.restype = c.c_int
    lib.bsfm_keydb_create.argtypes = [c.c_void_p, c.c_void_p, c.c_int]
    lib.bsfm_keydb_create.restype = c.c_void_p
    lib.bsfm_keydb_create_dev.argtypes = [c.c_void_p, c.c_void_p, c.c_int]
    lib.bsfm_keydb_create_dev.restype = c.c_void_p
    lib.bsfm_keydb_destroy.argtypes = [c.c_void_p]
    lib.bsfm_keydb_destroy.restype = None
    lib.bsfm_match_num_pairs.argtypes = [c.c_int, c.c_int]
    lib.bsfm_match_num_pairs.restype = c.c_int64
    lib.bsfm_match_run.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_double]
    lib.bsfm_match_run.restype = c.c_int64
    lib.bsfm_match_fetch.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p, c.c_int64]
    lib.bsfm_match_fetch.restype = c.c_int
    lib.bsfm_match_result_dev.argtypes = [c.c_void_p, c.POINTER(c.c_void_p), i64p, c.POINTER(c.c_void_p), i64p]
    lib.bsfm_match_result_dev.restype = c.c_int
    lib.bsfm_match_copy_result_dev.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p]
    lib.bsfm_match_copy_result_dev.restype = c.c_int
    lib.bsfm_match_shard_pairs.argtypes = [c.c_void_p]
    lib.bsfm_match_shard_pairs.restype = c.c_int64
    lib.bsfm_match_last_timing.argtypes = [c.c_void_p, c.POINTER(c.c_float), c.POINTER(c.c_int)]
    lib.bsfm_match_last_timing.restype = c.c_int
    lib.bsfm_match_all_pairs.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_double,
                                         c.c_void_p, c.c_int64, c.c_void_p, c.c_int64]
    lib.bsfm_match_all_pairs.restype = c.c_int64
    lib.bsfm_comm_unique_id.argtypes = [c.c_void_p]
    lib.bsfm_comm_unique_id.restype = c.c_int
    lib.bsfm_comm_create.argtypes = [c.c_void_p, c.c_int, c.c_int]
    lib.bsfm_comm_create.restype = c.c_void_p
    lib.bsfm_comm_destroy.argtypes = [c.c_void_p]
    lib.bsfm_comm_destroy.restype = None
    lib.bsfm_keydb_create_sharded.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int]
    lib.bsfm_keydb_create_sharded.restype = c.c_void_p
    lib.bsfm_match_shard_range.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.POINTER(c.c_int), c.POINTER(c.c_int)]
    lib.bsfm_match_shard_range.restype = c.c_int
    lib.bsfm_match_allgather.argtypes = [c.c_void_p, c.c_void_p]
    lib.bsfm_match_allgather.restype = c.c_int64
    lib.bsfm_match_gathered_fetch.argtypes = [c.c_void_p, i64p, i64p, c.c_void_p, c.c_int64, c.c_void_p, c.c_int64]
    lib.bsfm_match_gathered_fetch.restype = c.c_int
    lib.bsfm_match_all_pairs_multi.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_double, c.c_int, c.c_void_p,
                                               c.c_void_p, c.c_int64, c.c_void_p, c.c_int64]
    lib.bsfm_match_all_pairs_multi.restype = c.c_int64
    lib.bsfm_match_pair_cache_clear.restype = None
    lib.bsfm_measure_int8_peak.argtypes = [c.c_int, c.c_int]
    lib.bsfm_measure_int8_peak.restype = c.c_double
    lib.bsfm_measure_fp64_issue_cycles.argtypes = [c.c_int, c.c_int, c.c_int]
    lib.bsfm_measure_fp64_issue_cycles.restype = c.c_double
    _LIB = lib
    return lib


_NCCL_PRELOADED = False


def preload_nccl():
    """The library binds NCCL at run time with dlopen("libnccl.so.2").  Inside a Python process that also imports torch the
    copy that must answer is the one torch ships (nvidia/nccl/lib): if the system libnccl.so.2 were loaded first under the
    same SONAME, torch's own import would later resolve against it and fail on symbols newer than the system copy.
    So: load torch's copy first when there is one (found without importing torch); BSFM_NCCL_LIB overrides the path."""
    global _NCCL_PRELOADED
    if _NCCL_PRELOADED:
        return
    _NCCL_PRELOADED = True
    import importlib.util
    cands = []
    if os.environ.get("BSFM_NCCL_LIB"):
        cands.append(os.environ["BSFM_NCCL_LIB"])
    for mod in ("nvidia.nccl", "torch"):
        try:
            spec = importlib.util.find_spec(mod)
        except (ImportError, ValueError):
            spec = None
        if spec is None:
            continue
        for base in (list(spec.submodule_search_locations or []) or [os.path.dirname(spec.origin or "")]):
            cands.append(os.path.join(base, "lib", "libnccl.so.2"))
            cands.append(os.path.join(os.path.dirname(base), "nvidia", "nccl", "lib", "libnccl.so.2"))
    for path in cands:
        if path and os.path.exists(path):
            try:
                ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
                return
            except OSError:
                continue


def last_error():
    return load_library().bsfm_last_error().decode()


class BsfmError(RuntimeError):
    pass


def check(rc, what):
    if rc < 0:
        raise BsfmError(f"{what} failed ({rc}): {last_error()}")
    return rc
