"""Host-side mirror of the reference MATCH interface on top of the C ABI.

Reference (file:line under /root/reference):
  MatchKeys(num_keys1, k1, num_keys2, k2, ratio=0.6, max_pts_visit=200)   src/keys2a.h:99-107
  KeyMatchFull main loop + match-table writer                             src/KeyMatchFull.cpp:105-151

The GPU search is exact (== reference with max_pts_visit=0); see SURVEY.md F2.
"""
import ctypes

import numpy as np

from ._lib import load_library, check, preload_nccl

DESC_DIM = 128


def _as_keys(k):
    k = np.ascontiguousarray(k, dtype=np.uint8)
    if k.ndim != 2 or k.shape[1] != DESC_DIM:
        raise ValueError("keys must be [n,128] uint8")
    return k


def match_keys(k1, k2, ratio=0.6, max_pts_visit=200):
    """MatchKeys(k1 queries, k2 database) -> int32 [M,2] (idx1, idx2), ascending idx1.

    `max_pts_visit` is accepted for signature compatibility and ignored: the search is exact."""
    lib = load_library()
    k1, k2 = _as_keys(k1), _as_keys(k2)
    cap = max(int(k1.shape[0]), 1)
    out = np.empty((cap, 2), dtype=np.int32)
    n = check(lib.bsfm_match_pair(k1.ctypes.data, k1.shape[0], k2.ctypes.data, k2.shape[0],
                                  float(ratio), out.ctypes.data, cap), "bsfm_match_pair")
    return out[:n].copy()


def concat_keys(keys_list):
    """list of [n_i,128] uint8 -> (keys [sum n,128], key_off int64 [N+1])"""
    ns = [int(k.shape[0]) for k in keys_list]
    key_off = np.zeros(len(ns) + 1, dtype=np.int64)
    np.cumsum(ns, out=key_off[1:])
    if key_off[-1] == 0:
        return np.zeros((0, DESC_DIM), np.uint8), key_off
    keys = np.ascontiguousarray(np.concatenate([_as_keys(k) for k in keys_list if k.shape[0] > 0], axis=0))
    return keys, key_off


def start_image(i, window_radius):
    """KeyMatchFull.cpp:116-119"""
    return max(i - window_radius, 0) if window_radius > 0 else 0


def pair_list(num_images, window_radius=-1, img_begin=0, img_end=None):
    """(j, i) pairs in KeyMatchFull order for database images i in [img_begin, img_end)."""
    img_end = num_images if img_end is None else img_end
    return [(j, i) for i in range(img_begin, img_end) for j in range(start_image(i, window_radius), i)]


COMM_ID_BYTES = 128


class Comm:
    """One rank of the library's NCCL communicator (bsfm_comm).  `from_torch()` ships rank 0's unique id through the
    torch.distributed process group the caller already has (one process per GPU); `Comm(id, rank, world)` is the raw form."""

    def __init__(self, uid, rank, world):
        preload_nccl()
        self._lib = load_library()
        self.rank, self.world = int(rank), int(world)
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        assert uid.shape == (COMM_ID_BYTES,)
        self._h = self._lib.bsfm_comm_create(uid.ctypes.data, self.rank, self.world)
        if not self._h:
            check(-1, "bsfm_comm_create")

    @staticmethod
    def unique_id():
        preload_nccl()
        uid = np.zeros(COMM_ID_BYTES, np.uint8)
        check(load_library().bsfm_comm_unique_id(uid.ctypes.data), "bsfm_comm_unique_id")
        return uid

    @classmethod
    def from_torch(cls, device=None, group=None):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        t = torch.from_numpy(cls.unique_id() if rank == 0 else np.zeros(COMM_ID_BYTES, np.uint8))
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src=0, group=group)
        return cls(t.cpu().numpy(), rank, world)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bsfm_comm_destroy(self._h)
            self._h = None

    __del__ = close


class KeyDatabase:
    """Device-resident descriptors of all images (uploaded once, KeyMatchFull.cpp:93-99).  With `comm` the database is built
    cooperatively: every rank prepares 1/world of the images and the prepared rows are all-gathered over NVLink."""

    def __init__(self, keys, key_off, device_ptr=None, comm=None):
        self._lib = load_library()
        self.key_off = np.ascontiguousarray(key_off, dtype=np.int64)
        self.num_images = len(self.key_off) - 1
        if comm is not None:
            keys = np.ascontiguousarray(keys, dtype=np.uint8)
            self._h = self._lib.bsfm_keydb_create_sharded(comm._h, keys.ctypes.data, self.key_off.ctypes.data, self.num_images)
        elif device_ptr is not None:
            self._h = self._lib.bsfm_keydb_create_dev(ctypes.c_void_p(device_ptr), self.key_off.ctypes.data, self.num_images)
        else:
            keys = np.ascontiguousarray(keys, dtype=np.uint8)
            self._h = self._lib.bsfm_keydb_create(keys.ctypes.data, self.key_off.ctypes.data, self.num_images)
        if not self._h:
            check(-1, "bsfm_keydb_create")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bsfm_keydb_destroy(self._h)
            self._h = None

    __del__ = close

    def run(self, img_begin=0, img_end=None, window_radius=-1, ratio=0.6):
        img_end = self.num_images if img_end is None else img_end
        return check(self._lib.bsfm_match_run(self._h, img_begin, img_end, window_radius, float(ratio)), "bsfm_match_run")

    def fetch(self):
        npairs = self._lib.bsfm_match_shard_pairs(self._h)
        ptr_c, ptr_m = ctypes.c_void_p(), ctypes.c_void_p()
        n_p, n_m = ctypes.c_int64(), ctypes.c_int64()
        check(self._lib.bsfm_match_result_dev(self._h, ctypes.byref(ptr_c), ctypes.byref(n_p), ctypes.byref(ptr_m), ctypes.byref(n_m)),
              "bsfm_match_result_dev")
        counts = np.zeros(max(npairs, 1), dtype=np.int32)
        matches = np.zeros((max(n_m.value, 1), 2), dtype=np.int32)
        check(self._lib.bsfm_match_fetch(self._h, counts.ctypes.data, counts.shape[0], matches.ctypes.data, matches.shape[0]),
              "bsfm_match_fetch")
        return counts[:npairs], matches[:n_m.value]

    def allgather(self, comm):
        """NCCL all-gather of the ranks' tables inside the library (bsfm_match_allgather) -> total matches"""
        return check(self._lib.bsfm_match_allgather(comm._h, self._h), "bsfm_match_allgather")

    def gathered_fetch(self):
        """(pair_counts, matches) of the gathered table on the host"""
        n_p, n_m = ctypes.c_int64(), ctypes.c_int64()
        check(self._lib.bsfm_match_gathered_fetch(self._h, ctypes.byref(n_p), ctypes.byref(n_m), None, 0, None, 0), "bsfm_match_gathered_fetch")
        counts = np.zeros(max(n_p.value, 1), dtype=np.int32)
        matches = np.zeros((max(n_m.value, 1), 2), dtype=np.int32)
        check(self._lib.bsfm_match_gathered_fetch(self._h, None, None, counts.ctypes.data, counts.shape[0], matches.ctypes.data, matches.shape[0]),
              "bsfm_match_gathered_fetch")
        return counts[:n_p.value], matches[:n_m.value]

    def result_dev(self):
        """(pair_counts_ptr, num_pairs, matches_ptr, num_matches) -- device pointers for NCCL."""
        ptr_c, ptr_m = ctypes.c_void_p(), ctypes.c_void_p()
        n_p, n_m = ctypes.c_int64(), ctypes.c_int64()
        check(self._lib.bsfm_match_result_dev(self._h, ctypes.byref(ptr_c), ctypes.byref(n_p), ctypes.byref(ptr_m), ctypes.byref(n_m)),
              "bsfm_match_result_dev")
        return ptr_c.value, n_p.value, ptr_m.value, n_m.value

    def result_to_torch(self, device):
        """(pair_counts, matches) of the last run as torch int32 tensors on `device` (device-to-device copy)"""
        import torch
        _, n_p, _, n_m = self.result_dev()
        counts = torch.empty(max(n_p, 1), dtype=torch.int32, device=device)
        matches = torch.empty((max(n_m, 1), 2), dtype=torch.int32, device=device)
        check(self._lib.bsfm_match_copy_result_dev(self._h, counts.data_ptr(), matches.data_ptr()), "bsfm_match_copy_result_dev")
        return counts[:n_p], matches[:n_m]

    def timing(self):
        ms = (ctypes.c_float * 3)()
        launches = ctypes.c_int()
        check(self._lib.bsfm_match_last_timing(self._h, ms, ctypes.byref(launches)), "bsfm_match_last_timing")
        return {"search_ms": ms[0], "post_ms": ms[1], "total_ms": ms[2], "launches": launches.value}


def key_match_full(keys_list, window_radius=-1, ratio=0.6):
    """KeyMatchFull's pair loop over host descriptors -> (pairs [(j,i)], counts, matches)."""
    keys, key_off = concat_keys(keys_list)
    db = KeyDatabase(keys, key_off)
    try:
        db.run(0, len(keys_list), window_radius, ratio)
        counts, matches = db.fetch()
    finally:
        db.close()
    return pair_list(len(keys_list), window_radius), counts, matches


def shard_range(key_off, window_radius, world_size, rank):
    """the library's own split (bsfm_match_shard_range; same rule as shard_images)"""
    key_off = np.ascontiguousarray(key_off, dtype=np.int64)
    b, e = ctypes.c_int(), ctypes.c_int()
    check(load_library().bsfm_match_shard_range(key_off.ctypes.data, len(key_off) - 1, window_radius, world_size, rank, ctypes.byref(b), ctypes.byref(e)),
          "bsfm_match_shard_range")
    return b.value, e.value


def key_match_full_multi(keys_list, window_radius=-1, ratio=0.6, ngpus=2, devices=None):
    """bsfm_match_all_pairs_multi: the whole pair loop on `ngpus` devices of this process (one host thread per GPU,
    NCCL all-gather of prepared descriptors and of the match table) -> (pairs, counts, matches)"""
    preload_nccl()
    lib = load_library()
    keys, key_off = concat_keys(keys_list)
    npairs = lib.bsfm_match_num_pairs(len(keys_list), window_radius)
    counts = np.zeros(max(npairs, 1), np.int32)
    cap = int(sum(k.shape[0] for k in keys_list)) * 4 + 1024
    dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
    while True:
        matches = np.zeros((cap, 2), np.int32)
        total = lib.bsfm_match_all_pairs_multi(keys.ctypes.data, key_off.ctypes.data, len(keys_list), window_radius, float(ratio), ngpus,
                                               None if dev is None else dev.ctypes.data, counts.ctypes.data, counts.shape[0], matches.ctypes.data, cap)
        if total == -4 and cap < (1 << 30):      # BSFM_ERR_CAPACITY: grow and retry
            cap *= 4
            continue
        check(total, "bsfm_match_all_pairs_multi")
        return pair_list(len(keys_list), window_radius), counts[:npairs], matches[:total]


def format_match_table(pairs, counts, matches, min_matches=16):
    """The text KeyMatchFull writes (KeyMatchFull.cpp:131-142): only pairs with >= 16 matches."""
    out = []
    pos = 0
    for (j, i), c in zip(pairs, counts):
        c = int(c)
        if c >= min_matches:
            out.append(f"{j} {i}\n{c}\n")
            out.append("".join(f"{a} {b}\n" for a, b in matches[pos:pos + c]))
        pos += c
    return "".join(out)


# ------------------------------------------------------------------------------------------------
# multi-GPU: shard database images (contiguous ranges of the pair list) across ranks
# ------------------------------------------------------------------------------------------------
def shard_images(num_keys, window_radius, world_size):
    """Contiguous database-image ranges [b_r, e_r) with (nearly) equal work, work(i) = n_i * sum_{j in window} n_j.

    Contiguous ranges keep every rank's output a contiguous slice of the KeyMatchFull pair order, so the
    gathered table is the concatenation of the ranks' tables."""
    n = np.asarray(num_keys, dtype=np.float64)
    N = len(n)
    csum = np.concatenate([[0.0], np.cumsum(n)])
    work = np.array([n[i] * (csum[i] - csum[start_image(i, window_radius)]) for i in range(N)])
    total = work.sum()
    bounds = [0]
    acc = 0.0
    r = 1
    for i in range(N):
        acc += work[i]
        while r < world_size and acc >= total * r / world_size:
            bounds.append(i + 1)
            r += 1
    while len(bounds) < world_size:
        bounds.append(N)
    bounds.append(N)
    bounds = [min(b, N) for b in bounds]
    for k in range(1, len(bounds)):
        bounds[k] = max(bounds[k], bounds[k - 1])
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def gather_match_table(counts, matches, group=None, device=None):
    """all_gather of variable-length (pair_counts, matches) shards via torch.distributed.

    Two collectives: sizes, then payload padded to the max shard.  Works on NCCL
    (device tensors over NVLink) and gloo (CPU tensors; used by the world_size-2 CPU tests).
    `counts` / `matches` may be numpy arrays or torch tensors on `device`."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if not torch.is_tensor(counts):
        counts = torch.from_numpy(np.ascontiguousarray(counts, dtype=np.int32))
        matches = torch.from_numpy(np.ascontiguousarray(matches, dtype=np.int32).reshape(-1, 2))
    if device is not None:
        counts, matches = counts.to(device), matches.to(device)
    dev = counts.device
    sizes = torch.tensor([counts.numel(), matches.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    max_c, max_m = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    payload = torch.zeros(max_c + 2 * max_m, dtype=torch.int32, device=dev)
    payload[:counts.numel()] = counts
    payload[max_c:max_c + 2 * matches.shape[0]] = matches.reshape(-1)
    gathered = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    out_c, out_m = [], []
    for r in range(world):
        nc, nm = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        out_c.append(gathered[r][:nc])
        out_m.append(gathered[r][max_c:max_c + 2 * nm].reshape(-1, 2))
    return torch.cat(out_c), torch.cat(out_m)
