"""CPU tests (no GPU) of the multi-GPU host logic: image-range sharding of the KeyMatchFull pair list
and the variable-length all-gather of the match table, world_size 2 over gloo.  The per-rank compute
is supplied by the CPU oracle here (test infrastructure) -- on the GPU box the same code path runs
with libbsfm_b200.so and NCCL (tests/test_match_gpu.py::test_sharded_runs_concatenate, bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bundler_sfm_b200 import keymatch, synth


def test_shard_images_balanced_and_contiguous():
    sizes = [5000] * 500
    for world in (1, 2, 4, 8):
        sh = keymatch.shard_images(sizes, -1, world)
        assert sh[0][0] == 0 and sh[-1][1] == 500
        assert all(sh[r][1] == sh[r + 1][0] for r in range(world - 1))
        work = [sum(i for i in range(b, e)) for b, e in sh]
        assert max(work) / (sum(work) / world) < 1.02
    # windowed + ragged + empty images
    sizes = [100, 0, 300, 50, 1, 700, 20]
    sh = keymatch.shard_images(sizes, 2, 3)
    assert sh[0][0] == 0 and sh[-1][1] == len(sizes)
    assert [p for b, e in sh for p in keymatch.pair_list(len(sizes), 2, b, e)] == keymatch.pair_list(len(sizes), 2)


def _oracle_shard(imgs, b, e, window):
    from oracle import loader
    counts, matches = [], []
    for (j, i) in keymatch.pair_list(len(imgs), window, b, e):
        if imgs[j].shape[0] == 0 or imgs[i].shape[0] == 0:
            m = np.zeros((0, 2), np.int32)
        else:
            m = loader.match_pair_port(imgs[j], imgs[i], 0.6)
        counts.append(m.shape[0]); matches.append(m)
    return np.array(counts, np.int32), (np.concatenate(matches, 0) if matches else np.zeros((0, 2), np.int32))


def _worker(rank, world, port, sizes, window, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    imgs = synth.sift_like_descriptors(len(sizes), sizes, seed=13)
    b, e = keymatch.shard_images(sizes, window, world)[rank]
    c, m = _oracle_shard(imgs, b, e, window)
    gc, gm = keymatch.gather_match_table(c, m)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), counts=gc.numpy(), matches=gm.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("window", [-1, 2])
def test_gather_match_table_world2_gloo(tmp_path, window):
    sizes = [120, 90, 0, 150, 60, 110]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, sizes, window, str(tmp_path)), nprocs=2, join=True)
    imgs = synth.sift_like_descriptors(len(sizes), sizes, seed=13)
    c_all, m_all = _oracle_shard(imgs, 0, len(sizes), window)
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(got["counts"], c_all)
        assert np.array_equal(got["matches"], m_all)
    assert c_all.sum() > 0


def test_shard_range_equals_python_rule():
    sizes = [5000] * 37 + [0, 17, 4000]
    key_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    for world in (1, 2, 3, 8):
        for window in (-1, 4):
            want = keymatch.shard_images(sizes, window, world)
            got = [keymatch.shard_range(key_off, window, world, r) for r in range(world)]
            assert got == want, (world, window)
