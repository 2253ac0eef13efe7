"""Multi-GPU MATCH behind the C ABI (needs >= 2 GPUs, skipped otherwise; run with `gpurun --gpus 2`):
bsfm_match_all_pairs_multi (one host thread per GPU, NCCL inside the library: cooperative key-database build + all-gather
of the match table) must return the single-GPU table byte for byte, and so must the persistent KeyMatchFull CLI with
--gpus.  The one-GPU half of the sharded build (chunked device layout with a gap behind every group) is covered on any box
by test_sharded_layout_single_rank."""
import os
import subprocess

import numpy as np
import pytest

from bundler_sfm_b200 import keymatch, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_sharded_layout_single_rank():
    """a world-size-1 communicator exercises the cooperative build path (NCCL init, chunked layout) on one GPU"""
    sizes = [700, 0, 300, 17, 513, 256, 1, 640]
    imgs = synth.sift_like_descriptors(len(sizes), sizes, seed=21)
    keys, key_off = keymatch.concat_keys(imgs)
    ref = keymatch.key_match_full(imgs, -1, 0.6)
    comm = keymatch.Comm(keymatch.Comm.unique_id(), 0, 1)
    db = keymatch.KeyDatabase(keys, key_off, comm=comm)
    db.run(0, len(sizes), -1, 0.6)
    assert db.allgather(comm) == ref[2].shape[0]
    c, m = db.gathered_fetch()
    db.close(); comm.close()
    assert np.array_equal(c, ref[1]) and np.array_equal(m, ref[2])


@pytest.mark.parametrize("window", [-1, 3])
def test_all_pairs_multi_equals_single_gpu(window):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    sizes = [1500, 1400, 0, 1600, 700, 1550, 1480, 90, 1500, 1520, 1300, 1610, 5, 1450]
    imgs = synth.sift_like_descriptors(len(sizes), sizes, seed=33)
    ref = keymatch.key_match_full(imgs, window, 0.6)
    for g in sorted({2, min(n, 3), min(n, 8)}):
        got = keymatch.key_match_full_multi(imgs, window, 0.6, ngpus=g)
        assert got[1].tobytes() == ref[1].tobytes(), (g, window)
        assert got[2].tobytes() == ref[2].tobytes(), (g, window)
    assert ref[1].sum() > 1000


def test_persistent_cli_multi_gpu_writes_identical_table(tmp_path):
    n = _ngpu()
    exe = os.path.join(ROOT, "shim", "_build", "KeyMatchFull_b200_persistent")
    if n < 2 or not os.path.exists(exe):
        pytest.skip("needs two GPUs and shim/_build/KeyMatchFull_b200_persistent")
    sizes = [900, 850, 1000, 60, 940, 910]
    imgs = synth.sift_like_descriptors(len(sizes), sizes, seed=41)
    names = []
    for q, k in enumerate(imgs):
        path = tmp_path / f"img{q}.key"
        synth.write_key_file(str(path), k, seed=q)
        names.append(str(path))
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    outs = []
    for g in (1, 2):
        out = tmp_path / f"matches{g}.txt"
        r = subprocess.run([exe, str(tmp_path / "list.txt"), str(out)], capture_output=True, text=True, env=dict(os.environ, BSFM_MATCH_GPUS=str(g)))
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(out.read_bytes())
    assert outs[0] == outs[1] and len(outs[0]) > 1000
