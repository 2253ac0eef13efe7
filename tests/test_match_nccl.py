"""Two-rank NCCL test of the sharded MATCH path on real GPUs (skipped when fewer than two are visible): every rank runs its
shard of database images through libbsfm_b200.so on its own GPU, the match table is all-gathered over NCCL straight from
device memory (keymatch.gather_match_table), and the gathered table must equal the single-GPU table byte for byte."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, sizes, window, out_dir):
    import torch
    import torch.distributed as dist
    from bundler_sfm_b200 import _lib, keymatch, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    _lib.load_library().bsfm_set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    imgs = synth.sift_like_descriptors(len(sizes), sizes, seed=33)
    keys, key_off = keymatch.concat_keys(imgs)
    b, e = keymatch.shard_images(sizes, window, world)[rank]
    db = keymatch.KeyDatabase(keys, key_off)
    db.run(b, e, window, 0.6)
    counts, matches = db.result_to_torch(dev)
    gc, gm = keymatch.gather_match_table(counts, matches)
    out = {"counts": gc.cpu().numpy(), "matches": gm.cpu().numpy()}
    if rank == 0:
        db.run(0, len(sizes), window, 0.6)
        c1, m1 = db.fetch()
        out["single_counts"], out["single_matches"] = c1, m1
    db.close()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    dist.destroy_process_group()


@pytest.mark.parametrize("window", [-1, 3])
def test_nccl_gathered_table_equals_single_gpu_table(tmp_path, window):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    sizes = [1500, 1400, 0, 1600, 700, 1550, 1480, 90, 1500, 1520, 1300, 1610]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, sizes, window, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for r in (r0, r1):
        assert r["counts"].tobytes() == r0["single_counts"].tobytes()
        assert r["matches"].tobytes() == r0["single_matches"].tobytes()
    assert r0["single_counts"].sum() > 1000
