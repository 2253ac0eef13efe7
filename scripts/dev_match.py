"""dev timing of the MATCH kernels on the GPU box (not the bench): prints per-kernel times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundler_sfm_b200 import keymatch, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
imgs = synth.sift_like_descriptors(N, K, seed=7)
keys, key_off = keymatch.concat_keys(imgs)
res = {}
which = sys.argv[3] if len(sys.argv) > 3 else "both"
for name, sel in [x for x in (("tc", "0"), ("dp4a", "1")) if which in ("both", x[0])]:
    os.environ["BSFM_MATCH_KERNEL"] = sel
    db = keymatch.KeyDatabase(keys, key_off)
    for rep in range(3):
        t = time.time()
        total = db.run(0, N, -1, 0.6)
        wall = time.time() - t
        tm = db.timing()
    c, m = db.fetch()
    res[name] = (c, m)
    pairs = N * (N - 1) // 2
    dp = pairs * K * K
    print(f"{name}: matches={total} search_ms={tm['search_ms']:.3f} post_ms={tm['post_ms']:.3f} total_ms={tm['total_ms']:.3f} "
          f"wall_ms={wall*1e3:.2f} launches={tm['launches']} desc-pairs/s={dp/(tm['total_ms']*1e-3):.3e} "
          f"TOPS(search)={dp*256/(tm['search_ms']*1e-3)/1e12:.1f}", flush=True)
    db.close()
if which == "both":
    print("tc == dp4a:", np.array_equal(res["tc"][0], res["dp4a"][0]) and np.array_equal(res["tc"][1], res["dp4a"][1]))
