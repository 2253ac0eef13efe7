"""dev timing of the BA solve on the GPU box (not the bench)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BSFM_BA_VERBOSE", "0")
from bundler_sfm_b200 import bundle, synth

m, n, L = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (50, 20000, 5)
scene = synth.ba_scene(m, n, L, seed=1234)
for rep in range(3):
    if rep == 2:
        os.environ["BSFM_BA_TIMING"] = "1"
    t = time.time()
    out = bundle.run_sfm(scene)
    wall = time.time() - t
    tm = bundle.last_timing()
    nvis = scene["projections"].shape[0]
    print(f"rep{rep}: iters={int(out['info'][5])} stop={int(out['info'][6])} rmse={np.sqrt(out['info'][1]/nvis):.6f} wall={wall*1e3:.2f}ms "
          f"dev_total={tm['total_ms']:.2f}ms iters/s(wall)={out['info'][5]/wall:.1f} launches={tm['launches']}", flush=True)
print(tm)
