import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BSFM_BA_VERBOSE"] = "0"; os.environ["BSFM_VERBOSE"] = "1"
from bundler_sfm_b200 import bundle, synth
from bundler_sfm_b200._lib import load_library
for (m, n, L) in [(300, 30000, 6), (600, 60000, 6), (1000, 100000, 6)]:
    scene = synth.ba_scene(m, n, L, seed=1)
    t = time.time(); out = bundle.run_sfm(scene); dt = time.time() - t
    print("RESULT", m, n, "iters", out["info"][5], "stop", out["info"][6], "rmse", np.sqrt(out["info"][1] / scene["projections"].shape[0]), "t", dt, load_library().bsfm_last_error(), flush=True)
