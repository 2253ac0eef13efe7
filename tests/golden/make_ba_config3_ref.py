"""Generates tests/golden/ba_config3_ref.json: the UNMODIFIED reference run_sfm (oracle/_ref) on
BASELINE.json configs[2] (synthetic BA, 1000 cameras / 500,000 points / 3,000,000 observations,
synth.ba_scene(1000, 500000, 6, seed=1234)).  Takes ~14 minutes of CPU (835 s measured, 8 OpenBLAS
threads for dpotrf, everything else single-threaded).  Only a summary is stored: info[10], all camera
parameters and 400 sampled points.   python tests/golden/make_ba_config3_ref.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["OPENBLAS_NUM_THREADS"] = "8"
from bundler_sfm_b200 import bundle, synth  # noqa: E402
from oracle import loader  # noqa: E402

scene = synth.ba_scene(1000, 500000, 6, seed=1234)
t = time.time()
out = loader.run_sfm_ref(scene)
dt = time.time() - t
idx = np.random.default_rng(0).choice(500000, 400, replace=False)
res = {"seconds": dt, "info": out["info"].tolist(), "rc": int(out["rc"]), "pt_idx": idx.tolist(), "pts": out["pts"][idx].tolist(),
       "c": out["c"].tolist(), "f": out["f"].tolist(), "k": out["k"].tolist(), "R": out["R"].tolist(),
       "rmse": bundle.reprojection_rmse(scene, out)}
json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ba_config3_ref.json"), "w"))
print("done", dt, out["info"])
