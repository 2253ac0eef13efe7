"""Generates tests/golden/ba_config3_it4_ref.json: the UNMODIFIED reference on BASELINE.json configs[2]
(synth.ba_scene(1000, 500000, 6, seed=1234)) stopped after exactly 4 LM iterations (REF_SBA_ITMAX=4, opts[5] = 0 so that
Snavely's stop-8 rule cannot fire): the "equal iteration count" comparison of SURVEY.md H1 / section 8(d), free of the
eps4 = 0 rounding knife-edge that decides where the full solve stops.  ~3 minutes of CPU.
   python tests/golden/make_ba_config3_fixed_it_ref.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["OPENBLAS_NUM_THREADS"] = "8"
os.environ["REF_SBA_ITMAX"] = "4"
os.environ["REF_SBA_EPS5"] = "0"
from bundler_sfm_b200 import bundle, synth  # noqa: E402
from oracle import loader  # noqa: E402

scene = synth.ba_scene(1000, 500000, 6, seed=1234)
t = time.time()
out = loader.run_sfm_ref(scene)
dt = time.time() - t
idx = np.random.default_rng(0).choice(500000, 400, replace=False)
res = {"seconds": dt, "itmax": 4, "eps5": 0.0, "info": out["info"].tolist(), "rc": int(out["rc"]), "pt_idx": idx.tolist(),
       "pts": out["pts"][idx].tolist(), "c": out["c"].tolist(), "f": out["f"].tolist(), "k": out["k"].tolist(), "R": out["R"].tolist(),
       "rmse": bundle.reprojection_rmse(scene, out)}
json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ba_config3_it4_ref.json"), "w"))
print("done", dt, out["info"])
