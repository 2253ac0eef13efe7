"""Generates tests/golden/ba_mot_golden.npz: motion-only bundle adjustment (run_sfm with fix_points = 1 ->
sba_mot_levmar, lib/sfm-driver/sfm.c:843-849) of the UNMODIFIED reference (oracle/_ref/libref_sba.so) on
  kermit   the kermit example reconstruction (inputs taken from tests/golden/ba_golden.npz) with the cameras
           perturbed (seed 11), points fixed
  syn10    synthetic 10 cameras / 500 points / 2000 observations (synth.ba_scene seed 3)
  syn10c   the same with the RunBundler.sh camera constraints (focal prior 1e-4, distortion weight 100)
  syn6nf   6 cameras, focal length not estimated (cnp = 8)
Stored: the inputs and the reference's final cameras and info[10].  Run in the build container:
    python tests/golden/make_ba_mot_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from bundler_sfm_b200 import synth  # noqa: E402

KEYS = ("vmask", "projections", "R", "c", "f", "k", "pts")


def main():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ba_golden.npz"))
    out = {}
    cases = []
    rng = np.random.default_rng(11)
    kermit = {k: gold[f"kermit_ref_{k}"].copy() if k in ("R", "c", "f", "k", "pts") else gold[f"kermit_{k}"] for k in KEYS}
    kermit["c"] = kermit["c"] + rng.normal(0, 0.01, kermit["c"].shape)      # start the cameras off the optimum
    kermit["f"] = kermit["f"] * (1 + rng.normal(0, 0.01, kermit["f"].shape))
    cases.append(("kermit", kermit, {}))
    syn = synth.ba_scene(10, 500, 4, seed=3)
    cases.append(("syn10", syn, {}))
    m = 10
    cons = dict(use_constraints=1, constrained=np.tile(np.array([0, 0, 0, 0, 0, 0, 1, 1, 1], np.int8), (m, 1)),
                constraints=np.tile(np.array([0, 0, 0, 0, 0, 0, 800., 0, 0]), (m, 1)),
                weights=np.tile(np.array([0, 0, 0, 0, 0, 0, 1e-4, 100., 100.]), (m, 1)))
    cases.append(("syn10c", syn, cons))
    cases.append(("syn6nf", synth.ba_scene(6, 300, 4, seed=5), dict(est_focal_length=0)))
    for name, scene, kw in cases:
        ref = loader.run_sfm_ref(scene, fix_points=1, **kw)
        assert np.array_equal(ref["pts"], scene["pts"])
        for k in KEYS:
            out[f"{name}_{k}"] = scene[k]
        for k, v in kw.items():
            if isinstance(v, np.ndarray):
                out[f"{name}_{k}"] = v
        for k in ("R", "c", "f", "k", "info"):
            out[f"{name}_ref_{k}"] = ref[k]
        nvis = scene["projections"].shape[0]
        print(name, "iters", int(ref["info"][5]), "stop", int(ref["info"][6]), "rmse", np.sqrt(ref["info"][1] / nvis), "init", np.sqrt(ref["info"][0] / nvis))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ba_mot_golden.npz"), **out)


if __name__ == "__main__":
    main()
