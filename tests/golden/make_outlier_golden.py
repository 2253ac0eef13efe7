"""Generates tests/golden/outlier_golden.npz: inputs and outputs of the outlier statistics pass (src/Bundle.cpp:659-856)
computed by oracle/outlier_oracle.py with the UNMODIFIED reference sfm_project_rd and kth_element_copy
(oracle/_ref/libref_sba.so):
  kermit  the kermit example reconstruction as bundled (tests/golden/ba_golden.npz reference output), thresholds of
          bundler's defaults (min 8, max 16 px, BundlerApp.h) and a tight pair (0.5 / 2.0) that produces outliers
  syn     synthetic 8 cameras / 400 points with 25 gross measurement errors, 10 protected points, one camera that
          sees a single point and one that sees none (kth_element's k >= n rule, empty camera)
Run in the build container:  python tests/golden/make_outlier_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader, outlier_oracle  # noqa: E402
from bundler_sfm_b200 import synth  # noqa: E402

KEYS = ("vmask", "projections", "R", "c", "f", "k", "pts")


def syn_scene():
    rng = np.random.default_rng(21)
    sc = synth.ba_scene(8, 400, 4, seed=9)
    vmask = sc["vmask"].astype(bool).copy()
    proj_full = np.zeros(vmask.shape + (2,))
    proj_full[vmask] = sc["projections"]
    vmask[:, 7] = False                      # camera 7 sees nothing
    vmask[:, 6] = False; vmask[5, 6] = True  # camera 6 sees one point (k = iround(0.8) = 1 >= n = 1)
    if not np.any(proj_full[5, 6]):
        proj_full[5, 6] = (3.0, -2.0)
    keep = vmask.sum(axis=1) > 0
    assert keep.all()
    proj = proj_full[vmask]
    bad = rng.choice(len(proj), 25, replace=False)
    proj[bad] += rng.normal(0, 60.0, (25, 2))
    sc = dict(sc, vmask=vmask.astype(np.int8), projections=proj)
    prot = np.zeros(400, np.int8); prot[rng.choice(400, 10, replace=False)] = 1
    return sc, prot


def main():
    assert loader.ref_sba() is not None, "build oracle/_ref first"
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ba_golden.npz"))
    out = {}
    kermit = {k: gold[f"kermit_ref_{k}"] if k in ("R", "c", "f", "k", "pts") else gold[f"kermit_{k}"] for k in KEYS}
    syn, prot = syn_scene()
    cases = [("kermit_default", kermit, 8.0, 16.0, None), ("kermit_tight", kermit, 0.5, 2.0, None), ("syn", syn, 2.0, 16.0, prot)]
    for name, scene, lo, hi, pr in cases:
        r = outlier_oracle.reprojection_outliers(scene, 1, lo, hi, pr)
        for k in KEYS:
            out[f"{name}_{k}"] = scene[k]
        out[f"{name}_thresholds"] = np.array([lo, hi])
        if pr is not None:
            out[f"{name}_protected"] = pr
        for k in ("dist", "stats", "outliers", "errors"):
            out[f"{name}_ref_{k}"] = r[k]
        out[f"{name}_ref_global_mean"] = np.array(r["global_mean"])
        print(name, "observations", len(r["dist"]), "outliers", len(r["outliers"]), "global mean", r["global_mean"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "outlier_golden.npz"), **out)


if __name__ == "__main__":
    main()
