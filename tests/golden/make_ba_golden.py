"""Generates tests/golden/ba_golden.npz: (a) the kermit example reconstruction shipped with the
reference (examples/kermit/results.example/bundle.out, 11 cameras / 634 points / 2039 observations)
converted to run_sfm arguments, (b) small synthetic scenes, each with the output of the UNMODIFIED
reference run_sfm (oracle/_ref/libref_sba.so: lib/sba-1.5 + lib/sfm-driver compiled in place,
OpenBLAS 0.3.15) -- final parameters and info[10].  Run in the build container:
    python tests/golden/make_ba_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from bundler_sfm_b200 import synth  # noqa: E402

KERMIT = "/root/reference/examples/kermit/results.example/bundle.out"


def load_bundle_out(path):
    """bundle.out v0.3 (README.md:194-264): per camera f k1 k2 / R / t, per point XYZ / RGB / views"""
    tok = open(path).read().split("\n")
    assert tok[0].startswith("# Bundle file v0.3")
    vals = " ".join(tok[1:]).split()
    pos = 0
    ncam, npts = int(vals[0]), int(vals[1]); pos = 2
    f = np.zeros(ncam); k = np.zeros((ncam, 2)); R = np.zeros((ncam, 9)); t = np.zeros((ncam, 3))
    for j in range(ncam):
        f[j], k[j, 0], k[j, 1] = map(float, vals[pos:pos + 3]); pos += 3
        R[j] = list(map(float, vals[pos:pos + 9])); pos += 9
        t[j] = list(map(float, vals[pos:pos + 3])); pos += 3
    pts = np.zeros((npts, 3)); views = []
    for i in range(npts):
        pts[i] = list(map(float, vals[pos:pos + 3])); pos += 3
        pos += 3  # colour
        nv = int(vals[pos]); pos += 1
        for _ in range(nv):
            cam, key, x, y = int(vals[pos]), int(vals[pos + 1]), float(vals[pos + 2]), float(vals[pos + 3]); pos += 4
            views.append((i, cam, x, y))
    return f, k, R, t, pts, views


def kermit_scene():
    f, k, R, t, pts, views = load_bundle_out(KERMIT)
    m, n = len(f), len(pts)
    keep = f > 0                      # cameras that were registered (unregistered ones are all-zero)
    remap = -np.ones(m, int); remap[keep] = np.arange(keep.sum())
    f, k, R, t = f[keep], k[keep], R[keep], t[keep]
    m = len(f)
    c = np.stack([-R[j].reshape(3, 3).T @ t[j] for j in range(m)])     # t = -R c
    views = [(i, remap[cam], x, y) for (i, cam, x, y) in views if remap[cam] >= 0]
    views.sort(key=lambda v: (v[0], v[1]))
    vmask = np.zeros((n, m), np.int8)
    proj = []
    for (i, cam, x, y) in views:
        if vmask[i, cam]:
            continue
        vmask[i, cam] = 1
        proj.append((x, y))
    has = vmask.sum(1) >= 2
    proj = np.array(proj)[np.repeat(has, vmask.sum(1))]
    return {"vmask": vmask[has], "projections": proj, "R": R, "c": c, "f": f, "k": k, "pts": pts[has]}


def add(out, name, scene, **kw):
    for key in ("vmask", "projections", "R", "c", "f", "k", "pts"):
        out[f"{name}_{key}"] = scene[key]
    ref = loader.run_sfm_ref(scene, **kw)
    for key in ("R", "c", "f", "k", "pts", "info"):
        out[f"{name}_ref_{key}"] = ref[key]
    nvis = scene["projections"].shape[0]
    print(f"{name}: nvis={nvis} iters={int(ref['info'][5])} stop={int(ref['info'][6])} rmse={np.sqrt(ref['info'][1]/nvis):.6f}")


def main():
    out = {}
    ks = kermit_scene()
    # perturb the shipped optimum a little so that the solve does real work from a known start
    rng = np.random.default_rng(5)
    ks2 = dict(ks)
    ks2["pts"] = ks["pts"] + 0.002 * rng.standard_normal(ks["pts"].shape)
    ks2["c"] = ks["c"] + 0.001 * rng.standard_normal(ks["c"].shape)
    add(out, "kermit", ks2)
    # RunBundler.sh-style constraints: k1,k2 -> 0 with weight 100, focal prior weight 1e-4
    m = len(ks["f"])
    constrained = np.zeros((m, 9), np.int8); constrained[:, 6:9] = 1
    constraints = np.zeros((m, 9)); constraints[:, 6] = ks["f"] * 1.02
    weights = np.zeros((m, 9)); weights[:, 6] = 1e-4; weights[:, 7:9] = 100.0
    out["kermitc_constrained"] = constrained; out["kermitc_constraints"] = constraints; out["kermitc_weights"] = weights
    add(out, "kermitc", ks2, use_constraints=1, constrained=constrained, constraints=constraints, weights=weights)
    add(out, "syn10", synth.ba_scene(10, 500, 4, seed=3))
    add(out, "syn6nf", synth.ba_scene(6, 300, 3, seed=4), est_focal_length=0, undistort=0)
    add(out, "syn8nd", synth.ba_scene(8, 400, 4, seed=6), est_focal_length=1, undistort=0)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ba_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
