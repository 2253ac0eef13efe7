"""The reprojection statistics / outlier pass (src/Bundle.cpp:659-856): the CPU oracle
(oracle/outlier_oracle.py on the unmodified reference sfm_project_rd / kth_element_copy) against its committed golden
vectors, and bsfm_reprojection_outliers on the GPU against both."""
import os

import numpy as np
import pytest

from bundler_sfm_b200 import bundle, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "outlier_golden.npz")
KEYS = ("vmask", "projections", "R", "c", "f", "k", "pts")
CASES = ("kermit_default", "kermit_tight", "syn")


def case(g, name):
    scene = {k: g[f"{name}_{k}"] for k in KEYS}
    lo, hi = g[f"{name}_thresholds"]
    prot = g[f"{name}_protected"] if f"{name}_protected" in g.files else None
    ref = {k: g[f"{name}_ref_{k}"] for k in ("dist", "stats", "outliers", "errors", "global_mean")}
    return scene, float(lo), float(hi), prot, ref


def same(got, ref, what):
    # distances: same fp64 operations in the same order (the kernel is compiled without FMA contraction) -> the last-bit
    # differences of sqrt/div are the only slack; thresholds and k-th elements are picked, not computed
    assert np.allclose(got["dist"], ref["dist"], rtol=1e-12, atol=1e-12), what
    assert np.array_equal(got["stats"][:, 0], ref["stats"][:, 0]), what
    assert np.allclose(got["stats"][:, 1:], ref["stats"][:, 1:], rtol=1e-12, atol=1e-12, equal_nan=True), what
    assert np.array_equal(got["outliers"], ref["outliers"]), (what, got["outliers"], ref["outliers"])
    assert np.allclose(got["errors"], ref["errors"], rtol=1e-12), what
    assert abs(got["global_mean"] - float(ref["global_mean"])) <= 1e-12 * abs(float(ref["global_mean"])), what


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name, oracle):
    """pins the restatement: reference-backed and numpy-port variants both reproduce the committed vectors"""
    from oracle import outlier_oracle
    g = np.load(GOLD)
    scene, lo, hi, prot, ref = case(g, name)
    for use_ref in (True, False):
        got = outlier_oracle.reprojection_outliers(scene, 1, lo, hi, prot, use_reference=use_ref)
        same(got, ref, (name, use_ref))
    if name == "syn":
        assert ref["stats"][7, 0] == 0 and np.isnan(ref["stats"][7, 1]) and ref["stats"][7, 4] == lo     # empty camera
        assert ref["stats"][6, 0] == 1 and ref["stats"][6, 3] == 0.0 and ref["stats"][6, 4] == lo        # k >= n -> 0.0 -> clamp
        assert len(ref["outliers"]) > 10 and not np.any(prot[ref["outliers"]])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_equals_golden(name):
    g = np.load(GOLD)
    scene, lo, hi, prot, ref = case(g, name)
    got = bundle.reprojection_outliers(scene, 1, lo, hi, prot)
    same(got, ref, name)
    assert got["count"] == len(ref["outliers"])


@pytest.mark.gpu
def test_gpu_after_solve_vs_oracle(oracle):
    """the real sequence: run_sfm on the GPU, then the outlier pass on its output, against the oracle on the same output"""
    from oracle import outlier_oracle
    scene = synth.ba_scene(12, 1500, 4, seed=41)
    rng = np.random.default_rng(3)
    proj = scene["projections"].copy()
    bad = rng.choice(len(proj), 40, replace=False)
    proj[bad] += rng.normal(0, 25.0, (40, 2))
    scene = dict(scene, projections=proj)
    sol = bundle.run_sfm(scene)
    solved = dict(scene, R=sol["R"], c=sol["c"], f=sol["f"], k=sol["k"], pts=sol["pts"])
    got = bundle.reprojection_outliers(solved, 1, 2.0, 16.0)
    ref = outlier_oracle.reprojection_outliers(solved, 1, 2.0, 16.0)
    same(got, ref, "after_solve")
    assert got["count"] >= 20
    # capacity smaller than the count: the count is still returned, the prefix is filled
    few = bundle.reprojection_outliers(solved, 1, 2.0, 16.0, cap=5)
    assert few["count"] == got["count"] and np.array_equal(few["outliers"], got["outliers"][:5])
