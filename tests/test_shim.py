"""The reference-facing link shims (shim/): run_sfm with the reference signature forwarding to the GPU
solver, and the UNMODIFIED reference KeyMatchFull main (src/KeyMatchFull.cpp) linked against the GPU
MatchKeys shim.  The CPU tests only check the artefacts exist and export the reference symbols; the GPU
tests run them."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from bundler_sfm_b200 import bundle, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "shim", "_build")
KM = os.path.join(BUILD, "KeyMatchFull_b200")
SFMDRV = os.path.join(BUILD, "libsfmdrv_b200.so")


def test_sfmdrv_shim_exports_run_sfm():
    if not os.path.exists(SFMDRV):
        pytest.skip("shim/_build not built")
    lib = ctypes.CDLL(SFMDRV)
    assert hasattr(lib, "run_sfm")


def test_keymatchfull_shim_binary_links_reference_main():
    if not os.path.exists(KM):
        pytest.skip("shim/_build/KeyMatchFull_b200 not built (needs /root/reference at build time)")
    r = subprocess.run([KM], capture_output=True, text=True)
    assert "Usage:" in r.stdout and "<list.txt> <outfile> [window_radius]" in r.stdout   # KeyMatchFull.cpp:65


@pytest.mark.gpu
def test_unmodified_keymatchfull_main_on_gpu_matches_oracle_table(tmp_path, oracle):
    if not os.path.exists(KM):
        pytest.skip("shim/_build/KeyMatchFull_b200 not built")
    sizes = [400, 350, 0, 380, 300]
    imgs = synth.sift_like_descriptors(len(sizes), sizes, seed=17)
    names = []
    for i, d in enumerate(imgs):
        p = tmp_path / f"img{i}.key"
        synth.write_key_file(str(p), d, seed=i)
        names.append(str(p))
    lst = tmp_path / "list_keys.txt"
    lst.write_text("\n".join(names) + "\n")
    for window, extra in ((-1, []), (2, ["2"])):
        out = tmp_path / f"matches_{window}.txt"
        r = subprocess.run([KM, str(lst), str(out)] + extra, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        want, _ = oracle.match_all_pairs_port(imgs, window, 0.6, 16)
        assert out.read_text() == want


@pytest.mark.gpu
def test_run_sfm_through_reference_signature_shim(oracle):
    if not os.path.exists(SFMDRV):
        pytest.skip("shim/_build not built")
    lib = ctypes.CDLL(SFMDRV)
    fn = lib.run_sfm
    bundle._bind_run_sfm(fn)
    fn.restype = None
    scene = synth.ba_scene(8, 300, 4, seed=2)
    got = bundle.call_run_sfm(fn, scene)
    ref = oracle.run_sfm_oracle(scene)
    nvis = scene["projections"].shape[0]
    assert abs(bundle.reprojection_rmse(scene, got) - np.sqrt(ref["info"][1] / nvis)) <= 1e-5
    assert np.max(np.abs(got["pts"] - ref["pts"])) / np.max(np.abs(ref["pts"])) <= 1e-4
