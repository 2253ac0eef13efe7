"""Generates tests/golden/match_golden.npz from the UNMODIFIED reference matcher
(oracle/_ref/libref_match.so = lib/ann_1.1_char + src/keys2a.cpp compiled in place, exact mode
max_pts_visit=0, plus the stock 200-cap mode for the informational recall number).
Run in the build container (needs /root/reference):  python tests/golden/make_match_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from bundler_sfm_b200 import synth  # noqa: E402


def edge_cases():
    rng = np.random.default_rng(11)
    cases = {}
    base = synth.random_descriptors(40, 1)
    # identical descriptors (d0 = 0) + one far key
    q = base[:10].copy()
    db = np.concatenate([base[:10], 255 - base[:3]], 0)
    cases["identical"] = (q, db)
    # all zeros vs all 255 (d^2 = 8,323,200 saturating case) and mixtures
    cases["zeros_vs_255"] = (np.zeros((3, 128), np.uint8), np.concatenate([np.full((2, 128), 255, np.uint8), np.zeros((1, 128), np.uint8), base[:2]], 0))
    # database with a single key (d1 = INT_MAX) and with two keys
    cases["n2_is_1"] = (base[:7], base[20:21])
    cases["n2_is_2"] = (base[:7], base[20:22])
    # ratio boundary: craft d0/d1 around 0.36 exactly: d0 = 9k, d1 = 25k (integer form 25 d0 < 9 d1)
    q = np.zeros((6, 128), np.uint8)
    dbs = []
    for k, (a, b) in enumerate([(9, 25), (9, 26), (10, 25), (36, 100), (35, 100), (36, 101)]):
        # key with exactly `a` ones at positions [0,a) and key with `b` ones -> distances a and b to the zero query
        d0 = np.zeros(128, np.uint8); d0[:a] = 1
        d1 = np.zeros(128, np.uint8); d1[:b] = 1
        dbs.append((d0, d1))
    # one database per boundary query (queries are all zero): stack, each pair tested separately
    for k, (d0, d1) in enumerate(dbs):
        cases[f"ratio_boundary_{k}"] = (q[:1], np.stack([d1, d0], 0))
    # duplicates in the database (tie for best => never a match)
    cases["tie_best"] = (base[:5], np.concatenate([base[:5], base[:5], base[30:35]], 0))
    # random uniform, ragged sizes
    cases["uniform_37x300"] = (synth.random_descriptors(37, 2), synth.random_descriptors(300, 3))
    cases["uniform_257x129"] = (synth.random_descriptors(257, 4), synth.random_descriptors(129, 5))
    return cases


def main():
    out = {}
    imgs = synth.sift_like_descriptors(4, [700, 650, 513, 40], seed=7)
    for a in range(4):
        out[f"sift_img{a}"] = imgs[a]
    for i in range(4):
        for j in range(i):
            out[f"sift_exact_{j}_{i}"] = loader.match_pair_ref(imgs[j], imgs[i], 0.6, 0)
            out[f"sift_cap200_{j}_{i}"] = loader.match_pair_ref(imgs[j], imgs[i], 0.6, 200)
    for name, (q, db) in edge_cases().items():
        out[f"edge_{name}_q"] = q
        out[f"edge_{name}_db"] = db
        out[f"edge_{name}_m06"] = loader.match_pair_ref(q, db, 0.6, 0)
        out[f"edge_{name}_m09"] = loader.match_pair_ref(q, db, 0.9, 0)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "match_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
