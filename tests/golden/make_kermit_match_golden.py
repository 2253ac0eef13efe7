"""Generates tests/golden/kermit_match_golden.npz: REAL SIFT descriptors of the reference's example image set
(examples/kermit, BASELINE.json configs[0]) and the match lists of the UNMODIFIED reference on them.
`sift` (Lowe's binary) is not in this image, so the keys come from OpenCV SIFT (descriptors already 0..255 with
norm ~512 -> uint8), written and re-read as Lowe-format .key files by the reference's own ReadKeyFile; the 11 images give 55 pairs.
Stored: the descriptors (uint8), the exact-mode table of oracle/_ref (MatchKeys, max_pts_visit = 0, keys2a.cpp:347-372) as
per-pair counts + (idx1, idx2) lists, and the same for the in-bundler matcher (MatchKeysExhaustive, ratio 0.75, keys.cpp) on 3 pairs.
   python tests/golden/make_kermit_match_golden.py
"""
import glob
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402

files = sorted(glob.glob("/root/reference/examples/kermit/*.jpg"))
sift = cv2.SIFT_create()
descs = []
for f in files:
    _, d = sift.detectAndCompute(cv2.imread(f, 0), None)
    descs.append(np.clip(np.floor(d), 0, 255).astype(np.uint8))
out = {"num_images": np.array(len(descs))}
for i, d in enumerate(descs):
    out[f"desc{i}"] = d
counts, matches = [], []
for i in range(len(descs)):
    for j in range(i):
        m = loader.match_pair_ref(descs[j], descs[i], 0.6, 0)      # KeyMatchFull order: (j, i), queries = image j
        counts.append(m.shape[0]); matches.append(m)
out["counts"] = np.array(counts, np.int32)
out["matches"] = np.concatenate(matches, 0).astype(np.int32)
for (a, b) in ((0, 1), (3, 7), (10, 9)):
    out[f"keys_{a}_{b}"] = loader.keys_match_ref(descs[a], descs[b], None, False, 0.75, exhaustive=True)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kermit_match_golden.npz"), **out)
print("images", len(descs), "keys", [d.shape[0] for d in descs], "pairs", len(counts), "matches", int(out["counts"].sum()),
      "pairs >= 16:", int((out["counts"] >= 16).sum()))
