"""diagnostics for the tensor-core kernel on small cases (GPU box)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundler_sfm_b200 import keymatch, synth
from oracle import loader

for (n1, n2, ratio) in [(128, 256, 0.6), (128, 256, 0.999), (300, 700, 0.8), (1000, 5000, 0.6)]:
    imgs = synth.sift_like_descriptors(2, [n1, n2], seed=n1 + n2)
    want = loader.match_pair_port(imgs[0], imgs[1], ratio)
    for name, sel in (("dp4a", "1"), ("tc", "0")):
        os.environ["BSFM_MATCH_KERNEL"] = sel
        try:
            got = keymatch.match_keys(imgs[0], imgs[1], ratio)
        except Exception as e:
            print(name, n1, n2, ratio, "EXC", e, flush=True)
            continue
        ok = np.array_equal(got, want)
        print(name, n1, n2, ratio, "ok" if ok else "MISMATCH", got.shape[0], want.shape[0], flush=True)
        if not ok:
            gs, ws = {tuple(r) for r in got}, {tuple(r) for r in want}
            print("  missing", sorted(ws - gs)[:8], "extra", sorted(gs - ws)[:8], flush=True)
