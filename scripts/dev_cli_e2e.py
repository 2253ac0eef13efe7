"""dev/measurement: the KeyMatchFull command line end to end (key files on disk -> matches.init.txt) three ways:
  reference : oracle/_ref/KeyMatchFull (unmodified src/KeyMatchFull.cpp + ANN, CPU, 200-visit cap)
  per-pair  : shim/_build/KeyMatchFull_b200 (unmodified main, MatchKeys shim -> one GPU call per pair)
  persistent: shim/_build/KeyMatchFull_b200_persistent (parallel key reader + one device-resident database run)
usage: python scripts/dev_cli_e2e.py [num_images] [keys_per_image]"""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundler_sfm_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
imgs = synth.sift_like_descriptors(N, K, seed=7)
with tempfile.TemporaryDirectory() as d:
    t = time.time()
    names = []
    for i, x in enumerate(imgs):
        p = os.path.join(d, f"img{i:04d}.key")
        synth.write_key_file(p, x, seed=i)
        names.append(p)
    lst = os.path.join(d, "list_keys.txt")
    open(lst, "w").write("\n".join(names) + "\n")
    print(f"wrote {N} key files x {K} keys in {time.time() - t:.1f} s", flush=True)
    outs = {}
    for tag, exe in (("persistent", "shim/_build/KeyMatchFull_b200_persistent"), ("per-pair", "shim/_build/KeyMatchFull_b200"),
                     ("reference", "oracle/_ref/KeyMatchFull")):
        exe = os.path.join(ROOT, exe)
        if not os.path.exists(exe):
            print(tag, "missing"); continue
        out = os.path.join(d, f"matches_{tag}.txt")
        for rep in range(2 if tag != "reference" else 1):      # second run = warm CUDA context / page cache
            t = time.time()
            r = subprocess.run([exe, lst, out], capture_output=True, text=True)
            wall = time.time() - t
        assert r.returncode == 0, r.stdout + r.stderr
        outs[tag] = open(out).read()
        lines = [l for l in r.stdout.splitlines() if "Reading keys" in l or "b200]" in l or (tag == "persistent" and "Matching took" in l)]
        print(f"{tag:10s}: wall {wall:7.2f} s  pairs {N * (N - 1) // 2}  | " + " | ".join(lines[:3]), flush=True)
    if "persistent" in outs and "per-pair" in outs:
        print("persistent == per-pair table:", outs["persistent"] == outs["per-pair"])
    if "persistent" in outs and "reference" in outs:
        a, b = outs["persistent"].count("\n"), outs["reference"].count("\n")
        print(f"lines: exact (GPU) {a}, reference 200-visit cap {b}")
