import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BSFM_BA_VERBOSE"] = "0"
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
from bundler_sfm_b200 import bundle, synth
scene = synth.ba_scene(50, 20000, 5, seed=1234)
sys.stdout.flush()
saved = os.dup(1); dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 1)
ts = []
for rep in range(5):
    if rep == 4: os.environ["BSFM_BA_HOST_TIMING"] = "1"
    t0 = time.perf_counter()
    cams = bundle.make_cameras(scene["R"], scene["c"], scene["f"], scene["k"])
    t1 = time.perf_counter()
    out = bundle.run_sfm(scene)
    t2 = time.perf_counter()
    tm = bundle.last_timing()
    ts.append((t1 - t0, t2 - t1, tm["total_ms"]))
os.dup2(saved, 1)
for t in ts: print("make_cameras %.2f ms  run_sfm %.2f ms  device %.2f ms" % (t[0] * 1e3, t[1] * 1e3, t[2]))
