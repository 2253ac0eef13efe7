"""dev: measured fp64 issue intervals (vector DFMA vs DMMA) on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundler_sfm_b200 import _lib

lib = _lib.load_library()
for mode, name in ((0, "DFMA"), (1, "DMMA m8n8k4")):
    for warps in (1, 4, 8, 16):
        v = lib.bsfm_measure_fp64_issue_cycles(mode, warps, 4096)
        print(f"{name:12s} warps/CTA={warps:2d}: {v:7.2f} cycles per warp instruction per sub-partition", flush=True)
