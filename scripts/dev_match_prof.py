"""dev: cycle accounting of the MATCH pipeline roles (needs a library built with -DBSFM_TC_PROFILE, passed via
BSFM_LIB_PATH).  Prints average cycles per database tile for the MMA thread, the TMA producer and one epilogue warp."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundler_sfm_b200 import _lib, keymatch, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
imgs = synth.sift_like_descriptors(N, K, seed=7)
keys, key_off = keymatch.concat_keys(imgs)
db = keymatch.KeyDatabase(keys, key_off)
lib = _lib.load_library()
buf = (ctypes.c_ulonglong * 16)()
db.run(0, N, -1, 0.6)
lib.bsfm_debug_prof(buf)          # discard the warm-up
db.run(0, N, -1, 0.6)
tm = db.timing()
lib.bsfm_debug_prof(buf)
p = np.array(list(buf), dtype=np.float64)
print(f"search_ms={tm['search_ms']:.3f}")
print(f"MMA thread   per tile: wait b_full {p[0]/p[3]:.0f}, wait b_full+t_empty {p[1]/p[3]:.0f}, issue {p[2]/p[3]:.0f}  (tiles {p[3]:.0f})")
print(f"TMA producer per tile: wait b_empty {p[4]/p[6]:.0f}, issue {p[5]/p[6]:.0f}  (tiles {p[6]:.0f})")
print(f"epilogue warp per own tile: wait t_full {p[8]/p[11]:.0f}, wait->release {p[9]/p[11]:.0f}, wait->end {p[10]/p[11]:.0f}  (tiles {p[11]:.0f})")
print(f"per unit: MMA a_full wait {p[12]/p[13]:.0f} (units {p[13]:.0f}); epilogue warp: tiles phase {p[15]/p[13]:.0f}, merge+push {p[14]/p[13]:.0f}")
