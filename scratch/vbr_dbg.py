import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import lame_b200
from conftest import make_signal
lame_b200._lib = lame_b200.load_library("/root/repo/scratch/variants/lib_dbg.so")
S, F = 2, 8
pcm = np.stack([make_signal("click", F * 1152, seed=61 + 4 * s) for s in range(S)])
enc = lame_b200.BatchEncoder(S, 44100, 2, 2, -1, -1, frames_per_launch=4, vbr=4)
enc.encode(pcm); enc.flush(); enc.close()
