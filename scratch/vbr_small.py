import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import lame_b200, oracle
from conftest import make_signal
S, F = 2, 8
pcm = np.stack([make_signal("click", F * 1152, seed=61 + 4 * s) for s in range(S)])
enc = lame_b200.BatchEncoder(S, 44100, 2, 2, -1, -1, frames_per_launch=4, vbr=4)
_, a = enc.encode(pcm); _, b = enc.flush(); enc.close()
for s in range(S):
    want = oracle.PortEncoder(44100, 2, 2, -1, -1, vbr=4).encode_all(pcm[s, 0], pcm[s, 1])
    got = a[s] + b[s]
    d = [i for i in range(min(len(got), len(want))) if got[i] != want[i]]
    print("stream", s, len(got), len(want), "first diff", d[:5])
