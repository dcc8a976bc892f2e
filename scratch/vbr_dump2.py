import sys, numpy as np, zlib, ctypes
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import lame_b200
from conftest import make_signal
import os
if len(sys.argv) > 1: lame_b200._lib = lame_b200.load_library(sys.argv[1])
S, F = 2, 4
pcm = np.stack([make_signal("click", 8 * 1152, seed=61 + 4 * s) for s in range(S)])
enc = lame_b200.BatchEncoder(S, 44100, 2, 2, -1, -1, frames_per_launch=F, vbr=4)
L = lame_b200._lib if lame_b200._lib else lame_b200.load_library()
L.lamegpu_batch_debug_copy.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
enc.encode(pcm[:, :, :4 * 1152]); enc.encode(pcm[:, :, 4 * 1152:]); enc.flush()
for what, name, n in ((2, "psy", S * 2 * F * 2000), (3, "frm", S * F * 40), (4, "xr", S * 2 * F * 2 * 576 * 4), (1, "ana", S * 2 * F * 9432)):
    buf = np.zeros(n, dtype=np.uint8)
    got = L.lamegpu_batch_debug_copy(enc._h, what, buf.ctypes.data, buf.size)
    per = n // S
    for s in range(S):
        chunk = buf[s * per:(s + 1) * per]
        sub = per // (2 * F) if name != "frm" else per // F
        print(name, "s%d" % s, [("%08x" % zlib.crc32(chunk[i * sub:(i + 1) * sub].tobytes())) for i in range(per // sub)])
enc.close()
