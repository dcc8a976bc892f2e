import sys, numpy as np, zlib, ctypes
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import lame_b200
from conftest import make_signal
if len(sys.argv) > 1: lame_b200._lib = lame_b200.load_library(sys.argv[1])
S, F = 2, 4
pcm = np.stack([make_signal("click", 8 * 1152, seed=61 + 4 * s) for s in range(S)])
enc = lame_b200.BatchEncoder(S, 44100, 2, 2, -1, -1, frames_per_launch=F, vbr=4)
L = lame_b200._lib if lame_b200._lib else lame_b200.load_library()
gsz = L.lamegpu_sizeof_granule_out()
for launch in range(3):
    if launch < 2: enc.encode(pcm[:, :, launch * 4 * 1152:(launch + 1) * 4 * 1152])
    else: enc.flush()
    buf = np.zeros(S * 2 * F * 2 * gsz, dtype=np.uint8)
    L.lamegpu_batch_debug_copy.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    L.lamegpu_batch_debug_copy(enc._h, 5, buf.ctypes.data, buf.size)
    g = buf.reshape(S, 2 * F, 2, gsz)
    for s in range(S):
        for gb in range(2 * F):
            for ch in range(2):
                r = g[s, gb, ch]
                ix = r[:1152].tobytes(); sf = r[1152:1192].tobytes()
                side = r[1192:1192 + 8].view(np.int16)
                rest = r[1200:1216]
                if launch == 2 and s == 1 and gb < 2: print("   sf", list(r[1152:1192].view(np.int8)), "ix[:40]", list(r[:80].view(np.int16)))
                print("L%d s%d gb%d ch%d ix %08x sf %08x p23 %d p2 %d bv %d c1 %d gg %d sfc %d bt %d ts %d %d %d sbg %d %d %d r0 %d r1 %d pre %d sfs %d c1t %d" % (
                    launch, s, gb, ch, zlib.crc32(ix), zlib.crc32(sf), side[0], side[1], side[2], side[3], rest[0], rest[1], rest[2], rest[4], rest[5], rest[6], rest[7], rest[8], rest[9], rest[10], rest[11], rest[12], rest[13], rest[14]))
enc.close()
