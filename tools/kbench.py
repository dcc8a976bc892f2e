"""Time the four kernels of one library variant and hash its MP3 output (developer tool).
usage: python tools/kbench.py <lib.so> [S F reps]"""
import hashlib, sys, os, time, faulthandler
import numpy as np
faulthandler.dump_traceback_later(300, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lame_b200
lib = sys.argv[1]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
F = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
lame_b200._lib = lame_b200.load_library(os.path.abspath(lib))
rng = np.random.default_rng(1000)
pcm = rng.integers(-12000, 12001, size=(S, 2, F * 1152 + 224), dtype=np.int16)
enc = lame_b200.BatchEncoder(S, 44100, 2, 128, -1, -1, frames_per_launch=F)
enc.stage(rng.integers(-12000, 12001, size=(S, 2, 2 * F * 1152), dtype=np.int16), F)
for _ in range(3):
    enc.rerun_device(F)
k = np.zeros(5)
for _ in range(reps):
    enc.rerun_device(F)
    k += np.array(enc.kernel_ms())
k /= reps
enc.run_device_steps(F, 5)
pipe_ms = enc.run_device_steps(F, 4 * reps)       # persistent streams, steps back to back: A-B-C of step i+1 under D of step i
pk = enc.kernel_ms()
enc.close()
enc = lame_b200.BatchEncoder(S, 44100, 2, 128, -1, -1, frames_per_launch=F)
out = np.empty((S, int(1.25 * F * 1152) + 7200 + 4096), dtype=np.uint8)
nb = np.zeros(S, dtype=np.int32)
h = hashlib.sha1()
t0 = time.perf_counter()
for i in range(3):
    enc.encode_raw(pcm[:, :, :F * 1152], out, nb)
    for s in range(0, S, max(1, S // 64)):
        h.update(out[s, :nb[s]].tobytes())
e2e = (time.perf_counter() - t0) / 3
enc.close()
print("%-28s S=%d F=%d  A %.3f B %.3f C %.3f D %.3f E %.3f ms  total %.3f  | pipelined %.3f ms/step (A %.2f D %.2f) -> %.0f frames/s | e2e %.1f ms  sha %s" % (
    os.path.basename(lib), S, F, k[0], k[1], k[2], k[3], k[4], k.sum(), pipe_ms, pk[0], pk[3], S * F / (pipe_ms * 1e-3), e2e * 1e3, h.hexdigest()[:12]), flush=True)
