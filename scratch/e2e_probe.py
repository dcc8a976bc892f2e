import sys, time, faulthandler, numpy as np
faulthandler.dump_traceback_later(60, exit=True)
sys.path.insert(0, "/root/repo")
import lame_b200
mode = sys.argv[1]
if mode == "torch":
    import torch
    torch.cuda.set_device(0)
    fb = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda"); fb.zero_(); torch.cuda.synchronize()
S, F = 512, 8
rng = np.random.default_rng(1000)
pcm = rng.integers(-12000, 12001, size=(S, 2, F*1152+224), dtype=np.int16)
if mode != "fresh":
    enc = lame_b200.BatchEncoder(S, frames_per_launch=F)
    enc.stage(pcm, F); enc.rerun_device(F); print("first engine ok", enc.kernel_ms(), flush=True)
    enc.close()
enc = lame_b200.BatchEncoder(S, frames_per_launch=F)
out = np.empty((S, 30000), dtype=np.uint8); nb = np.zeros(S, dtype=np.int32)
t = time.time(); print("enc0", enc.encode_raw(pcm, out, nb), time.time()-t, flush=True)
for i in range(3):
    p = rng.integers(-12000, 12001, size=(S, 2, F*1152), dtype=np.int16)
    t = time.time(); print("enc", enc.encode_raw(p, out, nb), time.time()-t, nb[:3], flush=True)
enc.close()
