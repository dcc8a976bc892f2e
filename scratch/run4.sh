timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import lame_b200
S, F = 512, 8
rng = np.random.default_rng(1000)
pcm = rng.integers(-12000, 12001, size=(S, 2, F * 1152 + 224), dtype=np.int16)
for vbr, br, name in ((0, 128, "CBR128"), (3, 128, "ABR128"), (4, 2, "VBR-V2"), (4, 0, "VBR-V0")):
    enc = lame_b200.BatchEncoder(S, 44100, 2, br, -1, -1, frames_per_launch=F, vbr=vbr)
    enc.stage(pcm, F)
    for _ in range(3): enc.rerun_device(F)
    k = np.zeros(5)
    for _ in range(10):
        enc.rerun_device(F); k += np.array(enc.kernel_ms())
    k /= 10
    print("%-7s S=%d F=%d  A %.3f B %.3f C %.3f D %.3f E %.3f ms total %.3f -> %.0f frames/s" % (name, S, F, *k, k.sum(), S * F / (k.sum() * 1e-3)), flush=True)
    enc.close()
PY
