"""Build experiment variants of liblamegpu.so (extra -D flags for the kernels) into scratch/variants/.
usage: python tools/kvariants.py tag1:"-DLG_X=1 -DLG_Y=2" tag2:"..."    (developer tool, not part of the product)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deprecated-lame-mirror_b200")
sys.path.insert(0, PKG)
import build as B
out = os.path.join(ROOT, "scratch", "variants")
os.makedirs(out, exist_ok=True)
B.build_library()
for spec in sys.argv[1:]:
    tag, _, flags = spec.partition(":")
    o = os.path.join(out, "eng_%s.o" % tag)
    subprocess.run(["nvcc"] + B.NVCC_FLAGS + flags.split() + ["-c", os.path.join(B.CSRC, "lg_engine.cu"), "-o", o], check=True)
    objs = [o] + [os.path.join(B.OBJ, n + ".o") for n in ("lg_setup", "lg_bitstream", "lg_api", "lg_api_stubs")]
    so = os.path.join(out, "lib_%s.so" % tag)
    subprocess.run(["nvcc", "-shared", "-o", so] + objs + ["-lpthread"], check=True, stderr=subprocess.DEVNULL)
    os.remove(o)
    print("built", so)
