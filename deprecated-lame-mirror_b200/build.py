"""Builds liblamegpu.so (the C-ABI product library) in-tree.

    nvcc  -gencode arch=compute_100a,code=sm_100a  lg_engine.cu        (kernels + device engine)
    g++   -fno-fast-math -ffp-contract=off         lg_setup/lg_bitstream/lg_api.cpp  (host side)

-fmad=false / -ffp-contract=off are REQUIRED: the encoder must reproduce the reference's IEEE-754
operation sequence bit for bit (hard part 2).  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblamegpu.so")
OBJ = os.path.join(HERE, "build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-fno-fast-math,-ffp-contract=off,-fno-strict-aliasing"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fno-fast-math", "-ffp-contract=off", "-fno-strict-aliasing", "-Wall"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))


def build_library(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = sources() + [os.path.abspath(__file__)]
    if not force and _newer(OUT, srcs):
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    objs = []

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)

    o = os.path.join(OBJ, "lg_engine.o")
    run([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, "lg_engine.cu"), "-o", o])
    objs.append(o)
    for name in ("lg_setup", "lg_bitstream", "lg_api", "lg_api_stubs"):
        o = os.path.join(OBJ, name + ".o")
        run(["g++"] + CXX_FLAGS + ["-c", os.path.join(CSRC, name + ".cpp"), "-o", o])
        objs.append(o)
    run([nvcc, "-shared", "-o", OUT] + objs + ["-lpthread"])
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
