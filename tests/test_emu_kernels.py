"""CPU suite: the kernel SOURCES (the same .cu/.cuh files nvcc compiles) are built for the test-only SIMT emulator
(tests/emu) and run against the oracle port, so the kernels' logic is checked without a GPU.  This is a check of the
sources, not a product path - the emulator library is never loaded by the package."""
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def emu_bin(oracle_mod, tmp_path_factory):
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu")], check=True, capture_output=True)
    out = str(tmp_path_factory.mktemp("bin") / "emu_vs_port")
    subprocess.run(["gcc", "-O2", "-w", os.path.join(ROOT, "tests/c/gpu_vs_port.c"), "-o", out, "-L" + os.path.join(ROOT, "tests/emu"),
                    "-llamegpu_emu", "-L" + os.path.join(ROOT, "oracle"), "-llameport", "-lm",
                    "-Wl,-rpath," + os.path.join(ROOT, "tests/emu"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    return out


# nstreams frames_per_stream frames_per_launch brate mode quality samplerate chunk
@pytest.mark.parametrize("args", ["4 24 8 128 -1 -1 44100 1152", "3 20 5 320 1 -1 44100 3000", "3 16 16 192 0 5 48000 777",
                                  "2 12 4 160 -1 7 32000 1152"])
def test_emulated_kernels_match_port(emu_bin, args):
    r = subprocess.run([emu_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "IDENTICAL" in r.stdout
