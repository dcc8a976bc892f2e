"""CPU suite: lg_powf (csrc/lg_math.cuh), the device's restatement of glibc's powf, against the host's powf bit for bit.
The kernels call it where the reference calls libm at run time (athAdjust quantize_pvt.c:572, NS_INTERP psymodel.c:452); the
source is compiled for the host through the emulator shims (tests/c/powf_check.cpp).  Likewise lg_log10f against the host's log10f
(calc_scalefac vbrquantize.c:317, the quality-7 step guess of VBR-new) on every non-negative float (tests/c/log10f_check.cpp)."""
import os
import subprocess

from conftest import ROOT


def test_lg_powf_matches_host_powf(tmp_path):
    exe = str(tmp_path / "powf_check")
    subprocess.run(["g++", "-O2", "-fno-fast-math", "-ffp-contract=off", "-std=c++17", "-DLG_EMULATE", "-I" + os.path.join(ROOT, "tests/emu"),
                    "-I" + os.path.join(ROOT, "deprecated-lame-mirror_b200/csrc"), "-w", os.path.join(ROOT, "tests/c/powf_check.cpp"), "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe, "200"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "IDENTICAL" in r.stdout, r.stdout[-1000:]


def test_lg_log10f_matches_host_log10f_exhaustively(tmp_path):
    exe = str(tmp_path / "log10f_check")
    subprocess.run(["g++", "-O2", "-fno-fast-math", "-ffp-contract=off", "-std=c++17", "-DLG_EMULATE", "-I" + os.path.join(ROOT, "tests/emu"),
                    "-I" + os.path.join(ROOT, "deprecated-lame-mirror_b200/csrc"), "-w", os.path.join(ROOT, "tests/c/log10f_check.cpp"), "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=600)         # stride 1: all 2 139 095 041 floats in [+0, +inf]
    assert r.returncode == 0 and "IDENTICAL" in r.stdout, r.stdout[-1000:]


def test_lg_exp_and_lg_pow_match_host_libm(tmp_path):
    """the masking feedback of VBR-old (quantize.c:1419-1426) calls exp and pow in binary64 once per frame: lg_exp / lg_pow restate the
    host glibc's (x86-64 FMA build) operations; 4e7 arguments each over and beyond what the feedback can produce"""
    import platform
    if platform.machine() != "x86_64" or "fma" not in open("/proc/cpuinfo").read():
        import pytest
        pytest.skip("the restated build of glibc's exp/pow is the one x86-64 hosts with FMA run")
    exe = str(tmp_path / "exppow_check")
    subprocess.run(["g++", "-O2", "-fno-fast-math", "-ffp-contract=off", "-mfma", "-std=c++17", "-DLG_EMULATE", "-I" + os.path.join(ROOT, "tests/emu"),
                    "-I" + os.path.join(ROOT, "deprecated-lame-mirror_b200/csrc"), "-w", os.path.join(ROOT, "tests/c/exppow_check.cpp"), "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe, "40"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "IDENTICAL" in r.stdout, r.stdout[-1000:]
