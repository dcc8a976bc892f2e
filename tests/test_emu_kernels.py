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
                                  "2 12 4 160 -1 7 32000 1152", "3 10 4 128 2 -1 44100 1152"])
def test_emulated_kernels_match_port(emu_bin, args):
    r = subprocess.run([emu_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "IDENTICAL" in r.stdout


@pytest.mark.parametrize("args", ["4 20 8 128 -1 -1 44100 1152", "3 16 5 192 1 -1 44100 3000", "2 12 4 256 -1 7 32000 1152"])
def test_emulated_kernels_abr_match_port(emu_bin, args):
    """ABR: per-frame bitrate index chosen on the device, variable frame sizes through kernel E and the host splice"""
    r = subprocess.run([emu_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, timeout=600, env=dict(os.environ, LP_VBR="3"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "IDENTICAL" in r.stdout


@pytest.mark.parametrize("args", ["4 12 4 2 -1 -1 44100 1152", "4 16 8 0 -1 -1 44100 1152", "3 16 16 4 0 5 48000 777", "2 12 4 5 3 -1 44100 1152",
                                  # quality 7-9: guess_scalefac_x34 (vbrquantize.c:324) through lg_log10f
                                  "4 12 4 2 -1 7 44100 1152", "3 12 8 4 0 8 48000 777", "4 10 4 0 -1 9 44100 1152", "3 10 4 4 -1 7 22050 1152"])
def test_emulated_kernels_vbr_match_port(emu_bin, args):
    """VBR-new (vbr_mtrh; the 4th argument is VBR_q): lg_kernel_vbr incl. the out-of-bits path (click streams at -V0)"""
    r = subprocess.run([emu_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, timeout=900, env=dict(os.environ, LP_VBR="4"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "IDENTICAL" in r.stdout


@pytest.mark.parametrize("args,env", [
    ("3 10 4 128 -1 -1 48000 777", dict(LP_OUT_SR="44100")), ("2 8 8 192 -1 -1 44100 3000", dict(LP_OUT_SR="48000")),
    ("2 8 3 96 -1 -1 44100 100", {}), ("2 4 4 128 -1 -1 8000 50", dict(LP_OUT_SR="44100")), ("2 16 4 128 -1 -1 96000 1152", {}),
    ("2 8 4 2 -1 -1 44100 1152", dict(LP_VBR="4", LP_OUT_SR="32000")), ("2 4 4 128 -1 -1 48000 7", dict(LP_OUT_SR="44100")),
    # -V7 at 44.1 kHz (32 kHz output, quality 5.63) and a level between the presets
    ("3 8 4 7 -1 -1 44100 1152", dict(LP_VBR="4")), ("3 8 4 5 -1 -1 48000 1152", dict(LP_VBR="4", LP_VBRQ_FRAC="0.3")),
])
def test_emulated_resampler_matches_port(emu_bin, args, env):
    """kernel R (lg_kernel_resample) + the host's replay of the reference's per-call resampler bookkeeping (chunks, input
    clock, flush bunches), incl. calls of 7 samples (chunk list growth) and 8 kHz -> 44.1 kHz upsampling"""
    r = subprocess.run([emu_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, timeout=900, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "IDENTICAL" in r.stdout


@pytest.mark.parametrize("args,env", [
    ("4 12 4 128 -1 2 44100 1152", {}), ("3 12 4 128 -1 1 44100 1152", {}), ("3 12 8 320 1 0 44100 3000", {}), ("3 12 4 192 0 1 48000 777", {}),
    ("3 12 8 128 -1 0 44100 1152", dict(LP_VBR="3")),
])
def test_emulated_quality_0_to_2_matches_port(emu_bin, args, env):
    """kernel D with substep shaping (half-step flags as a bit mask, the post-quantisation drop in lg_substep_zero), one-band
    amplification and the full outer loop"""
    r = subprocess.run([emu_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, timeout=900, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "IDENTICAL" in r.stdout


@pytest.mark.parametrize("args,env", [
    ("3 10 4 64 -1 -1 22050 1152", {}), ("3 12 4 32 -1 -1 16000 777", {}), ("3 12 4 24 -1 -1 8000 1152", {}), ("3 12 4 32 3 -1 11025 1152", {}),
    ("3 12 4 64 -1 -1 44100 1152", {}), ("3 12 4 80 1 2 22050 1152", {}), ("3 12 4 64 -1 -1 22050 1152", dict(LP_VBR="3")),
    ("3 10 4 4 -1 -1 22050 1152", dict(LP_VBR="4")), ("3 10 4 9 -1 -1 44100 777", dict(LP_VBR="4")), ("3 10 4 5 -1 -1 8000 1152", dict(LP_VBR="4")),
])
def test_emulated_mpeg2_matches_port(emu_bin, args, env):
    """MPEG-2 / MPEG-2.5 output through all kernels: one granule per frame in B, D, D', E; LSF scalefactor partitions
    (lg_scale_bitcount_lsf), LSF side info and scalefactor packing in kernel E, the 8 kHz band limits; CBR, ABR, VBR"""
    r = subprocess.run([emu_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, timeout=900, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "IDENTICAL" in r.stdout


@pytest.mark.parametrize("args,env", [
    ("3 8 4 2 -1 -1 44100 1152", {}), ("3 10 4 0 -1 -1 44100 1152", {}), ("3 10 8 4 0 -1 48000 777", {}), ("3 10 4 4 -1 -1 22050 1152", {}),
    ("3 10 4 2 -1 0 44100 1152", {}), ("3 10 4 2 -1 -1 44100 1152", dict(LP_VBRQ_FRAC="0.5")),
])
def test_emulated_vbr_old_matches_port(emu_bin, args, env):
    """VBR-old (vbr_rh): lg_kernel_vbrold - bisection of the bit budget over kernel D's outer_loop with xrpow kept between runs, the
    sfb21 analog-silence cut, bit-pressure rounds, the entropy-driven masking feedback in the scan kernel"""
    r = subprocess.run([emu_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, timeout=1200, env=dict(os.environ, LP_VBR="2", **env))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "IDENTICAL" in r.stdout


TAG_SCRIPT = r"""
import json, os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import lame_b200
from conftest import make_signal
lame_b200._lib = lame_b200.load_library(os.path.join({root!r}, "tests", "emu", "liblamegpu_emu.so"))   # test-only build of the same sources
gold = os.path.join({root!r}, "tests", "golden")
for name, m in sorted(json.load(open(os.path.join(gold, "manifest_tag.json"))).items()):
    x = make_signal(m["signal"], m["frames"] * 1152)
    e = lame_b200.Encoder(m["samplerate"], 2, m["brate"], m["mode"] if m["mode"] >= 0 else lame_b200.NOT_SET, m["quality"], write_tag=True,
                          vbr=m.get("vbr", 0))
    mp3 = b""
    for pos in range(0, x.shape[1], 4000):
        mp3 += e.encode(x[0, pos:pos + 4000], x[1, pos:pos + 4000])
    mp3 += e.flush()
    tag = e.lametag_frame()
    e.close()
    assert mp3 == open(os.path.join(gold, name + ".mp3"), "rb").read(), name
    assert tag == open(os.path.join(gold, name + ".tagframe"), "rb").read(), name
print("TAG IDENTICAL")
"""


def test_emulated_info_tag_matches_reference_golden(emu_bin):
    """lame_set_bWriteVbrTag(1) (the reference's default): stream with the placeholder frame + lame_get_lametag_frame,
    byte for byte against vectors generated from the unmodified reference (tests/golden/make_golden.py)"""
    import sys
    r = subprocess.run([sys.executable, "-c", TAG_SCRIPT.format(root=ROOT)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "TAG IDENTICAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_emulated_header_bits_and_tag_file(emu_bin, oracle_mod):
    """lame_set_copyright / _original / _emphasis / _extension reach the frame headers and the Info tag; lame_mp3_tags_fid"""
    import sys
    if not oracle_mod.have_ref():
        pytest.skip("needs the reference build (oracle/_ref)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "header_bits_check.py"), os.path.join(ROOT, "tests", "emu", "liblamegpu_emu.so")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "HEADER BITS IDENTICAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_emulated_sample_type_entry_points(emu_bin, oracle_mod):
    """lame_encode_buffer_float/_ieee_float/_ieee_double/_int/_long/_long2/_interleaved* against libmp3lame itself"""
    import sys
    if not oracle_mod.have_ref():
        pytest.skip("needs the reference build (oracle/_ref)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sample_types_check.py"), os.path.join(ROOT, "tests", "emu", "liblamegpu_emu.so")],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, SAMPLE_TYPES_QUICK="1"))
    assert r.returncode == 0 and "SAMPLE TYPES IDENTICAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("args,env", [("5 12 4 128 -1 -1 44100 1152", dict(LAMEGPU_EMU_DEVICES="3")),
                                      ("4 10 4 2 -1 -1 44100 777", dict(LAMEGPU_EMU_DEVICES="2", LP_VBR="4", LP_PIPELINED="1"))])
def test_batch_spanning_several_devices(emu_bin, args, env):
    """lamegpu_batch_open(device = -1): one engine and one host thread per device, contiguous shares of the streams (here: emulated devices)"""
    r = subprocess.run([emu_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, timeout=900, env=dict(os.environ, LP_DEVICE="-1", **env))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "IDENTICAL" in r.stdout and "batch spans %s device(s)" % env["LAMEGPU_EMU_DEVICES"] in r.stdout


@pytest.mark.parametrize("args", ["4 24 8 128 -1 -1 44100 1152", "3 16 3 192 0 5 48000 777"])
def test_pipelined_batch_matches_port(emu_bin, args):
    """lamegpu_batch_set_pipelined: the newest step stays in flight between calls, its bytes come out of the next call or the flush"""
    r = subprocess.run([emu_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, timeout=900, env=dict(os.environ, LP_PIPELINED="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "IDENTICAL" in r.stdout


def test_fast_log2_table_is_the_c_one(emu_bin):
    """util.c:962 init_log_table computes log(1.0f + j/512.f) / log(2.0f) in C: the float argument goes to the double log().  The host
    setup is C++, where log(float) is the float overload - 141 of the 513 entries one ulp off, which showed as one stream in 512
    differing from the reference after 170 frames (white noise, handles_mt stream 495).  The table of the configuration must be C's."""
    import ctypes
    import sys
    import numpy as np
    code = r'''
import sys, ctypes, numpy as np
sys.path.insert(0, %r)
import lame_b200
lame_b200._lib = lame_b200.load_library(%r)
enc = lame_b200.BatchEncoder(1, 44100, 2, 128, -1, -1, frames_per_launch=1)
t = np.zeros(513, dtype=np.float32)
n = lame_b200._lib.lamegpu_batch_debug_copy(enc._h, 8, t.ctypes.data, t.nbytes)
assert n == t.nbytes, n
j = np.arange(513, dtype=np.float32)
want = (np.log((np.float32(1.0) + j / np.float32(512)).astype(np.float64)) / np.log(np.float64(np.float32(2.0)))).astype(np.float32)
bad = int((t != want).sum())
print("LOG TABLE", "IDENTICAL" if bad == 0 else "DIFFERENT %%d" %% bad)
''' % (ROOT, os.path.join(ROOT, "tests", "emu", "liblamegpu_emu.so"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "LOG TABLE IDENTICAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
