"""Handles of the libmp3lame face are lanes of one shared batch engine: application threads that each drive
their own lame_t concurrently - the reference's threading contract, HACKING:67-76 - share GPU launches.  tests/c/handles_mt.cpp runs
T threads x (lame_init .. lame_encode_buffer in small calls .. lame_encode_flush .. lame_close) and compares every stream byte for
byte with the same calls made to the reference library."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libmp3lame_ref.so")


@pytest.fixture(scope="module")
def mt_emu(oracle_mod, tmp_path_factory):
    if not os.path.exists(REF_SO):
        pytest.skip("needs the reference build (oracle/_ref)")
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu")], check=True, capture_output=True)
    exe = str(tmp_path_factory.mktemp("bin") / "handles_mt_emu")
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", os.path.join(ROOT, "tests", "c", "handles_mt.cpp"), "-o", exe, "-L" + os.path.join(ROOT, "tests", "emu"),
                    "-llamegpu_emu", "-ldl", "-lpthread", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "tests", "emu")], check=True)
    return exe


@pytest.mark.parametrize("args,lanes,depth,crowd", [("6 4 1152 128", 8, 1, "0 1"), ("5 4 700 112 {ref} 3 4", 2, 1, "2 32"), ("6 7 1152 128", 8, 3, "2 32"),
                                                    ("4 6 500 112 {ref} 4 4", 2, 2, "1 1"), ("6 9 4000 160 {ref} 4 2", 8, 5, "2 2"), ("5 8 1152 128", 8, 5, "1 3")])
def test_threads_with_own_handles_share_an_engine_emulated(mt_emu, args, lanes, depth, crowd):
    """more threads than lanes too: a second engine of the same configuration is made for the overflow; depth > 1 = calls return with
    up to depth - 1 frames of the lane still outstanding (LAMEGPU_HANDLE_DEPTH, default 5); crowd "c n" = a launch that gathers n lanes
    or more takes at most c frames of each (LAMEGPU_HANDLE_CROWD_CAP / _LANES, default 2 / 32) - the stream stays the same"""
    a = args.format(ref=REF_SO).split()
    if len(a) == 4:
        a.append(REF_SO)
    r = subprocess.run([mt_emu] + a, capture_output=True, text=True, cwd=ROOT, timeout=900,
                       env=dict(os.environ, LAMEGPU_LANES=str(lanes), LAMEGPU_HANDLE_DEPTH=str(depth), LAMEGPU_HANDLE_CROWD_CAP=crowd.split()[0],
                                LAMEGPU_HANDLE_CROWD_LANES=crowd.split()[1]))
    assert r.returncode == 0 and "IDENTICAL" in r.stdout, r.stdout[-1000:] + r.stderr[-1000:]


@pytest.fixture(scope="module")
def edges_emu(oracle_mod, tmp_path_factory):
    if not os.path.exists(REF_SO):
        pytest.skip("needs the reference build (oracle/_ref)")
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu")], check=True, capture_output=True)
    exe = str(tmp_path_factory.mktemp("bin") / "handle_edges_emu")
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", os.path.join(ROOT, "tests", "c", "handle_edges.cpp"), "-o", exe, "-L" + os.path.join(ROOT, "tests", "emu"),
                    "-llamegpu_emu", "-ldl", "-lpthread", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "tests", "emu")], check=True)
    return exe


@pytest.mark.parametrize("chunk,depth,frames", [(1152, 5, 8), (700, 5, 1)])
def test_calls_that_end_or_read_a_lane_with_frames_outstanding_emulated(edges_emu, chunk, depth, frames):
    """tests/c/handle_edges.cpp: encoding on after lame_encode_flush, lame_encode_flush_nogap + lame_init_bitstream in mid-stream
    (lame.c:1988, :2006), lame_close without a flush and the lane's next owner, lame_get_mf_samples_to_encode after every call, a flush right behind calls that
    follows others on a resampled stream (with one frame per launch - frames = 1 - two
    frames of the lane always wait for their launch when the flush begins: the case pad_for_flush once got wrong) - each sequence made to the product (emulated kernels) and to the reference library, bytes and values compared"""
    r = subprocess.run([edges_emu, REF_SO, str(chunk)], capture_output=True, text=True, cwd=ROOT, timeout=900,
                       env=dict(os.environ, LAMEGPU_LANES="2", LAMEGPU_HANDLE_DEPTH=str(depth), LAMEGPU_HANDLE_FRAMES=str(frames)))
    assert r.returncode == 0 and "IDENTICAL 5/5" in r.stdout, r.stdout[-1000:] + r.stderr[-1000:]


@pytest.mark.gpu
@pytest.mark.parametrize("chunk,frames", [(1152, 8), (700, 1)])
def test_calls_that_end_or_read_a_lane_with_frames_outstanding_on_the_gpu(chunk, frames):
    exe = os.path.join(ROOT, "tests", "c", "bin", "handle_edges")
    if not (os.path.exists(exe) and os.path.exists(REF_SO)):
        pytest.skip("tests/c/bin/handle_edges and oracle/_ref travel with the repository snapshot; not built here")
    r = subprocess.run([exe, REF_SO, str(chunk)], capture_output=True, text=True, cwd=ROOT, timeout=300,
                       env=dict(os.environ, LAMEGPU_LANES="8", LAMEGPU_HANDLE_FRAMES=str(frames)))
    assert r.returncode == 0 and "IDENTICAL 5/5" in r.stdout, r.stdout[-1000:] + r.stderr[-1000:]


@pytest.mark.gpu
@pytest.mark.parametrize("threads,frames,chunk,floor", [(512, 96, 1152, 8.0e4), (512, 96, 2304, 1.2e5), (64, 24, 4608, 0.0), (300, 16, 1000, 0.0)])
def test_512_threads_with_own_handles_on_the_gpu(threads, frames, chunk, floor):
    """512 threads x own lame_t x 1152-sample lame_encode_buffer calls: byte-identical per stream, and - the point of sharing the engine -
    an aggregate rate that one engine per handle cannot reach (round 1: 1.6e3 frames/s per handle = 8e5 only if 512 GPUs' worth of
    engines could run side by side; measured on a 16-core box over 256 frames per stream: 2.1e5 frames/s when every call waits for its
    own frames - a device round trip and a wake-up per call - and 3.8e5 with up to four frames of a lane outstanding and crowded
    launches capped at two frames per lane, the defaults; profiles/r2_handles.txt)"""
    exe = os.path.join(ROOT, "tests", "c", "bin", "handles_mt")
    if not (os.path.exists(exe) and os.path.exists(REF_SO)):
        pytest.skip("tests/c/bin/handles_mt and oracle/_ref travel with the repository snapshot; not built here")
    r = subprocess.run([exe, str(threads), str(frames), str(chunk), "128", REF_SO], capture_output=True, text=True, cwd=ROOT, timeout=900,
                       env=dict(os.environ, LAMEGPU_LANES="512"))
    assert r.returncode == 0 and "IDENTICAL" in r.stdout, r.stdout[-1000:] + r.stderr[-1000:]
    rate = float(re.search(r"= (\d+) frames/s", r.stdout).group(1))
    print(r.stdout.strip())
    assert rate >= floor, "aggregate %.0f frames/s" % rate


@pytest.mark.gpu
def test_long_streams_stay_identical():
    """24 streams x 300 frames through the handles (one frame per call), among them the white-noise stream whose frame 170 was the first
    to show the one-ulp error of a C++-computed fast_log2 table (handles_mt stream 495): a perceptual entropy one ulp off moves a
    bit target by one bit nine frames later.  Long streams find what 8-frame tests cannot."""
    exe = os.path.join(ROOT, "tests", "c", "bin", "handles_mt")
    if not (os.path.exists(exe) and os.path.exists(REF_SO)):
        pytest.skip("tests/c/bin/handles_mt and oracle/_ref travel with the repository snapshot; not built here")
    r = subprocess.run([exe, "24", "300", "1152", "128", REF_SO, "0", "4", "480"], capture_output=True, text=True, cwd=ROOT, timeout=900,
                       env=dict(os.environ, LAMEGPU_LANES="512"))
    assert r.returncode == 0 and "IDENTICAL" in r.stdout, r.stdout[-1000:] + r.stderr[-1000:]
