"""The header bits libmp3lame lets the caller set (lame_set_copyright / _original / _emphasis / _extension, lame.h:300-330) through
liblamegpu against the unmodified reference (oracle/_ref), stream and Info tag, plus lame_mp3_tags_fid on a real file:
run as  python tests/header_bits_check.py <path to liblamegpu(.so|_emu.so)>.  Prints HEADER BITS IDENTICAL on success."""
import ctypes
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from conftest import make_signal  # noqa: E402


def run(lib, x, sr=44100, brate=128, crc=0, vbr=0):
    lib.lame_init.restype = ctypes.c_void_p
    h = ctypes.c_void_p(lib.lame_init())
    for f, v in (("lame_set_in_samplerate", sr), ("lame_set_num_channels", 2), ("lame_set_brate", brate), ("lame_set_copyright", 1),
                 ("lame_set_original", 0), ("lame_set_emphasis", 1), ("lame_set_extension", 1), ("lame_set_bWriteVbrTag", 1),
                 ("lame_set_error_protection", crc), ("lame_set_VBR", vbr), ("lame_set_VBR_q", 3)):
        fn = getattr(lib, f)
        fn.argtypes, fn.restype = [ctypes.c_void_p, ctypes.c_int], ctypes.c_int
        assert fn(h, v) == 0, f
    lib.lame_init_params.argtypes = [ctypes.c_void_p]
    assert lib.lame_init_params(h) == 0
    buf = np.empty(400000, dtype=np.uint8)
    lib.lame_encode_buffer.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    left, right = np.ascontiguousarray(x[0]), np.ascontiguousarray(x[1])
    n = lib.lame_encode_buffer(h, left.ctypes.data, right.ctypes.data, len(left), buf.ctypes.data, buf.size)
    out = buf[:n].tobytes()
    lib.lame_encode_flush.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    n = lib.lame_encode_flush(h, buf.ctypes.data, buf.size)
    out += buf[:n].tobytes()
    lib.lame_get_lametag_frame.restype = ctypes.c_size_t
    lib.lame_get_lametag_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    n = lib.lame_get_lametag_frame(h, buf.ctypes.data, buf.size)
    return h, out, buf[:n].tobytes()


def main():
    ours = ctypes.CDLL(sys.argv[1])
    ref = ctypes.CDLL(oracle.REF_SO)
    x = make_signal("click", 12 * 1152, seed=3)
    # error_protection (lame.h:312): CRC-16 behind every header, MPEG-1 and MPEG-2, CBR and VBR
    for kw in (dict(crc=1), dict(crc=1, sr=22050, brate=64), dict(crc=1, vbr=4), dict(crc=1, vbr=2, sr=48000)):
        _, a, ta = run(ours, x, **kw)
        _, b, tb = run(ref, x, **kw)
        assert a == b, "stream differs with %r" % kw
        assert ta == tb, "tag frame differs with %r" % kw
    h, a, ta = run(ours, x)
    _, b, tb = run(ref, x)
    assert a == b, "stream differs"
    assert ta == tb, "tag frame differs"
    # lame_mp3_tags_fid (lame.h:950): rewrites the placeholder at the start of the file
    libc = ctypes.CDLL(None)
    libc.fopen.restype = ctypes.c_void_p
    libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    libc.fclose.argtypes = [ctypes.c_void_p]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "t.mp3")
        open(path, "wb").write(a)
        f = libc.fopen(path.encode(), b"r+b")
        ours.lame_mp3_tags_fid.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        ours.lame_mp3_tags_fid(h, f)
        libc.fclose(f)
        got = open(path, "rb").read()
    assert got == ta + a[len(ta):], "lame_mp3_tags_fid"
    print("HEADER BITS IDENTICAL")


if __name__ == "__main__":
    main()
