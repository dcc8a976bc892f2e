"""Every lame_encode_buffer_* sample type (lame.h:715-838) of liblamegpu against the unmodified reference (oracle/_ref):
run as  python tests/sample_types_check.py <path to liblamegpu(.so|_emu.so)>.  Prints SAMPLE TYPES IDENTICAL on success."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lame_b200  # noqa: E402
import oracle  # noqa: E402
from conftest import make_signal  # noqa: E402


def run(lib, fn, arrs, n, interleaved):
    lib.lame_init.restype = ctypes.c_void_p
    h = ctypes.c_void_p(lib.lame_init())
    lib.lame_set_bWriteVbrTag(h, 0)
    lib.lame_set_brate(h, 128)
    assert lib.lame_init_params(h) == 0
    buf = np.empty(200000, dtype=np.uint8)
    f = getattr(lib, fn)
    f.restype = ctypes.c_int
    p = [ctypes.c_void_p(a.ctypes.data) for a in arrs]
    rc = f(h, *p, ctypes.c_int(n), ctypes.c_void_p(buf.ctypes.data), ctypes.c_int(buf.size))
    assert rc >= 0, (fn, rc)
    out = buf[:rc].tobytes()
    rc = lib.lame_encode_flush(h, ctypes.c_void_p(buf.ctypes.data), ctypes.c_int(buf.size))
    out += buf[:rc].tobytes()
    lib.lame_close(h)
    return out


def main(path):
    L = lame_b200._lib = lame_b200.load_library(os.path.abspath(path))
    R = ctypes.CDLL(oracle.REF_SO)
    x = make_signal("click", 20 * 1152, seed=9)
    n = x.shape[1]
    cases = {
        "lame_encode_buffer_float": ([x[0].astype(np.float32) * 0.7, x[1].astype(np.float32) * 0.7], False),
        "lame_encode_buffer_ieee_float": ([(x[0] / 32768.0).astype(np.float32), (x[1] / 32768.0).astype(np.float32)], False),
        "lame_encode_buffer_interleaved_ieee_float": ([np.ascontiguousarray((x.T / 32768.0).astype(np.float32))], True),
        "lame_encode_buffer_ieee_double": ([x[0] / 32768.0 * 0.9, x[1] / 32768.0 * 0.9], False),
        "lame_encode_buffer_interleaved_ieee_double": ([np.ascontiguousarray(x.T / 32768.0)], True),
        "lame_encode_buffer_int": ([x[0].astype(np.int32) * 65536 + 1234, x[1].astype(np.int32) * 65536 - 77], False),
        "lame_encode_buffer_long": ([x[0].astype(np.int64), x[1].astype(np.int64)], False),
        "lame_encode_buffer_long2": ([x[0].astype(np.int64) * (1 << 48) + 12345, x[1].astype(np.int64) * (1 << 48)], False),
        "lame_encode_buffer_interleaved": ([np.ascontiguousarray(x.T)], True),
    }
    for fn, (arrs, inter) in cases.items():
        a, b = run(L, fn, arrs, n, inter), run(R, fn, arrs, n, inter)
        assert a == b, fn
        print(fn, len(a), "identical")
    print("SAMPLE TYPES IDENTICAL")


if __name__ == "__main__":
    main(sys.argv[1])
