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


def run_calls(lib, calls, in_rate=0, out_rate=0):
    """one handle, a list of (function name, arrays, nsamples) calls, then flush"""
    lib.lame_init.restype = ctypes.c_void_p
    h = ctypes.c_void_p(lib.lame_init())
    lib.lame_set_bWriteVbrTag(h, 0)
    lib.lame_set_brate(h, 128)
    if in_rate:
        lib.lame_set_in_samplerate(h, in_rate)
        lib.lame_set_out_samplerate(h, out_rate)
    assert lib.lame_init_params(h) == 0
    buf = np.empty(200000, dtype=np.uint8)
    out = b""
    for fn, arrs, n in calls:
        f = getattr(lib, fn)
        f.restype = ctypes.c_int
        p = [ctypes.c_void_p(a.ctypes.data) for a in arrs]
        rc = f(h, *p, ctypes.c_int(n), ctypes.c_void_p(buf.ctypes.data), ctypes.c_int(buf.size))
        assert rc >= 0, (fn, rc)
        out += buf[:rc].tobytes()
    rc = lib.lame_encode_flush(h, ctypes.c_void_p(buf.ctypes.data), ctypes.c_int(buf.size))
    out += buf[:rc].tobytes()
    lib.lame_close(h)
    return out


def chunked(fn, arrs, n, step, interleaved=False):
    calls = []
    for i in range(0, n, step):
        k = min(step, n - i)
        calls.append((fn, [np.ascontiguousarray(a[2 * i:2 * (i + k)] if interleaved else a[i:i + k]) for a in arrs], k))
    return calls


def main(path):
    L = lame_b200._lib = lame_b200.load_library(os.path.abspath(path))
    R = ctypes.CDLL(oracle.REF_SO)
    quick = os.environ.get("SAMPLE_TYPES_QUICK") == "1"      # the CPU suite's run over the SIMT emulator: shorter signal, fewer repeats
    x = make_signal("click", (7 if quick else 20) * 1152, seed=9)
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
    # the same types in many calls of odd sizes (the stream keeps the caller's type from call to call, the device converts)
    f32 = [(x[0] / 32768.0).astype(np.float32), (x[1] / 32768.0).astype(np.float32)]
    f64 = [x[0] / 32768.0 * 0.9, x[1] / 32768.0 * 0.9]
    i32 = [x[0].astype(np.int32) * 65536 + 1234, x[1].astype(np.int32) * 65536 - 77]
    i16 = [np.ascontiguousarray(x[0]), np.ascontiguousarray(x[1])]
    per_type = {"ieee_float x 1000": chunked("lame_encode_buffer_ieee_float", f32, n, 1000),
                "ieee_double x 777": chunked("lame_encode_buffer_ieee_double", f64, n, 777),
                "int x 1152": chunked("lame_encode_buffer_int", i32, n, 1152),
                "int16 x 1152": chunked("lame_encode_buffer", i16, n, 1152)}
    for name, calls in per_type.items():
        a, b = run_calls(L, calls), run_calls(R, calls)
        assert a == b, name
        print(name, len(a), "identical")
    # one stream whose calls mix sample types and normalisations
    third = n // 3
    mixed = (chunked("lame_encode_buffer", [a[:third] for a in i16], third, 1152)
             + chunked("lame_encode_buffer_ieee_float", [a[third:2 * third] for a in f32], third, 1000)
             + chunked("lame_encode_buffer_int", [a[2 * third:] for a in i32], n - 2 * third, 1152)
             + chunked("lame_encode_buffer_float", [x[0].astype(np.float32)[:2304], x[1].astype(np.float32)[:2304]], 2304, 2304))
    a, b = run_calls(L, mixed), run_calls(R, mixed)
    assert a == b, "mixed"
    print("mixed types in one stream", len(a), "identical")
    # the same through the resampler (32 kHz in, 44.1 kHz out): its input window keeps the caller's type as well (kernel R converts)
    for name, calls in (list(per_type.items())[1:3] if quick else list(per_type.items())) + [("mixed", mixed)]:
        a, b = run_calls(L, calls, 32000, 44100), run_calls(R, calls, 32000, 44100)
        assert a == b, "resampled " + name
        print("resampled", name, len(a), "identical")
    # handles of different sample types side by side: they share one engine, a launch carries rows of several types (and the native
    # window grows from 4-byte to 8-byte elements when the double stream joins)
    import threading
    got = {}
    def work(name, calls):
        got[name] = run_calls(L, calls)
    th = [threading.Thread(target=work, args=(k, v)) for k, v in per_type.items()]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for name, calls in per_type.items():
        assert got[name] == run_calls(R, calls), "concurrent " + name
    print("four sample types side by side identical")
    print("SAMPLE TYPES IDENTICAL")


if __name__ == "__main__":
    main(sys.argv[1])
