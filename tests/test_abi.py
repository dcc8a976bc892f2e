"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol include/lamegpu.h declares;
without a CUDA device it refuses to work (no CPU fallback)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT, cuda_available


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build_product()
    import lame_b200
    return lame_b200


def declared_functions():
    src = open(os.path.join(ROOT, "include", "lamegpu.h")).read() + open(os.path.join(ROOT, "include", "lamegpu_options.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)
    return sorted({n for n in names if n.startswith(("lame_", "lamegpu_", "get_lame", "get_psy"))})


def test_header_symbols_are_exported(lib):
    so = lib.LIB_PATH
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    decl = declared_functions()
    assert len(decl) >= 40
    missing = [n for n in decl if n not in exported]
    assert not missing, missing
    assert sorted(lib.EXPORTED_SYMBOLS) == decl


def test_library_contains_sm100a_kernels(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    for k in ("lg_kernel_analysis", "lg_kernel_scan", "lg_kernel_mdct", "lg_kernel_quant", "lg_kernel_pack"):
        assert k in sass
    assert "REDUX" in sass and "SHFL" in sass          # warp reductions of the bit counters / noise maxima


def test_carried_options_accept_defaults_only(lib):
    """the lame_set_X / lame_get_X pairs of libmp3lame that this library does not act on (free format, ReplayGain, the decoder, ...): the
    value is stored and read back, and lame_init_params refuses it when it is not what a fresh handle has - before any GPU work.  The
    honoured ones (filters, scaling, ATH, reservoir, ...) are checked against the reference on the GPU (tests/test_frontend_dropin.py)."""
    import ctypes
    L = lib.load_library()
    for fn, res, arg in (("lame_set_free_format", None, ctypes.c_int), ("lame_get_free_format", ctypes.c_int, None), ("lame_set_findReplayGain", None, ctypes.c_int),
                         ("lame_set_quant_comp", None, ctypes.c_int), ("lame_get_quant_comp", ctypes.c_int, None), ("lame_set_copyright", None, ctypes.c_int),
                         ("lame_get_copyright", ctypes.c_int, None), ("lame_get_original", ctypes.c_int, None), ("lame_get_useTemporal", ctypes.c_int, None),
                         ("lame_set_experimentalZ", None, ctypes.c_int)):
        f = getattr(L, fn)
        f.restype = res if res else ctypes.c_int
        f.argtypes = [ctypes.c_void_p] + ([arg] if arg else [])
    for setter, value in (("lame_set_free_format", 1), ("lame_set_findReplayGain", 1), ("lame_set_quant_comp", 3), ("lame_set_experimentalZ", 1)):
        h = ctypes.c_void_p(L.lame_init())
        assert L.lame_get_free_format(h) == 0 and L.lame_get_original(h) == 1 and L.lame_get_quant_comp(h) == -1 and L.lame_get_useTemporal(h) == -1
        assert L.lame_set_copyright(h, 1) == 0 and L.lame_get_copyright(h) == 1
        assert getattr(L, setter)(h, value) == 0
        assert L.lame_init_params(h) == -1           # refused, loudly - never a silently different stream
        L.lame_close(h)


def test_whole_libmp3lame_export_list_is_present(lib):
    """include/libmp3lame.sym of the reference (232 symbols; committed copy of the list in tests/golden/)"""
    want = {l.strip() for l in open(os.path.join(ROOT, "tests", "golden", "libmp3lame.sym")) if l.strip()}
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines()}
    assert len(want) == 232 and not (want - exported), sorted(want - exported)


def test_signatures_bind(lib):
    L = lib.load_library()
    assert L.get_lame_short_version() == b"3.99.5"
    assert L.lamegpu_sizeof_granule_out() % 16 == 0


def test_error_codes_without_init(lib):
    L = lib.load_library()
    h = L.lame_init()
    assert h
    # lame.h:687-691: -3 when lame_init_params was not called
    assert L.lame_encode_buffer(h, None, None, 10, None, 0) == -3
    assert L.lame_encode_flush(h, None, 0) == -3
    assert L.lame_set_num_channels(h, 3) == -1
    assert L.lame_set_brate(h, 128) == 0 and L.lame_get_brate(h) == 128
    assert L.lame_close(h) == 0


@pytest.mark.skipif(cuda_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(lib):
    with pytest.raises(lib.LameGpuError):
        lib.BatchEncoder(4)
    with pytest.raises(lib.LameGpuError):
        lib.Encoder()
