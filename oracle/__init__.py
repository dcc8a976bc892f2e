"""oracle - TEST INFRASTRUCTURE.  CPU checkers for the CUDA path; never imported by the product package.

* ``PortEncoder``  - oracle/port: our plain-C restatement of the reference's encode path (liblameport.so).
* ``RefEncoder``   - oracle/_ref: the UNMODIFIED reference compiled from /root/reference (libmp3lame_ref.so),
  built only where the reference sources exist; the built .so travels to the GPU box.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liblameport.so")
REF_SO = os.path.join(HERE, "_ref", "libmp3lame_ref.so")
REFDUMP_SO = os.path.join(HERE, "_ref", "refdump.so")
REFERENCE_ROOT = "/root/reference"


def build(verbose=False):
    """Compile the port (always) and the reference (only where its sources exist)."""
    targets = ["port"]
    if os.path.exists(os.path.join(REFERENCE_ROOT, "libmp3lame", "lame.c")):
        targets.append("ref")
        targets.append("frontend")      # the reference's own `lame` program linked against the reference and against the product library
    subprocess.run(["make", "-C", HERE, "-j8"] + targets, check=True,
                   stdout=None if verbose else subprocess.DEVNULL, stderr=None if verbose else subprocess.DEVNULL)
    return os.path.exists(PORT_SO), os.path.exists(REF_SO)


def have_ref():
    return os.path.exists(REF_SO)


class PortEncoder:
    def __init__(self, samplerate=44100, channels=2, brate=128, mode=4, quality=-1, vbr=0, out_samplerate=0):
        self.lib = ctypes.CDLL(PORT_SO)
        self.lib.lp_open_vq.restype = ctypes.c_void_p
        self.lib.lp_open_vq.argtypes = [ctypes.c_int] * 7 + [ctypes.c_float]
        self.lib.lp_encode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        self.lib.lp_flush.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        self.lib.lp_close.argtypes = [ctypes.c_void_p]
        frac = np.float32(brate) - np.float32(int(brate)) if vbr in (2, 4) else 0.0      # VBR quality = integer level + fraction
        self.h = self.lib.lp_open_vq(samplerate, out_samplerate, channels, int(brate), 4 if mode < 0 else mode, quality, vbr, float(frac))
        if not self.h:
            raise ValueError("port: unsupported configuration")

    def encode(self, left, right):
        l = np.ascontiguousarray(left, dtype=np.int16)
        r = np.ascontiguousarray(right, dtype=np.int16)
        n = int(l.shape[0])
        buf = np.empty(int(1.25 * n) + 7200 + 65536, dtype=np.uint8)
        rc = self.lib.lp_encode(self.h, l.ctypes.data, r.ctypes.data, n, buf.ctypes.data, buf.size)
        if rc < 0:
            raise RuntimeError("lp_encode %d" % rc)
        return buf[:rc].tobytes()

    def flush(self):
        buf = np.empty(65536, dtype=np.uint8)
        rc = self.lib.lp_flush(self.h, buf.ctypes.data, buf.size)
        return buf[:rc].tobytes()

    def close(self):
        if self.h:
            self.lib.lp_close(self.h)
            self.h = None

    def encode_all(self, left, right):
        out = self.encode(left, right) + self.flush()
        self.close()
        return out


class RefEncoder:
    """The real libmp3lame 3.99.5 through its own public API (include/lame.h)."""

    def __init__(self, samplerate=44100, channels=2, brate=128, mode=4, quality=-1, write_tag=False, vbr=0, out_samplerate=0):
        L = self.lib = ctypes.CDLL(REF_SO)
        L.lame_init.restype = ctypes.c_void_p
        for f in ("lame_set_in_samplerate", "lame_set_out_samplerate", "lame_set_num_channels", "lame_set_brate", "lame_set_mode", "lame_set_quality",
                  "lame_set_bWriteVbrTag"):
            getattr(L, f).argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.lame_init_params.argtypes = [ctypes.c_void_p]
        L.lame_encode_buffer.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.lame_encode_flush.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.lame_close.argtypes = [ctypes.c_void_p]
        self.h = L.lame_init()
        L.lame_set_in_samplerate(self.h, samplerate)
        L.lame_set_num_channels(self.h, channels)
        if out_samplerate:
            L.lame_set_out_samplerate(self.h, out_samplerate)
        if vbr == 3:                                          # vbr_abr: brate is the mean bitrate
            L.lame_set_VBR.argtypes = [ctypes.c_void_p, ctypes.c_int]
            L.lame_set_VBR_mean_bitrate_kbps.argtypes = [ctypes.c_void_p, ctypes.c_int]
            L.lame_set_VBR(self.h, 3)
            if brate:
                L.lame_set_VBR_mean_bitrate_kbps(self.h, brate)
        elif vbr in (1, 2, 4):                                # vbr_mt / vbr_rh / vbr_mtrh: brate is VBR_q
            L.lame_set_VBR.argtypes = [ctypes.c_void_p, ctypes.c_int]
            L.lame_set_VBR_q.argtypes = [ctypes.c_void_p, ctypes.c_int]
            L.lame_set_VBR_quality.argtypes = [ctypes.c_void_p, ctypes.c_float]
            L.lame_set_VBR(self.h, vbr)
            if float(brate) == int(brate):
                L.lame_set_VBR_q(self.h, int(brate))
            else:
                L.lame_set_VBR_quality(self.h, float(brate))
        elif brate:
            L.lame_set_brate(self.h, brate)
        if 0 <= mode < 4:
            L.lame_set_mode(self.h, mode)
        if quality >= 0:
            L.lame_set_quality(self.h, quality)
        L.lame_set_bWriteVbrTag(self.h, 1 if write_tag else 0)
        L.lame_get_lametag_frame.restype = ctypes.c_size_t
        L.lame_get_lametag_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        if L.lame_init_params(self.h) < 0:
            raise ValueError("reference: lame_init_params failed")

    def encode(self, left, right, chunk=1152):
        l = np.ascontiguousarray(left, dtype=np.int16)
        r = np.ascontiguousarray(right, dtype=np.int16)
        n = int(l.shape[0])
        buf = np.empty(int(1.25 * n) + 7200 + 65536, dtype=np.uint8)
        rc = self.lib.lame_encode_buffer(self.h, l.ctypes.data, r.ctypes.data, n, buf.ctypes.data, buf.size)
        if rc < 0:
            raise RuntimeError("lame_encode_buffer %d" % rc)
        return buf[:rc].tobytes()

    def flush(self):
        buf = np.empty(65536, dtype=np.uint8)
        rc = self.lib.lame_encode_flush(self.h, buf.ctypes.data, buf.size)
        return buf[:rc].tobytes()

    def lametag_frame(self):
        buf = np.empty(2880, dtype=np.uint8)
        n = self.lib.lame_get_lametag_frame(self.h, buf.ctypes.data, buf.size)
        return buf[:n].tobytes()

    def close(self):
        if self.h:
            self.lib.lame_close(self.h)
            self.h = None

    def encode_all(self, left, right):
        out = self.encode(left, right) + self.flush()
        self.close()
        return out
