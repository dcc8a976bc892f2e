"""B200 batched MP3 (MPEG-1 Layer III; CBR, ABR, VBR-new) encoder behind the libmp3lame API - Python host side.

The product is the C-ABI library ``liblamegpu.so`` (``csrc/``, declared in ``include/lamegpu.h``): hand-written
sm_100a CUDA kernels for the encode hot path plus the thin host code the reference keeps serial (PCM
buffering, bit packing).  This module is only the ctypes binding to it, mirroring the reference's operator
interface:

* :class:`Encoder` - one stream, same call sequence and return conventions as ``lame_init`` /
  ``lame_set_*`` / ``lame_init_params`` / ``lame_encode_buffer`` / ``lame_encode_flush`` / ``lame_close``
  (``include/lame.h`` of LAME 3.99.5).
* :class:`BatchEncoder` - many independent streams per call (``lamegpu_batch_*``), the form the GPU needs.

There is no CPU path: if ``liblamegpu.so`` is missing or no CUDA device is visible, construction raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblamegpu.so")

STEREO, JOINT_STEREO, DUAL_CHANNEL, MONO, NOT_SET = 0, 1, 2, 3, 4
VBR_OFF, VBR_MT, VBR_RH, VBR_ABR, VBR_MTRH = 0, 1, 2, 3, 4     # lame.h:94 vbr_mode (with the VBR modes, `brate` is VBR_q)

_lib = None


class LameGpuError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen liblamegpu.so and declare every prototype of include/lamegpu.h.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise LameGpuError("%s is not built - run `python __graft_entry__.py build` (nvcc, sm_100a); "
                           "there is no CPU fallback" % p)
    lib = ctypes.CDLL(p)
    c_int, c_long, c_void_p, c_char_p = ctypes.c_int, ctypes.c_long, ctypes.c_void_p, ctypes.c_char_p
    P = ctypes.POINTER
    sig = {
        "lame_init": (c_void_p, []),
        "lame_set_in_samplerate": (c_int, [c_void_p, c_int]), "lame_get_in_samplerate": (c_int, [c_void_p]),
        "lame_set_num_channels": (c_int, [c_void_p, c_int]), "lame_get_num_channels": (c_int, [c_void_p]),
        "lame_set_out_samplerate": (c_int, [c_void_p, c_int]), "lame_get_out_samplerate": (c_int, [c_void_p]),
        "lame_set_brate": (c_int, [c_void_p, c_int]), "lame_get_brate": (c_int, [c_void_p]),
        "lame_set_quality": (c_int, [c_void_p, c_int]), "lame_get_quality": (c_int, [c_void_p]),
        "lame_set_mode": (c_int, [c_void_p, c_int]), "lame_get_mode": (c_int, [c_void_p]),
        "lame_set_VBR": (c_int, [c_void_p, c_int]), "lame_get_VBR": (c_int, [c_void_p]),
        "lame_set_VBR_mean_bitrate_kbps": (c_int, [c_void_p, c_int]), "lame_get_VBR_mean_bitrate_kbps": (c_int, [c_void_p]),
        "lame_set_VBR_q": (c_int, [c_void_p, c_int]), "lame_get_VBR_q": (c_int, [c_void_p]),
        "lamegpu_batch_open_ex": (c_void_p, [c_int] * 9),
        "lamegpu_batch_open_rs": (c_void_p, [c_int] * 10),
        "lamegpu_batch_open_vq": (c_void_p, [c_int, c_int, c_int, ctypes.c_float] + [c_int] * 6),
        "lame_set_VBR_quality": (c_int, [c_void_p, ctypes.c_float]), "lame_get_VBR_quality": (ctypes.c_float, [c_void_p]),
        "lame_set_bWriteVbrTag": (c_int, [c_void_p, c_int]), "lame_get_bWriteVbrTag": (c_int, [c_void_p]),
        "lame_init_params": (c_int, [c_void_p]),
        "lame_get_framesize": (c_int, [c_void_p]), "lame_get_frameNum": (c_int, [c_void_p]),
        "lame_get_encoder_delay": (c_int, [c_void_p]),
        "lame_encode_buffer": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int]),
        "lame_encode_buffer_interleaved": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int]),
        "lame_encode_buffer_ieee_float": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int]),
        "lame_encode_flush": (c_int, [c_void_p, c_void_p, c_int]),
        "lame_close": (c_int, [c_void_p]),
        "lame_encode_buffer_float": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int]),
        "lame_encode_buffer_interleaved_ieee_float": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int]),
        "lame_encode_buffer_ieee_double": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int]),
        "lame_encode_buffer_interleaved_ieee_double": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int]),
        "lame_encode_buffer_long": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int]),
        "lame_encode_buffer_long2": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int]),
        "lame_encode_buffer_int": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int]),
        "lame_get_lametag_frame": (ctypes.c_size_t, [c_void_p, c_void_p, ctypes.c_size_t]),
        "get_lame_short_version": (c_char_p, []),
        "lame_bitrate_kbps": (None, [c_void_p, c_void_p]), "lame_bitrate_hist": (None, [c_void_p, c_void_p]),
        "lame_stereo_mode_hist": (None, [c_void_p, c_void_p]), "lame_bitrate_stereo_mode_hist": (None, [c_void_p, c_void_p]),
        "lame_block_type_hist": (None, [c_void_p, c_void_p]), "lame_bitrate_block_type_hist": (None, [c_void_p, c_void_p]),
        "lamegpu_batch_open": (c_void_p, [c_int] * 8),
        "lamegpu_batch_close": (None, [c_void_p]),
        "lamegpu_batch_encode": (c_long, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
        "lamegpu_batch_flush": (c_long, [c_void_p, c_void_p, c_void_p, c_void_p]),
        "lamegpu_batch_encode_packed": (c_long, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
        "lamegpu_batch_flush_packed": (c_long, [c_void_p, c_void_p, c_int, c_void_p]),
        "lamegpu_batch_rerun_device": (c_int, [c_void_p, c_int]),
        "lamegpu_batch_run_device_steps": (ctypes.c_float, [c_void_p, c_int, c_int]),
        "lamegpu_batch_set_pipelined": (c_int, [c_void_p, c_int]),
        "lamegpu_batch_devices": (c_int, [c_void_p]),
        "lamegpu_batch_stage_packed": (c_int, [c_void_p, c_void_p, c_int]),
        "lamegpu_batch_kernel_ms": (c_int, [c_void_p, P(ctypes.c_float)]),
        "lamegpu_batch_step_ms": (ctypes.c_float, [c_void_p]),
        "lamegpu_batch_kernel_launches": (c_long, [c_void_p]),
        "lamegpu_batch_set_threads": (c_int, [c_void_p, c_int]),
        "lamegpu_batch_debug_copy": (c_long, [c_void_p, c_int, c_void_p, ctypes.c_size_t]),
        "lamegpu_sizeof_granule_out": (ctypes.c_size_t, []),
        "lamegpu_batch_d2h_bytes": (c_long, [c_void_p]),
        "lamegpu_sizeof_analysis": (ctypes.c_size_t, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


def _declared_symbols():
    """every function include/lamegpu.h and include/lamegpu_options.h declare: the library's C ABI"""
    import re
    inc = os.path.join(os.path.dirname(_HERE), "include")
    src = "".join(open(os.path.join(inc, f)).read() for f in ("lamegpu.h", "lamegpu_options.h") if os.path.exists(os.path.join(inc, f)))
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)
    return sorted({n for n in names if n.startswith(("lame_", "lamegpu_", "get_lame", "get_psy"))})


EXPORTED_SYMBOLS = _declared_symbols()


def _as_i16(a):
    a = np.ascontiguousarray(a, dtype=np.int16)
    return a


def _histograms(L, h):
    """the six statistics calls of lame.h:909-929 on handle `h` of library `L` (ours or the reference's)"""
    out = {}
    for name, shape in (("lame_bitrate_kbps", (14,)), ("lame_bitrate_hist", (14,)), ("lame_stereo_mode_hist", (4,)),
                        ("lame_bitrate_stereo_mode_hist", (14, 4)), ("lame_block_type_hist", (6,)), ("lame_bitrate_block_type_hist", (14, 6))):
        a = np.zeros(shape, dtype=np.int32)
        fn = getattr(L, name)
        fn.restype, fn.argtypes = None, [ctypes.c_void_p, ctypes.c_void_p]
        fn(h, a.ctypes.data)
        out[name] = a
    return out


class Encoder:
    """One stream through the libmp3lame-compatible entry points (same semantics and error codes)."""

    def __init__(self, samplerate=44100, channels=2, brate=128, mode=NOT_SET, quality=-1, write_tag=False, vbr=VBR_OFF, out_samplerate=0):
        self._lib = load_library()
        self._h = self._lib.lame_init()
        if not self._h:
            raise LameGpuError("lame_init failed")
        L = self._lib
        L.lame_set_in_samplerate(self._h, samplerate)
        L.lame_set_num_channels(self._h, channels)
        if out_samplerate:
            L.lame_set_out_samplerate(self._h, out_samplerate)     # else chosen by lame_init_params; resampled on the device when it differs
        if vbr == VBR_ABR:
            L.lame_set_VBR(self._h, VBR_ABR)
            if brate:
                L.lame_set_VBR_mean_bitrate_kbps(self._h, brate)
        elif vbr in (VBR_MTRH, VBR_RH, VBR_MT):
            L.lame_set_VBR(self._h, vbr)
            if float(brate) == int(brate):
                L.lame_set_VBR_q(self._h, int(brate))
            else:
                L.lame_set_VBR_quality(self._h, float(brate))   # fractional level
        elif brate:
            L.lame_set_brate(self._h, brate)
        if mode != NOT_SET:
            L.lame_set_mode(self._h, mode)
        if quality >= 0:
            L.lame_set_quality(self._h, quality)
        L.lame_set_bWriteVbrTag(self._h, 1 if write_tag else 0)
        rc = L.lame_init_params(self._h)
        if rc < 0:
            L.lame_close(self._h)
            self._h = None
            raise LameGpuError("lame_init_params returned %d (unsupported configuration or no GPU)" % rc)
        self.channels = channels

    def encode(self, left, right=None):
        """lame_encode_buffer: int16 PCM in, bytes out (possibly empty)."""
        l = _as_i16(left)
        r = _as_i16(right) if right is not None else l
        n = int(l.shape[0])
        buf = np.empty(int(1.25 * n) + 7200 + 65536, dtype=np.uint8)
        rc = self._lib.lame_encode_buffer(self._h, l.ctypes.data, r.ctypes.data, n, buf.ctypes.data, buf.size)
        if rc < 0:
            raise LameGpuError("lame_encode_buffer returned %d" % rc)
        return buf[:rc].tobytes()

    def flush(self):
        buf = np.empty(65536 + 7200 * 8, dtype=np.uint8)
        rc = self._lib.lame_encode_flush(self._h, buf.ctypes.data, buf.size)
        if rc < 0:
            raise LameGpuError("lame_encode_flush returned %d" % rc)
        return buf[:rc].tobytes()

    def lametag_frame(self):
        """lame_get_lametag_frame: the finished Info tag frame (to be written over the placeholder at offset 0)"""
        buf = np.empty(2880, dtype=np.uint8)
        n = self._lib.lame_get_lametag_frame(self._h, buf.ctypes.data, buf.size)
        return buf[:n].tobytes()

    def histograms(self):
        """lame_bitrate_kbps / _hist / _stereo_mode_hist / _block_type_hist and the per-bitrate tables (lame.h:909-929)"""
        return _histograms(self._lib, self._h)

    def close(self):
        if self._h:
            self._lib.lame_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchEncoder:
    """`nstreams` independent streams with one configuration, encoded together on one GPU."""

    def __init__(self, nstreams, samplerate=44100, channels=2, brate=128, mode=-1, quality=-1,
                 frames_per_launch=8, device=0, vbr=VBR_OFF, out_samplerate=0):
        self._lib = load_library()
        self.nstreams, self.frames_per_launch = int(nstreams), int(frames_per_launch)
        self._h = self._lib.lamegpu_batch_open_vq(samplerate, out_samplerate, channels, float(brate), mode, quality, vbr, self.nstreams,
                                                  self.frames_per_launch, device)
        if not self._h:
            raise LameGpuError("lamegpu_batch_open failed (unsupported configuration, no CUDA device, or out of memory)")
        self._nbytes = np.zeros(self.nstreams, dtype=np.int32)

    def encode(self, pcm):
        """pcm: int16 array [nstreams, 2, nsamples].  Returns (frames_encoded, list of bytes per stream)."""
        pcm = _as_i16(pcm)
        assert pcm.ndim == 3 and pcm.shape[0] == self.nstreams and pcm.shape[1] == 2
        n = int(pcm.shape[2])
        stride = int(1.25 * n) + 7200 + 4096
        out = np.empty((self.nstreams, stride), dtype=np.uint8)
        done = self._lib.lamegpu_batch_encode_packed(self._h, pcm.ctypes.data, n, out.ctypes.data, stride,
                                                     self._nbytes.ctypes.data)
        if done < 0:
            raise LameGpuError("lamegpu_batch_encode_packed returned %d" % done)
        return int(done), [out[s, :self._nbytes[s]].tobytes() for s in range(self.nstreams)]

    def encode_raw(self, pcm, out, nbytes):
        """Same without Python-side copies: caller owns `out` [nstreams, stride] uint8 and `nbytes` int32."""
        n = int(pcm.shape[2])
        return int(self._lib.lamegpu_batch_encode_packed(self._h, pcm.ctypes.data, n, out.ctypes.data,
                                                         int(out.shape[1]), nbytes.ctypes.data))

    def flush(self):
        stride = 65536
        out = np.empty((self.nstreams, stride), dtype=np.uint8)
        done = self._lib.lamegpu_batch_flush_packed(self._h, out.ctypes.data, stride, self._nbytes.ctypes.data)
        if done < 0:
            raise LameGpuError("lamegpu_batch_flush_packed returned %d" % done)
        return int(done), [out[s, :self._nbytes[s]].tobytes() for s in range(self.nstreams)]

    def flush_raw(self, out, nbytes):
        return int(self._lib.lamegpu_batch_flush_packed(self._h, out.ctypes.data, int(out.shape[1]), nbytes.ctypes.data))

    # measurement hooks used by bench.py
    def stage(self, pcm, nframes):
        rc = self._lib.lamegpu_batch_stage_packed(self._h, _as_i16(pcm).ctypes.data, int(nframes))
        if rc != 0:
            raise LameGpuError("lamegpu_batch_stage_packed failed")

    def rerun_device(self, nframes):
        rc = self._lib.lamegpu_batch_rerun_device(self._h, int(nframes))
        if rc != 0:
            raise LameGpuError("lamegpu_batch_rerun_device failed")

    def run_device_steps(self, nframes, steps):
        """`steps` device-only steps back to back on persistent streams; device ms per step (CUDA events)"""
        ms = float(self._lib.lamegpu_batch_run_device_steps(self._h, int(nframes), int(steps)))
        if ms < 0:
            raise LameGpuError("lamegpu_batch_run_device_steps failed")
        return ms

    def set_pipelined(self, on=True):
        """leave the newest step in flight between calls: bytes come out one call later, host work overlaps the device"""
        if self._lib.lamegpu_batch_set_pipelined(self._h, 1 if on else 0) != 0:
            raise LameGpuError("lamegpu_batch_set_pipelined failed")

    def kernel_ms(self):
        ms = (ctypes.c_float * 5)()
        self._lib.lamegpu_batch_kernel_ms(self._h, ms)
        return [float(x) for x in ms]

    def step_ms(self):
        """device time of the last launch, first kernel start to last kernel end (CUDA events)"""
        return float(self._lib.lamegpu_batch_step_ms(self._h))

    def kernel_launches(self):
        return int(self._lib.lamegpu_batch_kernel_launches(self._h))

    def set_threads(self, n):
        self._lib.lamegpu_batch_set_threads(self._h, int(n))

    def close(self):
        if self._h:
            self._lib.lamegpu_batch_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
