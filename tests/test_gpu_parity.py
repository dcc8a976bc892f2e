"""GPU suite (-m gpu): the CUDA path, called through the C ABI (ctypes -> liblamegpu.so), against the oracle:
oracle/port (always) and the unmodified reference oracle/_ref (when its prebuilt .so travelled to the box), plus the
committed golden vectors.  Bar: byte-identical MP3 streams."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_signal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))


@pytest.fixture(scope="module")
def lib():
    import lame_b200
    lame_b200.load_library()          # raises if the CUDA extension is missing: no fallback
    return lame_b200


def oracle_bytes(oracle_mod, x, sr=44100, brate=128, mode=-1, q=-1):
    a = oracle_mod.PortEncoder(sr, 2, brate, mode, q).encode_all(x[0], x[1])
    if oracle_mod.have_ref():
        b = oracle_mod.RefEncoder(sr, 2, brate, mode if mode >= 0 else 4, q).encode_all(x[0], x[1])
        assert a == b, "oracle port and reference disagree"
    return a


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_golden_vectors(lib, name):
    """the reference's own output, committed as fixtures, reproduced by the GPU"""
    m = MANIFEST[name]
    x = make_signal(m["signal"], m["frames"] * 1152)
    enc = lib.BatchEncoder(1, m["samplerate"], 2, m["brate"], m["mode"], m["quality"], frames_per_launch=8, out_samplerate=m.get("out_samplerate", 0))
    _, a = enc.encode(x[None])
    _, b = enc.flush()
    enc.close()
    assert a[0] + b[0] == open(os.path.join(GOLD, name + ".mp3"), "rb").read()


@pytest.mark.parametrize("cfg", [
    dict(S=16, F=24, fpl=8, brate=128), dict(S=7, F=30, fpl=16, brate=320, mode=1), dict(S=5, F=20, fpl=3, brate=192, mode=0, q=5),
    dict(S=4, F=16, fpl=8, brate=256, sr=48000), dict(S=4, F=16, fpl=8, brate=128, sr=32000), dict(S=3, F=12, fpl=4, brate=160, q=7),
    dict(S=3, F=12, fpl=4, brate=112, q=9), dict(S=2, F=40, fpl=40, brate=224, q=4), dict(S=4, F=20, fpl=8, brate=160, mode=2),
    # quality 2 / 1 / 0: substep shaping, one-band amplification, full outer loop
    dict(S=8, F=24, fpl=8, brate=128, q=2), dict(S=8, F=24, fpl=8, brate=128, q=1), dict(S=8, F=24, fpl=8, brate=128, q=0),
    dict(S=4, F=20, fpl=4, brate=320, mode=1, q=0), dict(S=4, F=16, fpl=8, brate=192, mode=0, q=1, sr=48000), dict(S=4, F=16, fpl=8, brate=112, q=2, sr=32000),
])
def test_batch_matches_oracle(lib, oracle_mod, cfg):
    S, F = cfg["S"], cfg["F"]
    sr, brate, mode, q = cfg.get("sr", 44100), cfg["brate"], cfg.get("mode", -1), cfg.get("q", -1)
    kinds = ("noise", "click", "sine", "gap")
    pcm = np.stack([make_signal(kinds[s % 4], F * 1152, seed=10 + s) for s in range(S)])
    enc = lib.BatchEncoder(S, sr, 2, brate, mode, q, frames_per_launch=cfg["fpl"])
    got = [b""] * S
    pos = 0
    for chunk in (1000, 1152, 4000, 10 ** 9):              # ragged call sizes
        c = min(chunk, F * 1152 - pos)
        if c <= 0:
            break
        _, out = enc.encode(pcm[:, :, pos:pos + c])
        got = [g + o for g, o in zip(got, out)]
        pos += c
    _, out = enc.flush()
    got = [g + o for g, o in zip(got, out)]
    assert enc.kernel_launches() > 0
    enc.close()
    for s in range(S):
        assert got[s] == oracle_bytes(oracle_mod, pcm[s], sr, brate, mode, q), "stream %d (%s)" % (s, kinds[s % 4])


@pytest.mark.parametrize("cfg", [
    dict(S=16, F=24, fpl=8, brate=128), dict(S=6, F=30, fpl=16, brate=192, mode=1), dict(S=4, F=20, fpl=3, brate=150, mode=0, q=5),
    dict(S=4, F=16, fpl=8, brate=256, sr=48000), dict(S=3, F=16, fpl=8, brate=112, sr=32000, q=7),
    dict(S=4, F=20, fpl=8, brate=128, q=0), dict(S=4, F=20, fpl=8, brate=160, q=2),
])
def test_abr_batch_matches_oracle(lib, oracle_mod, cfg):
    """ABR (vbr_abr): the device chooses every frame's bitrate index (quantize.c:1962), frames of different sizes go
    through kernel E and the host splice; byte-identical to the port and to libmp3lame"""
    S, F = cfg["S"], cfg["F"]
    sr, brate, mode, q = cfg.get("sr", 44100), cfg["brate"], cfg.get("mode", -1), cfg.get("q", -1)
    kinds = ("noise", "click", "sine", "gap")
    pcm = np.stack([make_signal(kinds[s % 4], F * 1152, seed=40 + s) for s in range(S)])
    enc = lib.BatchEncoder(S, sr, 2, brate, mode, q, frames_per_launch=cfg["fpl"], vbr=lib.VBR_ABR)
    _, a = enc.encode(pcm)
    _, b = enc.flush()
    enc.close()
    sizes = set()
    for s in range(S):
        want = oracle_mod.PortEncoder(sr, 2, brate, mode, q, vbr=3).encode_all(pcm[s, 0], pcm[s, 1])
        if oracle_mod.have_ref():
            ref = oracle_mod.RefEncoder(sr, 2, brate, mode if mode >= 0 else 4, q, vbr=3).encode_all(pcm[s, 0], pcm[s, 1])
            assert want == ref, "oracle port and reference disagree"
        assert a[s] + b[s] == want, "stream %d (%s)" % (s, kinds[s % 4])
        sizes.add(len(want))
    assert len(sizes) > 1                                  # ABR: stream lengths depend on the signal


@pytest.mark.parametrize("cfg", [
    dict(S=16, F=24, fpl=8, q=2), dict(S=8, F=30, fpl=16, q=0), dict(S=4, F=20, fpl=3, q=4, mode=0, quality=5),
    dict(S=4, F=16, fpl=8, q=0, sr=48000), dict(S=3, F=16, fpl=8, q=5, mode=3), dict(S=4, F=20, fpl=20, q=6, quality=6),
    # levels between the presets (lame_set_VBR_quality), -V7 (32 kHz output through the resampler), VBR at 32 kHz input
    dict(S=4, F=20, fpl=8, q=2.5), dict(S=4, F=16, fpl=8, q=5.3, sr=48000), dict(S=4, F=20, fpl=8, q=7), dict(S=4, F=16, fpl=4, q=3, sr=32000),
    # quality 7-9: the step guess of vbrquantize.c:324 (log10f restated on the device as lg_log10f)
    dict(S=8, F=20, fpl=8, q=2, quality=7), dict(S=4, F=16, fpl=16, q=4, quality=9, sr=48000), dict(S=4, F=16, fpl=4, q=0, mode=0, quality=8),
    dict(S=4, F=16, fpl=8, q=5, quality=7, sr=22050),
])
def test_vbr_batch_matches_oracle(lib, oracle_mod, cfg):
    """VBR-new (vbr_mtrh, -V q): lg_kernel_vbr - per-band step search, fitting, and for the frames that do not
    fit (click at -V0) the out-of-bits strategy; byte-identical to the port and to libmp3lame"""
    S, F = cfg["S"], cfg["F"]
    sr, q, mode, quality = cfg.get("sr", 44100), cfg["q"], cfg.get("mode", -1), cfg.get("quality", -1)
    kinds = ("noise", "click", "sine", "gap")
    pcm = np.stack([make_signal(kinds[s % 4], F * 1152, seed=60 + s) for s in range(S)])
    enc = lib.BatchEncoder(S, sr, 2, q, mode, quality, frames_per_launch=cfg["fpl"], vbr=lib.VBR_MTRH)
    _, a = enc.encode(pcm)
    _, b = enc.flush()
    enc.close()
    for s in range(S):
        want = oracle_mod.PortEncoder(sr, 2, q, mode, quality, vbr=4).encode_all(pcm[s, 0], pcm[s, 1])
        if oracle_mod.have_ref():
            ref = oracle_mod.RefEncoder(sr, 2, q, mode if mode >= 0 else 4, quality, vbr=4).encode_all(pcm[s, 0], pcm[s, 1])
            assert want == ref, "oracle port and reference disagree"
        assert a[s] + b[s] == want, "stream %d (%s)" % (s, kinds[s % 4])


@pytest.mark.parametrize("cfg", [
    dict(S=8, F=20, fpl=8, sr=48000, out=44100, brate=128, chunk=4000), dict(S=4, F=16, fpl=4, sr=44100, out=48000, brate=192, chunk=1152),
    dict(S=4, F=16, fpl=8, sr=44100, out=0, brate=96, chunk=777), dict(S=4, F=6, fpl=4, sr=8000, out=44100, brate=128, chunk=100),
    dict(S=3, F=40, fpl=4, sr=192000, out=32000, brate=128, chunk=50000), dict(S=3, F=24, fpl=8, sr=96000, out=0, brate=160, chunk=10 ** 9),
    dict(S=4, F=16, fpl=8, sr=44100, out=32000, brate=2, vbr=4, chunk=3000), dict(S=4, F=16, fpl=8, sr=48000, out=0, brate=112, vbr=3, chunk=5000),
    dict(S=3, F=12, fpl=4, sr=44100, out=0, brate=64, mode=3, chunk=2000), dict(S=2, F=3, fpl=4, sr=48000, out=44100, brate=128, chunk=7),
])
def test_resampled_batch_matches_oracle(lib, oracle_mod, cfg):
    """input rate != MPEG output rate.  lg_kernel_resample makes the PCM window on the device from the reference's
    per-call chunk schedule (util.c:531), so the streams are fed to the oracle in the same call sizes; byte-identical to
    the port and to libmp3lame, incl. the low bitrates that imply a 32 kHz output and calls of 7 samples"""
    S, F, sr, out, brate, chunk = cfg["S"], cfg["F"], cfg["sr"], cfg["out"], cfg["brate"], cfg["chunk"]
    mode, vbr = cfg.get("mode", -1), cfg.get("vbr", 0)
    kinds = ("noise", "click", "sine", "gap")
    pcm = np.stack([make_signal(kinds[s % 4], F * 1152, seed=80 + s) for s in range(S)])
    enc = lib.BatchEncoder(S, sr, 2, brate, mode, -1, frames_per_launch=cfg["fpl"], vbr=vbr, out_samplerate=out)
    got = [b""] * S
    for pos in range(0, F * 1152, chunk):
        _, o = enc.encode(pcm[:, :, pos:pos + chunk])
        got = [g + x for g, x in zip(got, o)]
    _, o = enc.flush()
    got = [g + x for g, x in zip(got, o)]
    enc.close()
    for s in range(S):
        encs = [oracle_mod.PortEncoder(sr, 2, brate, mode, -1, vbr=vbr, out_samplerate=out)]
        if oracle_mod.have_ref():
            encs.append(oracle_mod.RefEncoder(sr, 2, brate, mode if mode >= 0 else 4, -1, vbr=vbr, out_samplerate=out))
        for e in encs:
            want = b""
            for pos in range(0, F * 1152, chunk):
                want += e.encode(pcm[s, 0, pos:pos + chunk], pcm[s, 1, pos:pos + chunk])
            want += e.flush()
            e.close()
            assert got[s] == want, "stream %d (%s) vs %s" % (s, kinds[s % 4], type(e).__name__)


@pytest.mark.parametrize("kw", [dict(samplerate=48000, brate=128, out_samplerate=44100), dict(samplerate=44100, brate=7, vbr=4),
                                dict(samplerate=44100, brate=4.6, vbr=4), dict(samplerate=44100, brate=96, quality=1),
                                dict(samplerate=22050, brate=64), dict(samplerate=44100, brate=9, vbr=4), dict(samplerate=8000, brate=16, mode=3),
                                dict(samplerate=24000, brate=80, vbr=3), dict(samplerate=44100, brate=2, vbr=2), dict(samplerate=44100, brate=4, vbr=1),
                                dict(samplerate=22050, brate=5, vbr=2)])
def test_resampled_lame_api_with_tag(lib, oracle_mod, kw):
    """the lame.h face with lame_set_out_samplerate / -V7 / a fractional level / quality 1, and the Info tag: source-rate field,
    encoder padding, quality and preset fields of the tag (VbrTag.c:775, lame.c:2083-2091)"""
    if not oracle_mod.have_ref():
        pytest.skip("needs the reference build")
    x = make_signal("click", 30 * 1152, seed=5)
    e = lib.Encoder(channels=2, write_tag=True, **kw)
    r = oracle_mod.RefEncoder(channels=2, write_tag=True, **kw)
    a = b_ = b""
    for pos in range(0, x.shape[1], 5000):
        a += e.encode(x[0, pos:pos + 5000], x[1, pos:pos + 5000])
        b_ += r.encode(x[0, pos:pos + 5000], x[1, pos:pos + 5000])
    a += e.flush()
    b_ += r.flush()
    assert a == b_
    assert e.lametag_frame() == r.lametag_frame()
    e.close()
    r.close()


@pytest.mark.parametrize("kw", [dict(brate=128), dict(brate=2, vbr=4), dict(brate=150, vbr=3), dict(brate=192, mode=0), dict(brate=96, mode=3),
                                dict(brate=64), dict(brate=9, vbr=4)])
def test_statistics_match_reference(lib, oracle_mod, kw):
    """lame_bitrate_hist / lame_stereo_mode_hist / lame_block_type_hist and the per-bitrate tables (lame.h:909-929, encoder.c:156):
    the host keeps them from what the device reports per frame; equal to the reference's after the same stream"""
    if not oracle_mod.have_ref():
        pytest.skip("needs the reference build")
    x = make_signal("click", 40 * 1152, seed=9)
    mode = kw.get("mode", -1)
    e = lib.Encoder(44100, 2, kw["brate"], mode if mode >= 0 else lib.NOT_SET, -1, vbr=kw.get("vbr", 0))
    r = oracle_mod.RefEncoder(44100, 2, kw["brate"], mode if mode >= 0 else 4, -1, vbr=kw.get("vbr", 0))
    a = e.encode(x[0], x[1]) + e.flush()
    b = r.encode(x[0], x[1]) + r.flush()
    assert a == b
    ours, ref = e.histograms(), lib._histograms(r.lib, r.h)
    for k in ref:
        assert (ours[k] == ref[k]).all(), k
    assert ours["lame_block_type_hist"][2] > 0        # the click stream has short blocks
    e.close()
    r.close()


@pytest.mark.parametrize("cfg", [
    dict(S=8, F=24, fpl=8, sr=22050, brate=64), dict(S=4, F=24, fpl=4, sr=16000, brate=32, chunk=777), dict(S=4, F=24, fpl=8, sr=24000, brate=160, mode=0),
    dict(S=4, F=24, fpl=4, sr=8000, brate=24), dict(S=4, F=24, fpl=4, sr=11025, brate=32, mode=3), dict(S=4, F=24, fpl=4, sr=12000, brate=8, mode=3, chunk=500),
    dict(S=8, F=20, fpl=8, sr=44100, brate=64), dict(S=4, F=20, fpl=4, sr=22050, brate=80, mode=1, q=2), dict(S=4, F=20, fpl=4, sr=16000, brate=56, q=0),
    dict(S=4, F=20, fpl=8, sr=22050, brate=64, vbr=3), dict(S=4, F=20, fpl=4, sr=8000, brate=24, vbr=3),
    dict(S=8, F=20, fpl=8, sr=22050, brate=4, vbr=4), dict(S=4, F=20, fpl=4, sr=44100, brate=8, vbr=4), dict(S=4, F=20, fpl=4, sr=44100, brate=9, vbr=4, chunk=777),
    dict(S=4, F=20, fpl=4, sr=16000, brate=2, vbr=4), dict(S=4, F=20, fpl=4, sr=8000, brate=5, vbr=4), dict(S=4, F=20, fpl=4, sr=48000, brate=9.5, vbr=4),
    dict(S=4, F=20, fpl=4, sr=44100, brate=48, out=16000, chunk=3000),
])
def test_mpeg2_batch_matches_oracle(lib, oracle_mod, cfg):
    """MPEG-2 (16 / 22.05 / 24 kHz) and MPEG-2.5 (8 / 11.025 / 12 kHz) output, native or through the resampler (64 kbps at
    44.1 kHz -> 24 kHz, -V8, -V9): one granule per frame, LSF scalefactor partitions and side info; CBR, ABR, VBR, quality 0 and 2;
    byte-identical to the port and to libmp3lame"""
    S, F, sr, brate, chunk = cfg["S"], cfg["F"], cfg["sr"], cfg["brate"], cfg.get("chunk", 1152 * 4)
    mode, vbr, q, out = cfg.get("mode", -1), cfg.get("vbr", 0), cfg.get("q", -1), cfg.get("out", 0)
    kinds = ("noise", "click", "sine", "gap")
    pcm = np.stack([make_signal(kinds[s % 4], F * 1152, seed=120 + s) for s in range(S)])
    enc = lib.BatchEncoder(S, sr, 2, brate, mode, q, frames_per_launch=cfg["fpl"], vbr=vbr, out_samplerate=out)
    got = [b""] * S
    for pos in range(0, F * 1152, chunk):
        _, o = enc.encode(pcm[:, :, pos:pos + chunk])
        got = [g + x for g, x in zip(got, o)]
    _, o = enc.flush()
    got = [g + x for g, x in zip(got, o)]
    enc.close()
    for s in range(S):
        encs = [oracle_mod.PortEncoder(sr, 2, brate, mode, q, vbr=vbr, out_samplerate=out)]
        if oracle_mod.have_ref():
            encs.append(oracle_mod.RefEncoder(sr, 2, brate, mode if mode >= 0 else 4, q, vbr=vbr, out_samplerate=out))
        for e in encs:
            want = b""
            for pos in range(0, F * 1152, chunk):
                want += e.encode(pcm[s, 0, pos:pos + chunk], pcm[s, 1, pos:pos + chunk])
            want += e.flush()
            e.close()
            assert got[s] == want, "stream %d (%s) vs %s" % (s, kinds[s % 4], type(e).__name__)


@pytest.mark.parametrize("cfg", [
    dict(S=8, F=20, fpl=8, q=2), dict(S=4, F=20, fpl=4, q=0), dict(S=4, F=16, fpl=8, q=4, mode=0, sr=48000), dict(S=4, F=16, fpl=4, q=5, mode=3),
    dict(S=4, F=16, fpl=4, q=9), dict(S=4, F=16, fpl=4, q=3, quality=5, sr=32000), dict(S=4, F=16, fpl=4, q=4, sr=22050), dict(S=4, F=16, fpl=4, q=2, quality=0),
    dict(S=4, F=16, fpl=4, q=7, sr=8000), dict(S=4, F=16, fpl=4, q=2.5),
])
def test_vbr_old_batch_matches_oracle(lib, oracle_mod, cfg):
    """VBR-old (vbr_rh, quantize.c:1491) on lg_kernel_vbrold; the masking feedback evaluates exp/pow in double on the device
    (rounded to float, equal to the host's unless a double lands within an ulp of a float rounding boundary)"""
    S, F = cfg["S"], cfg["F"]
    sr, q, mode, quality = cfg.get("sr", 44100), cfg["q"], cfg.get("mode", -1), cfg.get("quality", -1)
    kinds = ("noise", "click", "sine", "gap")
    pcm = np.stack([make_signal(kinds[s % 4], F * 1152, seed=140 + s) for s in range(S)])
    enc = lib.BatchEncoder(S, sr, 2, q, mode, quality, frames_per_launch=cfg["fpl"], vbr=lib.VBR_RH)
    _, a = enc.encode(pcm)
    _, b = enc.flush()
    enc.close()
    for s in range(S):
        want = oracle_mod.PortEncoder(sr, 2, q, mode, quality, vbr=2).encode_all(pcm[s, 0], pcm[s, 1])
        if oracle_mod.have_ref():
            ref = oracle_mod.RefEncoder(sr, 2, q, mode if mode >= 0 else 4, quality, vbr=2).encode_all(pcm[s, 0], pcm[s, 1])
            assert want == ref, "oracle port and reference disagree"
        assert a[s] + b[s] == want, "stream %d (%s)" % (s, kinds[s % 4])


def test_config4_vbr_v2_full_size(lib, oracle_mod):
    """BASELINE configs[3]: VBR -V2 (vbr_mtrh) on the sine + noise mix at 2048 streams x 16 frames; every stream structurally
    (frame sync, frame length from its own bitrate index), a sample of streams byte for byte against the oracle - exact,
    not tolerance-matched."""
    S, F = 2048, 16
    n = F * 1152
    t = np.arange(n) / 44100.0
    base = np.stack([8000 * np.sin(2 * np.pi * 440 * t) + 4000 * np.sin(2 * np.pi * 3300 * t),
                     8000 * np.sin(2 * np.pi * 554.37 * t) + 3000 * np.sin(2 * np.pi * 7000 * t)])
    rng = np.random.default_rng(2025)
    pcm = np.rint(rng.integers(-1000, 1001, size=(S, 2, n)) + base[None]).astype(np.int16)
    enc = lib.BatchEncoder(S, 44100, 2, 2, -1, -1, frames_per_launch=F, vbr=lib.VBR_MTRH)
    n1, a = enc.encode(pcm)
    n2, b = enc.flush()
    enc.close()
    assert n1 + n2 == S * (F + 1)
    kbps = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]
    for s in range(S):
        mp3 = a[s] + b[s]
        pos, nfr = 0, 0
        while pos < len(mp3):
            assert mp3[pos] == 0xFF and mp3[pos + 1] == 0xFB, (s, pos)
            idx = mp3[pos + 2] >> 4
            assert 1 <= idx <= 14 and ((mp3[pos + 2] >> 1) & 1) == 0             # no padding in VBR
            pos += 144000 * kbps[idx] // 44100
            nfr += 1
        assert pos == len(mp3) and nfr == F + 1, s
    for s in list(range(0, S, 293)) + [S - 1]:
        assert a[s] + b[s] == oracle_mod.PortEncoder(44100, 2, 2, -1, -1, vbr=4).encode_all(pcm[s, 0], pcm[s, 1]), s


def test_lame_api_handle_matches_oracle(lib, oracle_mod):
    """the libmp3lame-compatible face: lame_init .. lame_encode_buffer .. lame_encode_flush on one handle"""
    x = make_signal("click", 50 * 1152, seed=3)
    enc = lib.Encoder(44100, 2, 128)
    out, pos = b"", 0
    while pos < x.shape[1]:
        out += enc.encode(x[0, pos:pos + 1152], x[1, pos:pos + 1152])     # frontend/lame_main.c feeds 1152-sample chunks
        pos += 1152
    out += enc.flush()
    assert enc.flush() == b""                                             # lame.c:2067: second flush returns 0
    enc.close()
    assert out == oracle_bytes(oracle_mod, x)


def test_info_tag_default_settings(lib, oracle_mod):
    """bWriteVbrTag = 1 is the reference's default: placeholder frame ahead of the audio and the finished Info tag frame
    (VbrTag.c) - against the committed reference vectors and, where its .so travelled, the reference itself"""
    tags = json.load(open(os.path.join(GOLD, "manifest_tag.json")))
    for name, m in sorted(tags.items()):
        x = make_signal(m["signal"], m["frames"] * 1152)
        e = lib.Encoder(m["samplerate"], 2, m["brate"], m["mode"] if m["mode"] >= 0 else lib.NOT_SET, m["quality"], write_tag=True,
                        vbr=m.get("vbr", 0))
        mp3 = b""
        for pos in range(0, x.shape[1], 4000):
            mp3 += e.encode(x[0, pos:pos + 4000], x[1, pos:pos + 4000])
        mp3 += e.flush()
        tag = e.lametag_frame()
        e.close()
        assert mp3 == open(os.path.join(GOLD, name + ".mp3"), "rb").read(), name
        assert tag == open(os.path.join(GOLD, name + ".tagframe"), "rb").read(), name
    if oracle_mod.have_ref():
        x = make_signal("noise", 450 * 1152, seed=77)            # > 400 frames: the seek table is decimated (VbrTag.c:140)
        e = lib.Encoder(44100, 2, 160, lib.NOT_SET, 5, write_tag=True)
        r = oracle_mod.RefEncoder(44100, 2, 160, 4, 5, write_tag=True)
        a = e.encode(x[0], x[1]) + e.flush()
        b = r.encode(x[0], x[1]) + r.flush()
        assert a == b and e.lametag_frame() == r.lametag_frame()
        e.close()
        r.close()


def test_sample_type_entry_points(lib, oracle_mod):
    """all nine lame_encode_buffer_* variants (lame.h:715-838) against libmp3lame itself"""
    import sys
    if not oracle_mod.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sample_types_check.py"), lib.LIB_PATH], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SAMPLE TYPES IDENTICAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_header_bits_and_tag_file(lib, oracle_mod):
    """the header bits a caller may set (copyright, original, emphasis, extension) and lame_mp3_tags_fid, against libmp3lame"""
    import sys
    if not oracle_mod.have_ref():
        pytest.skip("needs the prebuilt reference (oracle/_ref)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "header_bits_check.py"), lib.LIB_PATH], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "HEADER BITS IDENTICAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_edge_cases(lib, oracle_mod):
    # fewer samples than one frame, then flush: the encoder delay padding still yields complete frames
    for n in (0, 1, 575, 1151, 1376, 1377):
        x = make_signal("noise", max(n, 1), seed=n)[:, :n]
        enc = lib.BatchEncoder(1, frames_per_launch=4)
        _, a = enc.encode(x[None]) if n else (0, [b""])
        _, b = enc.flush()
        enc.close()
        if n == 0:
            # lame_encode_flush on a fresh handle still emits the delay frames (lame.c:2076-2117)
            assert a[0] + b[0] == oracle_mod.PortEncoder().flush()
        else:
            assert a[0] + b[0] == oracle_bytes(oracle_mod, x), n


def test_unsupported_configurations_fail_loudly(lib):
    for kw in (dict(out_samplerate=20000), dict(channels=3), dict(brate=10, vbr=4), dict(samplerate=0)):
        with pytest.raises(lib.LameGpuError):
            lib.BatchEncoder(2, **kw)
    L = lib.load_library()
    h = L.lame_init()
    import ctypes
    L.lame_set_free_format.argtypes, L.lame_set_free_format.restype = [ctypes.c_void_p, ctypes.c_int], ctypes.c_int
    assert L.lame_set_free_format(h, 1) == 0   # free format is not implemented: the option is carried, a non-default value refused
    assert L.lame_init_params(h) == -1
    L.lame_close(h)


def test_full_size_config_properties(lib, oracle_mod):
    """BASELINE configs[1] at full size (512 streams x 8 frames): every stream is checked structurally (frame sync,
    CBR frame sizes, decodable side info sizes) and a sample of streams byte for byte against the oracle."""
    S, F = 512, 8
    rng = np.random.default_rng(99)
    pcm = rng.integers(-12000, 12001, size=(S, 2, F * 1152), dtype=np.int16)
    enc = lib.BatchEncoder(S, frames_per_launch=F)
    n1, a = enc.encode(pcm)
    n2, b = enc.flush()
    enc.close()
    assert n1 + n2 == S * (F + 1)                       # 8 input frames -> 9 output frames (encoder delay + padding)
    sizes = set()
    for s in range(S):
        mp3 = a[s] + b[s]
        sizes.add(len(mp3))
        pos, nfr = 0, 0
        while pos < len(mp3):
            assert mp3[pos] == 0xFF and mp3[pos + 1] == 0xFB, (s, pos)
            assert (mp3[pos + 2] >> 4) == 9              # 128 kbps
            pos += 417 + ((mp3[pos + 2] >> 1) & 1)
            nfr += 1
        assert pos == len(mp3) and nfr == F + 1
    assert len(sizes) == 1                               # CBR: identical length for every stream
    for s in list(range(0, S, 37)) + [S - 1]:
        assert a[s] + b[s] == oracle_bytes(oracle_mod, pcm[s]), s


def test_config2_cbr320_joint_stereo_full_size(lib, oracle_mod):
    """BASELINE configs[2] at full size: 65 536 frames = 2048 streams x 32 frames of the sine + noise mix, CBR 320 kbps
    joint stereo (escape Huffman tables, the masking_lower feedback of SURVEY section 7).  Every stream structurally,
    a sample of streams byte for byte against the oracle."""
    S, F = 2048, 32
    n = F * 1152
    t = np.arange(n) / 44100.0
    base_l = 8000 * np.sin(2 * np.pi * 440 * t) + 4000 * np.sin(2 * np.pi * 3300 * t)
    base_r = 8000 * np.sin(2 * np.pi * 554.37 * t) + 3000 * np.sin(2 * np.pi * 7000 * t)
    rng = np.random.default_rng(2024)
    noise = rng.integers(-1000, 1001, size=(S, 2, n))
    pcm = np.rint(noise + np.stack([base_l, base_r])[None]).astype(np.int16)
    enc = lib.BatchEncoder(S, 44100, 2, 320, 1, -1, frames_per_launch=F)
    n1, a = enc.encode(pcm)
    n2, b = enc.flush()
    enc.close()
    assert n1 + n2 == S * (F + 1)
    for s in range(S):
        mp3 = a[s] + b[s]
        pos, nfr = 0, 0
        while pos < len(mp3):
            assert mp3[pos] == 0xFF and mp3[pos + 1] == 0xFB and (mp3[pos + 2] >> 4) == 14, (s, pos)     # 320 kbps
            assert (mp3[pos + 3] >> 6) == 1                                                              # joint stereo
            pos += 1044 + ((mp3[pos + 2] >> 1) & 1)
            nfr += 1
        assert pos == len(mp3) and nfr == F + 1, s
    for s in list(range(0, S, 293)) + [S - 1]:
        assert a[s] + b[s] == oracle_bytes(oracle_mod, pcm[s], 44100, 320, 1, -1), s


def test_c_harness_ragged_streams():
    """tests/c/gpu_vs_port.c drives the pointer-array ABI (lamegpu_batch_encode) with per-stream buffers"""
    exe = os.path.join(ROOT, "tests", "c", "bin", "gpu_vs_port")
    if not os.path.exists(exe):
        import __graft_entry__ as ge
        ge.build_test_binaries()
    r = subprocess.run([exe, "24", "20", "6", "128", "-1", "-1", "44100", "2500"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "IDENTICAL" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_device_libm_restatements_match_host_libm(lib):
    """lg_powf / lg_log10f / lg_exp / lg_pow (csrc/lg_math.cuh) evaluated ON THE DEVICE against this host's glibc, argument by argument:
    powf in athAdjust / NS_INTERP, log10f in calc_scalefac (vbrquantize.c:317), exp and pow in the masking feedback of VBR-old
    (quantize.c:1419-1426).  The CPU suite checks the same sources compiled for the host on 2e8 .. 2e9 arguments (tests/test_powf.py)."""
    import ctypes
    import math
    L = lib.load_library()
    dp = ctypes.POINTER(ctypes.c_double)
    L.lamegpu_math_selftest.argtypes, L.lamegpu_math_selftest.restype = [ctypes.c_int, dp, dp, dp, ctypes.c_int], ctypes.c_int
    libm = ctypes.CDLL("libm.so.6")
    libm.powf.argtypes, libm.powf.restype = [ctypes.c_float, ctypes.c_float], ctypes.c_float
    libm.log10f.argtypes, libm.log10f.restype = [ctypes.c_float], ctypes.c_float
    rng = np.random.default_rng(5)
    n = 20000

    def device(fn, x, y):
        x, y, out = np.ascontiguousarray(x, np.float64), np.ascontiguousarray(y, np.float64), np.empty(len(x), np.float64)
        assert L.lamegpu_math_selftest(fn, x.ctypes.data_as(dp), y.ctypes.data_as(dp), out.ctypes.data_as(dp), len(x)) == 0
        return out

    # powf: base 10 with athAdjust's exponents, and ratios over the float range with exponents in (0, 1)
    x = np.concatenate([np.full(n // 2, 10.0, np.float32), rng.integers(1, 0x7f800000, n // 2).astype(np.uint32).view(np.float32)])
    y = np.concatenate([(rng.uniform(-40, 40, n // 2).astype(np.float32) * np.float32(0.1)), rng.uniform(0, 1, n // 2).astype(np.float32)])
    want = np.array([libm.powf(float(a), float(b)) for a, b in zip(x, y)], np.float32)
    assert np.array_equal(device(0, x, y).astype(np.float32).view(np.uint32), want.view(np.uint32))
    # log10f: positive floats of every exponent, subnormals included
    x = rng.integers(1, 0x7f800000, n).astype(np.uint32).view(np.float32)
    want = np.array([libm.log10f(float(a)) for a in x], np.float32)
    assert np.array_equal(device(1, x, x).astype(np.float32).view(np.uint32), want.view(np.uint32))
    # exp(3.5 - pe / 300.) and pow(10, db * 0.1) as VBR-old forms them, plus a wider range
    pe = rng.uniform(0, 6000, n).astype(np.float32)
    x = np.concatenate([3.5 - pe.astype(np.float64) / 300., rng.uniform(-500, 500, n)])
    want = np.array([math.exp(a) for a in x], np.float64)
    assert np.array_equal(device(2, x, x).view(np.uint64), want.view(np.uint64))
    db = rng.uniform(-12, 12, n).astype(np.float32)
    y = np.concatenate([db.astype(np.float64) * 0.1, rng.uniform(-100, 100, n)])
    x = np.concatenate([np.full(n, 10.0), rng.uniform(0.001, 1000, n)])
    want = np.array([math.pow(a, b) for a, b in zip(x, y)], np.float64)
    assert np.array_equal(device(3, x, y).view(np.uint64), want.view(np.uint64))


# ---------------------------------------------------------------- round 2
def reference_all(oracle_mod, pcm, **kw):
    """the reference on every stream of `pcm` [S][2][n], on all host cores (ctypes releases the GIL inside libmp3lame)"""
    from concurrent.futures import ThreadPoolExecutor
    Enc = oracle_mod.RefEncoder if oracle_mod.have_ref() else oracle_mod.PortEncoder
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        return list(pool.map(lambda s: Enc(**kw).encode_all(pcm[s, 0], pcm[s, 1]), range(pcm.shape[0])))


def tones(S, n, seed):
    t = np.arange(n) / 44100.0
    base = np.stack([8000 * np.sin(2 * np.pi * 440 * t) + 4000 * np.sin(2 * np.pi * 3300 * t), 8000 * np.sin(2 * np.pi * 554.37 * t) + 3000 * np.sin(2 * np.pi * 7000 * t)])
    rng = np.random.default_rng(seed)
    out = np.empty((S, 2, n), dtype=np.int16)
    for s0 in range(0, S, 256):
        s1 = min(S, s0 + 256)
        out[s0:s1] = np.rint(rng.integers(-1000, 1001, size=(s1 - s0, 2, n)) + base[None])
    return out


def test_config1_every_stream_byte_exact(lib, oracle_mod):
    """BASELINE configs[1] (512 streams x 8 frames, white noise, CBR 128): EVERY stream against the reference, two calls + flush, pipelined"""
    S, F = 512, 8
    rng = np.random.default_rng(4321)
    pcm = rng.integers(-12000, 12001, size=(S, 2, 2 * F * 1152), dtype=np.int16)
    enc = lib.BatchEncoder(S, frames_per_launch=F)
    enc.set_pipelined(True)
    _, a = enc.encode(pcm[:, :, :F * 1152])
    _, b = enc.encode(pcm[:, :, F * 1152:])
    _, c = enc.flush()
    enc.close()
    want = reference_all(oracle_mod, pcm, samplerate=44100, channels=2, brate=128)
    bad = [s for s in range(S) if a[s] + b[s] + c[s] != want[s]]
    assert not bad, bad[:8]


def test_config3_vbr_v2_262144_frames_every_stream_byte_exact(lib, oracle_mod):
    """BASELINE configs[3] at its stated size: 262 144 frames = 4096 streams x 64 frames, VBR -V2 (vbr_mtrh), the tone + noise mix; every stream
    byte for byte against the reference - exact, not tolerance-matched"""
    S, F = 4096, 64
    pcm = tones(S, F * 1152, 31)
    enc = lib.BatchEncoder(S, 44100, 2, 2, -1, -1, frames_per_launch=F, vbr=lib.VBR_MTRH)
    n1, a = enc.encode(pcm)
    n2, b = enc.flush()
    enc.close()
    assert n1 + n2 == S * (F + 1)
    want = reference_all(oracle_mod, pcm, samplerate=44100, channels=2, brate=2, vbr=4)
    bad = [s for s in range(S) if a[s] + b[s] != want[s]]
    assert not bad, (len(bad), bad[:8])


def test_config2_every_stream_byte_exact(lib, oracle_mod):
    """BASELINE configs[2]: 65 536 frames = 2048 streams x 32 frames, CBR 320 joint stereo; every stream against the reference"""
    S, F = 2048, 32
    pcm = tones(S, F * 1152, 77)
    enc = lib.BatchEncoder(S, 44100, 2, 320, 1, -1, frames_per_launch=F)
    _, a = enc.encode(pcm)
    _, b = enc.flush()
    enc.close()
    want = reference_all(oracle_mod, pcm, samplerate=44100, channels=2, brate=320, mode=1)
    bad = [s for s in range(S) if a[s] + b[s] != want[s]]
    assert not bad, (len(bad), bad[:8])


def test_testcase_wav_full_length_with_tag(lib, oracle_mod):
    """configs[0]: all 25 000 samples of the reference's testcase.wav through the libmp3lame face with the Info tag (the file `lame -b 128` writes)"""
    if not oracle_mod.have_ref():
        pytest.skip("needs the reference build")
    x = make_signal("testcase", 25000)
    r = oracle_mod.RefEncoder(44100, 2, 128, write_tag=True)
    want = r.encode(x[0], x[1]) + r.flush()
    want_tag = r.lametag_frame()
    r.close()
    e = lib.Encoder(44100, 2, 128, write_tag=True)
    got = b""
    for pos in range(0, 25000, 1152):
        got += e.encode(x[0, pos:pos + 1152], x[1, pos:pos + 1152])
    got += e.flush()
    tag = e.lametag_frame()
    e.close()
    assert got == want and tag == want_tag and len(got) == 10030


@pytest.mark.parametrize("kw", [dict(brate=128), dict(brate=320, mode=1), dict(brate=2, vbr=4), dict(brate=128, vbr=3), dict(brate=64)])
def test_encoding_continues_after_a_flush(lib, oracle_mod, kw):
    """lame_encode_flush ends the bit reservoir (flush_bitstream, bitstream.c:886-889: ResvSize = 0, main_data_begin = 0) - on the device too:
    samples fed after a flush are encoded as the reference encodes them (advisor finding of round 1: main_data_begin went stale)"""
    if not oracle_mod.have_ref():
        pytest.skip("needs the reference build")
    x = make_signal("click", 40 * 1152, seed=11)
    cut = 17 * 1152 + 300
    r = oracle_mod.RefEncoder(44100, 2, **kw)
    want = r.encode(x[0, :cut], x[1, :cut]) + r.flush() + r.encode(x[0, cut:], x[1, cut:]) + r.flush()
    r.close()
    e = lib.Encoder(44100, 2, kw["brate"], vbr=kw.get("vbr", 0), mode=kw.get("mode", -1))
    got = e.encode(x[0, :cut], x[1, :cut]) + e.flush() + e.encode(x[0, cut:], x[1, cut:]) + e.flush()
    e.close()
    assert got == want
    S = 6
    pcm = np.stack([make_signal(("click", "noise", "gap")[s % 3], 40 * 1152, seed=20 + s) for s in range(S)])
    b = lib.BatchEncoder(S, 44100, 2, kw["brate"], kw.get("mode", -1), -1, frames_per_launch=5, vbr=kw.get("vbr", 0))
    _, p1 = b.encode(pcm[:, :, :cut]); _, p2 = b.flush(); _, p3 = b.encode(pcm[:, :, cut:]); _, p4 = b.flush()
    b.close()
    for s in range(S):
        r = oracle_mod.RefEncoder(44100, 2, **kw)
        w = r.encode(pcm[s, 0, :cut], pcm[s, 1, :cut]) + r.flush() + r.encode(pcm[s, 0, cut:], pcm[s, 1, cut:]) + r.flush()
        r.close()
        assert p1[s] + p2[s] + p3[s] + p4[s] == w, s


def test_vbr_level_9_with_explicit_output_rate(lib, oracle_mod):
    """VBR-new level 9 kept at 44.1 kHz by lame_set_out_samplerate: the preset interpolates towards row 10 of vbr_mt_psy_switch_map (presets.c:124)"""
    if not oracle_mod.have_ref():
        pytest.skip("needs the reference build")
    x = make_signal("click", 20 * 1152, seed=5)
    for q in (9, 8.5):
        want = oracle_mod.RefEncoder(44100, 2, q, vbr=4, out_samplerate=44100).encode_all(x[0], x[1])
        enc = lib.BatchEncoder(1, 44100, 2, q, -1, -1, frames_per_launch=8, vbr=lib.VBR_MTRH, out_samplerate=44100)
        _, a = enc.encode(x[None]); _, b = enc.flush(); enc.close()
        assert a[0] + b[0] == want, q


def test_engines_driven_from_other_threads_and_side_by_side(lib, oracle_mod):
    """an engine made on one thread and used from another, and two 512-stream encoders running concurrently on one device (round 1: kernels
    waiting inside the kernel for flags could fill the SMs; now nothing waits on the device) - every entry point selects the engine's device"""
    import threading
    S, F = 512, 8
    rng = np.random.default_rng(77)
    pcm = [rng.integers(-12000, 12001, size=(S, 2, F * 1152), dtype=np.int16) for _ in range(2)]
    encs = [lib.BatchEncoder(S, frames_per_launch=F) for _ in range(2)]          # made on the main thread
    res = [None, None]

    def work(i):
        out = [b""] * S
        for rep in range(3):
            _, a = encs[i].encode(pcm[i])
            out = [o + x for o, x in zip(out, a)]
        _, b = encs[i].flush()
        res[i] = [o + x for o, x in zip(out, b)]

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in encs:
        e.close()
    for i in range(2):
        assert res[i] is not None
        for s in (0, 100, 511):
            x = np.concatenate([pcm[i][s]] * 3, axis=1)
            assert res[i][s] == oracle_bytes(oracle_mod, x), (i, s)


def test_batch_over_all_visible_gpus(lib, oracle_mod):
    """device = -1: one batch, every GPU of the process, streams in contiguous shares (needs two GPUs: gpurun --gpus 2)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible")
    S, F = 64, 6
    pcm = np.stack([make_signal(("noise", "click", "sine", "gap")[s % 4], F * 1152, seed=300 + s) for s in range(S)])
    enc = lib.BatchEncoder(S, frames_per_launch=4, device=-1)
    assert lib.load_library().lamegpu_batch_devices(enc._h) == torch.cuda.device_count()
    _, a = enc.encode(pcm); _, b = enc.flush(); enc.close()
    for s in range(S):
        assert a[s] + b[s] == oracle_bytes(oracle_mod, pcm[s]), s
