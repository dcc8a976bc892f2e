"""CPU suite: pins the oracle.  (a) oracle/port reproduces every committed golden vector (MP3 bytes produced by the
unmodified reference, tests/golden/make_golden.py); (b) where the reference build oracle/_ref exists (development
container and, as a prebuilt .so, the GPU box) the port is compared with the reference itself on more inputs and on
the init tables / per-frame state through tests/c/port_vs_ref.c."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_signal

GOLD = os.path.join(ROOT, "tests", "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_port_reproduces_golden(oracle_mod, name):
    m = MANIFEST[name]
    x = make_signal(m["signal"], m["frames"] * 1152)
    mp3 = oracle_mod.PortEncoder(m["samplerate"], 2, m["brate"], m["mode"], m["quality"], out_samplerate=m.get("out_samplerate", 0)).encode_all(x[0], x[1])
    want = open(os.path.join(GOLD, name + ".mp3"), "rb").read()
    assert len(mp3) == m["nbytes"]
    assert mp3 == want


def test_port_chunking_invariance(oracle_mod):
    """feeding the same PCM in ragged chunks gives the same stream (buffering logic, lame.c:1671)"""
    x = make_signal("click", 30 * 1152)
    whole = oracle_mod.PortEncoder().encode_all(x[0], x[1])
    e = oracle_mod.PortEncoder()
    out, pos = b"", 0
    for c in [1, 7, 1151, 1152, 1153, 5000, 333, 10 ** 6]:
        out += e.encode(x[0][pos:pos + c], x[1][pos:pos + c])
        pos += c
        if pos >= x.shape[1]:
            break
    out += e.flush()
    assert out == whole


def test_port_rejects_unsupported(oracle_mod):
    # not an MPEG rate, three channels, VBR level out of range, not a vbr_mode
    for kw in (dict(out_samplerate=20000), dict(channels=3), dict(brate=10, vbr=4), dict(vbr=5)):
        with pytest.raises(ValueError):
            oracle_mod.PortEncoder(**kw)


@pytest.fixture(scope="module")
def port_vs_ref_bin(oracle_mod, tmp_path_factory):
    if not oracle_mod.have_ref() or not os.path.exists(oracle_mod.REFDUMP_SO):
        pytest.skip("reference build oracle/_ref not present")
    out = str(tmp_path_factory.mktemp("bin") / "port_vs_ref")
    subprocess.run(["gcc", "-O2", "-fno-fast-math", "-ffp-contract=off", "-w", os.path.join(ROOT, "tests/c/port_vs_ref.c"), "-o", out,
                    "-L" + os.path.join(ROOT, "oracle"), "-llameport", oracle_mod.REFDUMP_SO, "-lm",
                    "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_ref")], check=True)
    return out


@pytest.mark.parametrize("args", [
    "noise 128 -1 -1 120", "sine 128 -1 -1 100", "click 128 -1 -1 150", "gap 128 -1 -1 60", "sine 320 1 -1 100",
    "click 320 1 -1 100", "noise 192 0 -1 60", "click 160 -1 5 80", "click 128 -1 7 60", "sine 128 -1 4 60",
    "click 256 -1 -1 60 48000", "click 128 -1 -1 60 32000", "silence 128 -1 -1 20", "click 224 0 6 60", "noise 112 -1 9 40",
    "click 128 2 -1 80", "click 192 2 5 60 48000", "click 64 2 -1 60 22050",        # dual channel (mode 2)
])
def test_port_vs_reference(port_vs_ref_bin, args):
    """byte-identical MP3 + identical init tables against the real libmp3lame (strict IEEE build)"""
    r = subprocess.run([port_vs_ref_bin] + args.split(), capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "setup tables: identical" in r.stdout and "IDENTICAL" in r.stdout.splitlines()[-1]


@pytest.mark.parametrize("args", [
    "noise 128 -1 -1 80", "click 128 -1 -1 120", "sine 160 -1 -1 60", "gap 192 0 -1 60", "noise 140 -1 5 40", "sine 256 1 -1 40",
    "silence 128 -1 -1 20", "click 320 -1 -1 40 48000", "click 112 -1 -1 40 32000", "click 200 -1 7 40",
])
def test_port_abr_vs_reference(port_vs_ref_bin, args):
    """ABR (lame_set_VBR(vbr_abr) + mean bitrate; quantize.c:1900): variable frame sizes, byte-identical to libmp3lame"""
    r = subprocess.run([port_vs_ref_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, LP_VBR="3"))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "setup tables: identical" in r.stdout and "IDENTICAL" in r.stdout.splitlines()[-1]


@pytest.mark.parametrize("args", [
    "noise 2 -1 -1 80", "click 2 -1 -1 150", "sine 2 -1 -1 80", "gap 2 -1 -1 60", "silence 2 -1 -1 20", "click 0 -1 -1 100", "noise 0 -1 -1 40",
    "sine 4 -1 -1 60", "click 5 -1 -1 80", "noise 6 -1 -1 40", "click 3 0 -1 60", "click 1 1 -1 60", "click 2 -1 5 60", "sine 2 -1 7 60",
    "click 2 -1 -1 60 48000", "noise 0 -1 -1 40 48000", "click 4 3 -1 60", "click 0 3 -1 100",
])
def test_port_vbr_new_vs_reference(port_vs_ref_bin, args):
    """VBR-new (vbr_mtrh, lame_set_VBR_q = the 2nd argument; vbrquantize.c + quantize.c:1645), including the frames that do not fit
    and go through outOfBitsStrategy (click at -V0): byte-identical to libmp3lame"""
    r = subprocess.run([port_vs_ref_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, LP_VBR="4"))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "setup tables: identical" in r.stdout and "IDENTICAL" in r.stdout.splitlines()[-1]


@pytest.mark.parametrize("args,env", [
    ("click 128 -1 -1 60 48000", dict(LP_OUT_SR="44100")), ("click 192 -1 -1 60 44100", dict(LP_OUT_SR="48000")),
    ("sine 128 -1 -1 60 44100", dict(LP_OUT_SR="32000", LP_CHUNK="777")), ("noise 128 -1 -1 30 8000", dict(LP_OUT_SR="44100", LP_CHUNK="100")),
    ("click 128 -1 -1 60 192000", dict(LP_OUT_SR="32000", LP_CHUNK="5000")), ("sine 128 -1 -1 60 96000", {}), ("click 112 -1 -1 60 48000", {}),
    ("click 96 -1 -1 60 44100", {}), ("noise 80 -1 -1 60 44100", {}), ("click 64 3 -1 60 44100", {}), ("click 160 -1 -1 40 37800", {}),
    ("click 48 -1 -1 40 44100", dict(LP_OUT_SR="32000")), ("click 2 -1 -1 60 44100", dict(LP_VBR="4", LP_OUT_SR="32000")),
    ("click 2 -1 -1 60 88200", dict(LP_VBR="4")), ("click 112 -1 -1 60 48000", dict(LP_VBR="3", LP_CHUNK="4000")),
    ("noise 128 -1 -1 8 48000", dict(LP_OUT_SR="44100", LP_CHUNK="7")),
])
def test_port_resampler_vs_reference(port_vs_ref_bin, args, env):
    """input rate != output rate (explicit lame_set_out_samplerate = LP_OUT_SR, or the rate lame_init_params picks): the
    polyphase resampler util.c:531, its per-call state (LP_CHUNK = samples per encode call) and the flush padding rule
    lame.c:2083-2100; also the low bitrates that only exist through it (96 kbps at 44.1 kHz -> 32 kHz).  Byte-identical."""
    r = subprocess.run([port_vs_ref_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "setup tables: identical" in r.stdout and "IDENTICAL" in r.stdout.splitlines()[-1]


@pytest.mark.parametrize("args,env", [
    ("noise 128 -1 2 60", {}), ("click 128 -1 2 100", {}), ("sine 128 -1 1 60", {}), ("click 128 -1 1 100", {}), ("click 128 -1 0 100", {}),
    ("noise 128 -1 0 40", {}), ("gap 192 0 0 60", {}), ("click 320 1 1 60", {}), ("click 160 -1 2 60 48000", {}), ("sine 112 -1 0 40 32000", {}),
    ("click 128 -1 0 60", dict(LP_VBR="3")), ("noise 160 -1 2 40", dict(LP_VBR="3")),
])
def test_port_quality_0_to_2_vs_reference(port_vs_ref_bin, args, env):
    """quality 2 / 1 / 0 (the 4th argument): substep shaping with the per-band half-step flags (quantize.c:131,781, takehiro.c:781),
    one-band-at-a-time amplification (noise_shaping_amp 2) and the full outer loop; CBR and ABR, byte-identical to libmp3lame"""
    r = subprocess.run([port_vs_ref_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "setup tables: identical" in r.stdout and "IDENTICAL" in r.stdout.splitlines()[-1]


@pytest.mark.parametrize("args", ["click 2 -1 7 60 44100", "noise 4 -1 8 40 44100", "sine 0 -1 9 40 48000", "gap 5 3 7 40 32000",
                                  "click 6 0 7 60 44100", "click 4 -1 7 40 22050"])
def test_port_vbr_new_quality_7_to_9_vs_reference(port_vs_ref_bin, args):
    """VBR-new with quality 7-9 (the 4th argument): guess_scalefac_x34 (vbrquantize.c:324) instead of the per-band step search -
    one log10f per band, the host's libm here as in the reference"""
    r = subprocess.run([port_vs_ref_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, LP_VBR="4"))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "setup tables: identical" in r.stdout and "IDENTICAL" in r.stdout.splitlines()[-1]


@pytest.mark.parametrize("args,env", [
    ("click 7 -1 -1 60 44100", {}), ("noise 2 -1 -1 40 32000", {}), ("click 5 -1 -1 60 32000", {}), ("click 2 -1 -1 60 44100", dict(LP_VBRQ_FRAC="0.5")),
    ("sine 5 -1 -1 60 48000", dict(LP_VBRQ_FRAC="0.3")), ("click 6 -1 -1 60 44100", dict(LP_VBRQ_FRAC="0.9")), ("gap 0 -1 -1 60 44100", dict(LP_VBRQ_FRAC="0.77")),
    ("click 3 0 -1 60 32000", dict(LP_VBRQ_FRAC="0.4")), ("click 7 -1 -1 60 96000", dict(LP_VBRQ_FRAC="0.999")),
    ("click 6 -1 -1 40 44100", dict(LP_VBRQ_FRAC="0.25", LP_OUT_SR="44100")),
])
def test_port_fractional_vbr_quality_vs_reference(port_vs_ref_bin, args, env):
    """VBR levels between the presets (lame_set_VBR_quality = 2nd argument + LP_VBRQ_FRAC; presets.c:143 interpolation) and the
    mapping of the -V scale to the output rate it implies (lame.c:661-698): -V7 at 44.1 kHz = quality 5.63 at 32 kHz through the
    resampler, levels at 32 kHz input shrunk to 0..5.2"""
    r = subprocess.run([port_vs_ref_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, LP_VBR="4", **env))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "setup tables: identical" in r.stdout and "IDENTICAL" in r.stdout.splitlines()[-1]


@pytest.mark.parametrize("args,env", [
    ("noise 64 -1 -1 40 24000", {}), ("click 64 -1 -1 80 22050", {}), ("sine 96 0 -1 40 24000", {}), ("click 32 -1 -1 40 16000", {}), ("click 160 -1 -1 40 24000", {}),
    ("click 24 -1 -1 40 8000", {}), ("noise 32 3 -1 40 11025", {}), ("gap 8 3 -1 40 12000", {}), ("click 40 -1 -1 60 12000", {}), ("sine 16 -1 5 40 8000", {}),
    ("click 80 1 2 40 22050", {}), ("sine 56 -1 0 40 16000", {}), ("click 64 -1 -1 60 44100", {}), ("sine 32 -1 -1 60 44100", {}), ("click 8 3 -1 60 44100", {}),
    ("click 48 -1 -1 60 44100", dict(LP_OUT_SR="16000", LP_CHUNK="777")), ("click 64 -1 -1 60 22050", dict(LP_VBR="3")), ("noise 24 -1 -1 40 8000", dict(LP_VBR="3")),
    ("click 8 -1 -1 60 44100", dict(LP_VBR="4")), ("click 9 -1 -1 60 44100", dict(LP_VBR="4")), ("click 4 -1 -1 60 22050", dict(LP_VBR="4")),
    ("noise 2 -1 -1 60 16000", dict(LP_VBR="4")), ("click 5 -1 -1 60 8000", dict(LP_VBR="4")), ("click 9 -1 -1 60 48000", dict(LP_VBR="4", LP_VBRQ_FRAC="0.5")),
])
def test_port_mpeg2_and_mpeg25_vs_reference(port_vs_ref_bin, args, env):
    """MPEG-2 (16/22.05/24 kHz) and MPEG-2.5 (8/11.025/12 kHz) output: one granule per frame, LSF scalefactor partitions
    (takehiro.c:1218), 8-bit main_data_begin side info, the 8 kHz band limits; CBR, ABR, VBR (-V8/-V9 through the resampler),
    byte-identical to libmp3lame"""
    r = subprocess.run([port_vs_ref_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "setup tables: identical" in r.stdout and "IDENTICAL" in r.stdout.splitlines()[-1]


@pytest.mark.parametrize("args,env", [
    ("noise 2 -1 -1 60", {}), ("click 2 -1 -1 100", {}), ("sine 4 -1 -1 60", {}), ("click 0 -1 -1 80", {}), ("gap 6 -1 -1 60", {}), ("click 2 0 -1 60 48000", {}),
    ("click 5 3 -1 60", {}), ("noise 9 -1 -1 40", {}), ("click 3 -1 5 60 32000", {}), ("click 4 -1 -1 60 22050", {}), ("click 2 -1 0 40", {}),
    ("click 7 -1 -1 60 8000", {}), ("click 2 -1 -1 60", dict(LP_VBRQ_FRAC="0.5")),
])
def test_port_vbr_old_vs_reference(port_vs_ref_bin, args, env):
    """VBR-old (vbr_rh, quantize.c:1491): per gr.ch bisection of the bit budget over outer_loop, masking lowered by the perceptual
    entropy, the sfb21 analog-silence cut, bit-pressure rounds; all MPEG versions, quality 0, a fractional level"""
    r = subprocess.run([port_vs_ref_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, LP_VBR="2", **env))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "setup tables: identical" in r.stdout and "IDENTICAL" in r.stdout.splitlines()[-1]


@pytest.mark.parametrize("args,env", [("click 128 -1 -1 60", {}), ("noise 128 3 -1 40", {}), ("click 64 -1 -1 60 22050", {}), ("click 2 -1 -1 60", dict(LP_VBR="4"))])
def test_port_crc_vs_reference(port_vs_ref_bin, args, env):
    """error_protection: the CRC-16 behind every frame header (bitstream.c:304), two more bytes of side info in every budget"""
    r = subprocess.run([port_vs_ref_bin] + args.split(), capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, LP_CRC="1", **env))
    assert r.returncode == 0, r.stdout[-2000:]
    assert "IDENTICAL" in r.stdout.splitlines()[-1]


def test_click_signal_has_short_blocks(oracle_mod):
    """the transient fixture must really exercise block switching, otherwise short-block parity is vacuous"""
    x = make_signal("click", 40 * 1152)
    mp3 = oracle_mod.PortEncoder().encode_all(x[0], x[1])
    # parse side info: window_switching_flag/block_type of gr0 ch0 in every frame (MPEG-1 stereo, no CRC)
    pos, types = 0, set()
    while pos + 40 < len(mp3):
        assert mp3[pos] == 0xFF and (mp3[pos + 1] & 0xE0) == 0xE0
        pad = (mp3[pos + 2] >> 1) & 1
        bits = int.from_bytes(mp3[pos + 4:pos + 36], "big")
        off = 9 + 3 + 8 + 12 + 9 + 8 + 4          # main_data_begin, private, scfsi, part2_3, big_values, global_gain, sf_compress
        ws = (bits >> (256 - off - 1)) & 1
        if ws:
            types.add((bits >> (256 - off - 3)) & 3)
        pos += 417 + pad
    assert 2 in types and 1 in types and 3 in types
