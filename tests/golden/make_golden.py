"""Generates the committed golden vectors from the UNMODIFIED reference (oracle/_ref, built from
/root/reference).  Run in the development container only:  python tests/golden/make_golden.py

For every case: the PCM comes from the seeded generators in tests/conftest.py (or testcase.wav shipped by
the reference, stored once as testcase_pcm.npy), the MP3 bytes come from the reference's own
lame_init / lame_encode_buffer / lame_encode_flush.  manifest.json records config, sizes and SHA-256.
"""
import hashlib
import json
import os
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from conftest import make_signal  # noqa: E402

CASES = [
    # name, signal, frames, samplerate, brate, mode, quality
    ("testcase_128", "testcase", 21, 44100, 128, -1, -1),
    ("noise_128", "noise", 24, 44100, 128, -1, -1),
    ("click_128", "click", 40, 44100, 128, -1, -1),
    ("gap_128", "gap", 30, 44100, 128, -1, -1),
    ("sine_320_js", "sine", 24, 44100, 320, 1, -1),
    ("click_192_stereo_q5", "click", 24, 44100, 192, 0, 5),
    ("noise_256_48k", "noise", 16, 48000, 256, -1, -1),
    ("click_160_q7", "click", 16, 44100, 160, -1, 7),
    # input rate != MPEG output rate: the reference's polyphase resampler (util.c:531) in front; an 8th field is an explicit
    # lame_set_out_samplerate, otherwise lame_init_params picks the rate (96 kbps at 44.1 kHz -> 32 kHz, 112 kbps at 48 kHz -> 44.1 kHz)
    ("click_128_48k_to_44k", "click", 20, 48000, 128, -1, -1, 44100),
    ("sine_96_44k_auto32k", "sine", 20, 44100, 96, -1, -1),
    ("noise_112_48k_auto44k", "noise", 16, 48000, 112, -1, -1),
    ("gap_160_37800_to_48k", "gap", 16, 37800, 160, -1, -1, 48000),
    # MPEG-2 (22.05 kHz input; 64 kbps at 44.1 kHz -> 24 kHz) and MPEG-2.5 (8 kHz, mono): one granule per frame, LSF scalefactors
    ("click_64_22k", "click", 20, 22050, 64, -1, -1),
    ("sine_64_44k_auto24k", "sine", 20, 44100, 64, -1, -1),
    ("noise_16_8k_mono", "noise", 12, 8000, 16, 3, -1),
]


TAG_CASES = [
    # name, signal, frames, samplerate, brate (CBR bitrate, ABR mean or VBR_q), mode, quality, vbr (0 = vbr_off, 3 = vbr_abr, 4 = vbr_mtrh)
    ("tag_noise_128", "noise", 30, 44100, 128, -1, -1, 0),
    ("tag_click_192_stereo", "click", 45, 44100, 192, 0, -1, 0),
    ("tag_sine_320_js", "sine", 20, 44100, 320, 1, -1, 0),
    ("tag_abr_click_128", "click", 45, 44100, 128, -1, -1, 3),
    ("tag_abr_gap_150_stereo", "gap", 30, 44100, 150, 0, -1, 3),
    # vbr 4 = vbr_mtrh: "brate" is VBR_q (-V2, -V0)
    ("tag_vbr_v2_noise", "noise", 30, 44100, 2, -1, -1, 4),
    ("tag_vbr_v0_click", "click", 45, 44100, 0, -1, -1, 4),
    ("tag_vbr_v4_sine_stereo", "sine", 24, 44100, 4, 0, -1, 4),
]


def main():
    assert oracle.build()[1], "reference not available"
    w = wave.open("/root/reference/testcase.wav", "rb")
    assert w.getnchannels() == 2 and w.getsampwidth() == 2
    pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).reshape(-1, 2).T.copy()
    np.save(os.path.join(HERE, "testcase_pcm.npy"), pcm)
    manifest = {}
    for name, sig, frames, sr, brate, mode, q, *rest in CASES:
        out_sr = rest[0] if rest else 0
        x = make_signal(sig, frames * 1152)
        mp3 = oracle.RefEncoder(sr, 2, brate, mode if mode >= 0 else 4, q, out_samplerate=out_sr).encode_all(x[0], x[1])
        with open(os.path.join(HERE, name + ".mp3"), "wb") as f:
            f.write(mp3)
        manifest[name] = dict(signal=sig, frames=frames, samplerate=sr, out_samplerate=out_sr, brate=brate, mode=mode, quality=q, nbytes=len(mp3),
                              pcm_sha256=hashlib.sha256(x.tobytes()).hexdigest(), mp3_sha256=hashlib.sha256(mp3).hexdigest())
        print(name, len(mp3))
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    # Info tag (lame_set_bWriteVbrTag(1), the reference's default): the stream with its placeholder frame, and the
    # finished tag frame of lame_get_lametag_frame
    tags = {}
    for name, sig, frames, sr, brate, mode, q, vbr in TAG_CASES:
        x = make_signal(sig, frames * 1152)
        r = oracle.RefEncoder(sr, 2, brate, mode if mode >= 0 else 4, q, write_tag=True, vbr=vbr)
        mp3 = b""
        for pos in range(0, x.shape[1], 4000):
            mp3 += r.encode(x[0, pos:pos + 4000], x[1, pos:pos + 4000])
        mp3 += r.flush()
        tag = r.lametag_frame()
        r.close()
        with open(os.path.join(HERE, name + ".mp3"), "wb") as f:
            f.write(mp3)
        with open(os.path.join(HERE, name + ".tagframe"), "wb") as f:
            f.write(tag)
        tags[name] = dict(signal=sig, frames=frames, samplerate=sr, brate=brate, mode=mode, quality=q, vbr=vbr, nbytes=len(mp3), tag_nbytes=len(tag),
                          mp3_sha256=hashlib.sha256(mp3).hexdigest(), tag_sha256=hashlib.sha256(tag).hexdigest())
        print(name, len(mp3), len(tag))
    with open(os.path.join(HERE, "manifest_tag.json"), "w") as f:
        json.dump(tags, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
