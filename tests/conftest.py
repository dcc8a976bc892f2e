import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# handles of the libmp3lame face are lanes of a shared engine (512 by default): the tests make a handle or two at a time
os.environ.setdefault("LAMEGPU_LANES", "8")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# ---- deterministic test signals, 44.1 kHz stereo int16 -------------------------
def sig_noise(nsamples, seed=1234, amp=12000):
    rng = np.random.default_rng(seed)
    return rng.integers(-amp, amp + 1, size=(2, nsamples)).astype(np.int16)


def sig_sine(nsamples, seed=1234, sr=44100):
    rng = np.random.default_rng(seed)
    t = np.arange(nsamples) / sr
    l = 8000 * np.sin(2 * np.pi * 440 * t) + 4000 * np.sin(2 * np.pi * 3300 * t) + rng.integers(-1000, 1001, nsamples)
    r = 8000 * np.sin(2 * np.pi * 554.37 * t) + 3000 * np.sin(2 * np.pi * 7000 * t) + rng.integers(-1000, 1001, nsamples)
    return np.stack([np.rint(l), np.rint(r)]).astype(np.int16)


def sig_click(nsamples, seed=1234):
    """quiet floor with loud decaying bursts: forces START/SHORT/STOP blocks (hard part 5)"""
    rng = np.random.default_rng(seed)
    i = np.arange(nsamples)
    ph, ph2 = i % 7919, (i + 3000) % 10007
    env = np.where(ph < 400, np.exp(-ph / 60.0), 0.0)
    env2 = np.where(ph2 < 300, np.exp(-ph2 / 40.0), 0.0)
    floor = rng.integers(-60, 61, size=(2, nsamples))
    b1 = rng.integers(-24000, 24001, nsamples)
    b2 = rng.integers(-16000, 16001, nsamples)
    l = floor[0] + env * b1
    r = floor[1] + 0.8 * env * b1 + env2 * b2
    return np.clip(np.rint(np.stack([l, r])), -32768, 32767).astype(np.int16)


def sig_gap(nsamples, seed=1234):
    """digital silence, then noise, then silence again: exercises the all-zero granule path and onsets"""
    x = sig_noise(nsamples, seed)
    x[:, :min(6000, nsamples)] = 0
    x[:, 20000:23000] = 0
    return x


SIGNALS = {"noise": sig_noise, "sine": sig_sine, "click": sig_click, "gap": sig_gap}


def make_signal(kind, nsamples, seed=1234):
    if kind == "testcase":
        pcm = np.load(os.path.join(ROOT, "tests", "golden", "testcase_pcm.npy"))
        return np.ascontiguousarray(pcm[:, :nsamples])
    return SIGNALS[kind](nsamples, seed)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
