"""The drop-in test of SURVEY.md section 8b: the reference's own command-line program (`/root/reference/frontend/*.c`, unmodified,
compiled where it lies by oracle/Makefile `frontend`) linked against the product library instead of libmp3lame, run over WAV files with
the reference's regression option lists (`/root/reference/test/*.op`, HACKING:3-15 `lametest.py`), and compared byte for byte with the
same program linked against the reference library: the MP3 FILE, Info/Xing tag frame included.

CPU suite: the link itself (every libmp3lame.sym symbol the frontend needs is exported), the option lists, and a few lines through the
emulator build of the kernels.  GPU suite: the option lists on the device."""
import os
import subprocess
import wave

import numpy as np
import pytest

from conftest import ROOT, make_signal

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
LAME_REF, LAME_GPU = os.path.join(REF_DIR, "lame_ref"), os.path.join(REF_DIR, "lame_gpu")
REFERENCE_TESTS = "/root/reference/test"


def op_lines(name):
    """The reference's option lists, regenerated (the files themselves stay in /root/reference; test_option_lists_match_reference checks
    them line by line where the reference tree is present)."""
    def qsweep(prefix, sep=" "):
        return ["%s%s-q%s%d" % (p, " " if p else "", sep, q) for p in prefix for q in range(10)]

    def modes(x):
        return [x, x + " -m m", x + " -m s"]

    if name == "CBRABR.op":
        mid = ["--abr 9", "-b 16", "--abr 20", "-b 24", "--abr 28", "-b 32", "--abr 36", "-b 40", "--abr 48", "-b 56", "--abr 60", "-b 64", "--abr 72",
               "-b 80", "--abr 88", "-b 96"]
        mid2 = ["-b 112", "--abr 120", "-b 128", "--abr 140", "-b 160", "--abr 180", "-b 192", "--abr 210", "-b 224", "--abr 235", "-b 256", "--abr 319"]
        return qsweep(modes("-b 8")) + mid + qsweep(modes("--abr 108")) + mid2 + qsweep(modes("-b 320"))
    if name == "VBR.op":
        out = []
        for pre in ("", "--vbr-new "):
            out += qsweep(modes(pre + "-V9"), "") + [pre + "-V%d" % v for v in (8, 7, 6, 5)] + qsweep(modes(pre + "-V4"), "")
            out += [pre + "-V%d" % v for v in (3, 2, 1)] + qsweep(modes(pre + "-V0"), "")
        return out
    if name == "shortCBRABR.op":
        def four(x):
            return [x, x + " -f", x + " -h", x + " -h -m m"]
        return four("-b 8") + ["--abr 9", "-b 56", "--abr 60"] + four("--abr 128") + ["-b 128", "--abr 319"] + four("-b 320") + ["-b 320 -h -m s"]
    if name == "shortVBR.op":
        out = []
        for pre in ("", "--vbr-new "):
            for v in (9, 4, 0):
                x = pre + "-V%d" % v
                out += [x, x + " -f", x + " -h", x + " -h -m m"] + ([x + " -h -m s"] if v == 0 else [])
        return out
    if name == "misc.op":
        return ["--athlower 10", "--abr 160 -b 128 -B 192 -F", "-V3 -b 128 -B 192 -F", "--freeformat -b 33", "--freeformat -b 330", "-k", "--lowpass 12", "--noath",
                "--noasm mmx --noasm 3dnow --noasm sse", "--noshort", "--notemp", "--nores", "-p", "--resample 48000", "--scale 0.8", "-t"]
    if name == "nores.op":
        return ["-h --nores", "--nores"]
    raise KeyError(name)


ALL_LISTS = ("CBRABR.op", "VBR.op", "shortCBRABR.op", "shortVBR.op", "misc.op", "nores.op")


@pytest.mark.skipif(not os.path.isdir(REFERENCE_TESTS), reason="the reference tree is only present in the development container")
@pytest.mark.parametrize("name", ALL_LISTS)
def test_option_lists_match_reference(name):
    want = [" ".join(l.split()) for l in open(os.path.join(REFERENCE_TESTS, name)) if l.strip()]
    assert [" ".join(l.split()) for l in op_lines(name)] == want


@pytest.fixture(scope="module")
def wav_files(tmp_path_factory):
    """testcase.wav of the reference (all 25 000 samples; the PCM is committed as tests/golden/testcase_pcm.npy) and a transient signal"""
    d = tmp_path_factory.mktemp("wav")
    out = {}
    for name, pcm in (("testcase", np.load(os.path.join(ROOT, "tests", "golden", "testcase_pcm.npy"))), ("click", make_signal("click", 30 * 1152 + 333, seed=7))):
        p = str(d / (name + ".wav"))
        w = wave.open(p, "wb")
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100)
        w.writeframes(np.ascontiguousarray(pcm.T).astype("<i2").tobytes())
        w.close()
        out[name] = p
    return out


def run_pair(gpu_exe, opts, wav, tmp, env=None):
    """both programs on the same input; returns (reference's exit code, ours, identical?)"""
    a, b = os.path.join(tmp, "ref.mp3"), os.path.join(tmp, "gpu.mp3")
    for f in (a, b):
        if os.path.exists(f):
            os.remove(f)
    # ReplayGain analysis is the one thing the CLI switches on that lies outside the hot path
    base = ["--quiet", "--noreplaygain"] + opts.split()
    r1 = subprocess.run([LAME_REF] + base + [wav, a], capture_output=True, timeout=300)
    r2 = subprocess.run([gpu_exe] + base + [wav, b], capture_output=True, timeout=600, env=dict(os.environ, LAMEGPU_LANES="2", **(env or {})))
    same = os.path.exists(a) and os.path.exists(b) and open(a, "rb").read() == open(b, "rb").read()
    return r1.returncode, r2.returncode, same, r2.stderr.decode(errors="replace")[-3000:]


def test_frontend_links_against_the_product_library(oracle_mod):
    """every libmp3lame symbol the reference's frontend uses is exported by liblamegpu.so: the link of oracle/Makefile `frontend` succeeds"""
    if not os.path.exists(LAME_GPU):
        pytest.skip("oracle/_ref/lame_gpu is built from /root/reference/frontend (development container)")
    r = subprocess.run([LAME_GPU, "--version"], capture_output=True, text=True)
    assert r.returncode == 0 and "3.99.5" in r.stdout
    undefined = subprocess.run(["nm", "-D", "--undefined-only", LAME_GPU], capture_output=True, text=True).stdout
    needed = {l.split()[-1] for l in undefined.splitlines() if " lame_" in l or " id3tag_" in l or " hip_" in l or " get_lame" in l}
    exported = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "deprecated-lame-mirror_b200", "liblamegpu.so")],
                                                      capture_output=True, text=True).stdout.splitlines()}
    assert needed and needed <= exported, sorted(needed - exported)


def test_every_libmp3lame_symbol_is_exported():
    """include/libmp3lame.sym of the reference, 232 symbols (the list is committed: tests/golden/libmp3lame.sym)"""
    want = {l.strip() for l in open(os.path.join(ROOT, "tests", "golden", "libmp3lame.sym")) if l.strip()}
    exported = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "deprecated-lame-mirror_b200", "liblamegpu.so")],
                                                      capture_output=True, text=True).stdout.splitlines()}
    assert len(want) == 232 and want <= exported, sorted(want - exported)


@pytest.fixture(scope="module")
def lame_emu(oracle_mod, tmp_path_factory):
    """the frontend linked against the emulator build of the kernels (tests/emu): the CPU suite's stand-in for lame_gpu"""
    fe = os.path.join(REF_DIR, "fe")
    if not os.path.isdir(fe):
        pytest.skip("frontend objects are built from /root/reference (development container)")
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu")], check=True, capture_output=True)
    exe = str(tmp_path_factory.mktemp("bin") / "lame_emu")
    objs = [os.path.join(fe, f) for f in sorted(os.listdir(fe)) if f.endswith(".o")]
    subprocess.run(["gcc", "-o", exe] + objs + ["-L" + os.path.join(ROOT, "tests", "emu"), "-llamegpu_emu", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "tests", "emu")], check=True)
    return exe


@pytest.mark.parametrize("opts", ["-V3 -b 128 -B 192 -F", "--abr 160 -b 128 -B 192 -F --lowpass 12", "--preset 192 --scale 0.8", "-b 128 --noshort --nores -p --athlower 10",
                                  "--resample 32 -b 96"])
def test_frontend_over_emulated_kernels(lame_emu, wav_files, tmp_path, opts):
    """(the two lines that resample - `--lowpass 12` picks 32 kHz - also pin the flush with frames of the lane still waiting for their
    launch: lamegpu_batch::pad_for_flush)"""
    rc_ref, rc_ours, same, err = run_pair(lame_emu, opts, wav_files["testcase"], str(tmp_path))
    assert rc_ref == 0 and rc_ours == 0, err
    assert same, "MP3 file differs from the reference's for `%s`" % opts


def test_unsupported_options_fail_loudly(lame_emu, wav_files, tmp_path):
    """free format is outside the supported corner: lame_init_params reports an error, the program exits non-zero - never a silently different file"""
    rc_ref, rc_ours, same, err = run_pair(lame_emu, "--freeformat -b 330", wav_files["testcase"], str(tmp_path))
    assert rc_ref == 0 and rc_ours != 0 and not same
    assert "not supported" in err


def _sweep(names, wav, tmp, step=1):
    bad, n = [], 0
    for name in names:
        for i, opts in enumerate(op_lines(name)):
            if i % step:
                continue
            n += 1
            rc_ref, rc_ours, same, err = run_pair(LAME_GPU, opts, wav, tmp)
            if "--freeformat" in opts:
                if rc_ours == 0:
                    bad.append((name, opts, "free format must be refused"))
                continue
            if rc_ref != 0 or rc_ours != 0 or not same:
                bad.append((name, opts, "exit codes %d / %d%s" % (rc_ref, rc_ours, "" if same else ", files differ"), err))
    return n, bad


@pytest.mark.gpu
def test_frontend_dropin_option_lists(wav_files, tmp_path):
    """`lame <options> testcase.wav` - the reference's procedure (Makefile.am:39-46, test/lametest.py) - for every line of shortCBRABR.op, misc.op and
    nores.op, every third line of shortVBR.op and every 16th of the two big matrices; LAMEGPU_FULL_SWEEP=1 runs every line of all six lists
    (378) plus the short lists on a transient signal (the result of the last full sweep is in profiles/)"""
    if not (os.path.exists(LAME_GPU) and os.path.exists(LAME_REF)):
        pytest.skip("oracle/_ref/lame_gpu and lame_ref travel with the repository snapshot; not built here")
    full = os.environ.get("LAMEGPU_FULL_SWEEP") == "1"
    n1, bad1 = _sweep(("shortCBRABR.op", "misc.op", "nores.op"), wav_files["testcase"], str(tmp_path))
    n2, bad2 = _sweep(("shortVBR.op",), wav_files["testcase"], str(tmp_path), 1 if full else 3)
    n3, bad3 = _sweep(("CBRABR.op", "VBR.op"), wav_files["testcase"], str(tmp_path), 1 if full else 16)
    n4, bad4 = _sweep(("shortCBRABR.op", "shortVBR.op", "misc.op"), wav_files["click"], str(tmp_path), 1 if full else 9)
    bad = bad1 + bad2 + bad3 + bad4
    print("%d option lines, %d mismatches" % (n1 + n2 + n3 + n4, len(bad)))
    assert not bad, bad[:5]
