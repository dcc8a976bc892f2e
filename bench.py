#!/usr/bin/env python
"""bench.py - MP3 frames/s of the batched encode hot path on N B200s (BASELINE.json metric).

A *step* is one pass of the hot path over one batch: BASELINE.json configs[1] = 4096 frames of 44.1 kHz stereo
white noise (uniform int16 in [-12000, 12000], L/R independent), CBR 128 kbps joint stereo, quality 3, run as
512 streams x 8 frames per GPU.  With N GPUs every rank encodes its own 512 streams (weak scaling, streams are
independent: no data-path collective).

  value     frames/s with the PCM already resident in HBM: K device steps (analysis, scan, mdct, quantise, pack) back to back on
            PERSISTENT streams (no reset between steps), consecutive steps overlapping on the engine's two CUDA streams as in
            production, timed with CUDA events from the first kernel's start to the last one's end, max over ranks.
  e2e       frames/s through the public C ABI (lamegpu_batch_encode_packed, pipelined) with HOST buffers: host->device copy of
            the PCM, kernels, device->host copy of the packed frame bytes, header splice to MP3 bytes on the host.
  roofline  dominant kernel (quantise): issued warp instructions per second against 592 schedulers x SM clock (the limit that
            binds: SURVEY 8d); roofline_hbm: algorithmic bytes (16 060 B/frame) / its CUDA-event time against the measured HBM peak.
  cpu_baseline  the unmodified reference libmp3lame (oracle/_ref) single-thread on the host, bounded sample.

`--impl reference` times the reference's own CPU implementation on all host threads for the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAMS, FRAMES = 512, 8                 # per GPU and per step: 4096 frames
ALG_BYTES_QUANT = 16060                  # SURVEY.md section 8d, quantisation kernel, per frame
ALG_BYTES_ANALYSIS = 15840               # SURVEY.md section 8d, MDCT+psy kernels, per frame
WORKLOAD = ("configs[1]: 4096 frames/GPU = 512 streams x 8 frames, 44.1 kHz stereo white noise U[-12000,12000], "
            "CBR 128 kbps joint stereo q3")


def shard_range(nstreams, rank, world):
    """contiguous block of streams owned by `rank` (balanced to within one stream)"""
    base, rem = divmod(nstreams, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_over_ranks(frames, ms, device="cuda"):
    """SUM of frames, MAX of milliseconds over the process group (no-op without one)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return frames, ms
    t = torch.tensor([frames], dtype=torch.float64, device=device)
    m = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(t.item()), float(m.item())


def noise_pcm(nstreams, nsamples, seed):
    rng = np.random.default_rng(seed)
    if SIGNAL == "sine":      # two tones per channel + U[-1000,1000], every stream at its own start time
        out = np.empty((nstreams, 2, nsamples), dtype=np.int16)
        for s0 in range(0, nstreams, 128):
            n = min(128, nstreams - s0)
            t = (np.arange(nsamples)[None, :] + rng.integers(0, 44100, size=(n, 1))) / 44100.0
            out[s0:s0 + n, 0] = np.rint(8000 * np.sin(2 * np.pi * 440 * t) + 4000 * np.sin(2 * np.pi * 3300 * t) + rng.integers(-1000, 1001, size=t.shape))
            out[s0:s0 + n, 1] = np.rint(8000 * np.sin(2 * np.pi * 554.37 * t) + 3000 * np.sin(2 * np.pi * 7000 * t) + rng.integers(-1000, 1001, size=t.shape))
        return out
    return rng.integers(-12000, 12001, size=(nstreams, 2, nsamples), dtype=np.int16)


SIGNAL, BRATE, VBR, QUALITY = "noise", 128, 0, -1       # configs[1]; --signal/--brate/--vbr select the other BASELINE configs


def set_workload(args):
    """configs[1] by default; `--signal sine --brate 320` = configs[2], `--signal sine --vbr 4 --brate 2` = configs[3] (VBR -V2)"""
    global SIGNAL, BRATE, VBR, STREAMS, FRAMES, WORKLOAD, QUALITY
    SIGNAL, BRATE, VBR, STREAMS, FRAMES, QUALITY = args.signal, args.brate, args.vbr, args.streams, args.frames, args.quality
    if (SIGNAL, BRATE, VBR, STREAMS, FRAMES, QUALITY) != ("noise", 128, 0, 512, 8, -1):
        rate = {0: "CBR %d kbps" % BRATE, 2: "VBR-old -V%d" % BRATE, 3: "ABR %d kbps" % BRATE, 4: "VBR-new -V%d" % BRATE}[VBR]
        sig = "white noise U[-12000,12000]" if SIGNAL == "noise" else "two tones per channel + U[-1000,1000]"
        WORKLOAD = "%d frames/GPU = %d streams x %d frames, 44.1 kHz stereo %s, %s joint stereo q%d" % (STREAMS * FRAMES, STREAMS, FRAMES, sig, rate, 3 if QUALITY < 0 else QUALITY)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.t0 = self.t1 = None

    def start(self):
        """start sampling (call BEFORE the warm-up: nvidia-smi needs a few hundred ms to produce its first row)"""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if len(r) > 8 and self.t0 is not None and self.t0 <= t <= self.t1]
        window = "timed region"
        if not rows:          # region shorter than the sampling period: the samples taken under load right around it
            rows = [r for t, r in self.rows if len(r) > 8 and self.t0 is not None and self.t0 - 0.25 <= t <= self.t1 + 0.05]
            window = "timed region +-0.25 s (warm-up load)"
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_profile(kernel):
    """per-launch figures of `kernel` from the committed `ncu --set full` capture of the headline workload (profiles/): DRAM bytes,
    executed warp instructions, issue-active %; newest round first"""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for name in ("r2_ncu_summary.json", "r1_ncu_summary.json"):
        try:
            k = json.load(open(os.path.join(ROOT, "profiles", name)))[kernel]
            return {"file": "profiles/" + name,
                    "traffic": sum(float(k[m]["value"]) * scale[k[m]["unit"]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum")),
                    "inst": float(k["smsp__inst_executed.sum"]["value"]),
                    "issue_active": float(k["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"])}
        except Exception:
            continue
    return None


def bench_config(S, F):
    """the `config` object, identical in both arms"""
    return {"workload": WORKLOAD, "streams_per_gpu": S, "frames_per_stream_per_step": F}


def cpu_baseline(seconds=12.0):
    """unmodified reference (or the port when the reference build is absent), one thread, bounded sample"""
    import oracle
    oracle.build()
    kind = "reference" if oracle.have_ref() else "port"
    Enc = oracle.RefEncoder if kind == "reference" else oracle.PortEncoder
    pcm = noise_pcm(1, 1152 * 256, 777)[0]
    enc = Enc(44100, 2, BRATE, 4, QUALITY, vbr=VBR)
    frames, t0 = 0, time.perf_counter()
    while True:
        enc.encode(pcm[0], pcm[1])
        frames += 256
        dt = time.perf_counter() - t0
        if dt > seconds:
            break
    enc.close()
    return {"value": frames / dt, "unit": "frames/s", "cores": 1, "kind": kind,
            "sample": "%d frames of the same workload, one stream, one host thread, %.1f s" % (frames, dt)}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on all host threads, same workload/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    oracle.build()
    kind = "reference" if oracle.have_ref() else "port"
    Enc = oracle.RefEncoder if kind == "reference" else oracle.PortEncoder
    cores = os.cpu_count() or 1
    nsamp = FRAMES * 1152
    pcm = noise_pcm(STREAMS, nsamp, 4242)
    encs = [Enc(44100, 2, BRATE, 4, QUALITY, vbr=VBR) for _ in range(STREAMS)]
    pool = ThreadPoolExecutor(max_workers=cores)

    def step():
        # ctypes releases the GIL inside lame_encode_buffer, so threads scale over cores
        list(pool.map(lambda s: encs[s].encode(pcm[s, 0], pcm[s, 1]), range(STREAMS)))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    frames = STREAMS * FRAMES * args.steps
    v = frames / dt
    line = {"impl": "reference", "metric": "mp3_frames_per_sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(STREAMS, FRAMES),
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind,
                             "sample": "the full step (%d persistent streams x %d frames) on %d host threads" % (STREAMS, FRAMES, cores)},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def log(msg):
    if os.environ.get("BENCH_VERBOSE"):
        print("[bench %.1fs] %s" % (time.perf_counter() - T0, msg), file=sys.stderr, flush=True)


T0 = time.perf_counter()


def main():
    import faulthandler
    # a hung run must not eat the GPU lease: dump all stacks and exit after the watchdog period
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG_S", "900")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--streams", type=int, default=STREAMS)
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--brate", type=int, default=128, help="kbps (CBR/ABR) or the -V level with --vbr 4")
    ap.add_argument("--vbr", type=int, default=0, choices=(0, 2, 3, 4), help="0 CBR, 2 VBR-old (vbr_rh), 3 ABR, 4 VBR-new (vbr_mtrh)")
    ap.add_argument("--signal", default="noise", choices=("noise", "sine"))
    ap.add_argument("--quality", type=int, default=-1, help="lame_set_quality 0..9, -1 = default (3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    set_workload(args)
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import lame_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, F = args.streams, args.frames
    nsamp = F * 1152
    # the global job is world*S independent streams; this rank owns a contiguous shard of them
    lo, hi = shard_range(world * S, rank, world)
    assert hi - lo == S
    # `python bench.py --gpus N` without torchrun: ONE process drives N GPUs through the library itself (device = -1: the batch spans the
    # devices, one engine and one host thread per device); under torchrun every rank has its own GPU
    ngpu_here = args.gpus if (world == 1 and args.gpus > 1) else 1
    dev = local
    if ngpu_here > 1:
        os.environ["LAMEGPU_DEVICES"] = str(ngpu_here)
        dev, S = -1, S * ngpu_here
    enc = lame_b200.BatchEncoder(S, 44100, 2, BRATE, -1, QUALITY, frames_per_launch=F, device=dev, vbr=VBR)
    pcm = noise_pcm(S, nsamp + 224, 1000 + rank)            # +224: the first launch needs 1152*F + 224 user samples
    ring = noise_pcm(S, 2 * nsamp, 3000 + rank)             # a periodic signal of 2F frames per stream for the device-only steps
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log('engine created')
    # ---------------- value: device pipeline, inputs resident in HBM, persistent streams, steps back to back
    enc.stage(ring, F)                                       # H2D once + one (untimed) full step per buffer set: the two sets hold the ring's halves
    log('staged')
    sampler = ClockSampler(local)
    sampler.start()
    enc.run_device_steps(F, args.warmup)
    enc.run_device_steps(F, 40)                              # untimed load so that the clock samples around a short timed region are under load
    log('warm')
    launches0 = enc.kernel_launches()
    barrier()
    sampler.mark_begin()
    # K steps on the same streams (reservoir, psycho-acoustic state and step-size memory carried from step to step); CUDA events on the
    # engine's streams from the first kernel's start to the last one's end.  Step i+1's analysis kernels run under step i's quantiser, as in
    # production; every step touches ~215 MB on alternating buffer sets, more than the 126 MB L2, so no step finds its inputs cached.
    dev_ms = enc.run_device_steps(F, args.steps) * args.steps
    kms = np.array(enc.kernel_ms()) * args.steps             # per-kernel CUDA-event times, mean over exactly these steps
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    launches = enc.kernel_launches() - launches0
    log('device timing done: %s' % (kms / args.steps))
    alone = np.zeros(5)                                      # outside the timed region: every kernel with the device to itself (two steps, one at a time)
    for _ in range(3):
        enc.rerun_device(F)
        alone += np.array(enc.kernel_ms()) / 3
    frames_dev, dev_ms_max = reduce_over_ranks(float(S * F * args.steps), dev_ms)
    value = frames_dev / (dev_ms_max * 1e-3)

    # ---------------- e2e: public API, host buffers in, MP3 bytes out
    enc.close()
    enc = lame_b200.BatchEncoder(S, 44100, 2, BRATE, -1, QUALITY, frames_per_launch=F, device=dev, vbr=VBR)
    # The batch is pipelined: a call stages its PCM from the caller's buffers into pinned memory and submits the step; while the device
    # encodes it, the previous step's bytes are spliced and handed over.  That host work is memcpy-class and hidden under the device
    # step, so a few host threads per rank do (round 1 needed 16 and lost 15 % at eight ranks on 32 cores).
    cores = os.cpu_count() or 1
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))) * ngpu_here
    host_threads = int(os.environ.get("LAMEGPU_THREADS", max(2, min(4, cores // local_world))))
    enc.set_threads(host_threads)
    enc.set_pipelined(True)
    step_pcm = [noise_pcm(S, nsamp, 5000 + 17 * i + rank) for i in range(4)]
    out = np.empty((S, int(1.25 * nsamp) + 7200 + 4096 + 1440 * F), dtype=np.uint8)
    nbytes = np.zeros(S, dtype=np.int32)
    empty = np.empty((S, 2, 0), dtype=np.int16)
    enc.encode_raw(pcm, out, nbytes)                          # primes the 528+... encoder delay: afterwards every call yields F frames
    for i in range(args.warmup):
        enc.encode_raw(step_pcm[i % 4], out, nbytes)
    enc.set_pipelined(False); enc.encode_raw(empty, out, nbytes); enc.set_pipelined(True)     # nothing in flight when the clock starts
    barrier()
    t0 = time.perf_counter()
    e2e_frames = 0
    total_bytes = 0
    for i in range(args.steps):
        e2e_frames += enc.encode_raw(step_pcm[i % 4], out, nbytes)    # H2D of this step's PCM; the bytes of step i-1 land in `out`
        total_bytes += int(nbytes.sum())
    enc.set_pipelined(False)                                  # the last step: wait for its D2H, splice
    enc.encode_raw(empty, out, nbytes)                        # ... and take its bytes: all K steps are in host memory inside the timed region
    total_bytes += int(nbytes.sum())
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    log('e2e done %.1f ms' % e2e_ms)
    e2e_frames_all, e2e_ms_max = reduce_over_ranks(float(e2e_frames), e2e_ms)
    lib = lame_b200.load_library()
    h2d = S * 2 * (F * 1152 + 1328) * 2 + S * 4             # all devices of this process
    d2h = int(lib.lamegpu_batch_d2h_bytes(enc._h))
    enc.close()

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        q_ms = kms[3] / args.steps
        qname = {4: "lg_kernel_vbr", 2: "lg_kernel_vbrold"}.get(VBR, "lg_kernel_quantg" if (QUALITY < 0 or 3 <= QUALITY <= 6) and S <= 592 else "lg_kernel_quant")
        S //= ngpu_here                                       # per GPU from here on: the kernel figures are one device's
        headline = (S, F, SIGNAL, BRATE, VBR, QUALITY) == (512, 8, "noise", 128, 0, -1)
        achieved = ALG_BYTES_QUANT * S * F / (q_ms * 1e-3) / 1e9
        a_ms = float(alone[0] + alone[1] + alone[2])               # the three kernels' own run times (in the pipeline kernel A's event time is mostly waiting)
        prof = ncu_profile(qname) if headline else None
        sm_hz = 1e6 * (clocks.get("sm_mhz") or 1965.0)
        issue_peak = 148 * 4 * sm_hz                          # one warp instruction per scheduler per cycle
        issue = None
        if prof and prof.get("inst"):
            issue = {"bound": "issue", "kernel": qname, "achieved": prof["inst"] / (q_ms * 1e-3), "peak": issue_peak, "unit": "warp-inst/s",
                     "frac": prof["inst"] / (q_ms * 1e-3) / issue_peak, "traffic": prof.get("traffic"),
                     "warp_instructions_per_launch": prof["inst"], "avg_launch_ms": q_ms, "issue_active_pct_ncu": prof.get("issue_active"),
                     "peak_source": "148 SMs x 4 schedulers x SM clock under load (%.0f MHz)" % (sm_hz / 1e6),
                     "source": "smsp__inst_executed.sum of one launch from the committed ncu --set full capture (%s); duration live, CUDA events over the timed region" % prof["file"],
                     "note": "the path is latency/issue-bound integer + table-lookup work (SURVEY 8d): the binding roofline is the issue rate; the HBM fraction is in roofline_hbm"}
        hbm = {"bound": "hbm", "kernel": qname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
               "traffic": prof.get("traffic") if prof else None, "peak_source": peak_src,
               "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/)",
               "algorithmic_bytes_per_launch": ALG_BYTES_QUANT * S * F, "avg_launch_ms": q_ms}
        line = {
            "metric": "mp3_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world * ngpu_here, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(S, F),
            "notes": {"process_model": "one process per GPU (torchrun)" if ngpu_here == 1 else "one process, %d GPUs through lamegpu_batch_open(device = -1)" % ngpu_here,
                      "l2": "no explicit flush: steps run back to back on persistent streams; a step touches ~215 MB (two alternating buffer sets) > 126 MB L2",
                      "parallelism": "streams sharded over %d GPU(s), no collective on the data path" % (world * ngpu_here),
                      "value": "K pipelined steps on persistent streams, first kernel start to last kernel end (CUDA events) / K",
                      "kernel_times": "kernel A of step i+1 is gated behind the launch of kernel D of step i and runs in the room D's last wave leaves, so its "
                                      "event-to-event time (kernels_ms_per_step.analysis) is mostly waiting; alone it takes 0.73 ms (profiles/)"},
            "e2e": {"value": e2e_frames_all / (e2e_ms_max * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_max / args.steps, "mp3_bytes": total_bytes,
                    "note": "lamegpu_batch_encode_packed, pipelined (lamegpu_batch_set_pipelined): caller's PCM -> pinned staging -> H2D -> kernels -> D2H of packed "
                            "bytes -> host header splice -> caller's buffer; all K steps' bytes received inside the timed region (host threads per rank: %d)" % host_threads},
            "gpu_launches": int(launches),
            "roofline": issue if issue else hbm,
            "roofline_hbm": hbm,
            "kernels_ms_per_step": {"analysis": kms[0] / args.steps, "scan": kms[1] / args.steps, "mdct": kms[2] / args.steps, "quant": q_ms,
                                    "pack": kms[4] / args.steps,
                                    "alone": {"analysis": float(alone[0]), "scan": float(alone[1]), "mdct": float(alone[2]), "quant": float(alone[3]),
                                              "pack": float(alone[4])},
                                    "note": "CUDA events around each kernel; analysis/scan/mdct of step i+1 run under the quantiser of step i, so ms_per_step is less than their sum"},
            "roofline_mdct_psy": {"bound": "hbm", "kernels": "analysis+scan+mdct", "achieved": ALG_BYTES_ANALYSIS * S * F / (a_ms * 1e-3) / 1e9,
                                  "peak": peak, "unit": "GB/s", "frac": ALG_BYTES_ANALYSIS * S * F / (a_ms * 1e-3) / 1e9 / peak},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 1, "kind": "reference", "sample": "measured at N=1 only"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
