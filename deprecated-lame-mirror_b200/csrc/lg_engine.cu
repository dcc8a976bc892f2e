// lg_engine.cu - device side of the batch engine: owns the HBM buffers of one configuration
// (S streams x up to F frames per launch), the four kernels and the copies.
//
//   H2D  pcm (int16 or float, pinned)  [->  R resample, when the input rate differs]  ->  A analysis  ->  B scan  ->  C mdct  ->  D quantise  ->  E pack  ->  D2H bytes
//
// A launch ("step") runs on one of two slots and two CUDA streams, so that consecutive steps overlap (see LgSlot); the host
// (lg_bitstream.cpp) only interleaves the packed payload bytes with the frame headers.  There is no CPU fallback: without a CUDA device lg_engine_create() fails and says so.
//
// This translation unit is also compiled by g++ with -DLG_EMULATE for tests/emu (see lg_compat.h); that
// build is test infrastructure and is never loaded by the product package.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <new>
#include "lg_compat.h"
#include "lg_types.h"
#include "lg_k_resample.cuh"
#include "lg_k_analysis.cuh"
#include "lg_k_scan.cuh"
#include "lg_k_mdct.cuh"
#include "lg_k_quant.cuh"
#include "lg_k_quantg.cuh"
#include "lg_k_vbr.cuh"
#include "lg_k_vbrold.cuh"
#include "lg_k_pack.cuh"
#include "lg_engine.h"

#ifdef LG_EMULATE
typedef int lgStream_t;
typedef struct { double t; } lgEvent_t;
#define LG_CHECK(x) (x)
static int lg_dev_malloc(void **p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : -1; }
static void lg_dev_free(void *p) { free(p); }
static int lg_host_malloc(void **p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : -1; }
static void lg_host_free(void *p) { free(p); }
#define LG_COPY_H2D(dst, src, n, st) memcpy(dst, src, n)
#define LG_COPY_D2H(dst, src, n, st) memcpy(dst, src, n)
#define LG_MEMSET(dst, v, n, st) memset(dst, v, n)
#else
typedef cudaStream_t lgStream_t;
#define LG_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "lamegpu: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while (0)
static int lg_dev_malloc(void **p, size_t n) { return cudaMalloc(p, n ? n : 1) == cudaSuccess ? 0 : -1; }
static void lg_dev_free(void *p) { if (p) cudaFree(p); }
static int lg_host_malloc(void **p, size_t n) { return cudaMallocHost(p, n ? n : 1) == cudaSuccess ? 0 : -1; }
static void lg_host_free(void *p) { if (p) cudaFreeHost(p); }
#define LG_COPY_H2D(dst, src, n, st) cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, st)
#define LG_COPY_D2H(dst, src, n, st) cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st)
#define LG_MEMSET(dst, v, n, st) cudaMemsetAsync(dst, v, n, st)
#endif

#define LG_SLOTS 2

/* RAII: every entry point runs with the engine's device current on the calling thread and puts the caller's back (CUDA's current
 * device is per host thread: an engine may be driven from a thread other than the one that made it, and a process may hold engines
 * on several GPUs) */
struct LgDeviceScope {
#ifndef LG_EMULATE
    int prev = -1; bool ok = true;
    explicit LgDeviceScope(int device) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != device) ok = cudaSetDevice(device) == cudaSuccess; }
    ~LgDeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
#else
    bool ok = true;
    explicit LgDeviceScope(int) {}
#endif
};

/* What one step (launch) owns while it is in flight.  Everything that crosses from the analysis stream to the quantiser stream, and
 * everything the host touches, exists once per slot, so that step i+1 can be staged, copied and analysed while step i is quantised,
 * packed and copied back.  Buffers that live on the analysis stream only (PCM, subband samples, analysis records) exist once. */
struct LgSlot {
    LgPsyOut *d_psy; LgFrameCtl *d_frm; float *d_xr;
    LgGranuleOut *d_gout; LgFrameOut *d_fout;
    unsigned char *d_pay, *d_hdr;
    int *d_nfr;
    int16_t *h_pcm16; int *h_nfr;
    char *h_pcmn; LgPcmKind *h_kind;  /* native sample types: rows of pcmn_esz-byte elements, every stream in its own type */
    LgFrameOut *h_fout; unsigned char *h_pay, *h_hdr;
    char *h_raw; LgRsChunk *h_rsc; LgRsStream *h_rss;      /* h_raw: the resampler's input, rows of raw_esz-byte elements in the stream's own sample type */
    int chunk_cap;
#ifndef LG_EMULATE
    cudaEvent_t evp[4][4];            /* per part of the analysis (lg_submit): 0 its kernel A done, 1 its kernel B starts, 2 B done, 3 C done */
    cudaEvent_t ev[10];               /* 0 start (analysis stream), 1 after A, 2 after B, 3 after C, 4 before D (quantiser stream), 5 after D, 6 after E, 7 results on the host, 8/9 around R */
#endif
    int in_flight, nframes, parts;
    float ms[8];                      /* A, B, C, D, E, R, -, start of A to end of E */
};

struct lg_engine {
    LgDevCfg hcfg;
    LgDevCfg *dcfg;
    int S, F, device;
    size_t pcm_stride;                /* samples per channel per stream */
    int16_t *d_pcm16; float *d_pcmf;  /* d_pcmf: the resampler's output */
    char *d_pcmn; LgPcmKind *d_kind; int pcmn_esz;
    float *d_sb;
    LgAnalysis *d_ana;
    size_t pay_stride;
    LgStreamState *d_state, *d_state0;   /* d_state0: S copies of the initial state, for one-copy resets */
    size_t raw_stride; int chunk_cap;
    char *d_raw; int raw_esz; LgRsChunk *d_rsc; LgRsStream *d_rss;
    LgSlot slot[LG_SLOTS];
    /* three CUDA streams: `stream` carries the copies in and the stateless/scan kernels (R, A, B, C), `stream2` the quantiser (D) and
     * `stream3` the packer and the copies out (E, D2H), so that the next step's kernel D follows this step's directly; events order slot
     * k's D behind its C, its E behind its D, and its next A-B-C behind its previous D2H.  Nothing waits inside a kernel. */
    lgStream_t stream, stream2, stream2b, stream3, stream4;   /* stream2b: kernel D of the odd slot (see lg_submit: independent steps) */
    int ana_split;                    /* parts the analysis of a step is cut into along the frames (lg_submit) */
    int dense;                        /* more than four streams per SM: the one-warp kernel D in its seven-CTAs-per-SM build */
    int group_nw;                     /* > 0: kernel D in its group form (lg_k_quantg.cuh) with this many warps per granule.channel */
    int gate;                         /* kernel A of a step waits for the launch of kernel D of the step before (lg_submit) */
#ifndef LG_EMULATE
    cudaEvent_t ev_mark[2];
#endif
    float last_ms[8];
    long launches;
};

extern "C" const LgDevCfg *lg_engine_config(const lg_engine *e) { return &e->hcfg; }
extern "C" int lg_engine_streams(const lg_engine *e) { return e->S; }
extern "C" int lg_engine_max_frames(const lg_engine *e) { return e->F; }
extern "C" int lg_engine_device(const lg_engine *e) { return e->device; }
extern "C" int lg_engine_slots(const lg_engine *) { return LG_SLOTS; }
extern "C" size_t lg_engine_pcm_stride(const lg_engine *e) { return e->pcm_stride; }
extern "C" int16_t *lg_engine_host_pcm16(lg_engine *e, int k) { return e->slot[k].h_pcm16; }
extern "C" void *lg_engine_host_pcmn(lg_engine *e, int k) { return e->slot[k].h_pcmn; }
extern "C" LgPcmKind *lg_engine_host_kinds(lg_engine *e, int k) { return e->slot[k].h_kind; }
extern "C" int lg_engine_native_esz(const lg_engine *e) { return e->pcmn_esz; }
extern "C" int *lg_engine_host_nfr(lg_engine *e, int k) { return e->slot[k].h_nfr; }
extern "C" const unsigned char *lg_engine_host_pay(const lg_engine *e, int k) { return e->slot[k].h_pay; }
extern "C" const unsigned char *lg_engine_host_hdr(const lg_engine *e, int k) { return e->slot[k].h_hdr; }
extern "C" size_t lg_engine_pay_stride(const lg_engine *e) { return e->pay_stride; }
extern "C" const LgFrameOut *lg_engine_host_fout(const lg_engine *e, int k) { return e->slot[k].h_fout; }
extern "C" const float *lg_engine_last_kernel_ms(const lg_engine *e) { return e->last_ms; }
extern "C" long lg_engine_launch_count(const lg_engine *e) { return e->launches; }
extern "C" size_t lg_engine_raw_stride(const lg_engine *e) { return e->raw_stride; }
extern "C" void *lg_engine_host_raw(lg_engine *e, int k) { return e->slot[k].h_raw; }
extern "C" int lg_engine_raw_esz(const lg_engine *e) { return e->raw_esz; }
extern "C" LgRsChunk *lg_engine_host_chunks(lg_engine *e, int k) { return e->slot[k].h_rsc; }
extern "C" int *lg_engine_host_rs_counts(lg_engine *e, int k) { return (int *) e->slot[k].h_rss; }
extern "C" int lg_engine_chunk_cap(const lg_engine *e, int k) { return e->slot[k].chunk_cap; }

/* initial per-stream state: lame.c:2274 lame_init_internal_flags, lame.c:962, psymodel.c:1897-1922 and :2075 */
static void lg_initial_state(const LgDevCfg *c, LgStreamState *s)
{
    memset(s, 0, sizeof *s);
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < LG_CBANDS; ++j) { s->nb_l1[i][j] = 1e20f; s->nb_l2[i][j] = 1e20f; }
        for (int sb = 0; sb < LG_SBMAX_L; sb++) { s->en[i].l[sb] = 1e20f; s->thm[i].l[sb] = 1e20f; }
        for (int j = 0; j < 3; ++j)
            for (int sb = 0; sb < LG_SBMAX_S; sb++) { s->en[i].s[sb][j] = 1e20f; s->thm[i].s[sb][j] = 1e20f; }
        for (int j = 0; j < 9; j++) s->last_en_subshort[i][j] = 10.f;
    }
    s->ath_adjust_factor = 0.01f;
    s->ath_adjust_limit = 1.0f;
    s->masking_lower = 1.f;
    for (int i = 0; i < 19; i++) s->pefirbuf[i] = (float) (700 * c->mode_gr * c->channels);
    s->slot_lag = c->frac_spf;
    s->old_value[0] = s->old_value[1] = 180;
    s->current_step[0] = s->current_step[1] = 4;
}

extern "C" void lg_engine_destroy(lg_engine *e)
{
    if (!e) return;
    LgDeviceScope dev(e->device);
#ifndef LG_EMULATE
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->stream2) cudaStreamSynchronize(e->stream2);
    if (e->stream2b) cudaStreamSynchronize(e->stream2b);
    if (e->stream3) cudaStreamSynchronize(e->stream3);
    if (e->stream4) cudaStreamSynchronize(e->stream4);
#endif
    lg_dev_free(e->dcfg); lg_dev_free(e->d_pcm16); lg_dev_free(e->d_pcmf); lg_dev_free(e->d_pcmn); lg_dev_free(e->d_kind); lg_dev_free(e->d_sb); lg_dev_free(e->d_ana);
    lg_dev_free(e->d_state); lg_dev_free(e->d_state0);
    lg_dev_free(e->d_raw); lg_dev_free(e->d_rsc); lg_dev_free(e->d_rss);
    for (int k = 0; k < LG_SLOTS; k++) {
        LgSlot &t = e->slot[k];
        lg_dev_free(t.d_psy); lg_dev_free(t.d_frm); lg_dev_free(t.d_xr); lg_dev_free(t.d_gout); lg_dev_free(t.d_fout); lg_dev_free(t.d_pay); lg_dev_free(t.d_hdr);
        lg_dev_free(t.d_nfr);
        lg_host_free(t.h_pcm16); lg_host_free(t.h_pcmn); lg_host_free(t.h_kind); lg_host_free(t.h_nfr); lg_host_free(t.h_pay); lg_host_free(t.h_hdr); lg_host_free(t.h_fout);
        lg_host_free(t.h_raw); lg_host_free(t.h_rsc); lg_host_free(t.h_rss);
#ifndef LG_EMULATE
        for (int i = 0; i < 10; i++) if (t.ev[i]) cudaEventDestroy(t.ev[i]);
        for (int i = 0; i < 16; i++) if (t.evp[i / 4][i % 4]) cudaEventDestroy(t.evp[i / 4][i % 4]);
#endif
    }
#ifndef LG_EMULATE
    for (int i = 0; i < 2; i++) if (e->ev_mark[i]) cudaEventDestroy(e->ev_mark[i]);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->stream2) cudaStreamDestroy(e->stream2);
    if (e->stream2b) cudaStreamDestroy(e->stream2b);
    if (e->stream3) cudaStreamDestroy(e->stream3);
    if (e->stream4) cudaStreamDestroy(e->stream4);
#endif
    free(e);
}

/* back to the state of a fresh stream (a lame_t that was closed and whose lane is handed out again; the bench's warm start).
 * Ordered behind everything in flight on both streams. */
extern "C" int lg_engine_reset_streams(lg_engine *e, int first, int count)
{
    if (first < 0 || count < 0 || first + count > e->S) return -1;
    LgDeviceScope dev(e->device);
#ifdef LG_EMULATE
    memcpy(e->d_state + first, e->d_state0 + first, (size_t) count * sizeof(LgStreamState));
#else
    LG_CHECK(cudaStreamSynchronize(e->stream2));
    LG_CHECK(cudaStreamSynchronize(e->stream2b));
    LG_CHECK(cudaStreamSynchronize(e->stream4));
    LG_CHECK(cudaMemcpyAsync(e->d_state + first, e->d_state0 + first, (size_t) count * sizeof(LgStreamState), cudaMemcpyDeviceToDevice, e->stream));
    LG_CHECK(cudaStreamSynchronize(e->stream));
#endif
    return 0;
}

/* lame_encode_flush / flush_bitstream (bitstream.c:863-890) ends a stream's bit reservoir: ResvSize = 0, main_data_begin = 0.  The
 * psycho-acoustic state and the step-size memory stay (the reference keeps encoding with them if more samples follow a flush). */
#ifndef LG_EMULATE
__global__ void lg_kernel_end_reservoir(LgStreamState *state, const int *which, int n)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    LgStreamState *s = state + which[i];
    s->resv_size = 0; s->main_data_begin = 0;
    s->ancillary_flag = which[n + i];           /* the stuffing bit's phase after the host's drain (bitstream.c:246-256) */
}
#endif
extern "C" int lg_engine_end_reservoir(lg_engine *e, const int *streams, const int *ancillary_flags, int n)
{
    if (n <= 0) return 0;
    LgDeviceScope dev(e->device);
#ifdef LG_EMULATE
    for (int i = 0; i < n; i++) { LgStreamState *s = e->d_state + streams[i]; s->resv_size = 0; s->main_data_begin = 0; s->ancillary_flag = ancillary_flags[i]; }
#else
    LG_CHECK(cudaStreamSynchronize(e->stream2));           /* the flushed frames have been quantised */
    LG_CHECK(cudaStreamSynchronize(e->stream2b));
    int *d = nullptr;
    LG_CHECK(cudaMalloc(&d, (size_t) 2 * n * sizeof(int)));
    cudaMemcpyAsync(d, streams, (size_t) n * sizeof(int), cudaMemcpyHostToDevice, e->stream2);
    cudaMemcpyAsync(d + n, ancillary_flags, (size_t) n * sizeof(int), cudaMemcpyHostToDevice, e->stream2);
    lg_kernel_end_reservoir<<<(n + 127) / 128, 128, 0, e->stream2>>>(e->d_state, d, n);
    cudaError_t const err = cudaStreamSynchronize(e->stream2);
    cudaFree(d);
    if (err != cudaSuccess) { fprintf(stderr, "lamegpu: CUDA error %s at %s:%d\n", cudaGetErrorString(err), __FILE__, __LINE__); return -1; }
#endif
    return 0;
}

extern "C" int lg_device_count(void)
{
#ifdef LG_EMULATE
    int n = 1;
    if (const char *e = getenv("LAMEGPU_EMU_DEVICES")) n = atoi(e);      /* tests: several emulated "devices" */
    return n;
#else
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
#endif
}

extern "C" lg_engine *lg_engine_create(const LgDevCfg *cfg, int nstreams, int max_frames, int device)
{
    if (nstreams < 1 || max_frames < 1) return NULL;
#ifndef LG_EMULATE
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        fprintf(stderr, "lamegpu: no CUDA device available - this library has no CPU path\n");
        return NULL;
    }
    if (device < 0 || device >= ndev) {
        fprintf(stderr, "lamegpu: no CUDA device %d (%d visible)\n", device, ndev);
        return NULL;
    }
#endif
    LgDeviceScope dev(device);
    if (!dev.ok) return NULL;
    lg_engine *e = (lg_engine *) calloc(1, sizeof *e);
    if (!e) return NULL;
    e->hcfg = *cfg;
    e->S = nstreams; e->F = max_frames; e->device = device;
    /* CBR/ABR at quality 3-6: kernel D in its group form (two warps per granule.channel) while the SMs have issue slots to spare */
    int const group_ok = (cfg->vbr == 0 || cfg->vbr == 3) && !(cfg->substep_shaping & 2) && cfg->noise_shaping != 0;
    e->group_nw = group_ok ? 2 : 0;
#ifndef LG_EMULATE
    {
        int nsm = 0;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
        /* more than four streams per SM: the one-warp kernel D in its seven-CTAs-per-SM build (measured 700..888 streams x 8 frames: 7.4-7.7 ms
         * against 11.1-11.7 ms; at 592 streams the two are equal), and no group form - it buys latency with issue slots (measured: 512 streams
         * 5.96 against 6.35 ms, 4096 streams 53 against 30 ms) */
        e->dense = nsm > 0 && nstreams > 4 * nsm;
        if (const char *de = getenv("LAMEGPU_DENSE")) e->dense = atoi(de);
        if (e->dense) e->group_nw = 0;
    }
#endif
    e->gate = 1;
    e->ana_split = 1;                      /* measured: 2 or 4 parts change nothing (6.39 / 6.42 / 6.39 ms per step): what stands behind kernel D is kernel A's own remainder */
    if (const char *sp = getenv("LAMEGPU_ANA_SPLIT")) e->ana_split = atoi(sp) < 1 ? 1 : (atoi(sp) > 4 ? 4 : atoi(sp));
    if (const char *ga = getenv("LAMEGPU_GATE")) e->gate = atoi(ga) != 0;
    if (const char *ge = getenv("LAMEGPU_GROUP_NW")) e->group_nw = group_ok ? atoi(ge) : 0;
    if (e->group_nw < 0 || e->group_nw > 3 || e->group_nw == 1) e->group_nw = group_ok ? 2 : 0;
    e->pcm_stride = (size_t) max_frames * 1152 + LG_PCM_HALO;
    {   /* largest frame (padded) minus its side info, per frame, plus what a full reservoir can add */
        int const max_kbps = cfg->vbr ? cfg->bitrate_kbps[cfg->vbr_max_bitrate_index] : cfg->brate;
        size_t const frame_bytes = (size_t) (cfg->version + 1) * 72000 * max_kbps / cfg->samplerate + 1;
        e->pay_stride = ((size_t) max_frames * (frame_bytes - cfg->sideinfo_len) + LG_PAY_SLACK + 15) & ~(size_t) 15;
    }
    size_t const S = nstreams, F = max_frames;
    int bad = 0;
    bad |= lg_dev_malloc((void **) &e->dcfg, sizeof(LgDevCfg));
    bad |= lg_dev_malloc((void **) &e->d_pcm16, S * 2 * e->pcm_stride * sizeof(int16_t));
    bad |= lg_dev_malloc((void **) &e->d_sb, S * (2 * F + 1) * 2 * 576 * sizeof(float));
    bad |= lg_dev_malloc((void **) &e->d_ana, S * 2 * F * sizeof(LgAnalysis));
    bad |= lg_dev_malloc((void **) &e->d_state, S * sizeof(LgStreamState));
    bad |= lg_dev_malloc((void **) &e->d_state0, S * sizeof(LgStreamState));
    bad |= lg_dev_malloc((void **) &e->d_kind, S * sizeof(LgPcmKind));
    if (cfg->resample) {
        /* inputs behind one window: its own samples, one more reference call (<= 1152 outputs) before it, the filter taps */
        e->raw_stride = (size_t) ceil((double) (e->pcm_stride + 1152) * cfg->rs_ratio) + 128;
        e->chunk_cap = 2 * max_frames + 16;
        e->raw_esz = 4;
        bad |= lg_dev_malloc((void **) &e->d_raw, S * 2 * e->raw_stride * (size_t) e->raw_esz);
        bad |= lg_dev_malloc((void **) &e->d_rsc, S * e->chunk_cap * sizeof(LgRsChunk));
        bad |= lg_dev_malloc((void **) &e->d_rss, S * sizeof(LgRsStream));
        bad |= lg_dev_malloc((void **) &e->d_pcmf, S * 2 * e->pcm_stride * sizeof(float));
    }
    for (int k = 0; k < LG_SLOTS; k++) {
        LgSlot &t = e->slot[k];
        bad |= lg_dev_malloc((void **) &t.d_xr, S * 2 * F * 2 * 576 * sizeof(float));
        bad |= lg_dev_malloc((void **) &t.d_psy, S * 2 * F * sizeof(LgPsyOut));
        bad |= lg_dev_malloc((void **) &t.d_frm, S * F * sizeof(LgFrameCtl));
        bad |= lg_dev_malloc((void **) &t.d_gout, S * 2 * F * 2 * sizeof(LgGranuleOut));
        bad |= lg_dev_malloc((void **) &t.d_fout, S * F * sizeof(LgFrameOut));
        bad |= lg_dev_malloc((void **) &t.d_pay, S * e->pay_stride);
        bad |= lg_dev_malloc((void **) &t.d_hdr, S * F * LG_HDR_STRIDE);
        bad |= lg_dev_malloc((void **) &t.d_nfr, S * sizeof(int));
        bad |= lg_host_malloc((void **) &t.h_pcm16, S * 2 * e->pcm_stride * sizeof(int16_t));
        bad |= lg_host_malloc((void **) &t.h_nfr, S * sizeof(int));
        bad |= lg_host_malloc((void **) &t.h_kind, S * sizeof(LgPcmKind));
        bad |= lg_host_malloc((void **) &t.h_pay, S * e->pay_stride);
        bad |= lg_host_malloc((void **) &t.h_hdr, S * F * LG_HDR_STRIDE);
        bad |= lg_host_malloc((void **) &t.h_fout, S * F * sizeof(LgFrameOut));
        if (cfg->resample) {
            t.chunk_cap = e->chunk_cap;
            bad |= lg_host_malloc((void **) &t.h_raw, S * 2 * e->raw_stride * (size_t) e->raw_esz);
            bad |= lg_host_malloc((void **) &t.h_rsc, S * e->chunk_cap * sizeof(LgRsChunk));
            bad |= lg_host_malloc((void **) &t.h_rss, S * sizeof(LgRsStream));
        }
    }
    if (bad) { fprintf(stderr, "lamegpu: out of memory (S=%d F=%d)\n", nstreams, max_frames); lg_engine_destroy(e); return NULL; }
#ifdef LG_EMULATE
    memcpy(e->dcfg, cfg, sizeof *cfg);
#else
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { lg_engine_destroy(e); return NULL; }
    if (cudaStreamCreateWithFlags(&e->stream4, cudaStreamNonBlocking) != cudaSuccess) { lg_engine_destroy(e); return NULL; }
    {   /* the quantiser is the critical path: its stream gets the highest priority, so its CTAs are placed before those of the next step's analysis */
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&e->stream2, cudaStreamNonBlocking, hi) != cudaSuccess) { lg_engine_destroy(e); return NULL; }
        if (cudaStreamCreateWithPriority(&e->stream2b, cudaStreamNonBlocking, hi) != cudaSuccess) { lg_engine_destroy(e); return NULL; }
        if (cudaStreamCreateWithPriority(&e->stream3, cudaStreamNonBlocking, hi) != cudaSuccess) { lg_engine_destroy(e); return NULL; }
    }
    for (int k = 0; k < LG_SLOTS; k++) for (int i = 0; i < 10; i++) cudaEventCreate(&e->slot[k].ev[i]);
    for (int k = 0; k < LG_SLOTS; k++) for (int i = 0; i < 16; i++) cudaEventCreate(&e->slot[k].evp[i / 4][i % 4]);
    for (int i = 0; i < 2; i++) cudaEventCreate(&e->ev_mark[i]);
    if (cudaMemcpy(e->dcfg, cfg, sizeof *cfg, cudaMemcpyHostToDevice) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { lg_engine_destroy(e); return NULL; }
    cudaFuncSetAttribute(lg_kernel_analysis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemA));
    cudaFuncSetAttribute(lg_kernel_quant<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemD));
    cudaFuncSetAttribute(lg_kernel_quant<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemD));
    cudaFuncSetAttribute(lg_kernel_quant<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemD));
    cudaFuncSetAttribute(lg_kernel_quant<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemD));
    cudaFuncSetAttribute(lg_kernel_quantg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemG<2>));
    cudaFuncSetAttribute(lg_kernel_quantg<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemG<3>));
    cudaFuncSetAttribute(lg_kernel_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemE));
    cudaFuncSetAttribute(lg_kernel_vbr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemV));
    cudaFuncSetAttribute(lg_kernel_vbrold<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemO));
    cudaFuncSetAttribute(lg_kernel_vbrold<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemO));
    /* The shared-memory carve-out is left to the driver.  Round 1 forced the largest one on every kernel so that a kernel could be placed
     * next to a resident one of the other stream without reconfiguring the SM; with whole steps overlapping (not pieces waiting inside
     * kernel D) that measures worse: 8.4 against 7.7 ms per pipelined 512 x 8 step - kernel D loses L1 for its tables. */
    if (const char *co = getenv("LAMEGPU_CARVEOUT")) {           /* experiment knob: one shared-memory carve-out (percent) for every kernel */
        int const pct = atoi(co);
        cudaFuncSetAttribute(lg_kernel_analysis, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(lg_kernel_scan, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(lg_kernel_mdct, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(lg_kernel_quantg<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(lg_kernel_pack, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    if (getenv("LAMEGPU_DEBUG_OCC")) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lg_kernel_quantg<2>, 128, sizeof(LgSmemG<2>));
        fprintf(stderr, "lamegpu: kernel D (group form, 2 warps) %zu B smem, %d CTAs per SM\n", sizeof(LgSmemG<2>), nb);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lg_kernel_quant<0>, 64, sizeof(LgSmemD));
        fprintf(stderr, "lamegpu: kernel D (one warp) %zu B smem, %d CTAs per SM\n", sizeof(LgSmemD), nb);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lg_kernel_analysis, 128, sizeof(LgSmemA));
        fprintf(stderr, "lamegpu: kernel A %zu B smem, %d CTAs per SM\n", sizeof(LgSmemA), nb);
    }
#endif
    {
        LgStreamState *h0 = (LgStreamState *) malloc(S * sizeof(LgStreamState));
        if (!h0) { lg_engine_destroy(e); return NULL; }
        lg_initial_state(&e->hcfg, &h0[0]);
        for (size_t i = 1; i < S; i++) h0[i] = h0[0];
#ifdef LG_EMULATE
        memcpy(e->d_state0, h0, S * sizeof(LgStreamState));
#else
        /* cudaMemcpy from pageable memory may return before the DMA has landed, and e->stream is a non-blocking
         * stream that does not order against the default stream: synchronise the device before anything reads it */
        if (cudaMemcpy(e->d_state0, h0, S * sizeof(LgStreamState), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaDeviceSynchronize() != cudaSuccess) { free(h0); lg_engine_destroy(e); return NULL; }
#endif
        free(h0);
    }
    if (lg_engine_reset_streams(e, 0, nstreams) != 0) { lg_engine_destroy(e); return NULL; }
    return e;
}

/* The native-type PCM window (lame_encode_buffer_int / _long / _float / _ieee_double ...: the caller's samples go to the device as they
 * are, kernel A converts) is allocated on first use, with rows of `esz`-byte elements (4, or 8 once a stream brings long or double
 * samples); growing it waits for the steps in flight.  The caller restages the slot it is filling. */
extern "C" int lg_engine_need_native_pcm(lg_engine *e, int esz)
{
    if (esz != 4 && esz != 8) return -1;
    if (e->pcmn_esz >= esz) return 0;
    LgDeviceScope dev(e->device);
#ifndef LG_EMULATE
    LG_CHECK(cudaStreamSynchronize(e->stream));           /* no H2D copy or kernel A may still read the old window */
#endif
    size_t const n = (size_t) e->S * 2 * e->pcm_stride * (size_t) esz;
    for (int k = 0; k < LG_SLOTS; k++) {
        lg_host_free(e->slot[k].h_pcmn); e->slot[k].h_pcmn = NULL;
        if (lg_host_malloc((void **) &e->slot[k].h_pcmn, n)) return -1;
    }
    lg_dev_free(e->d_pcmn); e->d_pcmn = NULL;
    if (lg_dev_malloc((void **) &e->d_pcmn, n)) return -1;
    e->pcmn_esz = esz;
    return 0;
}
/* the same for the resampler's input window (4-byte elements to begin with) */
extern "C" int lg_engine_need_raw(lg_engine *e, int esz)
{
    if (esz != 4 && esz != 8) return -1;
    if (!e->hcfg.resample) return -1;
    if (e->raw_esz >= esz) return 0;
    LgDeviceScope dev(e->device);
#ifndef LG_EMULATE
    LG_CHECK(cudaStreamSynchronize(e->stream));
#endif
    size_t const n = (size_t) e->S * 2 * e->raw_stride * (size_t) esz;
    for (int k = 0; k < LG_SLOTS; k++) {
        lg_host_free(e->slot[k].h_raw); e->slot[k].h_raw = NULL;
        if (lg_host_malloc((void **) &e->slot[k].h_raw, n)) return -1;
    }
    lg_dev_free(e->d_raw); e->d_raw = NULL;
    if (lg_dev_malloc((void **) &e->d_raw, n)) return -1;
    e->raw_esz = esz;
    return 0;
}

/* a launch whose streams were fed in many small calls has more chunks than usual: grow slot k's chunk list (the caller restages it) */
extern "C" int lg_engine_reserve_chunks(lg_engine *e, int k, int per_stream)
{
    LgSlot &t = e->slot[k];
    if (per_stream <= t.chunk_cap) return 0;
    LgDeviceScope dev(e->device);
    int const cap = per_stream + per_stream / 2;
    if (cap > e->chunk_cap) {
#ifndef LG_EMULATE
        LG_CHECK(cudaStreamSynchronize(e->stream));       /* nothing in flight may still read the old list */
#endif
        LgRsChunk *d = NULL;
        if (lg_dev_malloc((void **) &d, (size_t) e->S * cap * sizeof(LgRsChunk))) return -1;
        lg_dev_free(e->d_rsc);
        e->d_rsc = d; e->chunk_cap = cap;
    }
    LgRsChunk *h = NULL;
    if (lg_host_malloc((void **) &h, (size_t) e->S * cap * sizeof(LgRsChunk))) return -1;
    lg_host_free(t.h_rsc);
    t.h_rsc = h; t.chunk_cap = cap;
    return 0;
}

/* kernels D (or D', D'') and E of slot k on the quantiser stream */
static void lg_launch_quant_pack(lg_engine *e, LgSlot &t, lgStream_t ds)
{
    int const S = e->S, F = e->F;
#ifndef LG_EMULATE
    cudaEventRecord(t.ev[4], ds);
#endif
    if (e->hcfg.vbr == 4)
        LG_LAUNCH(lg_kernel_vbr, S, 128, sizeof(LgSmemV), ds, e->dcfg, t.d_xr, t.d_psy, t.d_frm, t.d_gout, t.d_fout, e->d_state, t.d_nfr, F);
    else if (e->hcfg.vbr == 2 && (e->hcfg.substep_shaping & 2))
        LG_LAUNCH(lg_kernel_vbrold<1>, S, 64, sizeof(LgSmemO), ds, e->dcfg, t.d_xr, t.d_psy, t.d_frm, t.d_gout, t.d_fout, e->d_state, t.d_nfr, F);
    else if (e->hcfg.vbr == 2)
        LG_LAUNCH(lg_kernel_vbrold<0>, S, 64, sizeof(LgSmemO), ds, e->dcfg, t.d_xr, t.d_psy, t.d_frm, t.d_gout, t.d_fout, e->d_state, t.d_nfr, F);
    else if (e->group_nw > 0) {
#define LG_LAUNCH_G(NWV) LG_LAUNCH(lg_kernel_quantg<NWV>, S, 64 * NWV, sizeof(LgSmemG<NWV>), ds, e->dcfg, t.d_xr, t.d_psy, t.d_frm, t.d_gout, t.d_fout, \
                                   e->d_state, t.d_nfr, F)
        if (e->group_nw == 2) LG_LAUNCH_G(2); else LG_LAUNCH_G(3);
#undef LG_LAUNCH_G
    }
    else {
        int const fl = ((e->hcfg.substep_shaping & 2) ? 1 : 0) | (e->dense ? 4 : 0);
#define LG_LAUNCH_D(FLV) LG_LAUNCH(lg_kernel_quant<FLV>, S, 64, sizeof(LgSmemD), ds, e->dcfg, t.d_xr, t.d_psy, t.d_frm, t.d_gout, t.d_fout, \
                                   e->d_state, t.d_nfr, F, 0, F)
        if (fl == 0) LG_LAUNCH_D(0); else if (fl == 1) LG_LAUNCH_D(1); else if (fl == 4) LG_LAUNCH_D(4); else LG_LAUNCH_D(5);
#undef LG_LAUNCH_D
    }
#ifndef LG_EMULATE
    cudaEventRecord(t.ev[5], ds);
    cudaStreamWaitEvent(e->stream3, t.ev[5], 0);
#endif
    LG_LAUNCH(lg_kernel_pack, S * F, 128, sizeof(LgSmemE), e->stream3, e->dcfg, t.d_gout, t.d_fout, t.d_pay, (int) e->pay_stride, t.d_hdr, t.d_nfr, F, 0, F);
#ifndef LG_EMULATE
    cudaEventRecord(t.ev[6], e->stream3);
#endif
    e->launches += 2;
}

/* One step on slot k, asynchronous: [H2D of the staged inputs,] kernels R A B C on the analysis stream, D E [and the D2H of the packed
 * frames] on the quantiser stream.  nframes = max over streams of the slot's frame counts.  with_copies = 0: the bench's device-only step
 * on what a previous lg_engine_submit left in device memory. */
static int lg_submit(lg_engine *e, int k, int nframes, int mode /* 0 int16 window, 1 the resampler's floats, 2 native types */, int with_copies,
                     int independent /* no stream of this step has frames in the step of the other slot */)
{
    if (k < 0 || k >= LG_SLOTS || nframes < 1 || nframes > e->F) return -1;
    LgDeviceScope dev(e->device);
    if (!dev.ok) return -1;
    LgSlot &t = e->slot[k];
    size_t const S = e->S, F = e->F;
    int const mgr = e->hcfg.mode_gr;
#ifndef LG_EMULATE
    /* the slot's previous step has left the quantiser stream (its buffers are free) - normally long ago */
    LG_CHECK(cudaStreamWaitEvent(e->stream, t.ev[7], 0));
#endif
    if (e->hcfg.resample) {
        if (with_copies) {
            LG_COPY_H2D(e->d_raw, t.h_raw, S * 2 * e->raw_stride * (size_t) e->raw_esz, e->stream);
            LG_COPY_H2D(e->d_kind, t.h_kind, S * sizeof(LgPcmKind), e->stream);
            if (t.chunk_cap == e->chunk_cap) LG_COPY_H2D(e->d_rsc, t.h_rsc, S * e->chunk_cap * sizeof(LgRsChunk), e->stream);
            else for (size_t s = 0; s < S; s++) LG_COPY_H2D(e->d_rsc + s * e->chunk_cap, t.h_rsc + s * t.chunk_cap, (size_t) t.chunk_cap * sizeof(LgRsChunk), e->stream);
            LG_COPY_H2D(e->d_rss, t.h_rss, S * sizeof(LgRsStream), e->stream);
        }
        /* kernel R turns the staged input samples + chunk lists into the float PCM window */
        int const tiles = (int) ((e->pcm_stride + 255) / 256);
#ifndef LG_EMULATE
        cudaEventRecord(t.ev[8], e->stream);
#endif
        LG_LAUNCH(lg_kernel_resample, (int) S * tiles, 256, 0, e->stream, e->dcfg, e->d_raw, (int) e->raw_stride, e->raw_esz, e->d_kind, e->d_rsc, e->chunk_cap, e->d_rss,
                  e->d_pcmf, (int) e->pcm_stride, tiles);
#ifndef LG_EMULATE
        cudaEventRecord(t.ev[9], e->stream);
#endif
        e->launches += 1;
        mode = 1;
    }
    else if (mode == 1) return -1;                            /* floats come from kernel R only */
    else if (with_copies) {
        if (mode == 2 && !t.h_pcmn) return -1;
        /* samples [0, 576*mgr*nframes + halo) of every channel row */
        size_t const n = (size_t) 576 * mgr * nframes + LG_PCM_HALO;
        size_t const esz = mode == 2 ? (size_t) e->pcmn_esz : sizeof(int16_t);
        char *dst = mode == 2 ? e->d_pcmn : (char *) e->d_pcm16;
        const char *src = mode == 2 ? t.h_pcmn : (const char *) t.h_pcm16;
#ifdef LG_EMULATE
        for (size_t r = 0; r < S * 2; r++) memcpy(dst + r * e->pcm_stride * esz, src + r * e->pcm_stride * esz, n * esz);
#else
        LG_CHECK(cudaMemcpy2DAsync(dst, e->pcm_stride * esz, src, e->pcm_stride * esz, n * esz, S * 2, cudaMemcpyHostToDevice, e->stream));
#endif
        if (mode == 2) LG_COPY_H2D(e->d_kind, t.h_kind, S * sizeof(LgPcmKind), e->stream);
    }
    if (with_copies) LG_COPY_H2D(t.d_nfr, t.h_nfr, S * sizeof(int), e->stream);
    const int16_t *p16 = mode == 0 ? e->d_pcm16 : NULL;
    const float *pf = mode == 1 ? e->d_pcmf : NULL;
    int const nslot = mgr * nframes + 1;
#ifndef LG_EMULATE
    /* Start gate: kernel A of this step may not reach the device before kernel D of the step before it (the other slot) has been launched.
     * Both hang on that step's kernel C; without the gate A - next in line on the analysis stream - takes the machine first and D's first
     * wave starts late by all of A's run time, so that the overlap gains nothing (step = sum of the kernels, measured).  With it D, on the
     * high-priority stream, fills every SM first and A runs in the room D's last, partial wave leaves (7.20 -> 6.57 ms per step at 512 x 8; LAMEGPU_GATE=0 switches it off). */
    if (e->gate) LG_CHECK(cudaStreamWaitEvent(e->stream, e->slot[k ^ 1].ev[4], 0));
    LG_CHECK(cudaStreamWaitEvent(e->stream, e->slot[k ^ 1].ev[3], 0));     /* the step before has read the analysis records and subband samples */
    cudaEventRecord(t.ev[0], e->stream);
#endif
    /* The analysis is cut into parts along the frames: kernel A part by part on the analysis stream, kernels B and C of a part on a
     * second stream behind that part's A - the scan (one warp per stream, all latency) of part p runs under kernel A of part p + 1, and
     * only the last part's B and C are left standing behind A (and, in the pipeline, behind the end of the step before's kernel D). */
    int const parts = nframes < e->ana_split ? nframes : e->ana_split;
    t.parts = parts;
    for (int p = 0; p < parts; p++) {
        int const f0 = (int) ((long) nframes * p / parts), f1 = (int) ((long) nframes * (p + 1) / parts);
        int const s0 = p == 0 ? 0 : 1 + mgr * f0, s1 = 1 + mgr * f1;        /* slot 0 = the granule before the launch */
        LG_LAUNCH(lg_kernel_analysis, (int) S * (s1 - s0), 128, sizeof(LgSmemA), e->stream,
                  e->dcfg, p16, (int) e->pcm_stride, pf, e->d_pcmn, e->pcmn_esz, e->d_kind, e->d_sb, e->d_ana, t.d_nfr, 2 * (int) F + 1, s0, s1 - s0);
#ifndef LG_EMULATE
        cudaEventRecord(t.evp[p][0], e->stream);
        if (p == parts - 1) cudaEventRecord(t.ev[1], e->stream);
        LG_CHECK(cudaStreamWaitEvent(e->stream4, t.evp[p][0], 0));
        cudaEventRecord(t.evp[p][1], e->stream4);
#endif
        LG_LAUNCH(lg_kernel_scan, (int) S, 32, sizeof(LgSmemB), e->stream4, e->dcfg, e->d_ana, t.d_psy, t.d_frm, e->d_state, t.d_nfr, (int) F, f0, f1);
#ifndef LG_EMULATE
        cudaEventRecord(t.evp[p][2], e->stream4);
#endif
        LG_LAUNCH(lg_kernel_mdct, (int) S * mgr * (f1 - f0), 64, sizeof(LgSmemC), e->stream4, e->dcfg, e->d_sb, t.d_psy, t.d_frm, t.d_xr, t.d_nfr, (int) F,
                  mgr * f0, mgr * (f1 - f0));
#ifndef LG_EMULATE
        cudaEventRecord(t.evp[p][3], e->stream4);
#endif
        e->launches += 3;
    }
#ifndef LG_EMULATE
    cudaEventRecord(t.ev[3], e->stream4);
#endif
    /* Kernel D of consecutive steps is ordered (a stream's reservoir and step-size memory pass from one to the next) - unless the two
     * steps share no stream, as happens behind the lame_t handles, where a launch gathers the lanes that are not in flight: then the
     * two kernels D, each all latency, run side by side on their own CUDA streams. */
    lgStream_t const ds = (k & 1) ? e->stream2b : e->stream2;
#ifndef LG_EMULATE
    LG_CHECK(cudaStreamWaitEvent(ds, t.ev[3], 0));
    if (!independent) LG_CHECK(cudaStreamWaitEvent(ds, e->slot[k ^ 1].ev[5], 0));
#endif
    lg_launch_quant_pack(e, t, ds);
    if (with_copies) {
        LG_COPY_D2H(t.h_fout, t.d_fout, S * F * sizeof(LgFrameOut), e->stream3);
        LG_COPY_D2H(t.h_pay, t.d_pay, S * e->pay_stride, e->stream3);
        LG_COPY_D2H(t.h_hdr, t.d_hdr, S * F * LG_HDR_STRIDE, e->stream3);
    }
#ifndef LG_EMULATE
    cudaEventRecord(t.ev[7], e->stream3);
    if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "lamegpu: kernel launch failed\n"); return -1; }
#endif
    t.in_flight = 1; t.nframes = nframes;
    return 0;
}
extern "C" int lg_engine_submit(lg_engine *e, int k, int nframes, int mode) { return lg_submit(e, k, nframes, mode, 1, 0); }
extern "C" int lg_engine_submit_independent(lg_engine *e, int k, int nframes, int mode) { return lg_submit(e, k, nframes, mode, 1, 1); }
extern "C" int lg_engine_run_device(lg_engine *e, int k, int nframes, int mode) { return lg_submit(e, k, nframes, mode, 0, 0); }

/* block until slot k's step has finished and its results are in the slot's host buffers; fills the kernel times of that step */
extern "C" int lg_engine_wait(lg_engine *e, int k)
{
    if (k < 0 || k >= LG_SLOTS) return -1;
    LgSlot &t = e->slot[k];
    if (!t.in_flight) return 0;
    LgDeviceScope dev(e->device);
#ifndef LG_EMULATE
    LG_CHECK(cudaEventSynchronize(t.ev[7]));
    static const int first[5] = { 0, 1, 2, 4, 5 };
    for (int i = 0; i < 5; i++) { float ms = 0; if (i != 1 && i != 2) cudaEventElapsedTime(&ms, t.ev[first[i]], t.ev[first[i] + 1]); t.ms[i] = ms; }
    for (int p = 0; p < t.parts; p++) {                     /* kernels B and C: the sum over the parts */
        float ms = 0;
        cudaEventElapsedTime(&ms, t.evp[p][1], t.evp[p][2]); t.ms[1] += ms;
        cudaEventElapsedTime(&ms, t.evp[p][2], t.evp[p][3]); t.ms[2] += ms;
    }
    t.ms[5] = 0.f;
    if (e->hcfg.resample) { float ms = 0; cudaEventElapsedTime(&ms, t.ev[8], t.ev[9]); t.ms[5] = ms; }
    { float ms = 0; cudaEventElapsedTime(&ms, t.ev[0], t.ev[6]); t.ms[7] = ms; }
    memcpy(e->last_ms, t.ms, sizeof e->last_ms);
#endif
    t.in_flight = 0;
    return 0;
}
extern "C" int lg_engine_in_flight(const lg_engine *e, int k) { return e->slot[k].in_flight; }

/* device-time marks for a run of pipelined steps (bench): which = 0 on the analysis stream in front of the next submit, which = 1 on the
 * quantiser stream behind the last one; the elapsed time is read after a wait on the last slot */
extern "C" int lg_engine_mark(lg_engine *e, int which)
{
#ifndef LG_EMULATE
    LgDeviceScope dev(e->device);
    LG_CHECK(cudaEventRecord(e->ev_mark[which ? 1 : 0], which ? e->stream3 : e->stream));
#else
    (void) e; (void) which;
#endif
    return 0;
}
extern "C" float lg_engine_marked_ms(lg_engine *e)
{
    float ms = 0.f;
#ifndef LG_EMULATE
    LgDeviceScope dev(e->device);
    if (cudaEventSynchronize(e->ev_mark[1]) != cudaSuccess || cudaEventElapsedTime(&ms, e->ev_mark[0], e->ev_mark[1]) != cudaSuccess) return -1.f;
#else
    (void) e;
#endif
    return ms;
}

/* test/debug hook: copy an intermediate device buffer to the host.
 * what: 0 sb, 1 ana, 2 psy, 3 frm, 4 xr, 5 gout, 6 fout, 7 state, 8 the configuration's fast_log2 table (host copy) */
extern "C" long lg_engine_debug_copy(lg_engine *e, int what, void *dst, size_t cap)
{
    size_t const S = e->S, F = e->F;
    LgDeviceScope dev(e->device);
    const LgSlot &t = e->slot[0];
    const void *src = NULL; size_t n = 0;
    if (what == 8) { n = sizeof e->hcfg.log_table < cap ? sizeof e->hcfg.log_table : cap; memcpy(dst, e->hcfg.log_table, n); return (long) n; }
    switch (what) {
    case 0: src = e->d_sb; n = S * (2 * F + 1) * 2 * 576 * sizeof(float); break;
    case 1: src = e->d_ana; n = S * 2 * F * sizeof(LgAnalysis); break;
    case 2: src = t.d_psy; n = S * 2 * F * sizeof(LgPsyOut); break;
    case 3: src = t.d_frm; n = S * F * sizeof(LgFrameCtl); break;
    case 4: src = t.d_xr; n = S * 2 * F * 2 * 576 * sizeof(float); break;
    case 5: src = t.d_gout; n = S * 2 * F * 2 * sizeof(LgGranuleOut); break;
    case 6: src = t.d_fout; n = S * F * sizeof(LgFrameOut); break;
    case 7: src = e->d_state; n = S * sizeof(LgStreamState); break;
    default: return -1;
    }
    if (n > cap) n = cap;
#ifdef LG_EMULATE
    memcpy(dst, src, n);
#else
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
#endif
    return (long) n;
}

#ifndef LG_EMULATE
/* Test hook (tests/test_gpu_parity.py): the restated libm functions of lg_math.cuh evaluated on the device, so that the GPU suite can
 * compare them with the host's libm argument by argument.  fn: 0 lg_powf(x, y), 1 lg_log10f(x), 2 lg_exp(x), 3 lg_pow(x, y); the
 * binary32 ones take and return their floats widened to double (exact). */
__global__ void lg_kernel_math_selftest(int fn, const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ out, int n)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (fn) {
    case 0: out[i] = (double) lg_powf((float) x[i], (float) y[i]); break;
    case 1: out[i] = (double) lg_log10f((float) x[i]); break;
    case 2: out[i] = lg_exp(x[i]); break;
    default: out[i] = lg_pow(x[i], y[i]); break;
    }
}
extern "C" int lamegpu_math_selftest(int fn, const double *x, const double *y, double *out, int n)
{
    if (fn < 0 || fn > 3 || n <= 0 || !x || !y || !out) return -1;
    double *d = nullptr;
    size_t const bytes = (size_t) n * sizeof(double);
    if (cudaMalloc(&d, 3 * bytes) != cudaSuccess) { fprintf(stderr, "lamegpu: no CUDA device for the math self-test\n"); return -1; }
    cudaMemcpy(d, x, bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(d + n, y, bytes, cudaMemcpyHostToDevice);
    lg_kernel_math_selftest<<<(n + 255) / 256, 256>>>(fn, d, d + n, d + 2 * (size_t) n, n);
    cudaError_t const err = cudaMemcpy(out, d + 2 * (size_t) n, bytes, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (err != cudaSuccess) { fprintf(stderr, "lamegpu: math self-test: %s\n", cudaGetErrorString(err)); return -1; }
    return 0;
}
#endif
