// lg_engine.cu - device side of the batch engine: owns the HBM buffers of one configuration
// (S streams x up to F frames per launch), the four kernels and the copies.
//
//   H2D  pcm (int16 or float, pinned)  [->  R resample, when the input rate differs]  ->  A analysis  ->  B scan  ->  C mdct  ->  D quantise  ->  E pack  ->  D2H bytes
//
// All work of one launch goes to one CUDA stream; the host (lg_bitstream.cpp) only interleaves the packed
// payload bytes with the frame headers.  There is no CPU fallback: without a CUDA device lg_engine_create() fails and says so.
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

#define LG_MAX_PIECES 8

struct lg_engine {
    LgDevCfg hcfg;
    LgDevCfg *dcfg;
    int S, F, device;
    size_t pcm_stride;                /* samples per channel per stream */
    int16_t *d_pcm16; float *d_pcmf;
    float *d_sb, *d_xr;
    LgAnalysis *d_ana; LgPsyOut *d_psy; LgFrameCtl *d_frm;
    LgGranuleOut *d_gout; LgFrameOut *d_fout;
    unsigned char *d_pay, *d_hdr;      /* kernel E: payload bytes [S][pay_stride], header + side info [S][F][LG_HDR_STRIDE] */
    size_t pay_stride;
    LgStreamState *d_state, *d_state0;   /* d_state0: S copies of the initial state, for one-copy resets */
    int *d_nfr;
    /* pinned host staging */
    int16_t *h_pcm16; float *h_pcmf; int *h_nfr;
    LgFrameOut *h_fout; unsigned char *h_pay, *h_hdr;
    /* kernel R (input-rate conversion): raw input samples, the chunk list and the per-stream header, device + pinned */
    size_t raw_stride; int chunk_cap;
    float *d_raw, *h_raw; LgRsChunk *d_rsc, *h_rsc; LgRsStream *d_rss, *h_rss;
    /* two CUDA streams: `stream` carries the copies and the stateless/scan kernels (A, B, C) of the batch piece by piece, `stream2` the
     * quantiser and the packer (D, E), each piece as soon as its A-B-C is done - kernel D is latency-bound and leaves most issue
     * slots free, so the next piece's A-B-C (and its share of the H2D copy) run underneath it */
    lgStream_t stream, stream2;
    int dense;                        /* more than six streams per SM: kernel D in its seven-CTAs-per-SM build */
    int group_nw;                     /* > 0: kernel D in its group form (lg_k_quantg.cuh) with this many warps per granule.channel */
    int pieces;                       /* how many pieces a launch is cut into along the frame axis (1 = no overlap) */
    int *d_ready;                     /* one flag per piece: raised on stream 1 behind the piece's kernel C, awaited by kernel D on stream 2 */
#ifndef LG_EMULATE
    cudaEvent_t ev[8];                /* 6, 7: around kernel R */
    cudaEvent_t pev[LG_MAX_PIECES][8];/* per piece: stream 1 before A, after A, after B, after C; stream 2 before D, after D, after E */
    cudaEvent_t ev_begin, ev_end;
    int pieces_used;
#endif
    float last_ms[8];                 /* A, B, C, D, E summed over the pieces, R, -, 7: begin of the first kernel to end of the last */
    long launches;
};

extern "C" const LgDevCfg *lg_engine_config(const lg_engine *e) { return &e->hcfg; }
extern "C" int lg_engine_streams(const lg_engine *e) { return e->S; }
extern "C" int lg_engine_max_frames(const lg_engine *e) { return e->F; }
extern "C" size_t lg_engine_pcm_stride(const lg_engine *e) { return e->pcm_stride; }
extern "C" int16_t *lg_engine_host_pcm16(lg_engine *e) { return e->h_pcm16; }
extern "C" float *lg_engine_host_pcmf(lg_engine *e) { return e->h_pcmf; }
extern "C" int *lg_engine_host_nfr(lg_engine *e) { return e->h_nfr; }
extern "C" const unsigned char *lg_engine_host_pay(const lg_engine *e) { return e->h_pay; }
extern "C" const unsigned char *lg_engine_host_hdr(const lg_engine *e) { return e->h_hdr; }
extern "C" size_t lg_engine_pay_stride(const lg_engine *e) { return e->pay_stride; }
extern "C" const LgFrameOut *lg_engine_host_fout(const lg_engine *e) { return e->h_fout; }
extern "C" const float *lg_engine_last_kernel_ms(const lg_engine *e) { return e->last_ms; }
extern "C" long lg_engine_launch_count(const lg_engine *e) { return e->launches; }
extern "C" void *lg_engine_device_pcm16(lg_engine *e) { return e->d_pcm16; }
extern "C" size_t lg_engine_raw_stride(const lg_engine *e) { return e->raw_stride; }
extern "C" float *lg_engine_host_raw(lg_engine *e) { return e->h_raw; }
extern "C" LgRsChunk *lg_engine_host_chunks(lg_engine *e) { return e->h_rsc; }
extern "C" int *lg_engine_host_rs_counts(lg_engine *e) { return (int *) e->h_rss; }
extern "C" int lg_engine_chunk_cap(const lg_engine *e) { return e->chunk_cap; }

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
#ifndef LG_EMULATE
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->stream2) cudaStreamSynchronize(e->stream2);
#endif
    lg_dev_free(e->dcfg); lg_dev_free(e->d_pcm16); lg_dev_free(e->d_pcmf); lg_dev_free(e->d_sb); lg_dev_free(e->d_xr);
    lg_dev_free(e->d_ana); lg_dev_free(e->d_psy); lg_dev_free(e->d_frm); lg_dev_free(e->d_gout); lg_dev_free(e->d_fout); lg_dev_free(e->d_pay); lg_dev_free(e->d_hdr);
    lg_dev_free(e->d_state); lg_dev_free(e->d_state0); lg_dev_free(e->d_nfr); lg_dev_free(e->d_ready);
    lg_dev_free(e->d_raw); lg_dev_free(e->d_rsc); lg_dev_free(e->d_rss); lg_host_free(e->h_raw); lg_host_free(e->h_rsc); lg_host_free(e->h_rss);
    lg_host_free(e->h_pcm16); lg_host_free(e->h_pcmf); lg_host_free(e->h_nfr); lg_host_free(e->h_pay); lg_host_free(e->h_hdr); lg_host_free(e->h_fout);
#ifndef LG_EMULATE
    for (int i = 0; i < 8; i++) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    for (int p = 0; p < LG_MAX_PIECES; p++) for (int i = 0; i < 8; i++) if (e->pev[p][i]) cudaEventDestroy(e->pev[p][i]);
    if (e->ev_begin) cudaEventDestroy(e->ev_begin);
    if (e->ev_end) cudaEventDestroy(e->ev_end);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->stream2) cudaStreamDestroy(e->stream2);
#endif
    free(e);
}

extern "C" int lg_engine_reset_streams(lg_engine *e, int first, int count)
{
    if (first < 0 || count < 0 || first + count > e->S) return -1;
#ifdef LG_EMULATE
    memcpy(e->d_state + first, e->d_state0 + first, (size_t) count * sizeof(LgStreamState));
#else
    LG_CHECK(cudaMemcpyAsync(e->d_state + first, e->d_state0 + first, (size_t) count * sizeof(LgStreamState), cudaMemcpyDeviceToDevice, e->stream));
#endif
    return 0;
}

/* Kernels of the two streams share SMs when a launch runs in pieces (A-B-C of a later piece next to the resident kernel D): then every
 * kernel asks for the same, largest shared-memory carve-out, so that placing one next to the other never needs an SM to be reconfigured
 * (which would wait for it to drain).  Without pieces the driver's own choice is better: it leaves more L1 for the tables. */
static void lg_set_carveout(int max_shared)
{
#ifndef LG_EMULATE
    int const v = max_shared ? (int) cudaSharedmemCarveoutMaxShared : (int) cudaSharedmemCarveoutDefault;
    cudaFuncSetAttribute(lg_kernel_analysis, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(lg_kernel_scan, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(lg_kernel_mdct, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(lg_kernel_quant<0>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(lg_kernel_quant<1>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(lg_kernel_pack, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    cudaFuncSetAttribute(lg_kernel_piece_ready, cudaFuncAttributePreferredSharedMemoryCarveout, v);
#else
    (void) max_shared;
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
    if (device < 0 || device >= ndev) device = 0;
    if (cudaSetDevice(device) != cudaSuccess) return NULL;
#endif
    lg_engine *e = (lg_engine *) calloc(1, sizeof *e);
    if (!e) return NULL;
    e->hcfg = *cfg;
    e->S = nstreams; e->F = max_frames; e->device = device;
    /* VBR kernels settle a frame's size over all its granules at once and run as one piece; CBR/ABR: up to 8 pieces (measured at 512 x 8: 2 pieces 8.02 ms, 4: 7.91, 8: 7.79; one piece 8.23) */
    e->pieces = (cfg->vbr == 0 || cfg->vbr == 3) ? LG_MAX_PIECES : 1;
    if (const char *pe = getenv("LAMEGPU_PIECES")) e->pieces = atoi(pe);
    if (e->pieces < 1) e->pieces = 1;
    if (e->pieces > LG_MAX_PIECES) e->pieces = LG_MAX_PIECES;
    if (cfg->vbr == 4 || cfg->vbr == 2) e->pieces = 1;
    if (cfg->noise_shaping == 0) e->pieces = 1;
    /* CBR/ABR at quality >= 3: kernel D in its group form (three warps per granule.channel, lines in registers); it runs behind A-B-C as one piece */
    e->group_nw = ((cfg->vbr == 0 || cfg->vbr == 3) && !(cfg->substep_shaping & 2)) ? 3 : 0;
    if (const char *ge = getenv("LAMEGPU_GROUP_NW")) e->group_nw = ((cfg->vbr == 0 || cfg->vbr == 3) && !(cfg->substep_shaping & 2)) ? atoi(ge) : 0;
    if (e->group_nw < 0 || e->group_nw > 3) e->group_nw = 3;
    if (e->group_nw > 0) e->pieces = 1;      /* quality 7-9: kernel D is too short to hide anything under (1.70e6 frames/s as one piece, 1.28e6 in pieces) */
    /* a profiler that serialises kernels (ncu replays each launch alone) would leave kernel D waiting for flags nobody can raise */
    if (getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("COMPUTE_SANITIZER_INJECTION")) e->pieces = 1;
#ifndef LG_EMULATE
    {   /* kernel D waits inside the kernel for the later pieces, whose kernels A-B-C need room on the SMs next to it: up to 7 of D's CTAs
         * fit on an SM (31 KB shared memory each), so with too many streams per SM the device fills up with waiting CTAs: measured on a B200, 640
         * streams (4.3 per SM) run, 740 (5 per SM) do not.  Above 4 per SM the batch runs as one piece.  (The wait is bounded in any case: lg_wait_piece traps after ~2 s.) */
        int nsm = 0;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
        if (nsm < 1 || nstreams > 4 * nsm) e->pieces = 1;
        /* without pieces and with more than four streams per SM: kernel D in its seven-CTAs-per-SM build (measured 700..888 streams x 8
         * frames: 7.4-7.7 ms against 11.1-11.7 ms; at 592 streams the two are equal) */
        e->dense = nsm > 0 && nstreams > 4 * nsm;
        if (const char *de = getenv("LAMEGPU_DENSE")) e->dense = atoi(de);
    }
#endif
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
    bad |= lg_dev_malloc((void **) &e->d_xr, S * 2 * F * 2 * 576 * sizeof(float));
    bad |= lg_dev_malloc((void **) &e->d_ana, S * 2 * F * sizeof(LgAnalysis));
    bad |= lg_dev_malloc((void **) &e->d_psy, S * 2 * F * sizeof(LgPsyOut));
    bad |= lg_dev_malloc((void **) &e->d_frm, S * F * sizeof(LgFrameCtl));
    bad |= lg_dev_malloc((void **) &e->d_gout, S * 2 * F * 2 * sizeof(LgGranuleOut));
    bad |= lg_dev_malloc((void **) &e->d_fout, S * F * sizeof(LgFrameOut));
    bad |= lg_dev_malloc((void **) &e->d_pay, S * e->pay_stride);
    bad |= lg_dev_malloc((void **) &e->d_hdr, S * F * LG_HDR_STRIDE);
    bad |= lg_dev_malloc((void **) &e->d_state, S * sizeof(LgStreamState));
    bad |= lg_dev_malloc((void **) &e->d_state0, S * sizeof(LgStreamState));
    bad |= lg_dev_malloc((void **) &e->d_nfr, S * sizeof(int));
    bad |= lg_dev_malloc((void **) &e->d_ready, LG_MAX_PIECES * sizeof(int));
    bad |= lg_host_malloc((void **) &e->h_pcm16, S * 2 * e->pcm_stride * sizeof(int16_t));
    bad |= lg_host_malloc((void **) &e->h_nfr, S * sizeof(int));
    bad |= lg_host_malloc((void **) &e->h_pay, S * e->pay_stride);
    bad |= lg_host_malloc((void **) &e->h_hdr, S * F * LG_HDR_STRIDE);
    bad |= lg_host_malloc((void **) &e->h_fout, S * F * sizeof(LgFrameOut));
    if (cfg->resample) {
        /* inputs behind one window: its own samples, one more reference call (<= 1152 outputs) before it, the filter taps */
        e->raw_stride = (size_t) ceil((double) (e->pcm_stride + 1152) * cfg->rs_ratio) + 128;
        e->chunk_cap = 2 * max_frames + 16;
        bad |= lg_dev_malloc((void **) &e->d_raw, S * 2 * e->raw_stride * sizeof(float));
        bad |= lg_host_malloc((void **) &e->h_raw, S * 2 * e->raw_stride * sizeof(float));
        bad |= lg_dev_malloc((void **) &e->d_rsc, S * e->chunk_cap * sizeof(LgRsChunk));
        bad |= lg_host_malloc((void **) &e->h_rsc, S * e->chunk_cap * sizeof(LgRsChunk));
        bad |= lg_dev_malloc((void **) &e->d_rss, S * sizeof(LgRsStream));
        bad |= lg_host_malloc((void **) &e->h_rss, S * sizeof(LgRsStream));
        bad |= lg_dev_malloc((void **) &e->d_pcmf, S * 2 * e->pcm_stride * sizeof(float));
    }
    if (bad) { fprintf(stderr, "lamegpu: out of memory (S=%d F=%d)\n", nstreams, max_frames); lg_engine_destroy(e); return NULL; }
#ifdef LG_EMULATE
    memcpy(e->dcfg, cfg, sizeof *cfg);
#else
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { lg_engine_destroy(e); return NULL; }
    {   /* kernel D is the critical path: its stream gets the highest priority, so its CTAs are placed before those of the pieces' kernels */
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (getenv("LAMEGPU_NO_PRIO")) hi = lo;
        if (cudaStreamCreateWithPriority(&e->stream2, cudaStreamNonBlocking, hi) != cudaSuccess) { lg_engine_destroy(e); return NULL; }
    }
    for (int i = 0; i < 8; i++) cudaEventCreate(&e->ev[i]);
    for (int p = 0; p < LG_MAX_PIECES; p++) for (int i = 0; i < 8; i++) cudaEventCreate(&e->pev[p][i]);
    cudaEventCreate(&e->ev_begin); cudaEventCreate(&e->ev_end);
    if (cudaMemcpy(e->dcfg, cfg, sizeof *cfg, cudaMemcpyHostToDevice) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { lg_engine_destroy(e); return NULL; }
    cudaFuncSetAttribute(lg_kernel_analysis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemA));
    cudaFuncSetAttribute(lg_kernel_quant<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemD));
    cudaFuncSetAttribute(lg_kernel_quant<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemD));
    cudaFuncSetAttribute(lg_kernel_quant<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemD));
    cudaFuncSetAttribute(lg_kernel_quant<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemD));
    cudaFuncSetAttribute(lg_kernel_quantg<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemG<1>));
    cudaFuncSetAttribute(lg_kernel_quantg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemG<2>));
    cudaFuncSetAttribute(lg_kernel_quantg<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemG<3>));
    cudaFuncSetAttribute(lg_kernel_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemE));
    cudaFuncSetAttribute(lg_kernel_vbr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemV));
    if (getenv("LAMEGPU_DEBUG_OCC")) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lg_kernel_quant<0>, 64, sizeof(LgSmemD));
        fprintf(stderr, "lamegpu: kernel D %zu B smem, %d CTAs per SM\n", sizeof(LgSmemD), nb);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lg_kernel_analysis, 128, sizeof(LgSmemA));
        fprintf(stderr, "lamegpu: kernel A %zu B smem, %d CTAs per SM\n", sizeof(LgSmemA), nb);
    }
    cudaFuncSetAttribute(lg_kernel_vbrold<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemO));
    cudaFuncSetAttribute(lg_kernel_vbrold<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LgSmemO));
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
    lg_set_carveout(e->pieces > 1);
    if (lg_engine_reset_streams(e, 0, nstreams) != 0) { lg_engine_destroy(e); return NULL; }
    return e;
}

/* float staging is allocated on first use (only the float/int32/double entry points need it) */
extern "C" int lg_engine_need_float_pcm(lg_engine *e)
{
    if (e->h_pcmf) return 0;
    size_t const n = (size_t) e->S * 2 * e->pcm_stride * sizeof(float);
    if (lg_host_malloc((void **) &e->h_pcmf, n)) return -1;
    if (!e->d_pcmf && lg_dev_malloc((void **) &e->d_pcmf, n)) return -1;
    return 0;
}

/* a launch whose streams were fed in many small calls has more chunks than usual: grow the chunk list (contents are
 * restaged by the caller afterwards) */
extern "C" int lg_engine_reserve_chunks(lg_engine *e, int per_stream)
{
    if (per_stream <= e->chunk_cap) return 0;
#ifndef LG_EMULATE
    LG_CHECK(cudaStreamSynchronize(e->stream));       /* nothing in flight may still read the old list */
#endif
    int const cap = per_stream + per_stream / 2;
    LgRsChunk *d = NULL, *h = NULL;
    if (lg_dev_malloc((void **) &d, (size_t) e->S * cap * sizeof(LgRsChunk)) || lg_host_malloc((void **) &h, (size_t) e->S * cap * sizeof(LgRsChunk))) {
        lg_dev_free(d); lg_host_free(h);
        return -1;
    }
    lg_dev_free(e->d_rsc); lg_host_free(e->h_rsc);
    e->d_rsc = d; e->h_rsc = h; e->chunk_cap = cap;
    return 0;
}

/* Launch the kernels on what is already in device memory (bench "value": inputs resident in HBM).  nframes = max over streams of
 * d_nfr[].  The batch is cut into `pieces` along the frame axis: A, B, C of piece i on stream 1, D and E of piece i on stream 2 behind an
 * event, so that A-B-C of piece i+1 run under kernel D of piece i.  When h2d_pcm is set, each piece's share of the staged PCM is copied in
 * front of its kernel A (the 1328-sample halo travels with the first piece). */
/* kernels D (or D', D'') and E over the whole batch on stream 2; D waits for the pieces' flags as it reaches their frames */
static void lg_launch_quant_pack(lg_engine *e, int nframes, int P)
{
    int const S = e->S, F = e->F;
#ifndef LG_EMULATE
    cudaEventRecord(e->pev[0][4], e->stream2);
#endif
    if (e->hcfg.vbr == 4)
        LG_LAUNCH(lg_kernel_vbr, S, 128, sizeof(LgSmemV), e->stream2, e->dcfg, e->d_xr, e->d_psy, e->d_frm, e->d_gout, e->d_fout,
                  e->d_state, e->d_nfr, F);
    else if (e->hcfg.vbr == 2 && (e->hcfg.substep_shaping & 2))
        LG_LAUNCH(lg_kernel_vbrold<1>, S, 64, sizeof(LgSmemO), e->stream2, e->dcfg, e->d_xr, e->d_psy, e->d_frm, e->d_gout, e->d_fout,
                  e->d_state, e->d_nfr, F);
    else if (e->hcfg.vbr == 2)
        LG_LAUNCH(lg_kernel_vbrold<0>, S, 64, sizeof(LgSmemO), e->stream2, e->dcfg, e->d_xr, e->d_psy, e->d_frm, e->d_gout, e->d_fout,
                  e->d_state, e->d_nfr, F);
    else if (e->group_nw > 0) {
#define LG_LAUNCH_G(NWV) LG_LAUNCH(lg_kernel_quantg<NWV>, S, 64 * NWV, sizeof(LgSmemG<NWV>), e->stream2, e->dcfg, e->d_xr, e->d_psy, e->d_frm, e->d_gout, e->d_fout, \
                                   e->d_state, e->d_nfr, F)
        if (e->group_nw == 1) LG_LAUNCH_G(1); else if (e->group_nw == 2) LG_LAUNCH_G(2); else LG_LAUNCH_G(3);
#undef LG_LAUNCH_G
    }
    else {
        int const fl = ((e->hcfg.substep_shaping & 2) ? 1 : 0) | (e->dense ? 4 : 0);
#define LG_LAUNCH_D(FLV) LG_LAUNCH(lg_kernel_quant<FLV>, S, 64, sizeof(LgSmemD), e->stream2, e->dcfg, e->d_xr, e->d_psy, e->d_frm, e->d_gout, e->d_fout, \
                                   e->d_state, e->d_nfr, F, 0, F, e->d_ready, P, nframes)
        if (fl == 0) LG_LAUNCH_D(0); else if (fl == 1) LG_LAUNCH_D(1); else if (fl == 4) LG_LAUNCH_D(4); else LG_LAUNCH_D(5);
#undef LG_LAUNCH_D
    }
#ifndef LG_EMULATE
    cudaEventRecord(e->pev[0][5], e->stream2);
#endif
    LG_LAUNCH(lg_kernel_pack, S * F, 128, sizeof(LgSmemE), e->stream2, e->dcfg, e->d_gout, e->d_fout, e->d_pay, (int) e->pay_stride, e->d_hdr,
              e->d_nfr, F, 0, F);
#ifndef LG_EMULATE
    cudaEventRecord(e->pev[0][6], e->stream2);
#endif
    e->launches += 2;
}

static int lg_run_pieces(lg_engine *e, int nframes, int use_float, int h2d_pcm)
{
    int const S = e->S, F = e->F, mgr = e->hcfg.mode_gr;
    if (nframes < 1 || nframes > F) return -1;
    const int16_t *p16 = use_float ? NULL : e->d_pcm16;
    const float *pf = use_float ? e->d_pcmf : NULL;
    int P = e->pieces;
    if (P > nframes) P = nframes;
#ifndef LG_EMULATE
    e->pieces_used = P;
    if (P > 1) LG_CHECK(cudaMemsetAsync(e->d_ready, 0, LG_MAX_PIECES * sizeof(int), e->stream));
    cudaEventRecord(e->ev_begin, e->stream);
#endif
    for (int p = 0; p < P; p++) {
        int const f0 = (int) ((long) nframes * p / P), f1 = (int) ((long) nframes * (p + 1) / P);
        if (h2d_pcm) {
            /* samples [576*mgr*f0 (+ the halo, for the first piece: from 0), 576*mgr*f1 + halo) of every channel row */
            size_t const a = (p == 0) ? 0 : (size_t) 576 * mgr * f0 + LG_PCM_HALO, b = (size_t) 576 * mgr * f1 + LG_PCM_HALO;
            size_t const esz = use_float ? sizeof(float) : sizeof(int16_t);
            char *dst = use_float ? (char *) e->d_pcmf : (char *) e->d_pcm16;
            const char *src = use_float ? (const char *) e->h_pcmf : (const char *) e->h_pcm16;
#ifdef LG_EMULATE
            for (size_t r = 0; r < (size_t) S * 2; r++) memcpy(dst + (r * e->pcm_stride + a) * esz, src + (r * e->pcm_stride + a) * esz, (b - a) * esz);
#else
            LG_CHECK(cudaMemcpy2DAsync(dst + a * esz, e->pcm_stride * esz, src + a * esz, e->pcm_stride * esz, (b - a) * esz, (size_t) S * 2,
                                       cudaMemcpyHostToDevice, e->stream));
#endif
        }
        int const slot0 = (f0 == 0) ? 0 : mgr * f0 + 1, nslot = mgr * f1 - slot0 + 1;
#ifndef LG_EMULATE
        cudaEventRecord(e->pev[p][0], e->stream);
#endif
        LG_LAUNCH(lg_kernel_analysis, S * nslot, 128, sizeof(LgSmemA), e->stream,
                  e->dcfg, p16, (int) e->pcm_stride, pf, e->d_sb, e->d_ana, e->d_nfr, 2 * F + 1, slot0, nslot);
#ifndef LG_EMULATE
        cudaEventRecord(e->pev[p][1], e->stream);
#endif
        LG_LAUNCH(lg_kernel_scan, S, 32, sizeof(LgSmemB), e->stream, e->dcfg, e->d_ana, e->d_psy, e->d_frm, e->d_state, e->d_nfr, F, f0, f1);
#ifndef LG_EMULATE
        cudaEventRecord(e->pev[p][2], e->stream);
#endif
        LG_LAUNCH(lg_kernel_mdct, S * mgr * (f1 - f0), 64, sizeof(LgSmemC), e->stream, e->dcfg, e->d_sb, e->d_psy, e->d_frm, e->d_xr, e->d_nfr, F,
                  mgr * f0, mgr * (f1 - f0));
#ifndef LG_EMULATE
        cudaEventRecord(e->pev[p][3], e->stream);
        if (P > 1) { lg_kernel_piece_ready<<<1, 1, 0, e->stream>>>(e->d_ready, p); e->launches += 1; }
        if (p == 0) { cudaStreamWaitEvent(e->stream2, e->pev[0][3], 0); lg_launch_quant_pack(e, nframes, P); }
#endif
        e->launches += 3;
    }
#ifdef LG_EMULATE
    lg_launch_quant_pack(e, nframes, P);          /* the emulator runs launches one after the other: quantise when everything is there */
#endif
#ifndef LG_EMULATE
    cudaEventRecord(e->ev_end, e->stream2);
#endif
    return 0;
}

extern "C" int lg_engine_run_device(lg_engine *e, int nframes, int use_float) { return lg_run_pieces(e, nframes, use_float, 0); }

extern "C" int lg_engine_sync(lg_engine *e)
{
#ifndef LG_EMULATE
    LG_CHECK(cudaStreamSynchronize(e->stream));
    LG_CHECK(cudaStreamSynchronize(e->stream2));
    for (int i = 0; i < 5; i++) e->last_ms[i] = 0.f;
    static const int first[5] = { 0, 1, 2, 4, 5 };
    for (int p = 0; p < e->pieces_used; p++)
        for (int i = 0; i < (p == 0 ? 5 : 3); i++) { float ms = 0; cudaEventElapsedTime(&ms, e->pev[p][first[i]], e->pev[p][first[i] + 1]); e->last_ms[i] += ms; }
    { float ms = 0; cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end); e->last_ms[7] = ms; }
#endif
    return 0;
}

/* Full step: H2D of the staged PCM + frame counts, kernels, D2H of the packed frames, sync. */
extern "C" int lg_engine_encode(lg_engine *e, int nframes, int use_float)
{
    size_t const S = e->S, F = e->F;
    int h2d_pcm = 1;
    if (e->hcfg.resample) {
        /* kernel R turns the staged input samples + chunk lists into the float PCM window */
        int const tiles = (int) ((e->pcm_stride + 255) / 256);
        LG_COPY_H2D(e->d_raw, e->h_raw, S * 2 * e->raw_stride * sizeof(float), e->stream);
        LG_COPY_H2D(e->d_rsc, e->h_rsc, S * e->chunk_cap * sizeof(LgRsChunk), e->stream);
        LG_COPY_H2D(e->d_rss, e->h_rss, S * sizeof(LgRsStream), e->stream);
#ifndef LG_EMULATE
        cudaEventRecord(e->ev[6], e->stream);
#endif
        LG_LAUNCH(lg_kernel_resample, (int) S * tiles, 256, 0, e->stream, e->dcfg, e->d_raw, (int) e->raw_stride, e->d_rsc, e->chunk_cap, e->d_rss,
                  e->d_pcmf, (int) e->pcm_stride, tiles);
#ifndef LG_EMULATE
        cudaEventRecord(e->ev[7], e->stream);
#endif
        e->launches += 1;
        use_float = 1;
        h2d_pcm = 0;
    }
    else if (use_float && !e->h_pcmf) return -1;
    LG_COPY_H2D(e->d_nfr, e->h_nfr, S * sizeof(int), e->stream);
    if (lg_run_pieces(e, nframes, use_float, h2d_pcm) != 0) return -1;
    LG_COPY_D2H(e->h_fout, e->d_fout, S * F * sizeof(LgFrameOut), e->stream2);
    LG_COPY_D2H(e->h_pay, e->d_pay, S * e->pay_stride, e->stream2);
    LG_COPY_D2H(e->h_hdr, e->d_hdr, S * F * LG_HDR_STRIDE, e->stream2);
    if (lg_engine_sync(e) != 0) return -1;
#ifndef LG_EMULATE
    if (e->hcfg.resample) { float ms = 0; cudaEventElapsedTime(&ms, e->ev[6], e->ev[7]); e->last_ms[5] = ms; }
#endif
    return 0;
}

/* test/debug hook: copy an intermediate device buffer to the host.
 * what: 0 sb, 1 ana, 2 psy, 3 frm, 4 xr, 5 gout, 6 fout, 7 state */
extern "C" long lg_engine_debug_copy(lg_engine *e, int what, void *dst, size_t cap)
{
    size_t const S = e->S, F = e->F;
    const void *src = NULL; size_t n = 0;
    switch (what) {
    case 0: src = e->d_sb; n = S * (2 * F + 1) * 2 * 576 * sizeof(float); break;
    case 1: src = e->d_ana; n = S * 2 * F * sizeof(LgAnalysis); break;
    case 2: src = e->d_psy; n = S * 2 * F * sizeof(LgPsyOut); break;
    case 3: src = e->d_frm; n = S * F * sizeof(LgFrameCtl); break;
    case 4: src = e->d_xr; n = S * 2 * F * 2 * 576 * sizeof(float); break;
    case 5: src = e->d_gout; n = S * 2 * F * 2 * sizeof(LgGranuleOut); break;
    case 6: src = e->d_fout; n = S * F * sizeof(LgFrameOut); break;
    case 7: src = e->d_state; n = S * sizeof(LgStreamState); break;
    default: return -1;
    }
    if (n > cap) n = cap;
#ifdef LG_EMULATE
    memcpy(dst, src, n);
#else
    if (cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
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
