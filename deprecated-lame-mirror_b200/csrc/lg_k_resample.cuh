// lg_k_resample.cuh - kernel R: input-rate conversion in front of kernel A (only when the input rate is not
// the MPEG output rate).
//
// The reference converts on the fly while it fills its frame buffer (fill_buffer_resample util.c:531): one call
// makes up to 1152 output samples from what the caller just handed in, output sample k of the call sits at input
// time k*ratio - itime, and a (filter_l + 1)-tap Blackman-windowed sinc - the nearest of 2*bpc + 1 precomputed
// fractional-offset filters - is applied around it.  The only sequential parts are `itime` and the number of input
// samples each call consumes; the host (lg_api.cpp) replays exactly that bookkeeping per call and hands the device
// a list of *chunks* (one per reference call).  Every output sample is then independent: thread = one sample of the
// launch's PCM window, both channels, so the kernel is a plain 32/33-tap FIR with a per-sample filter choice.
//
//   raw   [S][2][raw_stride] elements of raw_esz bytes: the stream's input samples in the caller's own sample type
//                                  (kinds[stream]; converted and put through s * pcm_transform here, lame.c:1797-1834, or floats
//                                  the host already transformed for a stream that mixed types), element 0 = the oldest sample any
//                                  chunk of this launch can touch (zeros before the stream began)
//   chunk LgRsChunk [S][chunk_cap] chunks overlapping the window, ascending; positions relative to the window/raw
//   out   f32 [S][2][pcm_stride]   the float PCM window kernel A reads (timeline samples, zero where no chunk covers)
//
// Algorithmic HBM bytes per output sample pair: 8 B written, 8*ratio B read (the 33-tap windows overlap in L1).
#pragma once
#include "lg_types.h"

typedef struct {
    int32_t nchunks, win_n;            /* chunks of this stream in this launch; window elements to produce */
} LgRsStream;

__global__ void __launch_bounds__(256)
lg_kernel_resample(const LgDevCfg *__restrict__ cfg, const void *__restrict__ raw, int raw_stride, int raw_esz, const LgPcmKind *__restrict__ kinds,
                   const LgRsChunk *__restrict__ chunks, int chunk_cap, const LgRsStream *__restrict__ rss,
                   float *__restrict__ out, int pcm_stride, int tiles)
{
    int const stream = blockIdx.x / tiles, tile = blockIdx.x % tiles;
    int const w = tile * 256 + threadIdx.x;
    LgRsStream const rs = rss[stream];
    if (w >= rs.win_n) return;
    const LgRsChunk *ck = chunks + (size_t) stream * chunk_cap;
    float *o0 = out + (size_t) stream * 2 * pcm_stride, *o1 = o0 + pcm_stride;
    /* last chunk that starts at or before w */
    int lo = 0, hi = rs.nchunks;
    while (lo < hi) {
        int const mid = (lo + hi) >> 1;
        if (ck[mid].out_pos <= w) lo = mid + 1; else hi = mid;
    }
    float x0 = 0.f, x1 = 0.f;
    if (lo > 0 && w < ck[lo - 1].out_pos + ck[lo - 1].count) {
        LgRsChunk const c = ck[lo - 1];
        int const filter_l = cfg->rs_filter_l, bpc = cfg->rs_bpc;
        int const k = w - c.out_pos;
        /* util.c:591-610, operation by operation: time0 and the difference in double, the fractional offset and the
         * filter choice in float until the final + .5 */
        double const time0 = (double) k * cfg->rs_ratio;
        double const d = time0 - c.itime;
        int const j = (int) floor(d);
        float const offset = (float) (d - ((double) j + .5 * (double) (filter_l % 2)));
        float t = offset * 2.f;
        t = t * (float) bpc;
        t = t + (float) bpc;
        int const joff = (int) floor((double) t + .5);
        const float *f = cfg->rs_filt + joff * LG_RS_TAPS;
        LgPcmKind const kd = kinds[stream];
        float const sc = kd.scale;
        float const n00 = sc * cfg->pcm_transform[0][0], n01 = sc * cfg->pcm_transform[0][1];
        float const n10 = sc * cfg->pcm_transform[1][0], n11 = sc * cfg->pcm_transform[1][1];
        const char *b0 = (const char *) raw + (size_t) stream * 2 * raw_stride * raw_esz, *b1 = b0 + (size_t) raw_stride * raw_esz;
        long const at = (long) c.in_base + j - filter_l / 2;
        for (int i = 0; i <= filter_l; ++i) {
            float const fi = f[i];
            float xl, xr;
            switch (kd.kind) {
            case LG_PCM_S16: xl = (float) ((const int16_t *) b0)[at + i]; xr = (float) ((const int16_t *) b1)[at + i]; break;
            case LG_PCM_S32: xl = (float) ((const int32_t *) b0)[at + i]; xr = (float) ((const int32_t *) b1)[at + i]; break;
            case LG_PCM_S64: xl = (float) ((const long long *) b0)[at + i]; xr = (float) ((const long long *) b1)[at + i]; break;
            case LG_PCM_F64: xl = (float) ((const double *) b0)[at + i]; xr = (float) ((const double *) b1)[at + i]; break;
            default:         xl = ((const float *) b0)[at + i]; xr = ((const float *) b1)[at + i]; break;
            }
            if (kd.kind != LG_PCM_DONE) {
                float const u = xl * n00 + xr * n01, v = xl * n10 + xr * n11;
                xl = u; xr = v;
            }
            x0 = x0 + xl * fi;
            x1 = x1 + xr * fi;
        }
    }
    o0[w] = x0;
    o1[w] = x1;
}
