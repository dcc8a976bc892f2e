// lg_k_mdct.cuh - kernel C: windowed 36/12-point MDCT + alias-reduction butterflies + the two
// rearrangements the quantiser expects (L/R -> M/S, short-block reordering).
//
// One CTA (2 warps = 2 channels) per (stream, granule); lane = polyphase subband (32 bands x 18 lines).
// Inputs are the subband samples of this and the previous granule (kernel A) and the block type /
// mode_ext decided by the scan (kernel B).  Reference: mdct_sub48 newmdct.c:944 (band loop), mdct_long
// :869, mdct_short :832, ms_convert quantize.c:48, init_outer_loop quantize.c:298-317 (reorder).
//
// Algorithmic HBM bytes per gr.ch: read 2 x 2304 B subband samples, write 2304 B of MDCT lines.
#pragma once
#include "lg_math.cuh"

struct LgSmemC { float xr[2][576]; float tmp[2][576]; };

/* newmdct.c:869 mdct_long */
__device__ __forceinline__ void lg_mdct_long(float *out, const float *in, const float *__restrict__ cx)
{
    float ct, st;
    {
        float tc1, tc2, tc3, tc4, ts5, ts6, ts7, ts8;
        tc1 = in[17] - in[9]; tc3 = in[15] - in[11]; tc4 = in[14] - in[12];
        ts5 = in[0] + in[8]; ts6 = in[1] + in[7]; ts7 = in[2] + in[6]; ts8 = in[3] + in[5];
        out[17] = (ts5 + ts7 - ts8) - (ts6 - in[4]);
        st = (ts5 + ts7 - ts8) * cx[7] + (ts6 - in[4]);
        ct = (tc1 - tc3 - tc4) * cx[6];
        out[5] = ct + st; out[6] = ct - st;
        tc2 = (in[16] - in[10]) * cx[6];
        ts6 = ts6 * cx[7] + in[4];
        ct = tc1 * cx[0] + tc2 + tc3 * cx[1] + tc4 * cx[2];
        st = -ts5 * cx[4] + ts6 - ts7 * cx[5] + ts8 * cx[3];
        out[1] = ct + st; out[2] = ct - st;
        ct = tc1 * cx[1] - tc2 - tc3 * cx[2] + tc4 * cx[0];
        st = -ts5 * cx[5] + ts6 - ts7 * cx[3] + ts8 * cx[4];
        out[9] = ct + st; out[10] = ct - st;
        ct = tc1 * cx[2] - tc2 + tc3 * cx[0] - tc4 * cx[1];
        st = ts5 * cx[3] - ts6 + ts7 * cx[4] - ts8 * cx[5];
        out[13] = ct + st; out[14] = ct - st;
    }
    {
        float ts1, ts2, ts3, ts4, tc5, tc6, tc7, tc8;
        ts1 = in[8] - in[0]; ts3 = in[6] - in[2]; ts4 = in[5] - in[3];
        tc5 = in[17] + in[9]; tc6 = in[16] + in[10]; tc7 = in[15] + in[11]; tc8 = in[14] + in[12];
        out[0] = (tc5 + tc7 + tc8) + (tc6 + in[13]);
        ct = (tc5 + tc7 + tc8) * cx[7] - (tc6 + in[13]);
        st = (ts1 - ts3 + ts4) * cx[6];
        out[11] = ct + st; out[12] = ct - st;
        ts2 = (in[7] - in[1]) * cx[6];
        tc6 = in[13] - tc6 * cx[7];
        ct = tc5 * cx[3] - tc6 + tc7 * cx[4] + tc8 * cx[5];
        st = ts1 * cx[2] + ts2 + ts3 * cx[0] + ts4 * cx[1];
        out[3] = ct + st; out[4] = ct - st;
        ct = -tc5 * cx[5] + tc6 - tc7 * cx[3] - tc8 * cx[4];
        st = ts1 * cx[1] + ts2 - ts3 * cx[2] - ts4 * cx[0];
        out[7] = ct + st; out[8] = ct - st;
        ct = -tc5 * cx[4] + tc6 - tc7 * cx[5] - tc8 * cx[3];
        st = ts1 * cx[0] - ts2 + ts3 * cx[1] - ts4 * cx[2];
        out[15] = ct + st; out[16] = ct - st;
    }
}

/* newmdct.c:832 mdct_short on 18 interleaved values; the literal scale factors are doubles */
__device__ __forceinline__ void lg_mdct_short(float *io, const float *__restrict__ win_s)
{
#pragma unroll
    for (int l = 0; l < 3; l++) {
        float *inout = io + l;
        float tc0, tc1, tc2, ts0, ts1, ts2;
        ts0 = inout[2 * 3] * win_s[0] - inout[5 * 3];
        tc0 = inout[0 * 3] * win_s[2] - inout[3 * 3];
        tc1 = ts0 + tc0;
        tc2 = ts0 - tc0;
        ts0 = inout[5 * 3] * win_s[0] + inout[2 * 3];
        tc0 = inout[3 * 3] * win_s[2] + inout[0 * 3];
        ts1 = ts0 + tc0;
        ts2 = -ts0 + tc0;
        tc0 = (float) ((inout[1 * 3] * win_s[1] - inout[4 * 3]) * 2.069978111953089e-11);
        ts0 = (float) ((inout[4 * 3] * win_s[1] + inout[1 * 3]) * 2.069978111953089e-11);
        inout[3 * 0] = (float) (tc1 * 1.907525191737280e-11 + tc0);
        inout[3 * 5] = (float) (-ts1 * 1.907525191737280e-11 + ts0);
        tc2 = (float) (tc2 * 0.86602540378443870761 * 1.907525191737281e-11);
        ts1 = (float) (ts1 * 0.5 * 1.907525191737281e-11 + ts0);
        inout[3 * 1] = tc2 - ts1;
        inout[3 * 2] = tc2 + ts1;
        tc1 = (float) (tc1 * 0.5 * 1.907525191737281e-11 - tc0);
        ts2 = (float) (ts2 * 0.86602540378443870761 * 1.907525191737281e-11);
        inout[3 * 3] = tc1 + ts2;
        inout[3 * 4] = tc1 - ts2;
    }
}

__global__ void __launch_bounds__(64)
lg_kernel_mdct(const LgDevCfg *__restrict__ cfg, const float *__restrict__ sb, const LgPsyOut *__restrict__ psy,
               const LgFrameCtl *__restrict__ frm, float *__restrict__ xr_out,
               const int *__restrict__ nfr, int nframes, int g0, int cnt /* this launch: granules g0 .. g0+cnt-1 */)
{
    LG_DYN_SMEM(LgSmemC, sm);
    int const lane = threadIdx.x & 31, ch = threadIdx.x >> 5;
    int const ngr = 2 * nframes;
    int const stream = blockIdx.x / cnt, gb = g0 + blockIdx.x % cnt;
    int const nch = cfg->channels;
    if (gb >= cfg->mode_gr * nfr[stream]) return;
    const LgPsyOut *P = psy + (size_t) stream * ngr + gb;
    const LgFrameCtl *F = frm + (size_t) stream * nframes + (gb / cfg->mode_gr);
    int const type = (ch < nch) ? P->block_type[ch] : LG_NORM;
    const float *win = cfg->mdctwin;
    const float *tantab_l = win + 2 * 36 + 3, *cx = win + 2 * 36 + 12, *ca = win + 2 * 36 + 20, *cs = win + 2 * 36 + 28;

    if (ch < nch) {
        int const band = lane;
        /* subband samples: slot gb is the previous granule, slot gb+1 the current one */
        const float *prev = sb + (((size_t) stream * (ngr + 1) + gb) * 2 + ch) * 576;
        const float *cur = prev + 2 * 576;
        int const ord = (band & 1) | ((band & 2) << 3) | ((band & 4) << 1) | ((band & 8) >> 1) | ((band & 16) >> 3);
        /* newmdct.c:418 order[]: bit0 stays, bits 1..4 are reversed */
        const float *band0 = prev + ord, *band1 = cur + ord;
        float enc[18];
        if (__ldg(&cfg->amp_filter[band]) < 1e-12) {
#pragma unroll
            for (int k = 0; k < 18; k++) enc[k] = 0.f;
        }
        else if (type == LG_SHORT) {
#pragma unroll
            for (int k = -3; k < 0; k++) {
                float const w = win[2 * 36 + k + 3];
                enc[k * 3 + 9] = band0[(9 + k) * 32] * w - band0[(8 - k) * 32];
                enc[k * 3 + 18] = band0[(14 - k) * 32] * w + band0[(15 + k) * 32];
                enc[k * 3 + 10] = band0[(15 + k) * 32] * w - band0[(14 - k) * 32];
                enc[k * 3 + 19] = band1[(2 - k) * 32] * w + band1[(3 + k) * 32];
                enc[k * 3 + 11] = band1[(3 + k) * 32] * w - band1[(2 - k) * 32];
                enc[k * 3 + 20] = band1[(8 - k) * 32] * w + band1[(9 + k) * 32];
            }
            lg_mdct_short(enc, win + 2 * 36);
        }
        else {
            float work[18];
            const float *wt = win + type * 36;
#pragma unroll
            for (int k = -9; k < 0; k++) {
                float a, b;
                a = wt[k + 27] * band1[(k + 9) * 32] + wt[k + 36] * band1[(8 - k) * 32];
                b = wt[k + 9] * band0[(k + 9) * 32] - wt[k + 18] * band0[(8 - k) * 32];
                work[k + 9] = a - b * tantab_l[k + 9];
                work[k + 18] = a * tantab_l[k + 9] + b;
            }
            lg_mdct_long(enc, work, cx);
        }
#pragma unroll
        for (int k = 0; k < 18; k++) sm->xr[ch][band * 18 + k] = enc[k];
        __syncwarp();
        /* alias reduction between band-1 and band (newmdct.c:1022-1031) */
        if (type != LG_SHORT && band != 0) {
            float *e = &sm->xr[ch][band * 18];
#pragma unroll
            for (int k = 7; k >= 0; --k) {
                float const bu = e[k] * ca[k] + e[-1 - k] * cs[k];
                float const bd = e[k] * cs[k] - e[-1 - k] * ca[k];
                e[-1 - k] = bu;
                e[k] = bd;
            }
        }
    }
    __syncthreads();
    /* quantize.c:48 ms_convert */
    if (F->mode_ext == 2 && nch == 2) {
        float const c = (float) (LG_SQRT2_D * 0.5);
        for (int i = threadIdx.x; i < 576; i += 64) {
            float const l = sm->xr[0][i], r = sm->xr[1][i];
            sm->xr[0][i] = (l + r) * c;
            sm->xr[1][i] = (l - r) * c;
        }
    }
    __syncthreads();
    if (ch < nch) {
        float *o = xr_out + (((size_t) stream * ngr + gb) * 2 + ch) * 576;
        if (type == LG_SHORT) {
            /* quantize.c:298-317: within each short sfb, window-major order */
            for (int sfb = 0; sfb < LG_SBMAX_S; sfb++) {
                int const start = cfg->sfb_s[sfb], end = cfg->sfb_s[sfb + 1], w = end - start;
                for (int i = lane; i < 3 * w; i += 32) {
                    int const window = i / w, l = start + i % w;
                    o[3 * start + i] = sm->xr[ch][3 * l + window];
                }
            }
        }
        else for (int i = lane; i < 576; i += 32) o[i] = sm->xr[ch][i];
    }
}
