/* oracle/port/port_mdct.c - TEST INFRASTRUCTURE (see lame_port.h).
 * Restates newmdct.c: the 512-tap polyphase analysis filterbank with its fast 32-point cosine
 * transform (window_subband, newmdct.c:430), the 36->18 and 3x(12->6) MDCTs (mdct_long :869,
 * mdct_short :832) and the framing around them (mdct_sub48 :944).  The butterfly network is the
 * reference's factorisation: every output must see the same float operations in the same order, so
 * the data flow is reproduced exactly; sqrt(2) factors are double-precision products rounded once
 * to float, as in the reference. */
#include <string.h>
#include "lame_port.h"

#define SQRT2_D 1.41421356237309504880
#define NS 12
#define NL 36

/* t = a[p]-a[q]; sum goes to q, scaled difference to p */
#define BF_A(p, q, c) { float t_ = a[p] - a[q]; a[q] += a[p]; a[p] = t_ * (c); }
/* t = a[p]-a[q]; sum goes to p, scaled difference to q */
#define BF_B(p, q, c) { float t_ = a[p] - a[q]; a[p] += a[q]; a[q] = t_ * (c); }
/* t = a[p]; a[p] = a[q]-t; a[q] += t */
#define XADD(p, q) { float t_ = a[p]; a[p] = a[q] - t_; a[q] = a[q] + t_; }
/* t = a[p]; a[p] += a[q]; a[q] -= t */
#define XSUB(p, q) { float t_ = a[p]; a[p] += a[q]; a[q] -= t_; }

/* newmdct.c:430 window_subband: 32 subband samples from the 512 PCM samples around x1 */
static void window_subband(const float *x1, float a[32])
{
    const float *wp = lp_tab_enwindow() + 10;
    const float *x2 = &x1[238 - 14 - 286];
    int i;
    for (i = -15; i < 0; i++) {
        float w, s, t;
        w = wp[-10]; s = x2[-224] * w; t = x1[224] * w;
        w = wp[-9]; s += x2[-160] * w; t += x1[160] * w;
        w = wp[-8]; s += x2[-96] * w; t += x1[96] * w;
        w = wp[-7]; s += x2[-32] * w; t += x1[32] * w;
        w = wp[-6]; s += x2[32] * w; t += x1[-32] * w;
        w = wp[-5]; s += x2[96] * w; t += x1[-96] * w;
        w = wp[-4]; s += x2[160] * w; t += x1[-160] * w;
        w = wp[-3]; s += x2[224] * w; t += x1[-224] * w;
        w = wp[-2]; s += x1[-256] * w; t -= x2[256] * w;
        w = wp[-1]; s += x1[-192] * w; t -= x2[192] * w;
        w = wp[0]; s += x1[-128] * w; t -= x2[128] * w;
        w = wp[1]; s += x1[-64] * w; t -= x2[64] * w;
        w = wp[2]; s += x1[0] * w; t -= x2[0] * w;
        w = wp[3]; s += x1[64] * w; t -= x2[-64] * w;
        w = wp[4]; s += x1[128] * w; t -= x2[-128] * w;
        w = wp[5]; s += x1[192] * w; t -= x2[-192] * w;
        s *= wp[6];
        w = t - s;
        a[30 + i * 2] = t + s;
        a[31 + i * 2] = wp[7] * w;
        wp += 18;
        x1--;
        x2++;
    }
    {
        float s, t, u, v;
        t = x1[-16] * wp[-10]; s = x1[-32] * wp[-2];
        t += (x1[-48] - x1[16]) * wp[-9]; s += x1[-96] * wp[-1];
        t += (x1[-80] + x1[48]) * wp[-8]; s += x1[-160] * wp[0];
        t += (x1[-112] - x1[80]) * wp[-7]; s += x1[-224] * wp[1];
        t += (x1[-144] + x1[112]) * wp[-6]; s -= x1[32] * wp[2];
        t += (x1[-176] - x1[144]) * wp[-5]; s -= x1[96] * wp[3];
        t += (x1[-208] + x1[176]) * wp[-4]; s -= x1[160] * wp[4];
        t += (x1[-240] - x1[208]) * wp[-3]; s -= x1[224];
        u = s - t; v = s + t;
        t = a[14]; s = a[15] - t;
        a[31] = v + t; a[30] = u + s; a[15] = u - s; a[14] = v - t;
    }
    {
        float const w2 = wp[-2 * 18 + 7], w4 = wp[-4 * 18 + 7], w6 = wp[-6 * 18 + 7];
        float const w10 = wp[-10 * 18 + 7], w12 = wp[-12 * 18 + 7], w14 = wp[-14 * 18 + 7];
        float xr;
        static const uint8_t chain_a[6] = { 9, 25, 5, 21, 13, 29 };
        static const uint8_t chain_b[14] = { 16, 17, 8, 9, 24, 25, 4, 5, 20, 21, 12, 13, 28, 29 };
        static const uint8_t fin[16][2] = { {0, 31}, {1, 30}, {16, 15}, {17, 14}, {8, 23}, {9, 22}, {24, 7}, {25, 6},
            {4, 27}, {5, 26}, {20, 11}, {21, 10}, {12, 19}, {13, 18}, {28, 3}, {29, 2} };
        int k;
        BF_A(28, 0, w2) BF_A(29, 1, w2) BF_A(26, 2, w4) BF_A(27, 3, w4) BF_A(24, 4, w6) BF_A(25, 5, w6)
        BF_A(22, 6, SQRT2_D)
        xr = a[23] - a[7]; a[7] += a[23]; a[23] = xr * SQRT2_D - a[7];     /* one double expression */
        a[7] -= a[6]; a[22] -= a[7]; a[23] -= a[22];
        XADD(6, 31) XADD(7, 30) XADD(22, 15) XADD(23, 14)
        BF_A(20, 8, w10) BF_A(21, 9, w10) BF_A(18, 10, w12) BF_A(19, 11, w12) BF_A(16, 12, w14) BF_A(17, 13, w14)
        BF_A(24, 20, w12) BF_A(25, 21, w12) BF_B(4, 8, w12) BF_B(5, 9, w12)
        BF_B(0, 12, w4) BF_B(1, 13, w4) BF_B(16, 28, w4) BF_A(29, 17, w4)
        BF_B(2, 10, SQRT2_D) BF_B(3, 11, SQRT2_D)
        xr = SQRT2_D * (-a[18] + a[26]); a[18] += a[26]; a[26] = xr - a[18];
        xr = SQRT2_D * (-a[19] + a[27]); a[19] += a[27]; a[27] = xr - a[19];
        xr = a[2]; a[19] -= a[3]; a[3] -= xr; a[2] = a[31] - xr; a[31] += xr;
        xr = a[3]; a[11] -= a[19]; a[18] -= xr; a[3] = a[30] - xr; a[30] += xr;
        xr = a[18]; a[27] -= a[11]; a[19] -= xr; a[18] = a[15] - xr; a[15] += xr;
        xr = a[19]; a[10] -= xr; a[19] = a[14] - xr; a[14] += xr;
        xr = a[10]; a[11] -= xr; a[10] = a[23] - xr; a[23] += xr;
        xr = a[11]; a[26] -= xr; a[11] = a[22] - xr; a[22] += xr;
        xr = a[26]; a[27] -= xr; a[26] = a[7] - xr; a[7] += xr;
        xr = a[27]; a[27] = a[6] - xr; a[6] += xr;
        BF_B(0, 4, SQRT2_D) BF_B(1, 5, SQRT2_D) BF_B(16, 20, SQRT2_D) BF_B(17, 21, SQRT2_D)
        xr = -SQRT2_D * (a[8] - a[12]); a[8] += a[12]; a[12] = xr - a[8];
        xr = -SQRT2_D * (a[9] - a[13]); a[9] += a[13]; a[13] = xr - a[9];
        xr = -SQRT2_D * (a[25] - a[29]); a[25] += a[29]; a[29] = xr - a[25];
        xr = -SQRT2_D * (a[24] + a[28]); a[24] -= a[28]; a[28] = xr - a[24];
        xr = a[24] - a[16]; a[24] = xr; xr = a[20] - xr; a[20] = xr; xr = a[28] - xr; a[28] = xr;
        xr = a[25] - a[17]; a[25] = xr; xr = a[21] - xr; a[21] = xr; xr = a[29] - xr; a[29] = xr;
        xr = a[17] - a[1]; a[17] = xr;
        for (k = 0; k < 6; k++) { xr = a[chain_a[k]] - xr; a[chain_a[k]] = xr; }
        xr = a[1] - a[0]; a[1] = xr;
        for (k = 0; k < 14; k++) { xr = a[chain_b[k]] - xr; a[chain_b[k]] = xr; }
        for (k = 0; k < 16; k++) XSUB(fin[k][0], fin[k][1])
    }
}

/* newmdct.c:832 mdct_short: three interleaved 12->6 transforms; literal scale factors are doubles */
static void mdct_short(float *inout)
{
    const float *win_s = lp_tab_mdctwin() + 2 * 36;
    int l;
    for (l = 0; l < 3; l++) {
        float tc0, tc1, tc2, ts0, ts1, ts2;
        ts0 = inout[2 * 3] * win_s[0] - inout[5 * 3];
        tc0 = inout[0 * 3] * win_s[2] - inout[3 * 3];
        tc1 = ts0 + tc0;
        tc2 = ts0 - tc0;
        ts0 = inout[5 * 3] * win_s[0] + inout[2 * 3];
        tc0 = inout[3 * 3] * win_s[2] + inout[0 * 3];
        ts1 = ts0 + tc0;
        ts2 = -ts0 + tc0;
        tc0 = (inout[1 * 3] * win_s[1] - inout[4 * 3]) * 2.069978111953089e-11;
        ts0 = (inout[4 * 3] * win_s[1] + inout[1 * 3]) * 2.069978111953089e-11;
        inout[3 * 0] = tc1 * 1.907525191737280e-11 + tc0;
        inout[3 * 5] = -ts1 * 1.907525191737280e-11 + ts0;
        tc2 = tc2 * 0.86602540378443870761 * 1.907525191737281e-11;
        ts1 = ts1 * 0.5 * 1.907525191737281e-11 + ts0;
        inout[3 * 1] = tc2 - ts1;
        inout[3 * 2] = tc2 + ts1;
        tc1 = tc1 * 0.5 * 1.907525191737281e-11 - tc0;
        ts2 = ts2 * 0.86602540378443870761 * 1.907525191737281e-11;
        inout[3 * 3] = tc1 + ts2;
        inout[3 * 4] = tc1 - ts2;
        inout++;
    }
}

/* newmdct.c:869 mdct_long: 18 pre-rotated inputs -> 18 coefficients (all float) */
static void mdct_long(float *out, const float *in)
{
    const float *cx = lp_tab_mdctwin() + 2 * 36 + 12;
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

/* newmdct.c:944 mdct_sub48 */
void lp_mdct_sub48(lp_encoder *e, const float *w0, const float *w1)
{
    static const uint8_t order[32] = { 0, 1, 16, 17, 8, 9, 24, 25, 4, 5, 20, 21, 12, 13, 28, 29,
        2, 3, 18, 19, 10, 11, 26, 27, 6, 7, 22, 23, 14, 15, 30, 31 };
    const float (*win)[NL] = (const float (*)[NL]) lp_tab_mdctwin();
    const float *tantab_l = win[LP_SHORT] + 3, *ca = win[LP_SHORT] + 20, *cs = win[LP_SHORT] + 28;
    const lp_config *cfg = &e->cfg;
    int gr, k, ch;
    const float *wk = w0 + 286;
    for (ch = 0; ch < cfg->channels; ch++) {
        for (gr = 0; gr < cfg->mode_gr; gr++) {
            int band;
            lp_granule *const gi = &e->tt[gr][ch];
            float *mdct_enc = gi->xr;
            float *samp = e->sb_sample[ch][1 - gr][0];
            for (k = 0; k < 18 / 2; k++) {
                window_subband(wk, samp);
                window_subband(wk + 32, samp + 32);
                samp += 64;
                wk += 64;
                for (band = 1; band < 32; band += 2) samp[band - 32] *= -1;
            }
            for (band = 0; band < 32; band++, mdct_enc += 18) {
                int type = gi->block_type;
                float const *const band0 = e->sb_sample[ch][gr][0] + order[band];
                float *const band1 = e->sb_sample[ch][1 - gr][0] + order[band];
                if (gi->mixed_block_flag && band < 2) type = 0;
                if (cfg->amp_filter[band] < 1e-12) memset(mdct_enc, 0, 18 * sizeof(float));
                else {
                    if (cfg->amp_filter[band] < 1.0)
                        for (k = 0; k < 18; k++) band1[k * 32] *= cfg->amp_filter[band];
                    if (type == LP_SHORT) {
                        for (k = -NS / 4; k < 0; k++) {
                            float const w = win[LP_SHORT][k + 3];
                            mdct_enc[k * 3 + 9] = band0[(9 + k) * 32] * w - band0[(8 - k) * 32];
                            mdct_enc[k * 3 + 18] = band0[(14 - k) * 32] * w + band0[(15 + k) * 32];
                            mdct_enc[k * 3 + 10] = band0[(15 + k) * 32] * w - band0[(14 - k) * 32];
                            mdct_enc[k * 3 + 19] = band1[(2 - k) * 32] * w + band1[(3 + k) * 32];
                            mdct_enc[k * 3 + 11] = band1[(3 + k) * 32] * w - band1[(2 - k) * 32];
                            mdct_enc[k * 3 + 20] = band1[(8 - k) * 32] * w + band1[(9 + k) * 32];
                        }
                        mdct_short(mdct_enc);
                    }
                    else {
                        float work[18];
                        for (k = -NL / 4; k < 0; k++) {
                            float a, b;
                            a = win[type][k + 27] * band1[(k + 9) * 32] + win[type][k + 36] * band1[(8 - k) * 32];
                            b = win[type][k + 9] * band0[(k + 9) * 32] - win[type][k + 18] * band0[(8 - k) * 32];
                            work[k + 9] = a - b * tantab_l[k + 9];
                            work[k + 18] = a * tantab_l[k + 9] + b;
                        }
                        mdct_long(mdct_enc, work);
                    }
                }
                if (type != LP_SHORT && band != 0) {
                    for (k = 7; k >= 0; --k) {
                        float bu, bd;
                        bu = mdct_enc[k] * ca[k] + mdct_enc[-1 - k] * cs[k];
                        bd = mdct_enc[k] * cs[k] - mdct_enc[-1 - k] * ca[k];
                        mdct_enc[-1 - k] = bu;
                        mdct_enc[k] = bd;
                    }
                }
            }
        }
        wk = w1 + 286;
        if (cfg->mode_gr == 1) memcpy(e->sb_sample[ch][0], e->sb_sample[ch][1], 576 * sizeof(float));    /* newmdct.c:1037 */
    }
}
