// lg_k_vbr.cuh - kernel D for VBR-new (vbr_mtrh): VBR_new_iteration_loop quantize.c:1645 + vbrquantize.c.
//
// Unlike the CBR/ABR search (lg_k_quant.cuh), VBR-new has no noise-shaping iteration: for every scalefactor band it
// searches the largest step whose quantisation noise stays below the band's allowed noise (find_scalefac_x34, eight
// bisection steps), then fits global_gain / scalefactors / subblock gains to those steps, quantises once and counts
// bits; only a frame that does not fit its bit budget is re-quantised with flattened steps (outOfBitsStrategy).  The
// four granule.channels of a frame are independent until the frame's bit budget is checked, so one CTA per stream runs
// them as FOUR warps side by side (warp = gr*2 + ch) and walks the stream's frames in order (the reservoir couples
// consecutive frames).
//
// Inside a warp: the band search runs all bands in lockstep - every lane takes quadruples of lines (the reference's
// calc_sfb_noise_x34 adds the squared errors in groups of four, in double, vbrquantize.c:218-270) of whatever band they
// belong to and writes the group's partial sum; one lane per band then adds its groups in the reference's order.  The
// band-level fitting (short/long_block_constrain) is one or two bands per lane with integer warp reductions.
// Quantisation and bit counting reuse the line-parallel code of the CBR kernel.
//
// Algorithmic HBM bytes per granule.channel: as kernel D (read 2304 B lines + 488 B ratios, write sizeof(LgGranuleOut)).
#pragma once
#include "lg_k_quant.cuh"

#define LG_VGROUPS 192                   /* quadruples per granule.channel: 144 full ones + at most one remainder per band */

struct __attribute__((aligned(16))) LgVWarp {
    LgQWarp q;
    double gsum[3][LG_VGROUPS];          /* group partial sums for step sf, sf + 1, sf - 1 */
    int    gband[LG_VGROUPS], gstart[LG_VGROUPS], gcnt[LG_VGROUPS];
    int    bgrp0[41];                    /* first group of every band */
    int    vbrsf[40], vbrsfmin[40], sftmp[40], wrk[40];
    int    bsf[40];                      /* step the band wants evaluated in this bisection round, or -1 */
    int    blen[40];                     /* lines of the band up to max_nonzero_coeff */
    float  bfac[40];
    int    mingain_l, mingain_s[3], ngroups, nb_active;
};
struct LgSmemV {
    LgVWarp w[4];
    int use_bits[4], max_bits[4], ath_over[4], nonzero[4];
    int sf_gr0[2][40];
    int bt_gr0[2];
    int scfsi[2][4];
};

struct LgVCtx { int mingain_l, mingain_s[3], is_short, guess_only; };

__constant__ uint8_t LG_VRANGE_SHORT[40] = { 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 0, 0, 0, 0 };
/* vbrquantize.c:579 max_range_long_lsf_pretab: MPEG-2/2.5 with preflag (scalefactor partition table 2) */
__constant__ uint8_t LG_VRANGE_LONG_LSF[40] = { 7, 7, 7, 7, 7, 7, 3, 3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
__constant__ uint8_t LG_VRANGE_LONG[40] = { 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };

/* vbrquantize.c:170 k_34_4 for one value */
__device__ __forceinline__ int lg_vquant1(const float *__restrict__ adj, float xs)
{
    double const d = (double) xs + 8388608.0;
    int const idx = __float_as_int((float) d) - 0x4b000000;
    return __float_as_int((float) (d + (double) __ldg(&adj[idx]))) - 0x4b000000;
}

/* squared-error sum of one quadruple at step sf (the body of calc_sfb_noise_x34's loop) */
__device__ __forceinline__ double lg_vgroup_noise(const LgDevCfg *__restrict__ c, const float *xr, const float *xr34, int cnt, int sf)
{
    float const sfpow = __ldg(&c->pow20[sf + LG_QMAX2]), sfpow34 = __ldg(&c->ipow20[sf]);
    double x[4] = { 0, 0, 0, 0 };
    for (int k = 0; k < 4; k++)
        if (k < cnt) {
            int const l3 = lg_vquant1(c->adj43asm, sfpow34 * xr34[k]);
            x[k] = (double) (fabsf(xr[k]) - sfpow * __ldg(&c->pow43[l3]));
        }
    return (x[0] * x[0] + x[1] * x[1]) + (x[2] * x[2] + x[3] * x[3]);
}

/* vbrquantize.c:148 find_lowest_scalefac */
__device__ __forceinline__ int lg_vlowest_sf(const LgDevCfg *__restrict__ c, float xr34)
{
    int sf_ok = 255, sf = 128, delsf = 64;
    float const ixmax_val = LG_IXMAX;
    for (int i = 0; i < 8; ++i) {
        float const xfsf = __ldg(&c->ipow20[sf]) * xr34;
        if (xfsf <= ixmax_val) { sf_ok = sf; sf -= delsf; }
        else sf += delsf;
        delsf >>= 1;
    }
    return sf_ok;
}

/* vbrquantize.c:395 block_sf.  Returns vbrmax; fills v->vbrsf / v->vbrsfmin and ctx.mingain_*. */
__device__ __noinline__ int lg_vblock_sf(const LgDevCfg *__restrict__ c, LgVWarp *v, const LgQConst &qc, LgVCtx &ctx, int lane)
{
    LgQWarp *w = &v->q;
    int const mnz = qc.max_nonzero_coeff;
    int const nbands = (qc.block_type == LG_SHORT) ? 39 : 22;
    /* bands that start at or below max_nonzero_coeff, their useful length, their groups */
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        if (sfb < 40) {
            int len = 0;
            if (sfb < nbands && w->lstart[sfb] <= mnz) {
                int const m = mnz - w->lstart[sfb] + 1;
                len = w->width[sfb] < m ? w->width[sfb] : m;
            }
            v->blen[sfb] = len;
        }
    }
    __syncwarp();
    if (lane == 0) {
        int g = 0, nact = 0;
        for (int sfb = 0; sfb < nbands; sfb++) {
            v->bgrp0[sfb] = g;
            int const len = v->blen[sfb];
            if (len > 0) nact = sfb + 1;
            for (int k = 0; k < len; k += 4) { v->gband[g] = sfb; v->gstart[g] = w->lstart[sfb] + k; v->gcnt[g] = len - k < 4 ? len - k : 4; g++; }
        }
        v->bgrp0[nbands] = g;
        v->ngroups = g;
        v->nb_active = nact;
    }
    __syncwarp();
    int const nact = v->nb_active;
    /* per band: largest xr34, smallest usable step */
    int mgl = 0, mg0 = 0, mg1 = 0, mg2 = 0;
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        if (sfb < nact) {
            float mx = 0.f;
            int const j0 = w->lstart[sfb], len = v->blen[sfb];
            for (int k = 0; k < len; k++) { float const t = w->xrpow[j0 + k]; if (mx < t) mx = t; }
            int const m1 = lg_vlowest_sf(c, mx);
            v->vbrsfmin[sfb] = m1;
            mgl = max(mgl, m1);
            int const i3 = sfb % 3;
            if (i3 == 0) mg0 = max(mg0, m1); else if (i3 == 1) mg1 = max(mg1, m1); else mg2 = max(mg2, m1);
        }
    }
    ctx.mingain_l = lg_wmax_i(mgl);
    ctx.mingain_s[0] = lg_wmax_i(mg0); ctx.mingain_s[1] = lg_wmax_i(mg1); ctx.mingain_s[2] = lg_wmax_i(mg2);
    __syncwarp();
    /* find_scalefac_x34 for every searched band, all bands in lockstep (vbrquantize.c:347) */
    int sfc[2] = { 128, 128 }, sfok[2] = { 255, 255 }, seen[2] = { 0, 0 }, srch[2] = { 0, 0 };
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        srch[r] = (sfb < nact && sfb < qc.psymax && w->width[sfb] > 2 && w->eac[sfb]);
    }
    if (ctx.guess_only) {
        /* vbrquantize.c:324 guess_scalefac_x34 (quality 7) */
        for (int r = 0; r < 2; r++) {
            int const sfb = lane + 32 * r;
            if (srch[r]) {
                float const cc = 5.799142446f;
                int const guess = 210 + (int) (cc * lg_log10f(w->l3_xmin[sfb] / v->blen[sfb]) - .5f);
                int const sf_min = v->vbrsfmin[sfb];
                sfc[r] = guess < sf_min ? sf_min : (guess >= 255 ? 255 : guess);
            }
        }
    }
    else {
        int delsf = 128;
        for (int it = 0; it < 8; ++it) {
            delsf >>= 1;
            int want[2];
            for (int r = 0; r < 2; r++) {
                int const sfb = lane + 32 * r;
                want[r] = -1;
                if (srch[r]) {
                    if (sfc[r] <= v->vbrsfmin[sfb]) sfc[r] += delsf;
                    else want[r] = sfc[r];
                }
                if (sfb < 40) v->bsf[sfb] = want[r];
            }
            __syncwarp();
            for (int g = lane; g < v->ngroups; g += 32) {
                int const sf = v->bsf[v->gband[g]];
                if (sf >= 0) {
                    const float *xr = &w->xr[v->gstart[g]], *x34 = &w->xrpow[v->gstart[g]];
                    int const cnt = v->gcnt[g];
                    v->gsum[0][g] = lg_vgroup_noise(c, xr, x34, cnt, sf);
                    if (sf < 255) v->gsum[1][g] = lg_vgroup_noise(c, xr, x34, cnt, sf + 1);
                    if (sf > 0) v->gsum[2][g] = lg_vgroup_noise(c, xr, x34, cnt, sf - 1);
                }
            }
            __syncwarp();
            for (int r = 0; r < 2; r++) {
                int const sfb = lane + 32 * r;
                if (want[r] >= 0) {
                    int const sf = want[r], g0 = v->bgrp0[sfb], g1 = v->bgrp0[sfb + 1];
                    float const xmin = w->l3_xmin[sfb];
                    /* tri_calc_sfb_noise_x34: too noisy at sf, sf + 1 or sf - 1 (xfsf is a float that takes double sums) */
                    int bad = 0;
                    for (int k = 0; k < 3 && !bad; k++) {
                        if ((k == 1 && sf >= 255) || (k == 2 && sf <= 0)) continue;
                        float xfsf = 0;
                        for (int g = g0; g < g1; g++) xfsf = (float) ((double) xfsf + v->gsum[k][g]);
                        if (xmin < xfsf) bad = 1;
                    }
                    if (bad) sfc[r] -= delsf;
                    else { sfok[r] = sf; sfc[r] += delsf; seen[r] = 1; }
                }
            }
            __syncwarp();
        }
        for (int r = 0; r < 2; r++) {
            int const sfb = lane + 32 * r;
            if (srch[r]) {
                if (seen[r] > 0) sfc[r] = sfok[r];
                if (sfc[r] <= v->vbrsfmin[sfb]) sfc[r] = v->vbrsfmin[sfb];
            }
        }
    }
    for (int r = 0; r < 2; r++) { int const sfb = lane + 32 * r; if (sfb < 40) v->bsf[sfb] = srch[r] ? sfc[r] : -1; }
    __syncwarp();
    /* the running maximum of block_sf's band loop is order dependent: one lane replays it (<= 39 bands) */
    int maxsf = 0;
    if (lane == 0) {
        int m_o = -1;
        int sfb = 0;
        for (; sfb < nact; sfb++) {
            int const m1 = v->vbrsfmin[sfb];
            int m2;
            if (sfb < qc.psymax && w->width[sfb] > 2) {
                if (w->eac[sfb]) {
                    m2 = v->bsf[sfb];
                    if (maxsf < m2) maxsf = m2;
                    if (m_o < m2 && m2 < 255) m_o = m2;
                }
                else { m2 = 255; maxsf = 255; }
            }
            else {
                if (maxsf < m1) maxsf = m1;
                m2 = maxsf;
            }
            v->vbrsf[sfb] = m2;
        }
        for (; sfb < LG_SFBMAX; ++sfb) { v->vbrsf[sfb] = maxsf; v->vbrsfmin[sfb] = 0; }
        if (m_o > -1) {
            maxsf = m_o;
            for (sfb = 0; sfb < LG_SFBMAX; ++sfb) if (v->vbrsf[sfb] == 255) v->vbrsf[sfb] = m_o;
        }
    }
    maxsf = __shfl_sync(LG_FULL, maxsf, 0);
    __syncwarp();
    return maxsf;
}

/* vbrquantize.c:689 set_scalefacs: sf[] (in v->sftmp) -> w->sfw */
__device__ __forceinline__ void lg_vset_scalefacs(LgVWarp *v, const LgQInfo &gi, const LgQConst &qc, const uint8_t *max_range, int lane)
{
    LgQWarp *w = &v->q;
    int const ifqstep = (gi.scalefac_scale == 0) ? 2 : 4, ifqstepShift = (gi.scalefac_scale == 0) ? 1 : 2;
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        if (sfb < LG_SFBMAX) {
            int sc = 0;
            if (sfb < qc.sfbmax) {
                int sf = v->sftmp[sfb];
                int const pre = gi.preflag ? lg_pretab(sfb) : 0;
                if (gi.preflag && sfb >= 11) sf += pre * ifqstep;
                int const gain = gi.global_gain - (((gi.sbg >> (4 * w->window[sfb])) & 15) * 8) - pre * ifqstep;
                if (sf < 0) {
                    int const m = gain - v->vbrsfmin[sfb];
                    sc = (ifqstep - 1 - sf) >> ifqstepShift;
                    if (sc > (int) max_range[sfb]) sc = max_range[sfb];
                    if (sc > 0 && (sc << ifqstepShift) > m) sc = m >> ifqstepShift;
                }
            }
            w->sfw[sfb] = sc;
        }
    }
    __syncwarp();
}

/* vbrquantize.c:770 short_block_constrain / :848 long_block_constrain (+ :596 set_subblock_gain) on the steps in sfin[] */
__device__ __noinline__ void lg_valloc(const LgDevCfg *__restrict__ c, LgVWarp *v, LgQInfo &gi, const LgQConst &qc, const LgVCtx &ctx,
                                          const int *sfin, int vbrmax, int lane)
{
    int const maxminsfb = ctx.mingain_l, psymax = qc.psymax;
    int delta = 0, mover;
    if (ctx.is_short) {
        int maxover0 = 0, maxover1 = 0;
        for (int sfb = lane; sfb < psymax; sfb += 32) {
            int const d = vbrmax - sfin[sfb];
            delta = max(delta, d);
            maxover0 = max(maxover0, d - (4 * 14 + 2 * (int) LG_VRANGE_SHORT[sfb]));
            maxover1 = max(maxover1, d - (4 * 14 + 4 * (int) LG_VRANGE_SHORT[sfb]));
        }
        delta = lg_wmax_i(delta); maxover0 = lg_wmax_i(maxover0); maxover1 = lg_wmax_i(maxover1);
        if (c->noise_shaping == 2) mover = min(maxover0, maxover1);
        else mover = maxover0;
        if (delta > mover) delta = mover;
        vbrmax -= delta;
        maxover0 -= mover;
        maxover1 -= mover;
        if (maxover0 == 0) gi.scalefac_scale = 0;
        else if (maxover1 == 0) gi.scalefac_scale = 1;
        if (vbrmax < maxminsfb) vbrmax = maxminsfb;
        gi.global_gain = vbrmax;
        if (gi.global_gain < 0) gi.global_gain = 0;
        else if (gi.global_gain > 255) gi.global_gain = 255;
        for (int sfb = lane; sfb < 40; sfb += 32) v->sftmp[sfb] = (sfb < LG_SFBMAX) ? sfin[sfb] - vbrmax : 0;
        __syncwarp();
        /* set_subblock_gain: lanes 0..2 = the three windows */
        int sbgv = 0;
        if (lane < 3) {
            int const ifqstepShift = (gi.scalefac_scale == 0) ? 1 : 2;
            int const psydiv = psymax < 18 ? psymax : 18;
            int maxsf1 = 0, maxsf2 = 0, minsf = 1000;
            /* the window's bands below psydiv (part 1) and from psydiv on (part 2), vbrquantize.c:613-632.
             * One signed loop on purpose: two back-to-back loops with an unsigned band counter came out of
             * nvcc 12.9 with a wrong minimum on sm_100a (the emulator build of the same source was right). */
            for (int sfb = lane; sfb < LG_SFBMAX; sfb += 3) {
                int const t = -v->sftmp[sfb];
                if (sfb < psydiv) maxsf1 = max(maxsf1, t); else maxsf2 = max(maxsf2, t);
                minsf = min(minsf, t);
            }
            int const m1 = maxsf1 - (15 << ifqstepShift), m2 = maxsf2 - (7 << ifqstepShift);
            maxsf1 = max(m1, m2);
            sbgv = (minsf > 0) ? (minsf >> 3) : 0;
            if (maxsf1 > 0) sbgv = max(sbgv, (maxsf1 + 7) >> 3);
            if (sbgv > 0 && ctx.mingain_s[lane] > (gi.global_gain - sbgv * 8)) sbgv = (gi.global_gain - ctx.mingain_s[lane]) >> 3;
            if (sbgv > 7) sbgv = 7;
        }
        int const s0 = __shfl_sync(LG_FULL, sbgv, 0), s1 = __shfl_sync(LG_FULL, sbgv, 1), s2 = __shfl_sync(LG_FULL, sbgv, 2);
        int const min_sbg = min(7, min(s0, min(s1, s2)));
        for (int sfb = lane; sfb < LG_SFBMAX; sfb += 32) { int const wn = sfb % 3; v->sftmp[sfb] += 8 * (wn == 0 ? s0 : (wn == 1 ? s1 : s2)); }
        gi.sbg = s0 | (s1 << 4) | (s2 << 8);
        if (min_sbg > 0) {
            gi.sbg = (s0 - min_sbg) | ((s1 - min_sbg) << 4) | ((s2 - min_sbg) << 8);
            gi.global_gain -= min_sbg * 8;
        }
        __syncwarp();
        lg_vset_scalefacs(v, gi, qc, LG_VRANGE_SHORT, lane);
    }
    else {
        int maxover0 = 0, maxover1 = 0, maxover0p = 0, maxover1p = 0;
        for (int sfb = lane; sfb < psymax; sfb += 32) {
            int const d = vbrmax - sfin[sfb];
            int const rng = LG_VRANGE_LONG[sfb], pre = lg_pretab(sfb);
            int const rngp = (c->mode_gr == 2) ? rng : (int) LG_VRANGE_LONG_LSF[sfb];      /* vbrquantize.c:861 */
            delta = max(delta, d);
            maxover0 = max(maxover0, d - 2 * rng);
            maxover1 = max(maxover1, d - 4 * rng);
            maxover0p = max(maxover0p, d - 2 * (rngp + pre));
            maxover1p = max(maxover1p, d - 4 * (rngp + pre));
        }
        delta = lg_wmax_i(delta); maxover0 = lg_wmax_i(maxover0); maxover1 = lg_wmax_i(maxover1);
        maxover0p = lg_wmax_i(maxover0p); maxover1p = lg_wmax_i(maxover1p);
        int vm0p = 1, vm1p = 1;
        {
            int gain = vbrmax - maxover0p;
            if (gain < maxminsfb) gain = maxminsfb;
            int fail = 0;
            for (int sfb = lane; sfb < psymax; sfb += 32) if ((gain - v->vbrsfmin[sfb]) - 2 * lg_pretab(sfb) <= 0) fail = 1;
            if (__any_sync(LG_FULL, fail)) { vm0p = 0; vm1p = 0; }
        }
        if (vm1p == 1) {
            int gain = vbrmax - maxover1p;
            if (gain < maxminsfb) gain = maxminsfb;
            int fail = 0;
            for (int sfb = lane; sfb < psymax; sfb += 32) if ((gain - v->vbrsfmin[sfb]) - 4 * lg_pretab(sfb) <= 0) fail = 1;
            if (__any_sync(LG_FULL, fail)) vm1p = 0;
        }
        if (vm0p == 0) maxover0p = maxover0;
        if (vm1p == 0) maxover1p = maxover1;
        if (c->noise_shaping != 2) { maxover1 = maxover0; maxover1p = maxover0p; }
        mover = min(min(maxover0, maxover0p), min(maxover1, maxover1p));
        if (delta > mover) delta = mover;
        vbrmax -= delta;
        if (vbrmax < maxminsfb) vbrmax = maxminsfb;
        maxover0 -= mover; maxover0p -= mover; maxover1 -= mover; maxover1p -= mover;
        if (maxover0 == 0) { gi.scalefac_scale = 0; gi.preflag = 0; }
        else if (maxover0p == 0) { gi.scalefac_scale = 0; gi.preflag = 1; }
        else if (maxover1 == 0) { gi.scalefac_scale = 1; gi.preflag = 0; }
        else if (maxover1p == 0) { gi.scalefac_scale = 1; gi.preflag = 1; }
        gi.global_gain = vbrmax;
        if (gi.global_gain < 0) gi.global_gain = 0;
        else if (gi.global_gain > 255) gi.global_gain = 255;
        for (int sfb = lane; sfb < 40; sfb += 32) v->sftmp[sfb] = (sfb < LG_SFBMAX) ? sfin[sfb] - vbrmax : 0;
        __syncwarp();
        lg_vset_scalefacs(v, gi, qc, (c->mode_gr == 2 || !gi.preflag) ? LG_VRANGE_LONG : LG_VRANGE_LONG_LSF, lane);
    }
}

/* vbrquantize.c:1000 quantizeAndCountBits: quantize_x34 (:501) + noquant_count_bits */
__device__ __noinline__ int lg_vquantize_count(const LgDevCfg *__restrict__ c, LgVWarp *v, LgQInfo &gi, const LgQConst &qc, int lane)
{
    LgQWarp *w = &v->q;
    int const mnz = qc.max_nonzero_coeff;
    int const ifqstep = (gi.scalefac_scale == 0) ? 2 : 4;
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        if (sfb < 40) {
            float f = 0.f;
            if (sfb < LG_SFBMAX && w->lstart[sfb] <= mnz) {
                int const s = (w->sfw[sfb] + (gi.preflag ? lg_pretab(sfb) : 0)) * ifqstep + ((gi.sbg >> (4 * w->window[sfb])) & 15) * 8;
                f = __ldg(&c->ipow20[(gi.global_gain - s) & 255]);
            }
            v->bfac[sfb] = f;
        }
    }
    __syncwarp();
    int const ilim = (mnz + 2) & ~1;
    int hi_nz = -1, hi_big = -1;
    for (int j = 0; j < qc.jn; j++) {
        int const P = lane + 32 * j, i = 2 * P;
        if (i < mnz) {                                /* max_nonzero_coeff is odd: the pair (i, i + 1) is inside or outside as a whole */
            float const f = v->bfac[w->line_sfb[i]];
            float2 const xp = *reinterpret_cast<const float2 *>(&w->xrpow[i]);
            int const v0 = lg_vquant1(c->adj43asm, f * xp.x), v1 = lg_vquant1(c->adj43asm, f * xp.y);
            unsigned const nv = (unsigned) v0 | ((unsigned) v1 << 16);
            *reinterpret_cast<unsigned *>(&w->ixw[i]) = nv;
            if (i < ilim) {
                if (nv != 0u) hi_nz = P;
                if ((nv & 0xfffefffeu) != 0u) hi_big = P;
            }
        }
    }
    LgPrev pv; pv.valid = 0; pv.global_gain = 0; pv.sfb_count1 = 0;
    gi.part2_3_length = lg_noquant_tail(c, w, gi, qc, pv, hi_nz, hi_big, lane);
    return gi.part2_3_length;
}

/* bitcount (vbrquantize.c:985): scale_bitcount must succeed for scalefactors chosen this way */
__device__ __forceinline__ void lg_vbitcount(const LgDevCfg *__restrict__ c, LgVWarp *v, LgQInfo &gi, const LgQConst &qc, int lane)
{
    unsigned const r = lg_scale_bitcount(&v->q, qc.block_type, qc.sfbmax, qc.sfbdivide, gi.preflag, gi.scalefac_compress, LG_LSF_ARG(c, gi), lane);
    LG_APPLY_SCALE_BITCOUNT(gi, r);
    if ((r >> 30) & 1u) lg_runaway();
}

/* vbrquantize.c:1141 tryThatOne / :1012 tryGlobalStepsize */
__device__ __noinline__ int lg_vtry(const LgDevCfg *__restrict__ c, LgVWarp *v, LgQInfo &gi, const LgQConst &qc, const LgVCtx &ctx,
                                      const int *sft, int vbrmax, int add_part2, int lane)
{
    lg_valloc(c, v, gi, qc, ctx, sft, vbrmax, lane);
    lg_vbitcount(c, v, gi, qc, lane);
    int nbits = lg_vquantize_count(c, v, gi, qc, lane);
    if (add_part2) nbits += gi.part2_length;
    return nbits;
}

/* vbrquantize.c:1105 flattenDistribution: sfwork (v->vbrsf) -> v->wrk, returns the largest step */
__device__ __forceinline__ int lg_vflatten(LgVWarp *v, int dm, int k, int p, int lane)
{
    int sfmax = 0;
    for (int i = lane; i < LG_SFBMAX; i += 32) {
        int x = v->vbrsf[i];
        if (dm > 0) {
            x = x + (k * (p - x)) / dm;
            if (x < 0) x = 0;
            else if (x > 255) x = 255;
        }
        v->wrk[i] = x;
        sfmax = max(sfmax, x);
    }
    sfmax = lg_wmax_i(sfmax);
    __syncwarp();
    return sfmax;
}

/* vbrquantize.c:1155 outOfBitsStrategy (+ :1041 searchGlobalStepsizeMax) */
__device__ __noinline__ void lg_vout_of_bits(const LgDevCfg *__restrict__ c, LgVWarp *v, LgQInfo &gi, const LgQConst &qc, const LgVCtx &ctx, int target, int lane)
{
    int dm = 0;
    for (int i = lane; i < LG_SFBMAX; i += 32) dm = max(dm, 255 - v->vbrsf[i]);          /* sfDepth */
    dm = lg_wmax_i(dm);
    int const p = gi.global_gain;
    for (int part = 0; part < 2; part++) {
        int bi = part == 0 ? dm / 2 : (255 + p) / 2, bi_ok = -1, bu = part == 0 ? 0 : p, bo = part == 0 ? dm : 255;
        for (int guard = 0;; guard++) {
            if (guard > 600) lg_runaway();
            int const sfmax = part == 0 ? lg_vflatten(v, dm, bi, p, lane) : lg_vflatten(v, dm, dm, bi, lane);
            int const nbits = lg_vtry(c, v, gi, qc, ctx, v->wrk, sfmax, 1, lane);
            if (nbits <= target) { bi_ok = bi; bo = bi - 1; }
            else bu = bi + 1;
            if (bu <= bo) bi = (bu + bo) / 2;
            else break;
        }
        if (bi_ok >= 0) {
            if (bi != bi_ok) {
                int const sfmax = part == 0 ? lg_vflatten(v, dm, bi_ok, p, lane) : lg_vflatten(v, dm, dm, bi_ok, lane);
                (void) lg_vtry(c, v, gi, qc, ctx, v->wrk, sfmax, 1, lane);
            }
            return;
        }
    }
    {   /* searchGlobalStepsizeMax on the last tried distribution (v->wrk) */
        int const gain = gi.global_gain;
        int curr = gain, gain_ok = 1024, l = gain, r = 512;
        for (int pass = 0;; pass++) {
            if (pass > 64) lg_runaway();
            int const last = !(l <= r);
            if (last) {
                if (gain_ok == curr) break;
                curr = gain_ok;
            }
            else curr = (l + r) >> 1;
            int vbrmax = 0;
            for (int i = lane; i < 40; i += 32) {
                int g = 0;
                if (i < LG_SFBMAX) {
                    g = v->wrk[i] + (curr - gain);
                    if (g < v->vbrsfmin[i]) g = v->vbrsfmin[i];
                    if (g > 255) g = 255;
                    vbrmax = max(vbrmax, g);
                }
                v->bsf[i] = g;
            }
            vbrmax = lg_wmax_i(vbrmax);
            __syncwarp();
            int const nbits = lg_vtry(c, v, gi, qc, ctx, v->bsf, vbrmax, 0, lane);
            if (last) break;
            if (nbits == 0 || (nbits + gi.part2_length) < target) { r = curr - 1; gain_ok = curr; }
            else { l = curr + 1; if (gain_ok == 1024) gain_ok = curr; }
        }
    }
}

/* ---------------------------------------------------------------- the kernel: one CTA per stream, warp = gr * 2 + ch */
__global__ void __launch_bounds__(128)
lg_kernel_vbr(const LgDevCfg *__restrict__ cfg, const float *__restrict__ xr_in, const LgPsyOut *__restrict__ psy,
              const LgFrameCtl *__restrict__ frm, LgGranuleOut *__restrict__ gout, LgFrameOut *__restrict__ fout,
              LgStreamState *__restrict__ state, const int *__restrict__ nfr, int nframes)
{
    LG_DYN_SMEM(LgSmemV, sm);
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int const gr = warp >> 1, ch = warp & 1;
    int const stream = blockIdx.x;
    int const nch = cfg->channels;
    const LgDevCfg *c = cfg;
    LgStreamState *st = state + stream;
    LgVWarp *v = &sm->w[warp];
    LgQWarp *w = &v->q;
    int resv_size = st->resv_size, main_data_begin = st->main_data_begin;
    int anc_flag = st->ancillary_flag, pay_off = 0;
    int const mgr = cfg->mode_gr;                 /* MPEG-2/2.5: one granule per frame, warps 2 and 3 stay idle */
    int const active = ch < nch && gr < mgr;
    int const my_frames = nfr[stream];
    if (my_frames <= 0) return;                        /* nothing of this stream in this step: its state is not ours to write back (another step's kernel may own it) */
    for (int frame = 0; frame < my_frames; frame++) {
        const LgFrameCtl *F = frm + (size_t) stream * nframes + frame;
        int const padding = F->padding, mode_ext = F->mode_ext;
        int const gb = mgr * frame + (active ? gr : 0);
        const LgPsyOut *P = psy + (size_t) stream * 2 * nframes + gb;
        /* ---- VBR_new_prepare (quantize.c:1582): frame sizes of every bitrate index, bit budget, allowed noise */
        int frameBits[16], avg, resv_max_m, dummy;
        (void) lg_resv_frame_begin(c, c->vbr_max_bitrate_index, padding, resv_size, &avg, &resv_max_m);
        int pad = resv_max_m;
        for (int i = 1; i <= c->vbr_max_bitrate_index; i++) frameBits[i] = lg_resv_frame_begin(c, i, padding, resv_size, &dummy, &dummy);
        int const maximum_framebits = frameBits[c->vbr_max_bitrate_index];
        int mb[2][2];
        mb[1][0] = mb[1][1] = 0;
        for (int g = 0; g < mgr; g++) {
            float pe[2] = { F->pe_use[g][0], F->pe_use[g][1] };
            int tb[2];
            (void) lg_on_pe(c, resv_size, resv_max_m, pe, tb, avg, 0);
            mb[g][0] = tb[0]; mb[g][1] = nch == 2 ? tb[1] : 0;
        }
        LgQInfo gi;
        LgQConst qc;
        LgVCtx ctx;
        gi.part2_3_length = 0; gi.big_values = 0; gi.count1 = 0; gi.global_gain = 210; gi.scalefac_compress = 0;
        gi.table_select[0] = gi.table_select[1] = gi.table_select[2] = 0; gi.sbg = 0;
        gi.region0_count = 0; gi.region1_count = 0; gi.preflag = 0; gi.scalefac_scale = 0;
        gi.count1table_select = 0; gi.part2_length = 0; gi.count1bits = 0; gi.xrpow_max = 0;
        qc.ath_over = 0; qc.block_type = LG_NORM; qc.max_nonzero_coeff = 575; qc.jn = 9;
        qc.sfb_lmax = LG_SBPSY_L; qc.sfb_smin = LG_SBPSY_S; qc.psy_lmax = LG_SBPSY_L; qc.psymax = LG_SBPSY_L; qc.sfbmax = LG_SBPSY_L; qc.sfbdivide = 11;
        ctx.mingain_l = 0; ctx.mingain_s[0] = ctx.mingain_s[1] = ctx.mingain_s[2] = 0; ctx.is_short = 0; ctx.guess_only = c->full_outer_loop < 0;
        int nonzero = 0;
        if (active) {
            /* quantize.c:226 init_outer_loop */
            qc.block_type = P->block_type[ch];
            qc.psy_lmax = c->sfb21_extra ? LG_SBMAX_L : LG_SBPSY_L;
            if (c->samplerate <= 8000) { qc.sfb_lmax = 17; qc.sfb_smin = 9; qc.psy_lmax = 17; qc.sfbmax = 17; }       /* quantize.c:252-256 */
            qc.psymax = qc.psy_lmax;
            if (qc.block_type == LG_SHORT) {
                qc.sfb_smin = 0; qc.sfb_lmax = 0;
                qc.psymax = 3 * (c->sfb21_extra ? LG_SBMAX_S : LG_SBPSY_S);
                qc.sfbmax = 3 * LG_SBPSY_S;
                if (c->samplerate <= 8000) qc.psymax = qc.sfbmax = 3 * 9;                                          /* quantize.c:284-289 */
                qc.sfbdivide = qc.sfbmax - 18;
                qc.psy_lmax = 0;
            }
            ctx.is_short = qc.block_type == LG_SHORT;
            for (int r = 0; r < 2; r++) {
                int const k = lane + 32 * r;
                if (k <= 40) {
                    int ws = 0, wn = 3, ls = 576;
                    if (qc.block_type == LG_SHORT) {
                        if (k < 39) {
                            int const sfb = k / 3;
                            ws = c->sfb_s[sfb + 1] - c->sfb_s[sfb];
                            wn = k % 3;
                            ls = 3 * c->sfb_s[sfb] + wn * ws;
                        }
                    }
                    else if (k < LG_SBMAX_L) { ws = c->sfb_l[k + 1] - c->sfb_l[k]; ls = c->sfb_l[k]; }
                    if (k < 40) { w->width[k] = ws; w->window[k] = wn; w->sfw[k] = 0; }
                    w->lstart[k] = ls;
                }
            }
            {
                const float *src = xr_in + (((size_t) stream * 2 * nframes + gb) * 2 + ch) * 576;
                for (int i = lane; i < 144; i += 32) reinterpret_cast<float4 *>(w->xr)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
                const unsigned *map = reinterpret_cast<const unsigned *>(qc.block_type == LG_SHORT ? c->line_sfb_s : c->line_sfb_l);
                for (int i = lane; i < 144; i += 32) reinterpret_cast<unsigned *>(w->line_sfb)[i] = __ldg(map + i);
                for (int i = lane; i < 288; i += 32) reinterpret_cast<unsigned *>(w->ixw)[i] = 0u;
            }
            __syncwarp();
            int const rch = (mode_ext == 2) ? ch + 2 : ch;
            {
                LgQConst qx = qc;
                lg_calc_xmin(c, w, qx, &P->en[rch], &P->thm[rch], F->ath_adjust_factor, lane);
                qc.max_nonzero_coeff = qx.max_nonzero_coeff; qc.jn = qx.jn; qc.ath_over = qx.ath_over;
            }
            /* quantize.c:110 init_xrpow with upper = max_nonzero_coeff */
            float mx = 0.f, amax = 0.f;
            for (int j = 0; j < 9; j++) {
                int const i = 2 * (lane + 32 * j);
                float2 pw; pw.x = 0.f; pw.y = 0.f;
                if (i <= qc.max_nonzero_coeff) {
                    float const t0 = fabsf(w->xr[i]);
                    pw.x = (float) sqrt((double) t0 * sqrt((double) t0));
                    if (t0 > amax) amax = t0;
                    if (i + 1 <= qc.max_nonzero_coeff) {
                        float const t1 = fabsf(w->xr[i + 1]);
                        pw.y = (float) sqrt((double) t1 * sqrt((double) t1));
                        if (t1 > amax) amax = t1;
                    }
                }
                *reinterpret_cast<float2 *>(&w->xrpow[i]) = pw;
                mx = fmaxf(mx, fmaxf(pw.x, pw.y));
            }
            gi.xrpow_max = lg_wmax_fpos(mx);
            amax = lg_wmax_fpos(amax);
            __syncwarp();
            nonzero = amax > (float) 1E-20;
            if (!nonzero && amax > 0.f) {
                float sum = 0;
                for (int i = 0; i <= qc.max_nonzero_coeff; ++i) sum += fabsf(w->xr[i]);
                nonzero = sum > (float) 1E-20;
            }
        }
        if (lane == 0) { sm->ath_over[warp] = active ? qc.ath_over : 0; sm->nonzero[warp] = nonzero; }
        __syncthreads();
        int analog_silence = 1;
        {
            int bits = 0;
            for (int g = 0; g < mgr; g++) for (int k = 0; k < nch; k++) { bits += mb[g][k]; if (sm->ath_over[g * 2 + k]) analog_silence = 0; }
            for (int g = 0; g < mgr; g++)
                for (int k = 0; k < nch; k++)
                    if (bits > maximum_framebits && bits > 0) { mb[g][k] *= maximum_framebits; mb[g][k] /= bits; }
            for (int g = 0; g < mgr; g++) for (int k = 0; k < nch; k++) if (!sm->nonzero[g * 2 + k]) mb[g][k] = 0;
        }
        if (analog_silence) pad = 0;
        /* ---- VBR_encode_frame (vbrquantize.c:1255): search, fit, quantise "as is" */
        int const my_max = active ? mb[gr][ch] : 0;
        if (active && my_max > 0) {
            int const vbrmax = lg_vblock_sf(c, v, qc, ctx, lane);
            lg_valloc(c, v, gi, qc, ctx, v->vbrsf, vbrmax, lane);
            lg_vbitcount(c, v, gi, qc, lane);
            (void) lg_vquantize_count(c, v, gi, qc, lane);
        }
        /* reduce_bit_usage: granule 1 reads granule 0's scalefactors (scfsi), so the two granules take turns */
        uint8_t scfsi[4] = { 0, 0, 0, 0 };
        for (int turn = 0; turn < 2; turn++) {
            if (active && gr == turn) {
                lg_best_scalefac_store(c, w, gi, qc, gr, sm->sf_gr0[ch], sm->bt_gr0[ch], scfsi, lane);
                if (c->use_best_huffman == 1) lg_best_huffman_divide(c, w, gi, qc, lane);
                if (gr == 0) {
                    for (int i = lane; i < 40; i += 32) sm->sf_gr0[ch][i] = w->sfw[i];
                    if (lane == 0) sm->bt_gr0[ch] = qc.block_type;
                }
                if (lane == 0) sm->use_bits[warp] = gi.part2_3_length + gi.part2_length;
            }
            else if (!active && lane == 0) sm->use_bits[warp] = 0;
            __syncthreads();
        }
        /* ---- the frame's bit budget (vbrquantize.c:1341-1530), computed by every thread alike */
        int use_ch[2][2], use_gr[2] = { 0, 0 }, use_fr = 0, max_fr = 0;
        for (int g = 0; g < mgr; g++) for (int k = 0; k < 2; k++) { use_ch[g][k] = k < nch ? sm->use_bits[g * 2 + k] : 0; use_gr[g] += use_ch[g][k]; if (k < nch) max_fr += mb[g][k]; }
        use_fr = use_gr[0] + use_gr[1];
        int fits = 0;
        if (use_fr <= max_fr) {
            fits = 1;
            for (int g = 0; g < mgr; g++) {
                if (use_gr[g] > LG_MAX_BITS_PER_GRANULE) fits = 0;
                for (int k = 0; k < nch; k++) if (use_ch[g][k] > LG_MAX_BITS_PER_CHANNEL) fits = 0;
            }
        }
        if (!fits) {
            int max_ch[2][2] = { { 0, 0 }, { 0, 0 } }, max_gr[2] = { 0, 0 }, ok = 1, sum_fr = 0;
            for (int g = 0; g < mgr; g++) {
                max_gr[g] = 0;
                for (int k = 0; k < nch; k++) {
                    max_ch[g][k] = use_ch[g][k] > LG_MAX_BITS_PER_CHANNEL ? LG_MAX_BITS_PER_CHANNEL : use_ch[g][k];
                    max_gr[g] += max_ch[g][k];
                }
                if (max_gr[g] > LG_MAX_BITS_PER_GRANULE) {
                    float f[2] = { 0.0f, 0.0f }, s = 0.0f;
                    for (int k = 0; k < nch; k++) {
                        if (max_ch[g][k] > 0) { f[k] = (float) sqrt(sqrt((double) max_ch[g][k])); s += f[k]; }
                        else f[k] = 0;
                    }
                    for (int k = 0; k < nch; k++) max_ch[g][k] = (s > 0) ? (int) (LG_MAX_BITS_PER_GRANULE * f[k] / s) : 0;
                    if (nch > 1) {
                        if (max_ch[g][0] > use_ch[g][0] + 32) { max_ch[g][1] += max_ch[g][0]; max_ch[g][1] -= use_ch[g][0] + 32; max_ch[g][0] = use_ch[g][0] + 32; }
                        if (max_ch[g][1] > use_ch[g][1] + 32) { max_ch[g][0] += max_ch[g][1]; max_ch[g][0] -= use_ch[g][1] + 32; max_ch[g][1] = use_ch[g][1] + 32; }
                        if (max_ch[g][0] > LG_MAX_BITS_PER_CHANNEL) max_ch[g][0] = LG_MAX_BITS_PER_CHANNEL;
                        if (max_ch[g][1] > LG_MAX_BITS_PER_CHANNEL) max_ch[g][1] = LG_MAX_BITS_PER_CHANNEL;
                    }
                    max_gr[g] = 0;
                    for (int k = 0; k < nch; k++) max_gr[g] += max_ch[g][k];
                }
                sum_fr += max_gr[g];
            }
            if (sum_fr > max_fr) {
                {
                    float f[2] = { 0.0f, 0.0f }, s = 0.0f;
                    for (int g = 0; g < mgr; g++) {
                        if (max_gr[g] > 0) { f[g] = (float) sqrt((double) max_gr[g]); s += f[g]; }
                        else f[g] = 0;
                    }
                    for (int g = 0; g < mgr; g++) max_gr[g] = (s > 0) ? (int) (max_fr * f[g] / s) : 0;
                }
                if (mgr > 1) {
                if (max_gr[0] > use_gr[0] + 125) { max_gr[1] += max_gr[0]; max_gr[1] -= use_gr[0] + 125; max_gr[0] = use_gr[0] + 125; }
                if (max_gr[1] > use_gr[1] + 125) { max_gr[0] += max_gr[1]; max_gr[0] -= use_gr[1] + 125; max_gr[1] = use_gr[1] + 125; }
                for (int g = 0; g < mgr; g++) if (max_gr[g] > LG_MAX_BITS_PER_GRANULE) max_gr[g] = LG_MAX_BITS_PER_GRANULE;
                }
                for (int g = 0; g < mgr; g++) {
                    float f[2] = { 0.0f, 0.0f }, s = 0.0f;
                    for (int k = 0; k < nch; k++) {
                        if (max_ch[g][k] > 0) { f[k] = (float) sqrt((double) max_ch[g][k]); s += f[k]; }
                        else f[k] = 0;
                    }
                    for (int k = 0; k < nch; k++) max_ch[g][k] = (s > 0) ? (int) (max_gr[g] * f[k] / s) : 0;
                    if (nch > 1) {
                        if (max_ch[g][0] > use_ch[g][0] + 32) { max_ch[g][1] += max_ch[g][0]; max_ch[g][1] -= use_ch[g][0] + 32; max_ch[g][0] = use_ch[g][0] + 32; }
                        if (max_ch[g][1] > use_ch[g][1] + 32) { max_ch[g][0] += max_ch[g][1]; max_ch[g][0] -= use_ch[g][1] + 32; max_ch[g][1] = use_ch[g][1] + 32; }
                        for (int k = 0; k < nch; k++) if (max_ch[g][k] > LG_MAX_BITS_PER_CHANNEL) max_ch[g][k] = LG_MAX_BITS_PER_CHANNEL;
                    }
                }
            }
            sum_fr = 0;
            for (int g = 0; g < mgr; g++) {
                int sum_gr = 0;
                for (int k = 0; k < nch; k++) { sum_gr += max_ch[g][k]; if (max_ch[g][k] > LG_MAX_BITS_PER_CHANNEL) ok = 0; }
                sum_fr += sum_gr;
                if (sum_gr > LG_MAX_BITS_PER_GRANULE) ok = 0;
            }
            if (sum_fr > max_fr) ok = 0;
            if (!ok) for (int g = 0; g < mgr; g++) for (int k = 0; k < nch; k++) max_ch[g][k] = mb[g][k];
            /* best_scalefac_store already ran once: reset what it left behind, then re-quantise until it fits */
            for (int i = 0; i < 4; i++) scfsi[i] = 0;
            gi.scalefac_compress = 0;
            if (active && my_max > 0) {
                int const cut = gi.global_gain;
                for (int i = lane; i < LG_SFBMAX; i += 32) if (v->vbrsf[i] > cut) v->vbrsf[i] = cut;      /* cutDistribution */
                __syncwarp();
                lg_vout_of_bits(c, v, gi, qc, ctx, max_ch[gr][ch], lane);
            }
            __syncthreads();
            for (int turn = 0; turn < 2; turn++) {
                if (active && gr == turn) {
                    lg_best_scalefac_store(c, w, gi, qc, gr, sm->sf_gr0[ch], sm->bt_gr0[ch], scfsi, lane);
                    if (c->use_best_huffman == 1) lg_best_huffman_divide(c, w, gi, qc, lane);
                    if (gr == 0) {
                        for (int i = lane; i < 40; i += 32) sm->sf_gr0[ch][i] = w->sfw[i];
                        if (lane == 0) sm->bt_gr0[ch] = qc.block_type;
                    }
                    if (lane == 0) sm->use_bits[warp] = gi.part2_3_length + gi.part2_length;
                }
                __syncthreads();
            }
            use_fr = 0;
            for (int g = 0; g < mgr; g++) for (int k = 0; k < nch; k++) use_fr += sm->use_bits[g * 2 + k];
        }
        int const used_bits = use_fr;
        /* ---- hand the granule to the bit packer */
        if (active) {
            LgGranuleOut *o = gout + (((size_t) stream * 2 * nframes + gb) * 2 + ch);
            for (int j = 0; j < 9; j++) {
                int const i = 2 * (lane + 32 * j);
                int v0 = w->ixw[i], v1 = w->ixw[i + 1];
                if (w->xr[i] < 0.0f) v0 = -v0;
                if (w->xr[i + 1] < 0.0f) v1 = -v1;
                *reinterpret_cast<unsigned *>(&o->ix[i]) = ((unsigned) v0 & 0xffffu) | ((unsigned) v1 << 16);
            }
            for (int i = lane; i < 40; i += 32) o->scalefac[i] = (int8_t) (i < 39 ? w->sfw[i] : 0);
            if (lane == 0) {
                o->part2_3_length = (int16_t) gi.part2_3_length; o->part2_length = (int16_t) gi.part2_length;
                o->big_values = (int16_t) gi.big_values; o->count1 = (int16_t) gi.count1;
                o->global_gain = (uint8_t) gi.global_gain; o->scalefac_compress = (uint8_t) gi.scalefac_compress;
                o->scalefac_compress_hi = (uint8_t) (gi.scalefac_compress >> 8);
                o->block_type = (uint8_t) qc.block_type; o->mixed_block_flag = 0;
                for (int i = 0; i < 3; i++) { o->table_select[i] = (uint8_t) gi.table_select[i]; o->subblock_gain[i] = (uint8_t) ((gi.sbg >> (4 * i)) & 15); }
                o->region0_count = (uint8_t) gi.region0_count; o->region1_count = (uint8_t) gi.region1_count;
                o->preflag = (uint8_t) gi.preflag; o->scalefac_scale = (uint8_t) gi.scalefac_scale;
                o->count1table_select = (uint8_t) gi.count1table_select;
                o->sfbmax = (uint8_t) qc.sfbmax; o->sfbdivide = (uint8_t) qc.sfbdivide;
            }
            if (lane == 0 && gr == 1) for (int i = 0; i < 4; i++) sm->scfsi[ch][i] = scfsi[i];
        }
        else if (lane == 0 && gr == 1) for (int i = 0; i < 4; i++) sm->scfsi[ch][i] = 0;
        __syncthreads();
        /* ---- the lowest bitrate that holds the frame (quantize.c:1700-1735), reservoir update */
        {
            int i = (analog_silence && !c->enforce_min_bitrate) ? 1 : c->vbr_min_bitrate_index, bitrate_index;
            for (; i < c->vbr_max_bitrate_index; i++) if (used_bits <= frameBits[i]) break;
            if (i > c->vbr_max_bitrate_index) i = c->vbr_max_bitrate_index;
            if (pad > 0) {
                int j;
                for (j = c->vbr_max_bitrate_index; j > i; --j) if (frameBits[j] - used_bits <= pad) break;
                bitrate_index = j;
            }
            else bitrate_index = i;
            if (used_bits > frameBits[bitrate_index]) lg_runaway();          /* "INTERNAL ERROR IN VBR NEW CODE" */
            int mean_bits, resv_max;
            (void) lg_resv_frame_begin(c, bitrate_index, padding, resv_size, &mean_bits, &resv_max);
            resv_size -= used_bits;                                          /* ResvAdjust for the four granule.channels */
            /* reservoir.c:239 ResvFrameEnd + the main_data_begin recurrence of format_bitstream (bitstream.c:937) */
            int stuffingBits = 0, over_bits, drain_pre = 0, drain_post = 0;
            resv_size += mean_bits * c->mode_gr;
            if ((over_bits = resv_size % 8) != 0) stuffingBits += over_bits;
            over_bits = (resv_size - stuffingBits) - resv_max;
            if (over_bits > 0) stuffingBits += over_bits;
            int const mdb_bytes = (main_data_begin * 8 < stuffingBits ? main_data_begin * 8 : stuffingBits) / 8;
            drain_pre += 8 * mdb_bytes;
            stuffingBits -= 8 * mdb_bytes;
            resv_size -= 8 * mdb_bytes;
            int const mdb_header = main_data_begin - mdb_bytes;
            drain_post += stuffingBits;
            resv_size -= stuffingBits;
            main_data_begin = resv_size / 8;
            int const pay_bits = drain_pre + used_bits + drain_post;
            if (pay_bits & 7) lg_runaway();
            int const anc_pre = anc_flag;
            if (!c->disable_reservoir) anc_flag ^= (lg_drain_tail_bits(drain_pre) + lg_drain_tail_bits(drain_post)) & 1;
            if (threadIdx.x == 0) {
                LgFrameOut *fo = fout + (size_t) stream * nframes + frame;
                fo->main_data_begin = mdb_header; fo->drain_pre = drain_pre; fo->drain_post = drain_post;
                fo->padding = padding; fo->mode_ext = mode_ext; fo->resv_size = resv_size;
                fo->pay_off = pay_off; fo->pay_bytes = pay_bits >> 3;
                fo->anc_pre = (uint8_t) anc_pre; fo->anc_post = (uint8_t) anc_flag; fo->pad_[0] = fo->pad_[1] = 0; fo->bitrate_index = bitrate_index;
                for (int k = 0; k < 2; k++) for (int i2 = 0; i2 < 4; i2++) fo->scfsi[k][i2] = (uint8_t) sm->scfsi[k][i2];
            }
            pay_off += pay_bits >> 3;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { st->resv_size = resv_size; st->main_data_begin = main_data_begin; st->ancillary_flag = anc_flag; }
}
