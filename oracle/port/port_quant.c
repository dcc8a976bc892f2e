/* oracle/port/port_quant.c - TEST INFRASTRUCTURE (see lame_port.h).
 * Restates the CBR quantisation path: CBR_iteration_loop (quantize.c:1988), outer_loop (:1010) and
 * its helpers, calc_xmin/calc_noise/on_pe/reduce_side (quantize_pvt.c), the quantiser and Huffman
 * bit counting (takehiro.c) and the bit reservoir budget (reservoir.c). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include "lame_port.h"
#include "port_tables.inc"

#define SQRT2_D 1.41421356237309504880
#define LOG2_D 0.69314718055994530942
#define LOG10_D 2.30258509299404568402
#define FAST_LOG10_D(c, x) (lp_fast_log2(c, x) * (LOG2_D / LOG10_D))
#define FAST_LOG10_X_D(c, x, y) (lp_fast_log2(c, x) * (LOG2_D / LOG10_D * (y)))

typedef struct { float over_noise, tot_noise, max_noise; int over_count, over_SSD, bits; } noise_result;

/* test aid: LP_TRACE_FRAME=<n> prints the bit targets and every noise-shaping round of frame n to stderr (compared by hand with
 * the same trace of the device code under the emulator, LG_TRACE_FRAME) */
static int lp_trace_on(const lp_encoder *e)
{
    static int want = -2;
    if (want == -2) { const char *v = getenv("LP_TRACE_FRAME"); want = v ? atoi(v) : -1; }
    return want >= 0 && e->frame_number == want;
}
typedef struct { int global_gain, sfb_count1, step[39]; float noise[39], noise_log[39]; } noise_cache;

static const uint8_t *hlen_of(int t) { return LGT_HUFF_LEN + LGT_HUFF_OFF[t]; }

/* ------------------------------------------------------------------ bit reservoir (reservoir.c) */
/* bitstream.c:65 getframebits */
int lp_getframebits(const lp_encoder *e)
{
    /* bit_rate = bitrate_table[version][bitrate_index] of the frame being coded */
    return 8 * ((e->cfg.version + 1) * 72000 * LGT_BITRATE[16 * e->cfg.version + e->bitrate_index] / e->cfg.samplerate + e->padding);
}
/* reservoir.c:83 ResvFrameBegin */
static int resv_frame_begin(lp_encoder *e, int *mean_bits)
{
    const lp_config *cfg = &e->cfg;
    int frameLength = lp_getframebits(e);
    int meanBits = (frameLength - cfg->sideinfo_len * 8) / cfg->mode_gr;
    int resvLimit = (8 * 256) * cfg->mode_gr - 8;
    int maxmp3buf = cfg->buffer_constraint;
    int fullFrameBits;
    e->resv_max = maxmp3buf - frameLength;
    if (e->resv_max > resvLimit) e->resv_max = resvLimit;
    if (e->resv_max < 0 || cfg->disable_reservoir) e->resv_max = 0;
    fullFrameBits = meanBits * cfg->mode_gr + (e->resv_size < e->resv_max ? e->resv_size : e->resv_max);
    if (fullFrameBits > maxmp3buf) fullFrameBits = maxmp3buf;
    e->drain_pre = 0;
    *mean_bits = meanBits;
    return fullFrameBits;
}
/* reservoir.c:175 ResvMaxBits */
static void resv_max_bits(lp_encoder *e, int mean_bits, int *targ_bits, int *extra_bits, int cbr)
{
    const lp_config *cfg = &e->cfg;
    int add_bits, targBits, extraBits;
    int ResvSize = e->resv_size, ResvMax = e->resv_max;
    if (cbr) ResvSize += mean_bits;
    targBits = mean_bits;
    if (ResvSize * 10 > ResvMax * 9) {
        add_bits = ResvSize - (ResvMax * 9) / 10;
        targBits += add_bits;
    }
    else {
        add_bits = 0;
        if (!cfg->disable_reservoir) targBits -= .1 * mean_bits;
    }
    extraBits = (ResvSize < (e->resv_max * 6) / 10 ? ResvSize : (e->resv_max * 6) / 10);
    extraBits -= add_bits;
    if (extraBits < 0) extraBits = 0;
    *targ_bits = targBits;
    *extra_bits = extraBits;
}
/* reservoir.c:239 ResvFrameEnd */
static void resv_frame_end(lp_encoder *e, int mean_bits)
{
    int stuffingBits = 0, over_bits, mdb_bytes;
    e->resv_size += mean_bits * e->cfg.mode_gr;
    e->drain_post = 0;
    e->drain_pre = 0;
    if ((over_bits = e->resv_size % 8) != 0) stuffingBits += over_bits;
    over_bits = (e->resv_size - stuffingBits) - e->resv_max;
    if (over_bits > 0) stuffingBits += over_bits;
    mdb_bytes = (e->main_data_begin * 8 < stuffingBits ? e->main_data_begin * 8 : stuffingBits) / 8;
    e->drain_pre += 8 * mdb_bytes;
    stuffingBits -= 8 * mdb_bytes;
    e->resv_size -= 8 * mdb_bytes;
    e->main_data_begin -= mdb_bytes;
    e->drain_post += stuffingBits;
    e->resv_size -= stuffingBits;
}

/* quantize_pvt.c:428 on_pe */
static int on_pe(lp_encoder *e, float pe[2][2], int targ_bits[2], int mean_bits, int gr, int cbr)
{
    const lp_config *cfg = &e->cfg;
    int extra_bits = 0, tbits, bits, add_bits[2] = { 0, 0 }, max_bits, ch;
    resv_max_bits(e, mean_bits, &tbits, &extra_bits, cbr);
    max_bits = tbits + extra_bits;
    if (max_bits > LP_MAX_BITS_PER_GRANULE) max_bits = LP_MAX_BITS_PER_GRANULE;
    for (bits = 0, ch = 0; ch < cfg->channels; ++ch) {
        targ_bits[ch] = LP_MAX_BITS_PER_CHANNEL < tbits / cfg->channels ? LP_MAX_BITS_PER_CHANNEL : tbits / cfg->channels;
        add_bits[ch] = targ_bits[ch] * pe[gr][ch] / 700.0 - targ_bits[ch];
        if (add_bits[ch] > mean_bits * 3 / 4) add_bits[ch] = mean_bits * 3 / 4;
        if (add_bits[ch] < 0) add_bits[ch] = 0;
        if (add_bits[ch] + targ_bits[ch] > LP_MAX_BITS_PER_CHANNEL)
            add_bits[ch] = 0 > LP_MAX_BITS_PER_CHANNEL - targ_bits[ch] ? 0 : LP_MAX_BITS_PER_CHANNEL - targ_bits[ch];
        bits += add_bits[ch];
    }
    if (bits > extra_bits && bits > 0)
        for (ch = 0; ch < cfg->channels; ++ch) add_bits[ch] = extra_bits * add_bits[ch] / bits;
    for (ch = 0; ch < cfg->channels; ++ch) {
        targ_bits[ch] += add_bits[ch];
        extra_bits -= add_bits[ch];
    }
    for (bits = 0, ch = 0; ch < cfg->channels; ++ch) bits += targ_bits[ch];
    if (bits > LP_MAX_BITS_PER_GRANULE)
        for (ch = 0; ch < cfg->channels; ++ch) {
            targ_bits[ch] *= LP_MAX_BITS_PER_GRANULE;
            targ_bits[ch] /= bits;
        }
    return max_bits;
}
/* quantize_pvt.c:492 reduce_side */
static void reduce_side(int targ_bits[2], float ms_ener_ratio, int mean_bits, int max_bits)
{
    int move_bits;
    float fac;
    fac = .33 * (.5 - ms_ener_ratio) / .5;
    if (fac < 0) fac = 0;
    if (fac > .5) fac = .5;
    move_bits = fac * .5 * (targ_bits[0] + targ_bits[1]);
    if (move_bits > LP_MAX_BITS_PER_CHANNEL - targ_bits[0]) move_bits = LP_MAX_BITS_PER_CHANNEL - targ_bits[0];
    if (move_bits < 0) move_bits = 0;
    if (targ_bits[1] >= 125) {
        if (targ_bits[1] - move_bits > 125) {
            if (targ_bits[0] < mean_bits) targ_bits[0] += move_bits;
            targ_bits[1] -= move_bits;
        }
        else {
            targ_bits[0] += targ_bits[1] - 125;
            targ_bits[1] = 125;
        }
    }
    move_bits = targ_bits[0] + targ_bits[1];
    if (move_bits > max_bits) {
        targ_bits[0] = (max_bits * targ_bits[0]) / move_bits;
        targ_bits[1] = (max_bits * targ_bits[1]) / move_bits;
    }
}

/* ------------------------------------------------------------------ allowed noise (quantize_pvt.c) */
/* quantize_pvt.c:554 athAdjust */
static float ath_adjust(const lp_config *c, float a, float x, float athFloor, float ATHfixpoint)
{
    float const o = 90.30873362f;
    float const p = (ATHfixpoint < 1.f) ? 94.82444863f : ATHfixpoint;
    float u = FAST_LOG10_X_D(c, x, 10.0f);
    float const v = a * a;
    float w = 0.0f;
    u -= athFloor;
    if (v > 1E-20f) w = 1.f + FAST_LOG10_X_D(c, v, 10.0f / o);
    if (w < 0) w = 0.f;
    u *= w;
    u += athFloor + o - p;
    return powf(10.f, 0.1f * u);
}

/* quantize_pvt.c:589 calc_xmin */
static int calc_xmin(lp_encoder *e, const lp_ratio *ratio, lp_granule *gi, float *pxmin)
{
    const lp_config *cfg = &e->cfg;
    int sfb, gsfb, j = 0, ath_over = 0, k, max_nonzero;
    const float *xr = gi->xr;
    for (gsfb = 0; gsfb < gi->psy_lmax; gsfb++) {
        float en0, xmin, rh1, rh2, rh3;
        int width, l;
        xmin = ath_adjust(cfg, e->ath_adjust_factor, cfg->ath_l[gsfb], cfg->ath_floor, cfg->athfixpoint);
        xmin *= cfg->longfact[gsfb];
        width = gi->width[gsfb];
        rh1 = xmin / width;
        rh2 = DBL_EPSILON;
        en0 = 0.0;
        for (l = 0; l < width; ++l) {
            float const xa = xr[j++];
            float const x2 = xa * xa;
            en0 += x2;
            rh2 += (x2 < rh1) ? x2 : rh1;
        }
        if (en0 > xmin) ath_over++;
        if (en0 < xmin) rh3 = en0;
        else if (rh2 < xmin) rh3 = xmin;
        else rh3 = rh2;
        xmin = rh3;
        {
            float const en = ratio->en.l[gsfb];
            if (en > 1e-12f) {
                float x = en0 * ratio->thm.l[gsfb] / en;
                x *= cfg->longfact[gsfb];
                if (xmin < x) xmin = x;
            }
        }
        xmin = (xmin > DBL_EPSILON) ? xmin : DBL_EPSILON;
        gi->energy_above_cutoff[gsfb] = (en0 > xmin + 1e-14f) ? 1 : 0;
        *pxmin++ = xmin;
    }
    max_nonzero = 0;
    for (k = 575; k > 0; --k)
        if (fabs(xr[k]) > 1e-12f) { max_nonzero = k; break; }
    if (gi->block_type != LP_SHORT) max_nonzero |= 1;
    else { max_nonzero /= 6; max_nonzero *= 6; max_nonzero += 5; }
    if (cfg->sfb21_extra == 0 && cfg->samplerate < 44000) {
        int limit;
        if (gi->block_type != LP_SHORT) limit = cfg->sfb_l[cfg->samplerate <= 8000 ? 17 : 21] - 1;
        else limit = 3 * cfg->sfb_s[cfg->samplerate <= 8000 ? 9 : 12] - 1;
        if (max_nonzero > limit) max_nonzero = limit;
    }
    gi->max_nonzero_coeff = max_nonzero;
    for (sfb = gi->sfb_smin; gsfb < gi->psymax; sfb++, gsfb += 3) {
        int width, b, l;
        float tmpATH;
        tmpATH = ath_adjust(cfg, e->ath_adjust_factor, cfg->ath_s[sfb], cfg->ath_floor, cfg->athfixpoint);
        tmpATH *= cfg->shortfact[sfb];
        width = gi->width[gsfb];
        for (b = 0; b < 3; b++) {
            float en0 = 0.0, xmin = tmpATH, rh1, rh2, rh3;
            rh1 = tmpATH / width;
            rh2 = DBL_EPSILON;
            for (l = 0; l < width; ++l) {
                float const xa = xr[j++];
                float const x2 = xa * xa;
                en0 += x2;
                rh2 += (x2 < rh1) ? x2 : rh1;
            }
            if (en0 > tmpATH) ath_over++;
            if (en0 < tmpATH) rh3 = en0;
            else if (rh2 < tmpATH) rh3 = tmpATH;
            else rh3 = rh2;
            xmin = rh3;
            {
                float const en = ratio->en.s[sfb][b];
                if (en > 1e-12f) {
                    float x = en0 * ratio->thm.s[sfb][b] / en;
                    x *= cfg->shortfact[sfb];
                    if (xmin < x) xmin = x;
                }
            }
            xmin = (xmin > DBL_EPSILON) ? xmin : DBL_EPSILON;
            gi->energy_above_cutoff[gsfb + b] = (en0 > xmin + 1e-14f) ? 1 : 0;
            *pxmin++ = xmin;
        }
        if (cfg->use_temporal) {
            if (pxmin[-3] > pxmin[-3 + 1]) pxmin[-3 + 1] += (pxmin[-3] - pxmin[-3 + 1]) * cfg->decay;
            if (pxmin[-3 + 1] > pxmin[-3 + 2]) pxmin[-3 + 2] += (pxmin[-3 + 1] - pxmin[-3 + 2]) * cfg->decay;
        }
    }
    return ath_over;
}

/* quantize_pvt.c:750 calc_noise_core_c */
static float noise_core(const lp_config *c, const lp_granule *gi, int *startline, int l, float step)
{
    float noise = 0;
    int j = *startline;
    const int *ix = gi->l3_enc;
    if (j > gi->count1) {
        while (l--) {
            float t;
            t = gi->xr[j]; j++; noise += t * t;
            t = gi->xr[j]; j++; noise += t * t;
        }
    }
    else if (j > gi->big_values) {
        float ix01[2];
        ix01[0] = 0; ix01[1] = step;
        while (l--) {
            float t;
            t = fabs(gi->xr[j]) - ix01[ix[j]]; j++; noise += t * t;
            t = fabs(gi->xr[j]) - ix01[ix[j]]; j++; noise += t * t;
        }
    }
    else {
        while (l--) {
            float t;
            t = fabs(gi->xr[j]) - c->pow43[ix[j]] * step; j++; noise += t * t;
            t = fabs(gi->xr[j]) - c->pow43[ix[j]] * step; j++; noise += t * t;
        }
    }
    *startline = j;
    return noise;
}

/* quantize_pvt.c:815 calc_noise */
static int calc_noise(const lp_config *c, const lp_granule *gi, const float *l3_xmin, float *distort,
                      noise_result *res, noise_cache *prev)
{
    int sfb, l, over = 0, j = 0;
    float over_noise_db = 0, tot_noise_db = 0, max_noise = -20.0;
    const int *scalefac = gi->scalefac;
    res->over_SSD = 0;
    for (sfb = 0; sfb < gi->psymax; sfb++) {
        int const s = gi->global_gain - (((*scalefac++) + (gi->preflag ? lp_pretab[sfb] : 0)) << (gi->scalefac_scale + 1))
            - gi->subblock_gain[gi->window[sfb]] * 8;
        float const r_l3_xmin = 1.f / *l3_xmin++;
        float distort_ = 0.0f, noise = 0.0f;
        if (prev && (prev->step[sfb] == s)) {
            j += gi->width[sfb];
            distort_ = r_l3_xmin * prev->noise[sfb];
            noise = prev->noise_log[sfb];
        }
        else {
            float const step = c->pow20[s + LP_QMAX2];
            l = gi->width[sfb] >> 1;
            if ((j + gi->width[sfb]) > gi->max_nonzero_coeff) {
                int usefullsize = gi->max_nonzero_coeff - j + 1;
                if (usefullsize > 0) l = usefullsize >> 1;
                else l = 0;
            }
            noise = noise_core(c, gi, &j, l, step);
            if (prev) { prev->step[sfb] = s; prev->noise[sfb] = noise; }
            distort_ = r_l3_xmin * noise;
            noise = FAST_LOG10_D(c, (distort_ > 1E-20f ? distort_ : 1E-20f));
            if (prev) prev->noise_log[sfb] = noise;
        }
        *distort++ = distort_;
        if (prev) prev->global_gain = gi->global_gain;
        tot_noise_db += noise;
        if (noise > 0.0) {
            int tmp = (int) (noise * 10 + .5);
            if (tmp < 1) tmp = 1;
            res->over_SSD += tmp * tmp;
            over++;
            over_noise_db += noise;
        }
        max_noise = max_noise > noise ? max_noise : noise;
    }
    res->over_count = over;
    res->tot_noise = tot_noise_db;
    res->over_noise = over_noise_db;
    res->max_noise = max_noise;
    return over;
}

/* ------------------------------------------------------------------ quantiser + bit counting (takehiro.c) */
/* takehiro.c:113 quantize_lines_xrpow_01 */
static void quant_lines_01(unsigned l, float istep, const float *xr, int *ix)
{
    float const compareval0 = (1.0f - 0.4054f) / istep;
    unsigned i;
    for (i = 0; i < l; i += 2) {
        ix[i + 0] = (compareval0 > xr[i + 0]) ? 0 : 1;
        ix[i + 1] = (compareval0 > xr[i + 1]) ? 0 : 1;
    }
}
/* takehiro.c:144 quantize_lines_xrpow (TAKEHIRO_IEEE754_HACK): double add of 2^23, float store, table
 * lookup on the integer part, second double add, float store, subtract the magic integer. */
static void quant_lines(const lp_config *c, unsigned l, float istep, const float *xp, int *pi)
{
    unsigned i;
    l = (l >> 1) << 1;
    for (i = 0; i < l; i++) {
        union { float f; int i; } fi;
        double x0 = istep * xp[i];
        x0 += 8388608.0;
        fi.f = x0;
        fi.f = x0 + c->adj43asm[fi.i - 0x4b000000];
        pi[i] = fi.i - 0x4b000000;
    }
}

/* takehiro.c:281 quantize_xrpow: only the scalefactor bands whose step changed are re-quantised */
static void quantize_xrpow(const lp_config *c, const float *xp, int *pi, float istep, const lp_granule *gi,
                           const noise_cache *prev)
{
    int sfb, sfbmax, j = 0, prev_data_use, accumulate = 0, accumulate01 = 0;
    int *iData = pi, *acc_iData = pi;
    const float *acc_xp = xp;
    prev_data_use = (prev && (gi->global_gain == prev->global_gain));
    sfbmax = (gi->block_type == LP_SHORT) ? 38 : 21;
    for (sfb = 0; sfb <= sfbmax; sfb++) {
        int step = -1;
        if (prev_data_use || gi->block_type == LP_NORM) {
            step = gi->global_gain - ((gi->scalefac[sfb] + (gi->preflag ? lp_pretab[sfb] : 0)) << (gi->scalefac_scale + 1))
                - gi->subblock_gain[gi->window[sfb]] * 8;
        }
        if (prev_data_use && (prev->step[sfb] == step)) {
            if (accumulate) { quant_lines(c, accumulate, istep, acc_xp, acc_iData); accumulate = 0; }
            if (accumulate01) { quant_lines_01(accumulate01, istep, acc_xp, acc_iData); accumulate01 = 0; }
        }
        else {
            int l = gi->width[sfb];
            int probe = sfb;
            if ((j + gi->width[sfb]) > gi->max_nonzero_coeff) {
                int usefullsize = gi->max_nonzero_coeff - j + 1;
                memset(&pi[gi->max_nonzero_coeff], 0, sizeof(int) * (576 - gi->max_nonzero_coeff));
                l = usefullsize;
                if (l < 0) l = 0;
                sfb = sfbmax + 1;
                probe = sfb;
            }
            if (!accumulate && !accumulate01) { acc_iData = iData; acc_xp = xp; }
            /* the reference indexes prev->step[] with the already-bumped sfb (takehiro.c:372): for a
             * long block that is entry 22 (never written, 0); for a short block entry 39 aliases
             * noise[0] reinterpreted as int - reproduced through the union below */
            {
                int use01 = 0;
                if (prev && prev->sfb_count1 > 0 && probe >= prev->sfb_count1) {
                    int pstep;
                    if (probe < 39) pstep = prev->step[probe];
                    else { union { float f; int i; } u; u.f = prev->noise[0]; pstep = u.i; }
                    if (pstep > 0 && step >= pstep) use01 = 1;
                }
                if (use01) {
                    if (accumulate) {
                        quant_lines(c, accumulate, istep, acc_xp, acc_iData);
                        accumulate = 0; acc_iData = iData; acc_xp = xp;
                    }
                    accumulate01 += l;
                }
                else {
                    if (accumulate01) {
                        quant_lines_01(accumulate01, istep, acc_xp, acc_iData);
                        accumulate01 = 0; acc_iData = iData; acc_xp = xp;
                    }
                    accumulate += l;
                }
            }
            if (l <= 0) {
                if (accumulate01) { quant_lines_01(accumulate01, istep, acc_xp, acc_iData); accumulate01 = 0; }
                if (accumulate) { quant_lines(c, accumulate, istep, acc_xp, acc_iData); accumulate = 0; }
                break;
            }
        }
        if (sfb <= sfbmax) {
            iData += gi->width[sfb];
            xp += gi->width[sfb];
            j += gi->width[sfb];
        }
    }
    if (accumulate) quant_lines(c, accumulate, istep, acc_xp, acc_iData);
    if (accumulate01) quant_lines_01(accumulate01, istep, acc_xp, acc_iData);
}

/* takehiro.c:618 choose_table_nonMMX and the count_bit_* helpers (:449-:573) */
static int choose_table(const int *ix, const int *end, int *_s)
{
    static const int huf_tbl_noESC[15] = { 1, 2, 5, 7, 7, 10, 10, 13, 13, 13, 13, 13, 13, 13, 13 };
    unsigned *s = (unsigned *) _s;
    unsigned max = 0;
    const int *p;
    int choice, choice2;
    for (p = ix; p < end; p++) if ((unsigned) *p > max) max = *p;
    if (max == 0) return 0;
    if (max == 1) {
        unsigned sum = 0;
        const uint8_t *h = hlen_of(1);
        for (p = ix; p < end; p += 2) sum += h[p[0] + p[0] + p[1]];
        *s += sum;
        return 1;
    }
    if (max <= 3) {
        int t1 = huf_tbl_noESC[max - 1];
        unsigned xlen = LGT_HUFF_XLEN[t1], sum = 0, sum2;
        const uint32_t *table = (t1 == 2) ? LGT_TABLE23 : LGT_TABLE56;
        for (p = ix; p < end; p += 2) sum += table[p[0] * xlen + p[1]];
        sum2 = sum & 0xffffu;
        sum >>= 16u;
        if (sum > sum2) { sum = sum2; t1++; }
        *s += sum;
        return t1;
    }
    if (max <= 15) {
        int t1 = huf_tbl_noESC[max - 1], t;
        unsigned sum1 = 0, sum2 = 0, sum3 = 0, xlen = LGT_HUFF_XLEN[t1];
        const uint8_t *h1 = hlen_of(t1), *h2 = hlen_of(t1 + 1), *h3 = hlen_of(t1 + 2);
        for (p = ix; p < end; p += 2) {
            unsigned x = p[0] * xlen + p[1];
            sum1 += h1[x]; sum2 += h2[x]; sum3 += h3[x];
        }
        t = t1;
        if (sum1 > sum2) { sum1 = sum2; t++; }
        if (sum1 > sum3) { sum1 = sum3; t = t1 + 2; }
        *s += sum1;
        return t;
    }
    if (max > LP_IXMAX) { *s = LP_LARGE_BITS; return -1; }
    max -= 15u;
    for (choice2 = 24; choice2 < 32; choice2++) if (LGT_HUFF_LINMAX[choice2] >= max) break;
    for (choice = choice2 - 8; choice < 24; choice++) if (LGT_HUFF_LINMAX[choice] >= max) break;
    {
        unsigned const linbits = LGT_HUFF_XLEN[choice] * 65536u + LGT_HUFF_XLEN[choice2];
        unsigned sum = 0, sum2;
        for (p = ix; p < end; p += 2) {
            unsigned x = p[0], y = p[1];
            if (x >= 15u) { x = 15u; sum += linbits; }
            if (y >= 15u) { y = 15u; sum += linbits; }
            sum += LGT_LARGETBL[(x << 4) + y];
        }
        sum2 = sum & 0xffffu;
        sum >>= 16u;
        if (sum > sum2) { sum = sum2; choice = choice2; }
        *s += sum;
        return choice;
    }
}

static void best_huffman_divide(const lp_config *c, lp_granule *gi);

/* takehiro.c:654 noquant_count_bits */
static int noquant_count_bits(const lp_config *c, lp_granule *gi, noise_cache *prev)
{
    const uint8_t *t32l = hlen_of(32), *t33l = hlen_of(33);
    int bits = 0, i, a1, a2;
    const int *ix = gi->l3_enc;
    i = ((gi->max_nonzero_coeff + 2) >> 1) << 1;
    if (i > 576) i = 576;
    if (prev) prev->sfb_count1 = 0;
    for (; i > 1; i -= 2) if (ix[i - 1] | ix[i - 2]) break;
    gi->count1 = i;
    a1 = a2 = 0;
    for (; i > 3; i -= 4) {
        int x4 = ix[i - 4], x3 = ix[i - 3], x2 = ix[i - 2], x1 = ix[i - 1], p;
        if ((unsigned) (x4 | x3 | x2 | x1) > 1) break;
        p = ((x4 * 2 + x3) * 2 + x2) * 2 + x1;
        a1 += t32l[p];
        a2 += t33l[p];
    }
    bits = a1;
    gi->count1table_select = 0;
    if (a1 > a2) { bits = a2; gi->count1table_select = 1; }
    gi->count1bits = bits;
    gi->big_values = i;
    if (i == 0) return bits;
    if (gi->block_type == LP_SHORT) {
        a1 = 3 * c->sfb_s[3];
        if (a1 > gi->big_values) a1 = gi->big_values;
        a2 = gi->big_values;
    }
    else if (gi->block_type == LP_NORM) {
        a1 = gi->region0_count = c->bv_scf[i - 2];
        a2 = gi->region1_count = c->bv_scf[i - 1];
        a2 = c->sfb_l[a1 + a2 + 2];
        a1 = c->sfb_l[a1 + 1];
        if (a2 < i) gi->table_select[2] = choose_table(ix + a2, ix + i, &bits);
    }
    else {
        gi->region0_count = 7;
        gi->region1_count = LP_SBMAX_L - 1 - 7 - 1;
        a1 = c->sfb_l[7 + 1];
        a2 = i;
        if (a1 > a2) a1 = a2;
    }
    a1 = a1 < i ? a1 : i;
    a2 = a2 < i ? a2 : i;
    if (0 < a1) gi->table_select[0] = choose_table(ix, ix + a1, &bits);
    if (a1 < a2) gi->table_select[1] = choose_table(ix + a1, ix + a2, &bits);
    if (c->use_best_huffman == 2) {
        gi->part2_3_length = bits;
        best_huffman_divide(c, gi);
        bits = gi->part2_3_length;
    }
    if (prev && gi->block_type == LP_NORM) {
        int sfb = 0;
        while (c->sfb_l[sfb] < gi->big_values) sfb++;
        prev->sfb_count1 = sfb;
    }
    return bits;
}

/* QntStateVar_t.pseudohalf (util.h:318): per scalefactor band, whether the "half step" of substep shaping (quality 0-2)
 * is on.  One array per encoder in the reference; init_xrpow resets every band a gr.ch uses before it is read, so a
 * per-thread array is equivalent. */
static __thread int pseudohalf[LP_SFBMAX];

/* takehiro.c:767 count_bits */
static int count_bits(const lp_config *c, const float *xr, lp_granule *gi, noise_cache *prev)
{
    float const w = (LP_IXMAX) / c->ipow20[gi->global_gain];
    if (gi->xrpow_max > w) return LP_LARGE_BITS;
    quantize_xrpow(c, xr, gi->l3_enc, c->ipow20[gi->global_gain], gi, prev);
    if (c->substep_shaping & 2) {
        /* bands in their half step drop the values that only just rounded up to 1 (takehiro.c:781-797) */
        int sfb, j = 0;
        int const gain = gi->global_gain + gi->scalefac_scale;
        const float roundfac = 0.634521682242439 / c->ipow20[gain];
        for (sfb = 0; sfb < gi->sfbmax; sfb++) {
            int const width = gi->width[sfb];
            if (!pseudohalf[sfb]) j += width;
            else {
                int k;
                for (k = j, j += width; k < j; ++k) gi->l3_enc[k] = (xr[k] >= roundfac) ? gi->l3_enc[k] : 0;
            }
        }
    }
    return noquant_count_bits(c, gi, prev);
}

/* takehiro.c:809 recalc_divide_init / :847 recalc_divide_sub / :884 best_huffman_divide */
static void recalc_divide_init(const lp_config *c, const lp_granule *gi, const int *ix, int r01_bits[], int r01_div[],
                               int r0_tbl[], int r1_tbl[])
{
    int r0, r1, bigv = gi->big_values, r0t, r1t, bits;
    for (r0 = 0; r0 <= 7 + 15; r0++) r01_bits[r0] = LP_LARGE_BITS;
    for (r0 = 0; r0 < 16; r0++) {
        int const a1 = c->sfb_l[r0 + 1];
        int r0bits;
        if (a1 >= bigv) break;
        r0bits = 0;
        r0t = choose_table(ix, ix + a1, &r0bits);
        for (r1 = 0; r1 < 8; r1++) {
            int const a2 = c->sfb_l[r0 + r1 + 2];
            if (a2 >= bigv) break;
            bits = r0bits;
            r1t = choose_table(ix + a1, ix + a2, &bits);
            if (r01_bits[r0 + r1] > bits) {
                r01_bits[r0 + r1] = bits;
                r01_div[r0 + r1] = r0;
                r0_tbl[r0 + r1] = r0t;
                r1_tbl[r0 + r1] = r1t;
            }
        }
    }
}
static void recalc_divide_sub(const lp_config *c, const lp_granule *gi2, lp_granule *gi, const int *ix,
                              const int r01_bits[], const int r01_div[], const int r0_tbl[], const int r1_tbl[])
{
    int bits, r2, a2, bigv = gi2->big_values, r2t;
    for (r2 = 2; r2 < LP_SBMAX_L + 1; r2++) {
        a2 = c->sfb_l[r2];
        if (a2 >= bigv) break;
        bits = r01_bits[r2 - 2] + gi2->count1bits;
        if (gi->part2_3_length <= bits) break;
        r2t = choose_table(ix + a2, ix + bigv, &bits);
        if (gi->part2_3_length <= bits) continue;
        memcpy(gi, gi2, sizeof(lp_granule));
        gi->part2_3_length = bits;
        gi->region0_count = r01_div[r2 - 2];
        gi->region1_count = r2 - 2 - r01_div[r2 - 2];
        gi->table_select[0] = r0_tbl[r2 - 2];
        gi->table_select[1] = r1_tbl[r2 - 2];
        gi->table_select[2] = r2t;
    }
}
static void best_huffman_divide(const lp_config *c, lp_granule *gi)
{
    const uint8_t *t32l = hlen_of(32), *t33l = hlen_of(33);
    int i, a1, a2;
    static __thread lp_granule gi2;
    const int *ix = gi->l3_enc;
    int r01_bits[7 + 15 + 1], r01_div[7 + 15 + 1], r0_tbl[7 + 15 + 1], r1_tbl[7 + 15 + 1];
    if (gi->block_type == LP_SHORT && c->mode_gr == 1) return;        /* takehiro.c:899: not for short blocks of MPEG-2 */
    memcpy(&gi2, gi, sizeof(lp_granule));
    if (gi->block_type == LP_NORM) {
        recalc_divide_init(c, gi, ix, r01_bits, r01_div, r0_tbl, r1_tbl);
        recalc_divide_sub(c, &gi2, gi, ix, r01_bits, r01_div, r0_tbl, r1_tbl);
    }
    i = gi2.big_values;
    if (i == 0 || (unsigned) (ix[i - 2] | ix[i - 1]) > 1) return;
    i = gi->count1 + 2;
    if (i > 576) return;
    memcpy(&gi2, gi, sizeof(lp_granule));
    gi2.count1 = i;
    a1 = a2 = 0;
    for (; i > gi2.big_values; i -= 4) {
        int const p = ((ix[i - 4] * 2 + ix[i - 3]) * 2 + ix[i - 2]) * 2 + ix[i - 1];
        a1 += t32l[p];
        a2 += t33l[p];
    }
    gi2.big_values = i;
    gi2.count1table_select = 0;
    if (a1 > a2) { a1 = a2; gi2.count1table_select = 1; }
    gi2.count1bits = a1;
    if (gi2.block_type == LP_NORM) recalc_divide_sub(c, &gi2, gi, ix, r01_bits, r01_div, r0_tbl, r1_tbl);
    else {
        gi2.part2_3_length = a1;
        a1 = c->sfb_l[7 + 1];
        if (a1 > i) a1 = i;
        if (a1 > 0) gi2.table_select[0] = choose_table(ix, ix + a1, &gi2.part2_3_length);
        if (i > a1) gi2.table_select[1] = choose_table(ix + a1, ix + i, &gi2.part2_3_length);
        if (gi->part2_3_length > gi2.part2_3_length) memcpy(gi, &gi2, sizeof(lp_granule));
    }
}

/* takehiro.c:1135 mpeg1_scale_bitcount */
static const int slen1_n[16] = { 1, 1, 1, 1, 8, 2, 2, 2, 4, 4, 4, 8, 8, 8, 16, 16 };
static const int slen2_n[16] = { 1, 2, 4, 8, 1, 2, 4, 8, 2, 4, 8, 2, 4, 8, 4, 8 };
static const int slen1_tab[16] = { 0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4 };
static const int slen2_tab[16] = { 0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3 };
/* takehiro.c:1218 mpeg2_scale_bitcount: MPEG-2/2.5 code the scalefactors in four partitions whose sizes come from
 * nr_of_sfb_block (tables.c) and whose widths slen[] are the bit lengths of the partitions' maxima.  When a maximum is out
 * of range nothing is written (part2_length keeps its value) and 1 comes back. */
static const int nr_of_sfb_block[6][3][4] = {
    { {6, 5, 5, 5}, {9, 9, 9, 9}, {6, 9, 9, 9} }, { {6, 5, 7, 3}, {9, 9, 12, 6}, {6, 9, 12, 6} }, { {11, 10, 0, 0}, {18, 18, 0, 0}, {15, 18, 0, 0} },
    { {7, 7, 7, 0}, {12, 12, 12, 0}, {6, 15, 12, 0} }, { {6, 6, 6, 3}, {12, 9, 9, 6}, {6, 12, 9, 6} }, { {8, 8, 5, 0}, {15, 12, 9, 0}, {6, 18, 9, 0} } };
static int scale_bitcount_lsf(lp_granule *gi)
{
    static const int max_range[6][4] = { {15, 15, 7, 7}, {15, 15, 7, 0}, {7, 3, 0, 0}, {15, 31, 31, 0}, {7, 7, 7, 0}, {3, 3, 0, 0} };
    static const int log2tab[16] = { 0, 1, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4 };
    int const table_number = gi->preflag ? 2 : 0, row = (gi->block_type == LP_SHORT) ? 1 : 0;
    const int *ptab = nr_of_sfb_block[table_number][row];
    const int *scalefac = gi->scalefac;
    int max_sfac[4] = { 0, 0, 0, 0 }, partition, sfb = 0, i, window, over = 0;
    for (partition = 0; partition < 4; partition++) {
        if (row) {
            for (i = 0; i < ptab[partition] / 3; i++, sfb++)
                for (window = 0; window < 3; window++)
                    if (scalefac[sfb * 3 + window] > max_sfac[partition]) max_sfac[partition] = scalefac[sfb * 3 + window];
        }
        else for (i = 0; i < ptab[partition]; i++, sfb++) if (scalefac[sfb] > max_sfac[partition]) max_sfac[partition] = scalefac[sfb];
    }
    for (partition = 0; partition < 4; partition++) if (max_sfac[partition] > max_range[table_number][partition]) over++;
    if (!over) {
        gi->sfb_partition_table = ptab;
        for (partition = 0; partition < 4; partition++) gi->slen[partition] = log2tab[max_sfac[partition]];
        if (table_number == 0) gi->scalefac_compress = (((gi->slen[0] * 5) + gi->slen[1]) << 4) + (gi->slen[2] << 2) + gi->slen[3];
        else gi->scalefac_compress = 500 + (gi->slen[0] * 3) + gi->slen[1];
        gi->part2_length = 0;
        for (partition = 0; partition < 4; partition++) gi->part2_length += gi->slen[partition] * ptab[partition];
    }
    return over;
}

static int scale_bitcount(const lp_config *c, lp_granule *gi)
{
    static const int scale_short[16] = { 0, 18, 36, 54, 54, 36, 54, 72, 54, 72, 90, 72, 90, 108, 108, 126 };
    static const int scale_mixed[16] = { 0, 18, 36, 54, 51, 35, 53, 71, 52, 70, 88, 69, 87, 105, 104, 122 };
    static const int scale_long[16] = { 0, 10, 20, 30, 33, 21, 31, 41, 32, 42, 52, 43, 53, 63, 64, 74 };
    int k, sfb, max_slen1 = 0, max_slen2 = 0;
    const int *tab;
    int *scalefac = gi->scalefac;
    if (c->mode_gr == 1) return scale_bitcount_lsf(gi);               /* takehiro.c:1318 */
    if (gi->block_type == LP_SHORT) {
        tab = scale_short;
        if (gi->mixed_block_flag) tab = scale_mixed;
    }
    else {
        tab = scale_long;
        if (!gi->preflag) {
            for (sfb = 11; sfb < LP_SBPSY_L; sfb++) if (scalefac[sfb] < lp_pretab[sfb]) break;
            if (sfb == LP_SBPSY_L) {
                gi->preflag = 1;
                for (sfb = 11; sfb < LP_SBPSY_L; sfb++) scalefac[sfb] -= lp_pretab[sfb];
            }
        }
    }
    for (sfb = 0; sfb < gi->sfbdivide; sfb++) if (max_slen1 < scalefac[sfb]) max_slen1 = scalefac[sfb];
    for (; sfb < gi->sfbmax; sfb++) if (max_slen2 < scalefac[sfb]) max_slen2 = scalefac[sfb];
    gi->part2_length = LP_LARGE_BITS;
    for (k = 0; k < 16; k++)
        if (max_slen1 < slen1_n[k] && max_slen2 < slen2_n[k] && gi->part2_length > tab[k]) {
            gi->part2_length = tab[k];
            gi->scalefac_compress = k;
        }
    return gi->part2_length == LP_LARGE_BITS;
}

/* takehiro.c:964 scfsi_calc */
static void scfsi_calc(lp_encoder *e, int ch)
{
    unsigned i;
    int s1, s2, c1, c2, sfb;
    lp_granule *gi = &e->tt[1][ch];
    const lp_granule *g0 = &e->tt[0][ch];
    for (i = 0; i < 4; i++) {
        for (sfb = LGT_SCFSI_BAND[i]; sfb < LGT_SCFSI_BAND[i + 1]; sfb++)
            if (g0->scalefac[sfb] != gi->scalefac[sfb] && gi->scalefac[sfb] >= 0) break;
        if (sfb == LGT_SCFSI_BAND[i + 1]) {
            for (sfb = LGT_SCFSI_BAND[i]; sfb < LGT_SCFSI_BAND[i + 1]; sfb++) gi->scalefac[sfb] = -1;
            e->scfsi[ch][i] = 1;
        }
    }
    s1 = c1 = 0;
    for (sfb = 0; sfb < 11; sfb++) {
        if (gi->scalefac[sfb] == -1) continue;
        c1++;
        if (s1 < gi->scalefac[sfb]) s1 = gi->scalefac[sfb];
    }
    s2 = c2 = 0;
    for (; sfb < LP_SBPSY_L; sfb++) {
        if (gi->scalefac[sfb] == -1) continue;
        c2++;
        if (s2 < gi->scalefac[sfb]) s2 = gi->scalefac[sfb];
    }
    for (i = 0; i < 16; i++)
        if (s1 < slen1_n[i] && s2 < slen2_n[i]) {
            int const c = slen1_tab[i] * c1 + slen2_tab[i] * c2;
            if (gi->part2_length > c) { gi->part2_length = c; gi->scalefac_compress = (int) i; }
        }
}

/* takehiro.c:1021 best_scalefac_store */
static void best_scalefac_store(lp_encoder *e, int gr, int ch)
{
    lp_granule *gi = &e->tt[gr][ch];
    int sfb, i, j, l, recalc = 0;
    j = 0;
    for (sfb = 0; sfb < gi->sfbmax; sfb++) {
        int const width = gi->width[sfb];
        for (l = j, j += width; l < j; ++l) if (gi->l3_enc[l] != 0) break;
        if (l == j) gi->scalefac[sfb] = recalc = -2;
    }
    if (!gi->scalefac_scale && !gi->preflag) {
        int s = 0;
        for (sfb = 0; sfb < gi->sfbmax; sfb++) if (gi->scalefac[sfb] > 0) s |= gi->scalefac[sfb];
        if (!(s & 1) && s != 0) {
            for (sfb = 0; sfb < gi->sfbmax; sfb++) if (gi->scalefac[sfb] > 0) gi->scalefac[sfb] >>= 1;
            gi->scalefac_scale = recalc = 1;
        }
    }
    if (!gi->preflag && gi->block_type != LP_SHORT && e->cfg.mode_gr == 2) {
        for (sfb = 11; sfb < LP_SBPSY_L; sfb++)
            if (gi->scalefac[sfb] < lp_pretab[sfb] && gi->scalefac[sfb] != -2) break;
        if (sfb == LP_SBPSY_L) {
            for (sfb = 11; sfb < LP_SBPSY_L; sfb++) if (gi->scalefac[sfb] > 0) gi->scalefac[sfb] -= lp_pretab[sfb];
            gi->preflag = recalc = 1;
        }
    }
    for (i = 0; i < 4; i++) e->scfsi[ch][i] = 0;
    if (e->cfg.mode_gr == 2 && gr == 1 && e->tt[0][ch].block_type != LP_SHORT && e->tt[1][ch].block_type != LP_SHORT) {
        scfsi_calc(e, ch);
        recalc = 0;
    }
    for (sfb = 0; sfb < gi->sfbmax; sfb++) if (gi->scalefac[sfb] == -2) gi->scalefac[sfb] = 0;
    if (recalc) (void) scale_bitcount(&e->cfg, gi);
}

/* ------------------------------------------------------------------ noise shaping loop (quantize.c) */
/* quantize.c:226 init_outer_loop */
static void init_outer_loop(const lp_config *c, lp_granule *gi)
{
    int sfb, j;
    gi->part2_3_length = 0; gi->big_values = 0; gi->count1 = 0; gi->global_gain = 210; gi->scalefac_compress = 0;
    gi->table_select[0] = gi->table_select[1] = gi->table_select[2] = 0;
    gi->subblock_gain[0] = gi->subblock_gain[1] = gi->subblock_gain[2] = gi->subblock_gain[3] = 0;
    gi->region0_count = 0; gi->region1_count = 0; gi->preflag = 0; gi->scalefac_scale = 0;
    gi->count1table_select = 0; gi->part2_length = 0;
    gi->sfb_partition_table = nr_of_sfb_block[0][0];                  /* quantize.c:331-335 */
    gi->slen[0] = gi->slen[1] = gi->slen[2] = gi->slen[3] = 0;
    gi->sfb_lmax = LP_SBPSY_L; gi->sfb_smin = LP_SBPSY_S;
    gi->psy_lmax = c->sfb21_extra ? LP_SBMAX_L : LP_SBPSY_L;
    if (c->samplerate <= 8000) { gi->sfb_lmax = 17; gi->sfb_smin = 9; gi->psy_lmax = 17; }       /* quantize.c:252-256 */
    gi->psymax = gi->psy_lmax;
    gi->sfbmax = gi->sfb_lmax;
    gi->sfbdivide = 11;
    for (sfb = 0; sfb < LP_SBMAX_L; sfb++) {
        gi->width[sfb] = c->sfb_l[sfb + 1] - c->sfb_l[sfb];
        gi->window[sfb] = 3;
    }
    if (gi->block_type == LP_SHORT) {
        float ixwork[576];
        float *ix;
        gi->sfb_smin = 0;
        gi->sfb_lmax = 0;
        if (gi->mixed_block_flag) { gi->sfb_smin = 3; gi->sfb_lmax = c->mode_gr * 2 + 4; }
        gi->psymax = gi->sfb_lmax + 3 * ((c->sfb21_extra ? LP_SBMAX_S : LP_SBPSY_S) - gi->sfb_smin);
        gi->sfbmax = gi->sfb_lmax + 3 * (LP_SBPSY_S - gi->sfb_smin);
        if (c->samplerate <= 8000) gi->psymax = gi->sfbmax = gi->sfb_lmax + 3 * (9 - gi->sfb_smin);     /* quantize.c:284-289 */
        gi->sfbdivide = gi->sfbmax - 18;
        gi->psy_lmax = gi->sfb_lmax;
        ix = &gi->xr[c->sfb_l[gi->sfb_lmax]];
        memcpy(ixwork, gi->xr, 576 * sizeof(float));
        for (sfb = gi->sfb_smin; sfb < LP_SBMAX_S; sfb++) {
            int const start = c->sfb_s[sfb], end = c->sfb_s[sfb + 1];
            int window, l;
            for (window = 0; window < 3; window++)
                for (l = start; l < end; l++) *ix++ = ixwork[3 * l + window];
        }
        j = gi->sfb_lmax;
        for (sfb = gi->sfb_smin; sfb < LP_SBMAX_S; sfb++) {
            gi->width[j] = gi->width[j + 1] = gi->width[j + 2] = c->sfb_s[sfb + 1] - c->sfb_s[sfb];
            gi->window[j] = 0; gi->window[j + 1] = 1; gi->window[j + 2] = 2;
            j += 3;
        }
    }
    gi->count1bits = 0;
    gi->max_nonzero_coeff = 575;
    memset(gi->scalefac, 0, sizeof gi->scalefac);
}

/* quantize.c:110 init_xrpow with :72 init_xrpow_core_c */
static int init_xrpow(const lp_config *c, lp_granule *gi, float xrpow[576])
{
    float sum = 0;
    int i;
    int const upper = gi->max_nonzero_coeff;
    gi->xrpow_max = 0;
    memset(&(xrpow[upper]), 0, (576 - upper) * sizeof(xrpow[0]));
    for (i = 0; i <= upper; ++i) {
        float tmp = fabs(gi->xr[i]);
        sum += tmp;
        xrpow[i] = sqrt(tmp * sqrt(tmp));       /* double sqrt of the float, float product, double sqrt */
        if (xrpow[i] > gi->xrpow_max) gi->xrpow_max = xrpow[i];
    }
    if (sum > (float) 1E-20) {
        int const j = (c->substep_shaping & 2) ? 1 : 0;
        for (i = 0; i < gi->psymax; i++) pseudohalf[i] = j;
        return 1;
    }
    memset(&gi->l3_enc[0], 0, sizeof(int) * 576);
    return 0;
}

/* quantize.c:367 bin_search_StepSize */
static int bin_search_stepsize(lp_encoder *e, lp_granule *gi, int desired_rate, int ch, const float xrpow[576])
{
    int nBits, CurrentStep = e->current_step[ch], flag_GoneOver = 0;
    int const start = e->old_value[ch];
    int Direction = 0;      /* 0 none, 1 up, 2 down */
    gi->global_gain = start;
    desired_rate -= gi->part2_length;
    for (;;) {
        int step;
        nBits = count_bits(&e->cfg, xrpow, gi, 0);
        if (CurrentStep == 1 || nBits == desired_rate) break;
        if (nBits > desired_rate) {
            if (Direction == 2) flag_GoneOver = 1;
            if (flag_GoneOver) CurrentStep /= 2;
            Direction = 1;
            step = CurrentStep;
        }
        else {
            if (Direction == 1) flag_GoneOver = 1;
            if (flag_GoneOver) CurrentStep /= 2;
            Direction = 2;
            step = -CurrentStep;
        }
        gi->global_gain += step;
        if (gi->global_gain < 0) { gi->global_gain = 0; flag_GoneOver = 1; }
        if (gi->global_gain > 255) { gi->global_gain = 255; flag_GoneOver = 1; }
    }
    while (nBits > desired_rate && gi->global_gain < 255) {
        gi->global_gain++;
        nBits = count_bits(&e->cfg, xrpow, gi, 0);
    }
    e->current_step[ch] = (start - gi->global_gain >= 4) ? 4 : 2;
    e->old_value[ch] = gi->global_gain;
    gi->part2_3_length = nBits;
    return nBits;
}

/* quantize.c:540 loop_break */
static int loop_break(const lp_granule *gi)
{
    int sfb;
    for (sfb = 0; sfb < gi->sfbmax; sfb++)
        if (gi->scalefac[sfb] + gi->subblock_gain[gi->window[sfb]] == 0) return 0;
    return 1;
}

/* quantize.c:585 quant_compare - only comparison mode 9 is reachable (every bitrate preset selects it,
 * presets.c:241) */
static int quant_compare(const noise_result *best, const noise_result *calc)
{
    int better;
    if (best->over_count > 0) {
        better = calc->over_SSD <= best->over_SSD;
        if (calc->over_SSD == best->over_SSD) better = calc->bits < best->bits;
    }
    else {
        better = ((calc->max_noise < 0) && ((calc->max_noise * 10 + calc->bits) <= (best->max_noise * 10 + best->bits)));
    }
    if (best->over_count == 0) better = better && calc->bits < best->bits;
    return better;
}

/* quantize.c:720 amp_scalefac_bands */
static void amp_scalefac_bands(const lp_config *c, lp_granule *gi, const float *distort, float xrpow[576], int bRefine)
{
    int j, sfb, noise_shaping_amp;
    float ifqstep34, trigger;
    if (gi->scalefac_scale == 0) ifqstep34 = 1.29683955465100964055;
    else ifqstep34 = 1.68179283050742922612;
    trigger = 0;
    for (sfb = 0; sfb < gi->sfbmax; sfb++) if (trigger < distort[sfb]) trigger = distort[sfb];
    noise_shaping_amp = c->noise_shaping_amp;
    if (noise_shaping_amp == 3) noise_shaping_amp = (bRefine == 1) ? 2 : 1;
    switch (noise_shaping_amp) {
    case 2: break;
    case 1:
        if (trigger > 1.0) trigger = pow(trigger, .5);
        else trigger *= .95;
        break;
    case 0:
    default:
        if (trigger > 1.0) trigger = 1.0;
        else trigger *= .95;
        break;
    }
    j = 0;
    for (sfb = 0; sfb < gi->sfbmax; sfb++) {
        int const width = gi->width[sfb];
        int l;
        j += width;
        if (distort[sfb] < trigger) continue;
        if (c->substep_shaping & 2) {
            pseudohalf[sfb] = !pseudohalf[sfb];
            if (!pseudohalf[sfb] && c->noise_shaping_amp == 2) return;
        }
        gi->scalefac[sfb]++;
        for (l = -width; l < 0; l++) {
            xrpow[j + l] *= ifqstep34;
            if (xrpow[j + l] > gi->xrpow_max) gi->xrpow_max = xrpow[j + l];
        }
        if (c->noise_shaping_amp == 2) return;
    }
}

/* quantize.c:808 inc_scalefac_scale */
static void inc_scalefac_scale(lp_granule *gi, float xrpow[576])
{
    int l, j, sfb;
    const float ifqstep34 = 1.29683955465100964055;
    j = 0;
    for (sfb = 0; sfb < gi->sfbmax; sfb++) {
        int const width = gi->width[sfb];
        int s = gi->scalefac[sfb];
        if (gi->preflag) s += lp_pretab[sfb];
        j += width;
        if (s & 1) {
            s++;
            for (l = -width; l < 0; l++) {
                xrpow[j + l] *= ifqstep34;
                if (xrpow[j + l] > gi->xrpow_max) gi->xrpow_max = xrpow[j + l];
            }
        }
        gi->scalefac[sfb] = s >> 1;
    }
    gi->preflag = 0;
    gi->scalefac_scale = 1;
}

/* quantize.c:847 inc_subblock_gain */
static int inc_subblock_gain(const lp_config *c, lp_granule *gi, float xrpow[576])
{
    int sfb, window;
    int *scalefac = gi->scalefac;
    for (sfb = 0; sfb < gi->sfb_lmax; sfb++) if (scalefac[sfb] >= 16) return 1;
    for (window = 0; window < 3; window++) {
        int s1, s2, l, j;
        s1 = s2 = 0;
        for (sfb = gi->sfb_lmax + window; sfb < gi->sfbdivide; sfb += 3) if (s1 < scalefac[sfb]) s1 = scalefac[sfb];
        for (; sfb < gi->sfbmax; sfb += 3) if (s2 < scalefac[sfb]) s2 = scalefac[sfb];
        if (s1 < 16 && s2 < 8) continue;
        if (gi->subblock_gain[window] >= 7) return 1;
        gi->subblock_gain[window]++;
        j = c->sfb_l[gi->sfb_lmax];
        for (sfb = gi->sfb_lmax + window; sfb < gi->sfbmax; sfb += 3) {
            float amp;
            int const width = gi->width[sfb];
            int s = scalefac[sfb];
            s = s - (4 >> gi->scalefac_scale);
            if (s >= 0) { scalefac[sfb] = s; j += width * 3; continue; }
            scalefac[sfb] = 0;
            amp = c->ipow20[210 + (s << (gi->scalefac_scale + 1))];
            j += width * (window + 1);
            for (l = -width; l < 0; l++) {
                xrpow[j + l] *= amp;
                if (xrpow[j + l] > gi->xrpow_max) gi->xrpow_max = xrpow[j + l];
            }
            j += width * (3 - window - 1);
        }
        {
            float const amp = c->ipow20[202];
            j += gi->width[sfb] * (window + 1);
            for (l = -gi->width[sfb]; l < 0; l++) {
                xrpow[j + l] *= amp;
                if (xrpow[j + l] > gi->xrpow_max) gi->xrpow_max = xrpow[j + l];
            }
        }
    }
    return 0;
}

/* quantize.c:940 balance_noise */
static int balance_noise(const lp_config *c, lp_granule *gi, const float *distort, float xrpow[576], int bRefine)
{
    int status;
    amp_scalefac_bands(c, gi, distort, xrpow, bRefine);
    status = loop_break(gi);
    if (status) return 0;
    status = scale_bitcount(c, gi);
    if (!status) return 1;
    if (c->noise_shaping > 1) {
        memset(pseudohalf, 0, sizeof pseudohalf);
        if (!gi->scalefac_scale) { inc_scalefac_scale(gi, xrpow); status = 0; }
        else if (gi->block_type == LP_SHORT && c->subblock_gain > 0)
            status = inc_subblock_gain(c, gi, xrpow) || loop_break(gi);
    }
    if (!status) status = scale_bitcount(c, gi);
    return !status;
}

/* VBR-old switches sfb21_extra off for the tries near the bit limit (quantize.c:1265-1268): -1 = the configuration's value */
static __thread int sfb21_now = -1;

/* quantize.c:1010 outer_loop */
static int outer_loop(lp_encoder *e, lp_granule *gi, const float *l3_xmin, float xrpow[576], int ch, int targ_bits)
{
    const lp_config *cfg = &e->cfg;
    static __thread lp_granule gi_w;
    float save_xrpow[576], distort[LP_SFBMAX];
    noise_result best_noise_info;
    int huff_bits, better, age;
    noise_cache prev_noise;
    int best_part2_3_length = 9999999, bEndOfSearch = 0, bRefine = 0, best_ggain_pass1 = 0;

    (void) bin_search_stepsize(e, gi, targ_bits, ch, xrpow);
    if (!cfg->noise_shaping) return 100;
    memset(&prev_noise, 0, sizeof prev_noise);
    (void) calc_noise(cfg, gi, l3_xmin, distort, &best_noise_info, &prev_noise);
    if (lp_trace_on(e)) fprintf(stderr, "T ch%d start targ %d gg %d p23 %d over %d tot %.9g ovn %.9g max %.9g\n", ch, targ_bits, gi->global_gain, gi->part2_3_length,
                                best_noise_info.over_count, best_noise_info.tot_noise, best_noise_info.over_noise, best_noise_info.max_noise);
    best_noise_info.bits = gi->part2_3_length;
    gi_w = *gi;
    age = 0;
    memcpy(save_xrpow, xrpow, sizeof(float) * 576);
    while (!bEndOfSearch) {
        do {
            noise_result noise_info;
            int const search_limit = (cfg->substep_shaping & 2) ? 20 : 3;
            int maxggain = 255;
            if (sfb21_now >= 0 ? sfb21_now : cfg->sfb21_extra) {
                if (distort[gi_w.sfbmax] > 1.0) break;
                if (gi_w.block_type == LP_SHORT && (distort[gi_w.sfbmax + 1] > 1.0 || distort[gi_w.sfbmax + 2] > 1.0)) break;
            }
            if (balance_noise(cfg, &gi_w, distort, xrpow, bRefine) == 0) break;
            if (gi_w.scalefac_scale) maxggain = 254;
            huff_bits = targ_bits - gi_w.part2_length;
            if (huff_bits <= 0) break;
            while ((gi_w.part2_3_length = count_bits(cfg, xrpow, &gi_w, &prev_noise)) > huff_bits && gi_w.global_gain <= maxggain)
                gi_w.global_gain++;
            if (gi_w.global_gain > maxggain) break;
            if (best_noise_info.over_count == 0) {
                while ((gi_w.part2_3_length = count_bits(cfg, xrpow, &gi_w, &prev_noise)) > best_part2_3_length
                       && gi_w.global_gain <= maxggain)
                    gi_w.global_gain++;
                if (gi_w.global_gain > maxggain) break;
            }
            (void) calc_noise(cfg, &gi_w, l3_xmin, distort, &noise_info, &prev_noise);
            noise_info.bits = gi_w.part2_3_length;
            better = quant_compare(&best_noise_info, &noise_info);
            if (lp_trace_on(e)) fprintf(stderr, "T ch%d round gg %d p23 %d over %d tot %.9g ovn %.9g max %.9g ssd %d better %d\n", ch, gi_w.global_gain, gi_w.part2_3_length,
                                        noise_info.over_count, noise_info.tot_noise, noise_info.over_noise, noise_info.max_noise, noise_info.over_SSD, better);
            if (better) {
                best_part2_3_length = gi->part2_3_length;
                best_noise_info = noise_info;
                *gi = gi_w;
                age = 0;
                memcpy(save_xrpow, xrpow, sizeof(float) * 576);
            }
            else {
                if (cfg->full_outer_loop == 0) {
                    if (++age > search_limit && best_noise_info.over_count == 0) break;
                    if ((cfg->noise_shaping_amp == 3) && bRefine && age > 30) break;
                    if ((cfg->noise_shaping_amp == 3) && bRefine && (gi_w.global_gain - best_ggain_pass1) > 15) break;
                }
            }
        } while ((gi_w.global_gain + gi_w.scalefac_scale) < 255);
        if (cfg->noise_shaping_amp == 3) {
            if (!bRefine) {
                gi_w = *gi;
                memcpy(xrpow, save_xrpow, sizeof(float) * 576);
                age = 0;
                best_ggain_pass1 = gi_w.global_gain;
                bRefine = 1;
            }
            else bEndOfSearch = 1;
        }
        else bEndOfSearch = 1;
    }
    if (cfg->vbr == 2) memcpy(xrpow, save_xrpow, sizeof(float) * 576);     /* quantize.c:1188: restore for the next try */
    return best_noise_info.over_count;
}

/* quantize.c:1988 CBR_iteration_loop (+ :48 ms_convert, :1213 iteration_finish_one) */
void lp_cbr_iteration_loop(lp_encoder *e, float pe[2][2], const float ms_ener_ratio[2], lp_ratio ratio[2][2])
{
    const lp_config *cfg = &e->cfg;
    float l3_xmin[LP_SFBMAX], xrpow[576];
    int targ_bits[2], mean_bits, max_bits, gr, ch, i;
    (void) resv_frame_begin(e, &mean_bits);
    for (gr = 0; gr < cfg->mode_gr; gr++) {
        max_bits = on_pe(e, pe, targ_bits, mean_bits, gr, gr);
        if (e->mode_ext == 2) {
            for (i = 0; i < 576; ++i) {
                float l = e->tt[gr][0].xr[i], r = e->tt[gr][1].xr[i];
                e->tt[gr][0].xr[i] = (l + r) * (float) (SQRT2_D * 0.5);
                e->tt[gr][1].xr[i] = (l - r) * (float) (SQRT2_D * 0.5);
            }
            reduce_side(targ_bits, ms_ener_ratio[gr], mean_bits, max_bits);
        }
        for (ch = 0; ch < cfg->channels; ch++) {
            lp_granule *gi = &e->tt[gr][ch];
            float masking_lower_db;
            if (gi->block_type != LP_SHORT) masking_lower_db = cfg->mask_adjust - 0;
            else masking_lower_db = cfg->mask_adjust_short - 0;
            e->masking_lower = pow(10.0, masking_lower_db * 0.1);
            init_outer_loop(cfg, gi);
            if (init_xrpow(cfg, gi, xrpow)) {
                (void) calc_xmin(e, &ratio[gr][ch], gi, l3_xmin);
                (void) outer_loop(e, gi, l3_xmin, xrpow, ch, targ_bits[ch]);
            }
            best_scalefac_store(e, gr, ch);
            if (cfg->use_best_huffman == 1) best_huffman_divide(cfg, gi);
            e->resv_size -= gi->part2_3_length + gi->part2_length;       /* reservoir.c:226 ResvAdjust */
        }
    }
    resv_frame_end(e, mean_bits);
}

/* quantize.c:1768 calc_target_bits */
static void calc_target_bits(lp_encoder *e, float pe[2][2], const float ms_ener_ratio[2], int targ_bits[2][2],
                             int *analog_silence_bits, int *max_frame_bits)
{
    const lp_config *cfg = &e->cfg;
    float res_factor;
    int gr, ch, totbits, mean_bits;
    int const framesize = 576 * cfg->mode_gr;
    e->bitrate_index = cfg->vbr_max_bitrate_index;
    *max_frame_bits = resv_frame_begin(e, &mean_bits);
    e->bitrate_index = 1;
    mean_bits = lp_getframebits(e) - cfg->sideinfo_len * 8;
    *analog_silence_bits = mean_bits / (cfg->mode_gr * cfg->channels);
    mean_bits = cfg->vbr_mean_kbps * framesize * 1000;
    if (cfg->substep_shaping & 1) mean_bits *= 1.09;
    mean_bits /= cfg->samplerate;
    mean_bits -= cfg->sideinfo_len * 8;
    mean_bits /= (cfg->mode_gr * cfg->channels);
    res_factor = .93 + .07 * (11.0 - cfg->compression_ratio) / (11.0 - 5.5);     /* double expression stored in a float */
    if (res_factor < .90) res_factor = .90;
    if (res_factor > 1.00) res_factor = 1.00;
    for (gr = 0; gr < cfg->mode_gr; gr++) {
        int sum = 0;
        for (ch = 0; ch < cfg->channels; ch++) {
            targ_bits[gr][ch] = res_factor * mean_bits;
            if (pe[gr][ch] > 700) {
                int add_bits = (pe[gr][ch] - 700) / 1.4;
                targ_bits[gr][ch] = res_factor * mean_bits;
                if (e->tt[gr][ch].block_type == LP_SHORT) {
                    if (add_bits < mean_bits / 2) add_bits = mean_bits / 2;
                }
                if (add_bits > mean_bits * 3 / 2) add_bits = mean_bits * 3 / 2;
                else if (add_bits < 0) add_bits = 0;
                targ_bits[gr][ch] += add_bits;
            }
            if (targ_bits[gr][ch] > LP_MAX_BITS_PER_CHANNEL) targ_bits[gr][ch] = LP_MAX_BITS_PER_CHANNEL;
            sum += targ_bits[gr][ch];
        }
        if (sum > LP_MAX_BITS_PER_GRANULE)
            for (ch = 0; ch < cfg->channels; ++ch) {
                targ_bits[gr][ch] *= LP_MAX_BITS_PER_GRANULE;
                targ_bits[gr][ch] /= sum;
            }
    }
    if (e->mode_ext == 2)
        for (gr = 0; gr < cfg->mode_gr; gr++)
            reduce_side(targ_bits[gr], ms_ener_ratio[gr], mean_bits * cfg->channels, LP_MAX_BITS_PER_GRANULE);
    totbits = 0;
    for (gr = 0; gr < cfg->mode_gr; gr++)
        for (ch = 0; ch < cfg->channels; ch++) {
            if (targ_bits[gr][ch] > LP_MAX_BITS_PER_CHANNEL) targ_bits[gr][ch] = LP_MAX_BITS_PER_CHANNEL;
            totbits += targ_bits[gr][ch];
        }
    if (totbits > *max_frame_bits && totbits > 0)
        for (gr = 0; gr < cfg->mode_gr; gr++)
            for (ch = 0; ch < cfg->channels; ch++) {
                targ_bits[gr][ch] *= *max_frame_bits;
                targ_bits[gr][ch] /= totbits;
            }
}

/* quantize.c:1900 ABR_iteration_loop */
void lp_abr_iteration_loop(lp_encoder *e, float pe[2][2], const float ms_ener_ratio[2], lp_ratio ratio[2][2])
{
    const lp_config *cfg = &e->cfg;
    float l3_xmin[LP_SFBMAX], xrpow[576];
    int targ_bits[2][2], mean_bits = 0, max_frame_bits, analog_silence_bits, gr, ch, i;
    calc_target_bits(e, pe, ms_ener_ratio, targ_bits, &analog_silence_bits, &max_frame_bits);
    for (gr = 0; gr < cfg->mode_gr; gr++) {
        if (e->mode_ext == 2)
            for (i = 0; i < 576; ++i) {
                float l = e->tt[gr][0].xr[i], r = e->tt[gr][1].xr[i];
                e->tt[gr][0].xr[i] = (l + r) * (float) (SQRT2_D * 0.5);
                e->tt[gr][1].xr[i] = (l - r) * (float) (SQRT2_D * 0.5);
            }
        for (ch = 0; ch < cfg->channels; ch++) {
            lp_granule *gi = &e->tt[gr][ch];
            float masking_lower_db;
            if (gi->block_type != LP_SHORT) masking_lower_db = cfg->mask_adjust - 0;
            else masking_lower_db = cfg->mask_adjust_short - 0;
            e->masking_lower = pow(10.0, masking_lower_db * 0.1);
            init_outer_loop(cfg, gi);
            if (init_xrpow(cfg, gi, xrpow)) {
                int const ath_over = calc_xmin(e, &ratio[gr][ch], gi, l3_xmin);
                if (0 == ath_over) targ_bits[gr][ch] = analog_silence_bits;
                (void) outer_loop(e, gi, l3_xmin, xrpow, ch, targ_bits[gr][ch]);
            }
            best_scalefac_store(e, gr, ch);
            if (cfg->use_best_huffman == 1) best_huffman_divide(cfg, gi);
            e->resv_size -= gi->part2_3_length + gi->part2_length;
        }
    }
    /* the smallest frame that brings the reservoir back to a non-negative size */
    for (e->bitrate_index = cfg->vbr_min_bitrate_index; e->bitrate_index <= cfg->vbr_max_bitrate_index; e->bitrate_index++)
        if (resv_frame_begin(e, &mean_bits) >= 0) break;
    resv_frame_end(e, mean_bits);
}

/* quantize.c:160 psfb21_analogsilence (vbr_rh only, the tail of init_outer_loop :342): from the top of the spectrum down, lines
 * of the six sub-bands of sfb21 (sfb12 of every window for short blocks) below the ATH become zero, until one is not */
static void psfb21_analog_silence(const lp_encoder *e, lp_granule *gi)
{
    const lp_config *cfg = &e->cfg;
    float *xr = gi->xr;
    int gsfb, j, block, stop = 0;
    if (gi->block_type != LP_SHORT) {
        for (gsfb = 6 - 1; gsfb >= 0 && !stop; gsfb--) {
            int const start = cfg->psfb21[gsfb], end = cfg->psfb21[gsfb + 1];
            float ath21 = ath_adjust(cfg, e->ath_adjust_factor, cfg->ath_psfb21[gsfb], cfg->ath_floor, 0);
            if (cfg->longfact[21] > 1e-12f) ath21 *= cfg->longfact[21];
            for (j = end - 1; j >= start; j--) {
                if (fabs(xr[j]) < ath21) xr[j] = 0;
                else { stop = 1; break; }
            }
        }
        return;
    }
    for (block = 0; block < 3; block++) {
        stop = 0;
        for (gsfb = 6 - 1; gsfb >= 0 && !stop; gsfb--) {
            int const start = cfg->sfb_s[12] * 3 + (cfg->sfb_s[13] - cfg->sfb_s[12]) * block + (cfg->psfb12[gsfb] - cfg->psfb12[0]);
            int const end = start + (cfg->psfb12[gsfb + 1] - cfg->psfb12[gsfb]);
            float ath12 = ath_adjust(cfg, e->ath_adjust_factor, cfg->ath_psfb12[gsfb], cfg->ath_floor, 0);
            if (cfg->shortfact[12] > 1e-12f) ath12 *= cfg->shortfact[12];
            for (j = end - 1; j >= start; j--) {
                if (fabs(xr[j]) < ath12) xr[j] = 0;
                else { stop = 1; break; }
            }
        }
    }
}

/* quantize.c:1246 VBR_encode_granule: bisection on the bit budget of one gr.ch - the smallest number of bits (within 32)
 * at which outer_loop leaves no band distorted */
static void vbr_old_encode_granule(lp_encoder *e, lp_granule *gi, const float *l3_xmin, float xrpow[576], int ch, int min_bits, int max_bits)
{
    static __thread lp_granule bst;
    float bst_xrpow[576];
    int const Max_bits = max_bits;
    int real_bits = max_bits + 1, this_bits = (max_bits + min_bits) / 2, dbits, over, found = 0;
    memset(bst.l3_enc, 0, sizeof bst.l3_enc);
    do {
        sfb21_now = (this_bits > Max_bits - 42) ? 0 : e->cfg.sfb21_extra;
        over = outer_loop(e, gi, l3_xmin, xrpow, ch, this_bits);
        if (over <= 0) {
            found = 1;
            real_bits = gi->part2_3_length;
            bst = *gi;
            memcpy(bst_xrpow, xrpow, sizeof(float) * 576);
            max_bits = real_bits - 32;
            dbits = max_bits - min_bits;
            this_bits = (max_bits + min_bits) / 2;
        }
        else {
            min_bits = this_bits + 32;
            dbits = max_bits - min_bits;
            this_bits = (max_bits + min_bits) / 2;
            if (found) {
                found = 2;
                *gi = bst;
                memcpy(xrpow, bst_xrpow, sizeof(float) * 576);
            }
        }
    } while (dbits > 12);
    sfb21_now = -1;
    if (found == 2) memcpy(gi->l3_enc, bst.l3_enc, sizeof(int) * 576);
}

/* quantize.c:1491 VBR_old_iteration_loop (+ :1370 VBR_old_prepare, :1326 get_framebits, :1455 bitpressure_strategy) */
void lp_vbr_old_iteration_loop(lp_encoder *e, float pe[2][2], const float ms_ener_ratio[2], lp_ratio ratio[2][2])
{
    const lp_config *cfg = &e->cfg;
    float l3_xmin[2][2][LP_SFBMAX], xrpow[576], adjust, masking_lower_db;
    int frameBits[16], min_bits[2][2], max_bits[2][2], bands[2][2], used_bits, bits = 0, mean_bits, gr, ch, i, sfb, analog_silence = 1, avg, mxb, dummy;
    /* prepare */
    e->bitrate_index = cfg->vbr_max_bitrate_index;
    avg = resv_frame_begin(e, &avg) / cfg->mode_gr;
    for (i = 1; i <= cfg->vbr_max_bitrate_index; i++) { e->bitrate_index = i; frameBits[i] = resv_frame_begin(e, &dummy); }
    for (gr = 0; gr < cfg->mode_gr; gr++) {
        mxb = on_pe(e, pe, max_bits[gr], avg, gr, 0);
        if (e->mode_ext == 2) {
            for (i = 0; i < 576; ++i) {
                float l = e->tt[gr][0].xr[i], r = e->tt[gr][1].xr[i];
                e->tt[gr][0].xr[i] = (l + r) * (float) (SQRT2_D * 0.5);
                e->tt[gr][1].xr[i] = (l - r) * (float) (SQRT2_D * 0.5);
            }
            reduce_side(max_bits[gr], ms_ener_ratio[gr], avg, mxb);
        }
        for (ch = 0; ch < cfg->channels; ++ch) {
            lp_granule *gi = &e->tt[gr][ch];
            if (gi->block_type != LP_SHORT) {
                adjust = 1.28 / (1 + exp(3.5 - pe[gr][ch] / 300.)) - 0.05;
                masking_lower_db = cfg->mask_adjust - adjust;
            }
            else {
                adjust = 2.56 / (1 + exp(3.5 - pe[gr][ch] / 300.)) - 0.14;
                masking_lower_db = cfg->mask_adjust_short - adjust;
            }
            e->masking_lower = pow(10.0, masking_lower_db * 0.1);
            init_outer_loop(cfg, gi);
            psfb21_analog_silence(e, gi);
            bands[gr][ch] = calc_xmin(e, &ratio[gr][ch], gi, l3_xmin[gr][ch]);
            if (bands[gr][ch]) analog_silence = 0;
            min_bits[gr][ch] = 126;
            bits += max_bits[gr][ch];
        }
    }
    for (gr = 0; gr < cfg->mode_gr; gr++)
        for (ch = 0; ch < cfg->channels; ch++) {
            if (bits > frameBits[cfg->vbr_max_bitrate_index] && bits > 0) {
                max_bits[gr][ch] *= frameBits[cfg->vbr_max_bitrate_index];
                max_bits[gr][ch] /= bits;
            }
            if (min_bits[gr][ch] > max_bits[gr][ch]) min_bits[gr][ch] = max_bits[gr][ch];
        }
    for (;;) {
        used_bits = 0;
        for (gr = 0; gr < cfg->mode_gr; gr++)
            for (ch = 0; ch < cfg->channels; ch++) {
                lp_granule *gi = &e->tt[gr][ch];
                int const ret = init_xrpow(cfg, gi, xrpow);
                if (ret == 0 || max_bits[gr][ch] == 0) continue;
                vbr_old_encode_granule(e, gi, l3_xmin[gr][ch], xrpow, ch, min_bits[gr][ch], max_bits[gr][ch]);
                used_bits += gi->part2_3_length + gi->part2_length;
            }
        e->bitrate_index = analog_silence ? 1 : cfg->vbr_min_bitrate_index;
        for (; e->bitrate_index < cfg->vbr_max_bitrate_index; e->bitrate_index++) if (used_bits <= frameBits[e->bitrate_index]) break;
        bits = resv_frame_begin(e, &mean_bits);
        if (used_bits <= bits) break;
        /* bitpressure_strategy: allow more noise towards the high bands and shrink the budgets */
        for (gr = 0; gr < cfg->mode_gr; gr++)
            for (ch = 0; ch < cfg->channels; ch++) {
                const lp_granule *gi = &e->tt[gr][ch];
                float *pxmin = l3_xmin[gr][ch];
                for (sfb = 0; sfb < gi->psy_lmax; sfb++) *pxmin++ *= 1. + .029 * sfb * sfb / LP_SBMAX_L / LP_SBMAX_L;
                if (gi->block_type == LP_SHORT)
                    for (sfb = gi->sfb_smin; sfb < LP_SBMAX_S; sfb++) {
                        *pxmin++ *= 1. + .029 * sfb * sfb / LP_SBMAX_S / LP_SBMAX_S;
                        *pxmin++ *= 1. + .029 * sfb * sfb / LP_SBMAX_S / LP_SBMAX_S;
                        *pxmin++ *= 1. + .029 * sfb * sfb / LP_SBMAX_S / LP_SBMAX_S;
                    }
                max_bits[gr][ch] = min_bits[gr][ch] > 0.9 * max_bits[gr][ch] ? min_bits[gr][ch] : 0.9 * max_bits[gr][ch];
            }
    }
    for (gr = 0; gr < cfg->mode_gr; gr++)
        for (ch = 0; ch < cfg->channels; ch++) {
            lp_granule *gi = &e->tt[gr][ch];
            best_scalefac_store(e, gr, ch);
            if (cfg->use_best_huffman == 1) best_huffman_divide(cfg, gi);
            e->resv_size -= gi->part2_3_length + gi->part2_length;
        }
    resv_frame_end(e, mean_bits);
}

/* ================================================================== VBR-new (vbr_mtrh / vbr_mt), vbrquantize.c
 * One search context per granule.channel: the band-wise scalefactor search works on xr34 = |xr|^(3/4). */
typedef struct {
    lp_granule *gi;
    const float *xr34;
    int  mingain_l, mingain_s[3];
    int  is_short, guess_only;
} vbr_ctx;

static const uint8_t vbr_range_short[LP_SBMAX_S * 3] = {
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 0, 0, 0 };
static const uint8_t vbr_range_long[LP_SBMAX_L] = { 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 0 };

/* vbrquantize.c:170 k_34_4 for n <= 4 values held in double: the 2^23 rounding trick with the adj43asm correction */
static void vbr_quant4(const lp_config *c, double x[4], int l3[4])
{
    int k;
    union { float f; int i; } fi[4];
    for (k = 0; k < 4; k++) { x[k] += 8388608.0; fi[k].f = x[k]; }
    for (k = 0; k < 4; k++) { fi[k].f = x[k] + c->adj43asm[fi[k].i - 0x4b000000]; l3[k] = fi[k].i - 0x4b000000; }
}

/* vbrquantize.c:218 calc_sfb_noise_x34: quantisation noise of one band at step sf */
static float vbr_band_noise(const lp_config *c, const float *xr, const float *xr34, unsigned bw, int sf)
{
    double x[4];
    int l3[4];
    float const sfpow = c->pow20[sf + LP_QMAX2], sfpow34 = c->ipow20[sf];
    float xfsf = 0;
    unsigned i = bw >> 2, k;
    unsigned const remaining = bw & 3u;
    while (i-- > 0) {
        for (k = 0; k < 4; k++) x[k] = sfpow34 * xr34[k];
        vbr_quant4(c, x, l3);
        for (k = 0; k < 4; k++) x[k] = fabsf(xr[k]) - sfpow * c->pow43[l3[k]];
        xfsf += (x[0] * x[0] + x[1] * x[1]) + (x[2] * x[2] + x[3] * x[3]);       /* double sum added to the float */
        xr += 4; xr34 += 4;
    }
    if (remaining) {
        x[0] = x[1] = x[2] = x[3] = 0;
        for (k = 0; k < remaining; k++) x[k] = sfpow34 * xr34[k];
        vbr_quant4(c, x, l3);
        x[0] = x[1] = x[2] = x[3] = 0;
        for (k = 0; k < remaining; k++) x[k] = fabsf(xr[k]) - sfpow * c->pow43[l3[k]];
        xfsf += (x[0] * x[0] + x[1] * x[1]) + (x[2] * x[2] + x[3] * x[3]);
    }
    return xfsf;
}

/* vbrquantize.c:148 find_lowest_scalefac: smallest step that keeps the band's largest line within IXMAX_VAL */
static int vbr_lowest_sf(const lp_config *c, float xr34)
{
    int sf_ok = 255, sf = 128, delsf = 64, i;
    float const ixmax_val = LP_IXMAX;
    for (i = 0; i < 8; ++i) {
        float const xfsf = c->ipow20[sf] * xr34;
        if (xfsf <= ixmax_val) { sf_ok = sf; sf -= delsf; }
        else sf += delsf;
        delsf >>= 1;
    }
    return sf_ok;
}

/* vbrquantize.c:278 tri_calc_sfb_noise_x34 without its memo table (the noise is a pure function of sf) */
static int vbr_too_noisy(const lp_config *c, const float *xr, const float *xr34, float l3_xmin, unsigned bw, int sf)
{
    if (l3_xmin < vbr_band_noise(c, xr, xr34, bw, sf)) return 1;
    if (sf < 255 && l3_xmin < vbr_band_noise(c, xr, xr34, bw, sf + 1)) return 1;
    if (sf > 0 && l3_xmin < vbr_band_noise(c, xr, xr34, bw, sf - 1)) return 1;
    return 0;
}

/* vbrquantize.c:347 find_scalefac_x34 / :324 guess_scalefac_x34 */
static int vbr_find_sf(const lp_config *c, const float *xr, const float *xr34, float l3_xmin, unsigned bw, int sf_min, int guess_only)
{
    int sf = 128, sf_ok = 255, delsf = 128, seen_good_one = 0, i;
    if (guess_only) {
        float const cc = 5.799142446;
        int const guess = 210 + (int) (cc * log10f(l3_xmin / bw) - .5f);
        if (guess < sf_min) return sf_min;
        if (guess >= 255) return 255;
        return guess;
    }
    for (i = 0; i < 8; ++i) {
        delsf >>= 1;
        if (sf <= sf_min) sf += delsf;
        else if (vbr_too_noisy(c, xr, xr34, l3_xmin, bw, sf)) sf -= delsf;
        else { sf_ok = sf; sf += delsf; seen_good_one = 1; }
        sf &= 255;                                  /* uint8_t arithmetic of the reference */
    }
    if (seen_good_one > 0) sf = sf_ok;
    if (sf <= sf_min) sf = sf_min;
    return sf;
}

/* vbrquantize.c:395 block_sf: per band the smallest usable step (vbrsfmin) and the step that just meets l3_xmin (vbrsf) */
static int vbr_block_sf(const lp_config *c, vbr_ctx *t, const float *l3_xmin, int vbrsf[LP_SFBMAX], int vbrsfmin[LP_SFBMAX])
{
    const lp_granule *gi = t->gi;
    unsigned const max_nonzero_coeff = (unsigned) gi->max_nonzero_coeff;
    int maxsf = 0, sfb = 0, m_o = -1;
    unsigned j = 0, i = 0;
    t->mingain_l = 0;
    t->mingain_s[0] = t->mingain_s[1] = t->mingain_s[2] = 0;
    while (j <= max_nonzero_coeff) {
        unsigned const w = (unsigned) gi->width[sfb], m = max_nonzero_coeff - j + 1;
        unsigned l = w, k;
        int m1, m2;
        float mx = 0;
        if (l > m) l = m;
        for (k = 0; k < l; k++) if (mx < t->xr34[j + k]) mx = t->xr34[j + k];      /* vec_max_c */
        m1 = vbr_lowest_sf(c, mx);
        vbrsfmin[sfb] = m1;
        if (t->mingain_l < m1) t->mingain_l = m1;
        if (t->mingain_s[i] < m1) t->mingain_s[i] = m1;
        if (++i > 2) i = 0;
        if (sfb < gi->psymax && w > 2) {
            if (gi->energy_above_cutoff[sfb]) {
                m2 = vbr_find_sf(c, &gi->xr[j], &t->xr34[j], l3_xmin[sfb], l, m1, t->guess_only);
                if (maxsf < m2) maxsf = m2;
                if (m_o < m2 && m2 < 255) m_o = m2;
            }
            else { m2 = 255; maxsf = 255; }
        }
        else {
            if (maxsf < m1) maxsf = m1;
            m2 = maxsf;
        }
        vbrsf[sfb] = m2;
        ++sfb;
        j += w;
    }
    for (; sfb < LP_SFBMAX; ++sfb) { vbrsf[sfb] = maxsf; vbrsfmin[sfb] = 0; }
    if (m_o > -1) {
        maxsf = m_o;
        for (sfb = 0; sfb < LP_SFBMAX; ++sfb) if (vbrsf[sfb] == 255) vbrsf[sfb] = m_o;
    }
    return maxsf;
}

/* vbrquantize.c:501 quantize_x34 */
static void vbr_quantize(const lp_config *c, const vbr_ctx *t)
{
    lp_granule *gi = t->gi;
    const float *xr34 = t->xr34;
    int const ifqstep = (gi->scalefac_scale == 0) ? 2 : 4;
    int *l3 = gi->l3_enc;
    unsigned j = 0, sfb = 0;
    unsigned const max_nonzero_coeff = (unsigned) gi->max_nonzero_coeff;
    while (j <= max_nonzero_coeff) {
        int const s = (gi->scalefac[sfb] + (gi->preflag ? lp_pretab[sfb] : 0)) * ifqstep + gi->subblock_gain[gi->window[sfb]] * 8;
        int const sfac = (gi->global_gain - s) & 255;                       /* (uint8_t) */
        float const sfpow34 = c->ipow20[sfac];
        unsigned const w = (unsigned) gi->width[sfb], m = max_nonzero_coeff - j + 1;
        unsigned n = (w <= m) ? w : m, k;
        j += w;
        ++sfb;
        while (n > 0) {
            double x[4] = { 0, 0, 0, 0 };
            int q[4];
            unsigned const take = n < 4 ? n : 4;
            for (k = 0; k < take; k++) x[k] = sfpow34 * xr34[k];
            vbr_quant4(c, x, q);
            for (k = 0; k < take; k++) l3[k] = q[k];
            l3 += take; xr34 += take; n -= take;
        }
    }
}

/* vbrquantize.c:596 set_subblock_gain */
static void vbr_set_subblock_gain(lp_granule *gi, const int mingain_s[3], int sf[])
{
    int const maxrange1 = 15, maxrange2 = 7;
    int const ifqstepShift = (gi->scalefac_scale == 0) ? 1 : 2;
    int *const sbg = gi->subblock_gain;
    unsigned const psymax = (unsigned) gi->psymax;
    unsigned psydiv = 18, sfb, i;
    int min_sbg = 7;
    if (psydiv > psymax) psydiv = psymax;
    for (i = 0; i < 3; ++i) {
        int maxsf1 = 0, maxsf2 = 0, minsf = 1000;
        for (sfb = i; sfb < psydiv; sfb += 3) {
            int const v = -sf[sfb];
            if (maxsf1 < v) maxsf1 = v;
            if (minsf > v) minsf = v;
        }
        for (; sfb < LP_SFBMAX; sfb += 3) {
            int const v = -sf[sfb];
            if (maxsf2 < v) maxsf2 = v;
            if (minsf > v) minsf = v;
        }
        {
            int const m1 = maxsf1 - (maxrange1 << ifqstepShift), m2 = maxsf2 - (maxrange2 << ifqstepShift);
            maxsf1 = m1 > m2 ? m1 : m2;
        }
        sbg[i] = (minsf > 0) ? (minsf >> 3) : 0;
        if (maxsf1 > 0) {
            int const m2 = (maxsf1 + 7) >> 3;
            if (sbg[i] < m2) sbg[i] = m2;
        }
        if (sbg[i] > 0 && mingain_s[i] > (gi->global_gain - sbg[i] * 8)) sbg[i] = (gi->global_gain - mingain_s[i]) >> 3;
        if (sbg[i] > 7) sbg[i] = 7;
        if (min_sbg > sbg[i]) min_sbg = sbg[i];
    }
    for (sfb = 0; sfb < LP_SFBMAX; sfb += 3) { sf[sfb + 0] += sbg[0] * 8; sf[sfb + 1] += sbg[1] * 8; sf[sfb + 2] += sbg[2] * 8; }
    if (min_sbg > 0) {
        for (i = 0; i < 3; ++i) sbg[i] -= min_sbg;
        gi->global_gain -= min_sbg * 8;
    }
}

/* vbrquantize.c:689 set_scalefacs */
static void vbr_set_scalefacs(lp_granule *gi, const int *vbrsfmin, int sf[], const uint8_t *max_range)
{
    int const ifqstep = (gi->scalefac_scale == 0) ? 2 : 4, ifqstepShift = (gi->scalefac_scale == 0) ? 1 : 2;
    int sfb;
    if (gi->preflag) for (sfb = 11; sfb < gi->sfbmax; ++sfb) sf[sfb] += lp_pretab[sfb] * ifqstep;
    for (sfb = 0; sfb < gi->sfbmax; ++sfb) {
        int const gain = gi->global_gain - (gi->subblock_gain[gi->window[sfb]] * 8) - ((gi->preflag ? lp_pretab[sfb] : 0) * ifqstep);
        if (sf[sfb] < 0) {
            int const m = gain - vbrsfmin[sfb];
            gi->scalefac[sfb] = (ifqstep - 1 - sf[sfb]) >> ifqstepShift;
            if (gi->scalefac[sfb] > max_range[sfb]) gi->scalefac[sfb] = max_range[sfb];
            if (gi->scalefac[sfb] > 0 && (gi->scalefac[sfb] << ifqstepShift) > m) gi->scalefac[sfb] = m >> ifqstepShift;
        }
        else gi->scalefac[sfb] = 0;
    }
    for (; sfb < LP_SFBMAX; ++sfb) gi->scalefac[sfb] = 0;
}

/* vbrquantize.c:770 short_block_constrain / :848 long_block_constrain: global_gain, scalefac_scale, preflag, subblock gains
 * and scalefactors that realise the wanted per-band steps as closely as the format allows */
static void vbr_alloc(const lp_config *c, const vbr_ctx *t, const int vbrsf[LP_SFBMAX], const int vbrsfmin[LP_SFBMAX], int vbrmax)
{
    lp_granule *gi = t->gi;
    int const maxminsfb = t->mingain_l, psymax = gi->psymax;
    int sf_temp[LP_SFBMAX], sfb, delta = 0, mover, v;
    if (t->is_short) {
        int maxover0 = 0, maxover1 = 0;
        for (sfb = 0; sfb < psymax; ++sfb) {
            int v0, v1;
            v = vbrmax - vbrsf[sfb];
            if (delta < v) delta = v;
            v0 = v - (4 * 14 + 2 * vbr_range_short[sfb]);
            v1 = v - (4 * 14 + 4 * vbr_range_short[sfb]);
            if (maxover0 < v0) maxover0 = v0;
            if (maxover1 < v1) maxover1 = v1;
        }
        if (c->noise_shaping == 2) mover = maxover0 < maxover1 ? maxover0 : maxover1;
        else mover = maxover0;
        if (delta > mover) delta = mover;
        vbrmax -= delta;
        maxover0 -= mover;
        maxover1 -= mover;
        if (maxover0 == 0) gi->scalefac_scale = 0;
        else if (maxover1 == 0) gi->scalefac_scale = 1;
        if (vbrmax < maxminsfb) vbrmax = maxminsfb;
        gi->global_gain = vbrmax;
        if (gi->global_gain < 0) gi->global_gain = 0;
        else if (gi->global_gain > 255) gi->global_gain = 255;
        for (sfb = 0; sfb < LP_SFBMAX; ++sfb) sf_temp[sfb] = vbrsf[sfb] - vbrmax;
        vbr_set_subblock_gain(gi, t->mingain_s, sf_temp);
        vbr_set_scalefacs(gi, vbrsfmin, sf_temp, vbr_range_short);
    }
    else {
        static const uint8_t range_long_lsf_pretab[LP_SBMAX_L] = { 7, 7, 7, 7, 7, 7, 3, 3, 3, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        const uint8_t *max_rangep = c->mode_gr == 2 ? vbr_range_long : range_long_lsf_pretab;      /* vbrquantize.c:861 */
        int maxover0 = 0, maxover1 = 0, maxover0p = 0, maxover1p = 0, vm0p = 1, vm1p = 1;
        for (sfb = 0; sfb < psymax; ++sfb) {
            int v0, v1, v0p, v1p;
            v = vbrmax - vbrsf[sfb];
            if (delta < v) delta = v;
            v0 = v - 2 * vbr_range_long[sfb];
            v1 = v - 4 * vbr_range_long[sfb];
            v0p = v - 2 * (max_rangep[sfb] + lp_pretab[sfb]);
            v1p = v - 4 * (max_rangep[sfb] + lp_pretab[sfb]);
            if (maxover0 < v0) maxover0 = v0;
            if (maxover1 < v1) maxover1 = v1;
            if (maxover0p < v0p) maxover0p = v0p;
            if (maxover1p < v1p) maxover1p = v1p;
        }
        if (vm0p == 1) {
            int gain = vbrmax - maxover0p;
            if (gain < maxminsfb) gain = maxminsfb;
            for (sfb = 0; sfb < psymax; ++sfb)
                if ((gain - vbrsfmin[sfb]) - 2 * lp_pretab[sfb] <= 0) { vm0p = 0; vm1p = 0; break; }
        }
        if (vm1p == 1) {
            int gain = vbrmax - maxover1p;
            if (gain < maxminsfb) gain = maxminsfb;
            for (sfb = 0; sfb < psymax; ++sfb)
                if ((gain - vbrsfmin[sfb]) - 4 * lp_pretab[sfb] <= 0) { vm1p = 0; break; }
        }
        if (vm0p == 0) maxover0p = maxover0;
        if (vm1p == 0) maxover1p = maxover1;
        if (c->noise_shaping != 2) { maxover1 = maxover0; maxover1p = maxover0p; }
        mover = maxover0 < maxover0p ? maxover0 : maxover0p;
        mover = mover < maxover1 ? mover : maxover1;
        mover = mover < maxover1p ? mover : maxover1p;
        if (delta > mover) delta = mover;
        vbrmax -= delta;
        if (vbrmax < maxminsfb) vbrmax = maxminsfb;
        maxover0 -= mover; maxover0p -= mover; maxover1 -= mover; maxover1p -= mover;
        if (maxover0 == 0) { gi->scalefac_scale = 0; gi->preflag = 0; max_rangep = vbr_range_long; }
        else if (maxover0p == 0) { gi->scalefac_scale = 0; gi->preflag = 1; }
        else if (maxover1 == 0) { gi->scalefac_scale = 1; gi->preflag = 0; max_rangep = vbr_range_long; }
        else if (maxover1p == 0) { gi->scalefac_scale = 1; gi->preflag = 1; }
        gi->global_gain = vbrmax;
        if (gi->global_gain < 0) gi->global_gain = 0;
        else if (gi->global_gain > 255) gi->global_gain = 255;
        for (sfb = 0; sfb < LP_SFBMAX; ++sfb) sf_temp[sfb] = vbrsf[sfb] - vbrmax;
        vbr_set_scalefacs(gi, vbrsfmin, sf_temp, max_rangep);
    }
}

/* vbrquantize.c:1000 quantizeAndCountBits */
static int vbr_quantize_and_count(const lp_config *c, const vbr_ctx *t)
{
    vbr_quantize(c, t);
    t->gi->part2_3_length = noquant_count_bits(c, t->gi, 0);
    return t->gi->part2_3_length;
}

/* vbrquantize.c:1141 tryThatOne (+ :1012 tryGlobalStepsize when add_part2 == 0) */
static int vbr_try(const lp_config *c, const vbr_ctx *t, const int sftemp[LP_SFBMAX], const int vbrsfmin[LP_SFBMAX], int vbrmax, int add_part2)
{
    float const xrpow_max = t->gi->xrpow_max;
    int nbits;
    vbr_alloc(c, t, sftemp, vbrsfmin, vbrmax);
    (void) scale_bitcount(c, t->gi);
    nbits = vbr_quantize_and_count(c, t);
    if (add_part2) nbits += t->gi->part2_length;
    t->gi->xrpow_max = xrpow_max;
    return nbits;
}

/* vbrquantize.c:1105 flattenDistribution */
static int vbr_flatten(const int sfwork[LP_SFBMAX], int sf_out[LP_SFBMAX], int dm, int k, int p)
{
    int i, x, sfmax = 0;
    for (i = 0; i < LP_SFBMAX; ++i) {
        if (dm > 0) {
            int const di = p - sfwork[i];
            x = sfwork[i] + (k * di) / dm;
            if (x < 0) x = 0;
            else if (x > 255) x = 255;
        }
        else x = sfwork[i];
        sf_out[i] = x;
        if (sfmax < x) sfmax = x;
    }
    return sfmax;
}

/* vbrquantize.c:1155 outOfBitsStrategy (+ :1041 searchGlobalStepsizeMax) */
static void vbr_out_of_bits(const lp_config *c, const vbr_ctx *t, const int sfwork[LP_SFBMAX], const int vbrsfmin[LP_SFBMAX], int target)
{
    int wrk[LP_SFBMAX], dm = 0, i, part, nbits;
    int const p = t->gi->global_gain;
    for (i = 0; i < LP_SFBMAX; ++i) { int const di = 255 - sfwork[i]; if (dm < di) dm = di; }      /* sfDepth */
    for (part = 0; part < 2; part++) {
        /* part 0: pull every band towards global_gain by k/dm; part 1: towards a level bi above it */
        int bi = part == 0 ? dm / 2 : (255 + p) / 2, bi_ok = -1, bu = part == 0 ? 0 : p, bo = part == 0 ? dm : 255;
        for (;;) {
            int const sfmax = part == 0 ? vbr_flatten(sfwork, wrk, dm, bi, p) : vbr_flatten(sfwork, wrk, dm, dm, bi);
            nbits = vbr_try(c, t, wrk, vbrsfmin, sfmax, 1);
            if (nbits <= target) { bi_ok = bi; bo = bi - 1; }
            else bu = bi + 1;
            if (bu <= bo) bi = (bu + bo) / 2;
            else break;
        }
        if (bi_ok >= 0) {
            if (bi != bi_ok) {
                int const sfmax = part == 0 ? vbr_flatten(sfwork, wrk, dm, bi_ok, p) : vbr_flatten(sfwork, wrk, dm, dm, bi_ok);
                (void) vbr_try(c, t, wrk, vbrsfmin, sfmax, 1);
            }
            return;
        }
    }
    {   /* searchGlobalStepsizeMax on the last tried distribution */
        int const gain = t->gi->global_gain;
        int curr = gain, gain_ok = 1024, l = gain, r = 512;
        while (l <= r) {
            int sft[LP_SFBMAX], vbrmax = 0, g;
            curr = (l + r) >> 1;
            for (i = 0; i < LP_SFBMAX; ++i) {
                g = wrk[i] + (curr - gain);
                if (g < vbrsfmin[i]) g = vbrsfmin[i];
                if (g > 255) g = 255;
                if (vbrmax < g) vbrmax = g;
                sft[i] = g;
            }
            nbits = vbr_try(c, t, sft, vbrsfmin, vbrmax, 0);
            if (nbits == 0 || (nbits + t->gi->part2_length) < target) { r = curr - 1; gain_ok = curr; }
            else { l = curr + 1; if (gain_ok == 1024) gain_ok = curr; }
        }
        if (gain_ok != curr) {
            int sft[LP_SFBMAX], vbrmax = 0, g;
            curr = gain_ok;
            for (i = 0; i < LP_SFBMAX; ++i) {
                g = wrk[i] + (curr - gain);
                if (g < vbrsfmin[i]) g = vbrsfmin[i];
                if (g > 255) g = 255;
                if (vbrmax < g) vbrmax = g;
                sft[i] = g;
            }
            (void) vbr_try(c, t, sft, vbrsfmin, vbrmax, 0);
        }
    }
}

/* vbrquantize.c:1232 reduce_bit_usage */
static int vbr_reduce_bits(lp_encoder *e, int gr, int ch)
{
    lp_granule *gi = &e->tt[gr][ch];
    best_scalefac_store(e, gr, ch);
    if (e->cfg.use_best_huffman == 1) best_huffman_divide(&e->cfg, gi);
    return gi->part2_3_length + gi->part2_length;
}

/* vbrquantize.c:1255 VBR_encode_frame */
static int vbr_encode_frame(lp_encoder *e, float xr34orig[2][2][576], float l3_xmin[2][2][LP_SFBMAX], int max_bits[2][2])
{
    const lp_config *c = &e->cfg;
    int sfwork_[2][2][LP_SFBMAX], vbrsfmin_[2][2][LP_SFBMAX];
    vbr_ctx that_[2][2];
    int const ngr = c->mode_gr, nch = c->channels;
    int max_nbits_ch[2][2] = { { 0, 0 }, { 0, 0 } }, max_nbits_gr[2] = { 0, 0 }, max_nbits_fr = 0;
    int use_nbits_ch[2][2], use_nbits_gr[2], use_nbits_fr, gr, ch, ok, sum_fr;
    for (gr = 0; gr < ngr; ++gr) {
        max_nbits_gr[gr] = 0;
        for (ch = 0; ch < nch; ++ch) {
            max_nbits_ch[gr][ch] = max_bits[gr][ch];
            use_nbits_ch[gr][ch] = 0;
            max_nbits_gr[gr] += max_bits[gr][ch];
            max_nbits_fr += max_bits[gr][ch];
            that_[gr][ch].gi = &e->tt[gr][ch];
            that_[gr][ch].xr34 = xr34orig[gr][ch];
            that_[gr][ch].is_short = e->tt[gr][ch].block_type == LP_SHORT;
            that_[gr][ch].guess_only = c->full_outer_loop < 0;
        }
    }
    for (gr = 0; gr < ngr; ++gr)
        for (ch = 0; ch < nch; ++ch)
            if (max_bits[gr][ch] > 0) {
                vbr_ctx *t = &that_[gr][ch];
                int const vbrmax = vbr_block_sf(c, t, l3_xmin[gr][ch], sfwork_[gr][ch], vbrsfmin_[gr][ch]);
                vbr_alloc(c, t, sfwork_[gr][ch], vbrsfmin_[gr][ch], vbrmax);
                (void) scale_bitcount(c, t->gi);
            }
    use_nbits_fr = 0;
    for (gr = 0; gr < ngr; ++gr) {
        use_nbits_gr[gr] = 0;
        for (ch = 0; ch < nch; ++ch) {
            if (max_bits[gr][ch] > 0) {
                memset(&e->tt[gr][ch].l3_enc[0], 0, sizeof e->tt[gr][ch].l3_enc);
                (void) vbr_quantize_and_count(c, &that_[gr][ch]);
            }
            use_nbits_ch[gr][ch] = vbr_reduce_bits(e, gr, ch);
            use_nbits_gr[gr] += use_nbits_ch[gr][ch];
        }
        use_nbits_fr += use_nbits_gr[gr];
    }
    if (use_nbits_fr <= max_nbits_fr) {
        ok = 1;
        for (gr = 0; gr < ngr; ++gr) {
            if (use_nbits_gr[gr] > LP_MAX_BITS_PER_GRANULE) ok = 0;
            for (ch = 0; ch < nch; ++ch) if (use_nbits_ch[gr][ch] > LP_MAX_BITS_PER_CHANNEL) ok = 0;
        }
        if (ok) return use_nbits_fr;
    }
    /* the frame does not fit: decide how many bits every granule.channel may use (vbrquantize.c:1370-1510) */
    ok = 1;
    sum_fr = 0;
    for (gr = 0; gr < ngr; ++gr) {
        max_nbits_gr[gr] = 0;
        for (ch = 0; ch < nch; ++ch) {
            max_nbits_ch[gr][ch] = use_nbits_ch[gr][ch] > LP_MAX_BITS_PER_CHANNEL ? LP_MAX_BITS_PER_CHANNEL : use_nbits_ch[gr][ch];
            max_nbits_gr[gr] += max_nbits_ch[gr][ch];
        }
        if (max_nbits_gr[gr] > LP_MAX_BITS_PER_GRANULE) {
            float f[2] = { 0.0f, 0.0f }, s = 0.0f;
            for (ch = 0; ch < nch; ++ch) {
                if (max_nbits_ch[gr][ch] > 0) { f[ch] = sqrt(sqrt(max_nbits_ch[gr][ch])); s += f[ch]; }
                else f[ch] = 0;
            }
            for (ch = 0; ch < nch; ++ch) max_nbits_ch[gr][ch] = (s > 0) ? LP_MAX_BITS_PER_GRANULE * f[ch] / s : 0;
            if (nch > 1) {
                if (max_nbits_ch[gr][0] > use_nbits_ch[gr][0] + 32) {
                    max_nbits_ch[gr][1] += max_nbits_ch[gr][0];
                    max_nbits_ch[gr][1] -= use_nbits_ch[gr][0] + 32;
                    max_nbits_ch[gr][0] = use_nbits_ch[gr][0] + 32;
                }
                if (max_nbits_ch[gr][1] > use_nbits_ch[gr][1] + 32) {
                    max_nbits_ch[gr][0] += max_nbits_ch[gr][1];
                    max_nbits_ch[gr][0] -= use_nbits_ch[gr][1] + 32;
                    max_nbits_ch[gr][1] = use_nbits_ch[gr][1] + 32;
                }
                if (max_nbits_ch[gr][0] > LP_MAX_BITS_PER_CHANNEL) max_nbits_ch[gr][0] = LP_MAX_BITS_PER_CHANNEL;
                if (max_nbits_ch[gr][1] > LP_MAX_BITS_PER_CHANNEL) max_nbits_ch[gr][1] = LP_MAX_BITS_PER_CHANNEL;
            }
            max_nbits_gr[gr] = 0;
            for (ch = 0; ch < nch; ++ch) max_nbits_gr[gr] += max_nbits_ch[gr][ch];
        }
        sum_fr += max_nbits_gr[gr];
    }
    if (sum_fr > max_nbits_fr) {
        {
            float f[2] = { 0.0f, 0.0f }, s = 0.0f;
            for (gr = 0; gr < ngr; ++gr) {
                if (max_nbits_gr[gr] > 0) { f[gr] = sqrt(max_nbits_gr[gr]); s += f[gr]; }
                else f[gr] = 0;
            }
            for (gr = 0; gr < ngr; ++gr) max_nbits_gr[gr] = (s > 0) ? max_nbits_fr * f[gr] / s : 0;
        }
        if (ngr > 1) {
            if (max_nbits_gr[0] > use_nbits_gr[0] + 125) {
                max_nbits_gr[1] += max_nbits_gr[0];
                max_nbits_gr[1] -= use_nbits_gr[0] + 125;
                max_nbits_gr[0] = use_nbits_gr[0] + 125;
            }
            if (max_nbits_gr[1] > use_nbits_gr[1] + 125) {
                max_nbits_gr[0] += max_nbits_gr[1];
                max_nbits_gr[0] -= use_nbits_gr[1] + 125;
                max_nbits_gr[1] = use_nbits_gr[1] + 125;
            }
            for (gr = 0; gr < ngr; ++gr) if (max_nbits_gr[gr] > LP_MAX_BITS_PER_GRANULE) max_nbits_gr[gr] = LP_MAX_BITS_PER_GRANULE;
        }
        for (gr = 0; gr < ngr; ++gr) {
            float f[2] = { 0.0f, 0.0f }, s = 0.0f;
            for (ch = 0; ch < nch; ++ch) {
                if (max_nbits_ch[gr][ch] > 0) { f[ch] = sqrt(max_nbits_ch[gr][ch]); s += f[ch]; }
                else f[ch] = 0;
            }
            for (ch = 0; ch < nch; ++ch) max_nbits_ch[gr][ch] = (s > 0) ? max_nbits_gr[gr] * f[ch] / s : 0;
            if (nch > 1) {
                if (max_nbits_ch[gr][0] > use_nbits_ch[gr][0] + 32) {
                    max_nbits_ch[gr][1] += max_nbits_ch[gr][0];
                    max_nbits_ch[gr][1] -= use_nbits_ch[gr][0] + 32;
                    max_nbits_ch[gr][0] = use_nbits_ch[gr][0] + 32;
                }
                if (max_nbits_ch[gr][1] > use_nbits_ch[gr][1] + 32) {
                    max_nbits_ch[gr][0] += max_nbits_ch[gr][1];
                    max_nbits_ch[gr][0] -= use_nbits_ch[gr][1] + 32;
                    max_nbits_ch[gr][1] = use_nbits_ch[gr][1] + 32;
                }
                for (ch = 0; ch < nch; ++ch) if (max_nbits_ch[gr][ch] > LP_MAX_BITS_PER_CHANNEL) max_nbits_ch[gr][ch] = LP_MAX_BITS_PER_CHANNEL;
            }
        }
    }
    sum_fr = 0;
    for (gr = 0; gr < ngr; ++gr) {
        int sum_gr = 0;
        for (ch = 0; ch < nch; ++ch) {
            sum_gr += max_nbits_ch[gr][ch];
            if (max_nbits_ch[gr][ch] > LP_MAX_BITS_PER_CHANNEL) ok = 0;
        }
        sum_fr += sum_gr;
        if (sum_gr > LP_MAX_BITS_PER_GRANULE) ok = 0;
    }
    if (sum_fr > max_nbits_fr) ok = 0;
    if (!ok) for (gr = 0; gr < ngr; ++gr) for (ch = 0; ch < nch; ++ch) max_nbits_ch[gr][ch] = max_bits[gr][ch];
    /* best_scalefac_store already ran once: reset what it left behind */
    for (ch = 0; ch < nch; ++ch) e->scfsi[ch][0] = e->scfsi[ch][1] = e->scfsi[ch][2] = e->scfsi[ch][3] = 0;
    for (gr = 0; gr < ngr; ++gr) for (ch = 0; ch < nch; ++ch) e->tt[gr][ch].scalefac_compress = 0;
    use_nbits_fr = 0;
    for (gr = 0; gr < ngr; ++gr) {
        use_nbits_gr[gr] = 0;
        for (ch = 0; ch < nch; ++ch) {
            const vbr_ctx *t = &that_[gr][ch];
            use_nbits_ch[gr][ch] = 0;
            if (max_bits[gr][ch] > 0) {
                int *sfwork = sfwork_[gr][ch], i;
                int const cut = t->gi->global_gain;
                for (i = 0; i < LP_SFBMAX; ++i) if (sfwork[i] > cut) sfwork[i] = cut;        /* cutDistribution */
                vbr_out_of_bits(c, t, sfwork, vbrsfmin_[gr][ch], max_nbits_ch[gr][ch]);
            }
            use_nbits_ch[gr][ch] = vbr_reduce_bits(e, gr, ch);
            use_nbits_gr[gr] += use_nbits_ch[gr][ch];
        }
        use_nbits_fr += use_nbits_gr[gr];
    }
    return use_nbits_fr;
}

/* quantize.c:1645 VBR_new_iteration_loop (+ :1582 VBR_new_prepare, :1341 get_framebits) */
void lp_vbr_new_iteration_loop(lp_encoder *e, float pe[2][2], const float ms_ener_ratio[2], lp_ratio ratio[2][2])
{
    const lp_config *cfg = &e->cfg;
    float l3_xmin[2][2][LP_SFBMAX], xrpow[2][2][576];
    int frameBits[16], max_bits[2][2], used_bits, gr, ch, i, j, analog_silence = 1, avg, bits = 0, maximum_framebits, pad, dummy;
    (void) ms_ener_ratio;
    memset(xrpow, 0, sizeof xrpow);
    /* VBR_new_prepare */
    e->bitrate_index = cfg->vbr_max_bitrate_index;
    (void) resv_frame_begin(e, &avg);
    pad = e->resv_max;
    for (i = 1; i <= cfg->vbr_max_bitrate_index; i++) {            /* get_framebits */
        e->bitrate_index = i;
        frameBits[i] = resv_frame_begin(e, &dummy);
    }
    maximum_framebits = frameBits[cfg->vbr_max_bitrate_index];
    for (gr = 0; gr < cfg->mode_gr; gr++) {
        (void) on_pe(e, pe, max_bits[gr], avg, gr, 0);
        if (e->mode_ext == 2)
            for (i = 0; i < 576; ++i) {
                float l = e->tt[gr][0].xr[i], r = e->tt[gr][1].xr[i];
                e->tt[gr][0].xr[i] = (l + r) * (float) (SQRT2_D * 0.5);
                e->tt[gr][1].xr[i] = (l - r) * (float) (SQRT2_D * 0.5);
            }
        for (ch = 0; ch < cfg->channels; ++ch) {
            lp_granule *gi = &e->tt[gr][ch];
            e->masking_lower = pow(10.0, cfg->mask_adjust * 0.1);
            init_outer_loop(cfg, gi);
            if (0 != calc_xmin(e, &ratio[gr][ch], gi, l3_xmin[gr][ch])) analog_silence = 0;
            bits += max_bits[gr][ch];
        }
    }
    for (gr = 0; gr < cfg->mode_gr; gr++)
        for (ch = 0; ch < cfg->channels; ch++)
            if (bits > maximum_framebits && bits > 0) { max_bits[gr][ch] *= maximum_framebits; max_bits[gr][ch] /= bits; }
    if (analog_silence) pad = 0;
    /* the loop itself */
    for (gr = 0; gr < cfg->mode_gr; gr++)
        for (ch = 0; ch < cfg->channels; ch++)
            if (0 == init_xrpow(cfg, &e->tt[gr][ch], xrpow[gr][ch])) max_bits[gr][ch] = 0;
    used_bits = vbr_encode_frame(e, xrpow, l3_xmin, max_bits);
    i = (analog_silence /* && !enforce_min_bitrate */) ? 1 : cfg->vbr_min_bitrate_index;
    for (; i < cfg->vbr_max_bitrate_index; i++) if (used_bits <= frameBits[i]) break;
    if (i > cfg->vbr_max_bitrate_index) i = cfg->vbr_max_bitrate_index;
    if (pad > 0) {
        for (j = cfg->vbr_max_bitrate_index; j > i; --j) {
            int const unused = frameBits[j] - used_bits;
            if (unused <= pad) break;
        }
        e->bitrate_index = j;
    }
    else e->bitrate_index = i;
    {
        int mean_bits;
        (void) resv_frame_begin(e, &mean_bits);
        for (gr = 0; gr < cfg->mode_gr; gr++)
            for (ch = 0; ch < cfg->channels; ch++) e->resv_size -= e->tt[gr][ch].part2_3_length + e->tt[gr][ch].part2_length;
        resv_frame_end(e, mean_bits);
    }
}
