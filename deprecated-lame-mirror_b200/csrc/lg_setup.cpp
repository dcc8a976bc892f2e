/* lg_setup.cpp - host-side, one-time configuration of the encoder (product code).
 * Resolves user parameters into the immutable LgDevCfg blob that is uploaded to the GPU.  It follows
 * the one-time configuration of the reference: lame_init_params (lame.c:538), the
 * bitrate preset (presets.c:216), lame_init_qval (lame.c:363), the polyphase low-pass gains
 * (lame.c:103), iteration_init (quantize_pvt.c:336), compute_ath (:226), huffman_init
 * (takehiro.c:1334), psymodel_init (psymodel.c:1867), init_fft (fft.c:297) and init_log_table
 * (util.c:959).  Only the CBR / MPEG-1 corner of the configuration space is implemented; anything
 * else makes lg_setup() fail.  All libm calls are the host's, as in the reference. */
#include <math.h>
#include <stdlib.h>
#include <float.h>
#include <string.h>
#include "lg_types.h"
#include "lg_engine.h"
#include "lg_tables_data.inc"

#define LG_FLOAT_MAX 1e37 /* machine.h:137 (FLT_MAX is not visible there) */
#define LG_PI 3.14159265358979323846
#define LG_LOG10 2.30258509299404568402

extern "C" const uint8_t lg_pretab[22] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0 };


/* util.c:336 nearestBitrateFullIndex */
static int nearest_full_index(int bitrate)
{
    static const int tab[17] = { 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320 };
    int lo = 16, hi = 16, lo_k = 320, hi_k = 320, b;
    for (b = 0; b < 16; b++) {
        if ((bitrate > tab[b + 1] ? bitrate : tab[b + 1]) != bitrate) {
            hi_k = tab[b + 1]; hi = b + 1; lo_k = tab[b]; lo = b;
            break;
        }
    }
    return (hi_k - bitrate) > (bitrate - lo_k) ? lo : hi;
}

/* util.c:243 ATHformula_GB / ATHformula */
static float ath_formula_gb(float f, float value, float f_min, float f_max)
{
    float ath;
    if (f < -.3) f = 3410;
    f /= 1000;
    f = f_min > f ? f_min : f;
    f = f_max < f ? f_max : f;
    ath = 3.640 * pow(f, -0.8) - 6.800 * exp(-0.6 * pow(f - 3.4, 2.0))
        + 6.000 * exp(-0.15 * pow(f - 8.7, 2.0)) + (0.6 + 0.04 * value) * 0.001 * pow(f, 4.0);
    return ath;
}
static float ath_formula(const LgDevCfg *c, float f)
{
    switch (c->athtype) {
    case 0: return ath_formula_gb(f, 9, 0.1f, 24.0f);
    case 1: return ath_formula_gb(f, -1, 0.1f, 24.0f);
    case 2: return ath_formula_gb(f, 0, 0.1f, 24.0f);
    case 3: return ath_formula_gb(f, 1, 0.1f, 24.0f) + 6;
    case 4: return ath_formula_gb(f, c->athcurve, 0.1f, 24.0f);
    case 5: return ath_formula_gb(f, c->athcurve, 3.41f, 16.1f);
    default: return ath_formula_gb(f, 0, 0.1f, 24.0f);
    }
}

/* util.c:306 freq2bark */
static float freq2bark(float freq)
{
    if (freq < 0) freq = 0;
    freq = freq * 0.001;
    return 13.0 * atan(.76 * freq) + 3.5 * atan(freq * freq / (7.5 * 7.5));
}

/* lame.c:91 filter_coef */
static float filter_coef(float x)
{
    if (x > 1.0) return 0.0;
    if (x <= 0.0) return 1.0;
    return cos(LG_PI / 2 * x);
}

/* lame.c:103 lame_init_params_ppflt */
static void setup_polyphase_gains(LgDevCfg *c)
{
    int band, maxband, minband, lowpass_band = 32, highpass_band = -1;
    float freq;
    if (c->lowpass1 > 0) {
        minband = 999;
        for (band = 0; band <= 31; band++) {
            freq = band / 31.0;
            if (freq >= c->lowpass2) lowpass_band = lowpass_band < band ? lowpass_band : band;
            if (c->lowpass1 < freq && freq < c->lowpass2) minband = minband < band ? minband : band;
        }
        if (minband == 999) c->lowpass1 = (lowpass_band - .75) / 31.0;
        else c->lowpass1 = (minband - .75) / 31.0;
        c->lowpass2 = lowpass_band / 31.0;
    }
    if (c->highpass2 > 0) {
        if (c->highpass2 < .9 * (.75 / 31.0)) { c->highpass1 = 0; c->highpass2 = 0; }
    }
    if (c->highpass2 > 0) {
        maxband = -1;
        for (band = 0; band <= 31; band++) {
            freq = band / 31.0;
            if (freq <= c->highpass1) highpass_band = highpass_band > band ? highpass_band : band;
            if (c->highpass1 < freq && freq < c->highpass2) maxband = maxband > band ? maxband : band;
        }
        c->highpass1 = highpass_band / 31.0;
        if (maxband == -1) c->highpass2 = (highpass_band + .75) / 31.0;
        else c->highpass2 = (maxband + .75) / 31.0;
    }
    for (band = 0; band < 32; band++) {
        float fc1, fc2;
        freq = band / 31.0f;
        if (c->highpass2 > c->highpass1)
            fc1 = filter_coef((c->highpass2 - freq) / (c->highpass2 - c->highpass1 + 1e-20));
        else fc1 = 1.0f;
        if (c->lowpass2 > c->lowpass1)
            fc2 = filter_coef((freq - c->lowpass1) / (c->lowpass2 - c->lowpass1 + 1e-20));
        else fc2 = 1.0f;
        c->amp_filter[band] = fc1 * fc2;
    }
}

/* quantize_pvt.c:207 ATHmdct */
static float ath_mdct(const LgDevCfg *c, float f)
{
    float ath = ath_formula(c, f);
    if (c->athfixpoint > 0) ath -= c->athfixpoint;
    else ath -= 100;
    ath += c->ath_offset_db;
    ath = powf(10.0f, ath * 0.1f);
    return ath;
}

/* quantize_pvt.c:226 compute_ath */
static void setup_ath_sfb(LgDevCfg *c)
{
    int sfb, i;
    float const samp_freq = c->samplerate;
    for (sfb = 0; sfb < 22; sfb++) {
        c->ath_l[sfb] = LG_FLOAT_MAX;
        for (i = c->sfb_l[sfb]; i < c->sfb_l[sfb + 1]; i++) {
            float const freq = i * samp_freq / (2 * 576);
            float const a = ath_mdct(c, freq);
            c->ath_l[sfb] = c->ath_l[sfb] < a ? c->ath_l[sfb] : a;
        }
    }
    for (sfb = 0; sfb < 6; sfb++) {
        c->ath_psfb21[sfb] = LG_FLOAT_MAX;
        for (i = c->psfb21[sfb]; i < c->psfb21[sfb + 1]; i++) {
            float const freq = i * samp_freq / (2 * 576);
            float const a = ath_mdct(c, freq);
            c->ath_psfb21[sfb] = c->ath_psfb21[sfb] < a ? c->ath_psfb21[sfb] : a;
        }
    }
    for (sfb = 0; sfb < 13; sfb++) {
        c->ath_s[sfb] = LG_FLOAT_MAX;
        for (i = c->sfb_s[sfb]; i < c->sfb_s[sfb + 1]; i++) {
            float const freq = i * samp_freq / (2 * 192);
            float const a = ath_mdct(c, freq);
            c->ath_s[sfb] = c->ath_s[sfb] < a ? c->ath_s[sfb] : a;
        }
        c->ath_s[sfb] *= (c->sfb_s[sfb + 1] - c->sfb_s[sfb]);
    }
    for (sfb = 0; sfb < 6; sfb++) {
        c->ath_psfb12[sfb] = LG_FLOAT_MAX;
        for (i = c->psfb12[sfb]; i < c->psfb12[sfb + 1]; i++) {
            float const freq = i * samp_freq / (2 * 192);
            float const a = ath_mdct(c, freq);
            c->ath_psfb12[sfb] = c->ath_psfb12[sfb] < a ? c->ath_psfb12[sfb] : a;
        }
        c->ath_psfb12[sfb] *= (c->sfb_s[13] - c->sfb_s[12]);
    }
    if (c->no_ath) {                                                   /* quantize_pvt.c:294: reduce the ATH to -200 dB */
        for (sfb = 0; sfb < 22; sfb++) c->ath_l[sfb] = 1E-20;
        for (sfb = 0; sfb < 6; sfb++) c->ath_psfb21[sfb] = 1E-20;
        for (sfb = 0; sfb < 13; sfb++) c->ath_s[sfb] = 1E-20;
        for (sfb = 0; sfb < 6; sfb++) c->ath_psfb12[sfb] = 1E-20;
    }
    c->ath_floor = 10. * log10(ath_mdct(c, -1.));
}

/* quantize_pvt.c:336 iteration_init + takehiro.c:1334 huffman_init */
static void setup_quantizer_tables(LgDevCfg *c)
{
    static const float payload_long[4] = { -0.500f, -0.250f, -0.025f, +0.500f };
    static const float payload_short[4] = { -2.000f, -1.000f, -0.050f, +0.500f };
    /* takehiro.c:36 subdv_table (region0,region1 per number of scalefactor bands) */
    static const uint8_t subdv[23][2] = { {0,0},{0,0},{0,0},{0,0},{0,0},{0,1},{1,1},{1,1},{1,2},{2,2},{2,3},
        {2,3},{3,4},{3,4},{3,4},{4,5},{4,5},{4,6},{5,6},{5,6},{5,7},{6,7},{6,7} };
    int i;
    float db, adjust;
    setup_ath_sfb(c);
    c->pow43[0] = 0.0;
    for (i = 1; i < LG_PRECALC; i++) c->pow43[i] = pow((float) i, 4.0 / 3.0);
    c->adj43asm[0] = 0.0;
    for (i = 1; i < LG_PRECALC; i++) c->adj43asm[i] = i - 0.5 - pow(0.5 * (c->pow43[i - 1] + c->pow43[i]), 0.75);
    for (i = 0; i < LG_QMAX; i++) c->ipow20[i] = pow(2.0, (double) (i - 210) * -0.1875);
    for (i = 0; i <= LG_QMAX + LG_QMAX2; i++) c->pow20[i] = pow(2.0, (double) (i - 210 - LG_QMAX2) * 0.25);

    for (i = 2; i <= 576; i += 2) {
        int scfb_anz = 0, bv_index;
        while (c->sfb_l[++scfb_anz] < i);
        bv_index = subdv[scfb_anz][0];
        while (c->sfb_l[bv_index + 1] > i) bv_index--;
        if (bv_index < 0) bv_index = subdv[scfb_anz][0];
        c->bv_scf[i - 2] = bv_index;
        bv_index = subdv[scfb_anz][1];
        while (c->sfb_l[bv_index + c->bv_scf[i - 2] + 2] > i) bv_index--;
        if (bv_index < 0) bv_index = subdv[scfb_anz][1];
        c->bv_scf[i - 1] = bv_index;
    }

    db = c->adjust_bass_db + payload_long[0]; adjust = powf(10.f, db * 0.1f);
    for (i = 0; i <= 6; ++i) c->longfact[i] = adjust;
    db = c->adjust_alto_db + payload_long[1]; adjust = powf(10.f, db * 0.1f);
    for (; i <= 13; ++i) c->longfact[i] = adjust;
    db = c->adjust_treble_db + payload_long[2]; adjust = powf(10.f, db * 0.1f);
    for (; i <= 20; ++i) c->longfact[i] = adjust;
    db = c->adjust_sfb21_db + payload_long[3]; adjust = powf(10.f, db * 0.1f);
    for (; i < 22; ++i) c->longfact[i] = adjust;
    db = c->adjust_bass_db + payload_short[0]; adjust = powf(10.f, db * 0.1f);
    for (i = 0; i <= 2; ++i) c->shortfact[i] = adjust;
    db = c->adjust_alto_db + payload_short[1]; adjust = powf(10.f, db * 0.1f);
    for (; i <= 6; ++i) c->shortfact[i] = adjust;
    db = c->adjust_treble_db + payload_short[2]; adjust = powf(10.f, db * 0.1f);
    for (; i <= 11; ++i) c->shortfact[i] = adjust;
    db = c->adjust_sfb21_db + payload_short[3]; adjust = powf(10.f, db * 0.1f);
    for (; i < 13; ++i) c->shortfact[i] = adjust;
}

/* psymodel.c:1605 s3_func */
static float spreading(float bark)
{
    float tempx, x, tempy, temp;
    tempx = bark;
    if (tempx >= 0) tempx *= 3;
    else tempx *= 1.5;
    if (tempx >= 0.5 && tempx <= 2.5) {
        temp = tempx - 0.5;
        x = 8.0 * (temp * temp - 2.0 * temp);
    }
    else x = 0.0;
    tempx += 0.474;
    tempy = 15.811389 + 7.5 * tempx - 17.5 * sqrt(1.0 + tempx * tempx);
    if (tempy <= -60.0) return 0.0;
    tempx = exp((x + tempy) * (LG_LOG10 / 10));
    tempx /= .6609193;
    return tempx;
}

/* psymodel.c:1690 stereo_demask */
static float stereo_demask(double f)
{
    double arg = freq2bark(f);
    arg = ((arg < 15.5 ? arg : 15.5) / 15.5);
    return pow(10.0, 1.25 * (1 - cos(LG_PI * arg)) - 2.5);
}

/* psymodel.c:1701 init_numline */
static void setup_partitions(LgBands *gd, float sfreq, int fft_size, int mdct_size, int sbmax, const int *scalepos)
{
    float b_frq[LG_CBANDS + 1];
    float const mdct_freq_frac = sfreq / (2.0f * mdct_size);
    float const deltafreq = fft_size / (2.0f * mdct_size);
    int partition[LG_HBLK] = { 0 };
    int i, j, ni, sfb;
    sfreq /= fft_size;
    j = 0;
    ni = 0;
    for (i = 0; i < LG_CBANDS; i++) {
        float bark1;
        int j2, nl;
        bark1 = freq2bark(sfreq * j);
        b_frq[i] = sfreq * j;
        for (j2 = j; freq2bark(sfreq * j2) - bark1 < .34 && j2 <= fft_size / 2; j2++);
        nl = j2 - j;
        gd->numlines[i] = nl;
        gd->rnumlines[i] = (nl > 0) ? (1.0f / nl) : 0;
        ni = i + 1;
        while (j < j2) partition[j++] = i;
        if (j > fft_size / 2) { j = fft_size / 2; ++i; break; }
    }
    b_frq[i] = sfreq * j;
    gd->n_sb = sbmax;
    gd->npart = ni;
    j = 0;
    for (i = 0; i < gd->npart; i++) {
        int const nl = gd->numlines[i];
        float const freq = sfreq * (j + nl / 2);
        gd->mld_cb[i] = stereo_demask(freq);
        j += nl;
    }
    for (; i < LG_CBANDS; ++i) gd->mld_cb[i] = 1;
    for (sfb = 0; sfb < sbmax; sfb++) {
        int i1, i2, bo;
        int start = scalepos[sfb], end = scalepos[sfb + 1];
        i1 = floor(.5 + deltafreq * (start - .5));
        if (i1 < 0) i1 = 0;
        i2 = floor(.5 + deltafreq * (end - .5));
        if (i2 > fft_size / 2) i2 = fft_size / 2;
        bo = partition[i2];
        gd->bo[sfb] = bo;
        {
            float const f_tmp = mdct_freq_frac * end;
            float bo_w = (f_tmp - b_frq[bo]) / (b_frq[bo + 1] - b_frq[bo]);
            if (bo_w < 0) bo_w = 0;
            else if (bo_w > 1) bo_w = 1;
            gd->bo_weight[sfb] = bo_w;
        }
    }
}

/* psymodel.c:1794 compute_bark_values */
static void bark_values(const LgBands *gd, float sfreq, int fft_size, float *bval, float *bval_width)
{
    int k, j = 0, ni = gd->npart;
    sfreq /= fft_size;
    for (k = 0; k < ni; k++) {
        int const w = gd->numlines[k];
        float bark1, bark2;
        bark1 = freq2bark(sfreq * (j));
        bark2 = freq2bark(sfreq * (j + w - 1));
        bval[k] = .5 * (bark1 + bark2);
        bark1 = freq2bark(sfreq * (j - .5));
        bark2 = freq2bark(sfreq * (j + w - .5));
        bval_width[k] = bark2 - bark1;
        j += w;
    }
}

/* psymodel.c:1816 init_s3_values (the dense matrix is local; only the non-zero band is kept) */
static void setup_spreading(LgBands *gd, const float *bval, const float *bval_width, const float *norm)
{
    static float s3[LG_CBANDS][LG_CBANDS];
    int i, j, k, npart = gd->npart;
    memset(s3, 0, sizeof s3);
    for (i = 0; i < npart; i++)
        for (j = 0; j < npart; j++) {
            float v = spreading(bval[i] - bval[j]) * bval_width[j];
            s3[i][j] = v * norm[i];
        }
    gd->n_s3 = 0;
    for (i = 0; i < npart; i++) {
        for (j = 0; j < npart; j++) if (s3[i][j] > 0.0f) break;
        gd->s3lo[i] = j;
        for (j = npart - 1; j > 0; j--) if (s3[i][j] > 0.0f) break;
        gd->s3hi[i] = j;
        gd->n_s3 += gd->s3hi[i] - gd->s3lo[i] + 1;
    }
    k = 0;
    for (i = 0; i < npart; i++) {
        gd->s3off[i] = k;
        for (j = gd->s3lo[i]; j <= gd->s3hi[i]; j++) gd->s3[k++] = s3[i][j];
    }
}

/* psymodel.c:1867 psymodel_init; fft.c:297 init_fft; psymodel.c:283 init_mask_add_max_values;
 * util.c:959 init_log_table */
static void setup_psy(LgDevCfg *c, float attackthre, float attackthre_s, int vbr_q, float vbr_q_frac)
{
    int i, j, b, k;
    float bvl_a = 13, bvl_b = 24, snr_l_a = 0, snr_l_b = 0, snr_s_a = -8.25, snr_s_b = -4.5;
    float bval[LG_CBANDS], bval_width[LG_CBANDS], norm[LG_CBANDS];
    float const sfreq = c->samplerate;
    float xav = 10, xbv = 12;
    float const minval_low = (0.f - c->minval);
    memset(norm, 0, sizeof norm);

    setup_partitions(&c->l, sfreq, LG_BLK, 576, LG_SBMAX_L, c->sfb_l);
    bark_values(&c->l, sfreq, LG_BLK, bval, bval_width);
    for (i = 0; i < c->l.npart; i++) {
        double snr = snr_l_a;
        if (bval[i] >= bvl_a)
            snr = snr_l_b * (bval[i] - bvl_a) / (bvl_b - bvl_a) + snr_l_a * (bvl_b - bval[i]) / (bvl_b - bvl_a);
        norm[i] = pow(10.0, snr / 10.0);
    }
    setup_spreading(&c->l, bval, bval_width, norm);
    j = 0;
    for (i = 0; i < c->l.npart; i++) {
        double x = LG_FLOAT_MAX;
        for (k = 0; k < c->l.numlines[i]; k++, j++) {
            float const freq = sfreq * j / (1000.0 * LG_BLK);
            float level;
            level = ath_formula(c, freq * 1000) - 20;
            level = pow(10., 0.1 * level);
            level *= c->l.numlines[i];
            if (x > level) x = level;
        }
        c->ath_cb_l[i] = x;
        x = 20.0 * (bval[i] / xav - 1.0);
        if (x > 6) x = 30;
        if (x < minval_low) x = minval_low;
        if (c->samplerate < 44000) x = 30;
        x -= 8.;
        c->l.minval[i] = pow(10.0, x / 10.) * c->l.numlines[i];
    }

    setup_partitions(&c->s, sfreq, LG_BLK_S, 192, LG_SBMAX_S, c->sfb_s);
    bark_values(&c->s, sfreq, LG_BLK_S, bval, bval_width);
    j = 0;
    for (i = 0; i < c->s.npart; i++) {
        double x;
        double snr = snr_s_a;
        if (bval[i] >= bvl_a)
            snr = snr_s_b * (bval[i] - bvl_a) / (bvl_b - bvl_a) + snr_s_a * (bvl_b - bval[i]) / (bvl_b - bvl_a);
        norm[i] = pow(10.0, snr / 10.0);
        x = LG_FLOAT_MAX;
        for (k = 0; k < c->s.numlines[i]; k++, j++) {
            float const freq = sfreq * j / (1000.0 * LG_BLK_S);
            float level;
            level = ath_formula(c, freq * 1000) - 20;
            level = pow(10., 0.1 * level);
            level *= c->s.numlines[i];
            if (x > level) x = level;
        }
        c->ath_cb_s[i] = x;
        x = 7.0 * (bval[i] / xbv - 1.0);
        if (bval[i] > xbv) x *= 1 + log(1 + x) * 3.1;
        if (bval[i] < xbv) x *= 1 + log(1 - x) * 2.3;
        if (x > 6) x = 30;
        if (x < minval_low) x = minval_low;
        if (c->samplerate < 44000) x = 30;
        x -= 8;
        c->s.minval[i] = pow(10.0, x / 10) * c->s.numlines[i];
    }
    setup_spreading(&c->s, bval, bval_width, norm);

    c->ma_max_i1 = pow(10, (8 + 1) / 16.0);
    c->ma_max_i2 = pow(10, (23 + 1) / 16.0);
    for (i = 0; i < LG_BLK; i++)
        c->window[i] = 0.42 - 0.5 * cos(2 * LG_PI * (i + .5) / LG_BLK) + 0.08 * cos(4 * LG_PI * (i + .5) / LG_BLK);
    for (i = 0; i < LG_BLK_S / 2; i++)
        c->window_s[i] = 0.5 * (1.0 - cos(2.0 * LG_PI * (i + 0.5) / LG_BLK_S));

    c->decay = exp(-1.0 * LG_LOG10 / (0.01 * sfreq / 192.0));
    {
        float msfix = 3.5;
        if (c->use_safe_joint_stereo) msfix = 1.0;
        if (fabs(c->msfix) > 0.0) msfix = c->msfix;
        c->msfix = msfix;
        for (b = 0; b < c->l.npart; b++)
            if (c->l.s3hi[b] > c->l.npart - 1) c->l.s3hi[b] = c->l.npart - 1;
    }
    c->ath_decay = pow(10., -12. / 10. * (576. * c->mode_gr / sfreq));
    {
        float freq;
        float const freq_inc = (float) c->samplerate / (float) (LG_BLK);
        float eql_balance = 0.0;
        freq = 0.0;
        for (i = 0; i < LG_BLK / 2; ++i) {
            freq += freq_inc;
            c->eql_w[i] = 1. / pow(10, ath_formula(c, freq) / 10);
            eql_balance += c->eql_w[i];
        }
        eql_balance = 1.0 / eql_balance;
        for (i = LG_BLK / 2; --i >= 0;) c->eql_w[i] *= eql_balance;
    }
    {
        float x = attackthre, y = attackthre_s;
        if (x < 0) x = 4.4;
        if (y < 0) y = 25;
        c->attack_threshold[0] = c->attack_threshold[1] = c->attack_threshold[2] = x;
        c->attack_threshold[3] = y;
    }
    {
        float sk_s, sk_l;
        static float const sk[] = { -7.4, -7.4, -7.4, -9.5, -7.4, -6.1, -5.5, -4.7, -4.7, -4.7, -4.7 };
        if (vbr_q < 4) sk_l = sk_s = sk[0];
        else sk_l = sk_s = sk[vbr_q] + vbr_q_frac * (sk[vbr_q] - sk[vbr_q + 1]);
        for (b = 0; b < c->s.npart; b++) {
            float m = (float) (c->s.npart - b) / c->s.npart;
            c->s.masking_lower[b] = powf(10.f, sk_s * m * 0.1f);
        }
        for (; b < LG_CBANDS; ++b) c->s.masking_lower[b] = 1.f;
        for (b = 0; b < c->l.npart; b++) {
            float m = (float) (c->l.npart - b) / c->l.npart;
            c->l.masking_lower[b] = powf(10.f, sk_l * m * 0.1f);
        }
        for (; b < LG_CBANDS; ++b) c->l.masking_lower[b] = 1.f;
    }
    memcpy(&c->l2s, &c->l, sizeof c->l2s);
    setup_partitions(&c->l2s, sfreq, LG_BLK, 192, LG_SBMAX_S, c->sfb_s);

    /* util.c:962 init_log_table: C's log() takes the float argument as a double (C++ would pick the float overload: 141 of the 513 entries
     * then come out one ulp off - seen as one stream in 512 differing after 170 frames, round 2) */
    for (j = 0; j < 513; j++) c->log_table[j] = (float) (::log((double) (1.0f + j / (float) 512)) / ::log((double) 2.0f));
}

/* lame.c:363 lame_init_qval */
static int setup_quality(LgDevCfg *c, int quality)
{
    if (c->vbr == 4 && quality == 0) {
        /* lame.c:458-470 case 0; substep_shaping = 2 only matters to the CBR/ABR loops */
        if (c->noise_shaping == 0) c->noise_shaping = 1;
        c->noise_shaping_amp = 2; c->noise_shaping_stop = 1;
        if (c->subblock_gain == -1) c->subblock_gain = 1;
        c->use_best_huffman = 1; c->full_outer_loop = 1;
        return 0;
    }
    switch (quality) {
    default:
    case 9:
        c->noise_shaping = 0; c->noise_shaping_amp = 0; c->noise_shaping_stop = 0;
        c->use_best_huffman = 0; c->full_outer_loop = 0;
        break;
    case 8:
    case 7:
        c->noise_shaping = 0; c->noise_shaping_amp = 0; c->noise_shaping_stop = 0;
        c->use_best_huffman = 0; c->full_outer_loop = 0;
        if (c->vbr == 4) c->full_outer_loop = -1;          /* lame.c:389-391 */
        break;
    case 6:
    case 5:
        if (c->noise_shaping == 0) c->noise_shaping = 1;
        c->noise_shaping_amp = 0; c->noise_shaping_stop = 0;
        if (c->subblock_gain == -1) c->subblock_gain = 1;
        c->use_best_huffman = 0; c->full_outer_loop = 0;
        break;
    case 4:
        if (c->noise_shaping == 0) c->noise_shaping = 1;
        c->noise_shaping_amp = 0; c->noise_shaping_stop = 0;
        if (c->subblock_gain == -1) c->subblock_gain = 1;
        c->use_best_huffman = 1; c->full_outer_loop = 0;
        break;
    case 3:
        if (c->noise_shaping == 0) c->noise_shaping = 1;
        c->noise_shaping_amp = 1; c->noise_shaping_stop = 1;
        if (c->subblock_gain == -1) c->subblock_gain = 1;
        c->use_best_huffman = 1; c->full_outer_loop = 0;
        break;
    case 2:
        if (c->noise_shaping == 0) c->noise_shaping = 1;
        if (c->substep_shaping == 0) c->substep_shaping = 2;
        c->noise_shaping_amp = 1; c->noise_shaping_stop = 1;
        if (c->subblock_gain == -1) c->subblock_gain = 1;
        c->use_best_huffman = 1; c->full_outer_loop = 0;
        break;
    case 1:
    case 0:
        if (c->noise_shaping == 0) c->noise_shaping = 1;
        if (c->substep_shaping == 0) c->substep_shaping = 2;
        c->noise_shaping_amp = 2; c->noise_shaping_stop = 1;
        if (c->subblock_gain == -1) c->subblock_gain = 1;
        c->use_best_huffman = 1; c->full_outer_loop = (quality == 0);
        break;
    }
    return 0;
}

/* util.c:393 map2MP3Frequency */
static int map_to_mp3_frequency(int freq)
{
    static const int f[8] = { 8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100 };
    int i;
    for (i = 0; i < 8; i++) if (freq <= f[i]) return f[i];
    return 48000;
}

/* lame.c:274 optimum_samplefreq: the MPEG rate that suits the low-pass, never far below the input rate */
static int optimum_samplefreq(int lowpassfreq, int in)
{
    static const int rate[9] = { 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000 };
    static const int cut[8] = { 15960, 15250, 11220, 9970, 7230, 5420, 4510, 3970 };    /* low-pass at or below cut[i] -> rate[i + 1] */
    int i, suggested = 44100;
    for (i = 0; i < 9; i++) if (in >= rate[i]) { suggested = rate[i]; break; }
    if (lowpassfreq == -1) return suggested;
    for (i = 0; i < 8; i++) if (lowpassfreq <= cut[i]) suggested = rate[i + 1];
    if (in < suggested) {
        for (i = 1; i < 9; i++) if (in > rate[i]) return rate[i - 1];
        return 8000;
    }
    return suggested;
}

/* util.c:490 blackman + the table part of util.c:531 fill_buffer_resample: (2*bpc + 1) windowed-sinc filters of
 * filter_l + 1 taps, one per fractional offset, each normalised to unit sum */
static float rs_blackman(float x, float fcn, int l)
{
    float bkwn, x2;
    float const wcn = (M_PI * fcn);
    x /= l;
    if (x < 0) x = 0;
    if (x > 1) x = 1;
    x2 = x - .5;
    bkwn = 0.42 - 0.5 * cos(2 * x * M_PI) + 0.08 * cos(4 * x * M_PI);
    if (fabs((double) x2) < 1e-9) return wcn / M_PI;
    return (bkwn * sin((double) (l * wcn * x2)) / (M_PI * l * x2));
}

static int rs_gcd(int i, int j) { return j ? rs_gcd(j, i % j) : i; }

static int setup_resampler(LgDevCfg *c)
{
    int const lo = c->samplerate * 0.9995f, hi = c->samplerate * 1.0005f;      /* util.c:654 isResamplingNecessary */
    int i, j, bpc;
    double ratio;
    float fcn;
    c->resample = (c->samplerate_in < lo) || (hi < c->samplerate_in) ? 1 : 0;
    if (!c->resample) return 0;
    ratio = (double) c->samplerate_in / (double) c->samplerate;
    bpc = c->samplerate / rs_gcd(c->samplerate, c->samplerate_in);
    if (bpc > LG_RS_BPC) bpc = LG_RS_BPC;
    fcn = 1.00 / ratio;
    if (fcn > 1.00) fcn = 1.00;
    c->rs_filter_l = 31 + (fabs(ratio - floor(.5 + ratio)) < FLT_EPSILON ? 1 : 0);
    c->rs_bpc = bpc;
    c->rs_ratio = ratio;
    for (j = 0; j <= 2 * bpc; j++) {
        float sum = 0.;
        float const offset = (j - bpc) / (2. * bpc);
        float *f = c->rs_filt + j * LG_RS_TAPS;
        for (i = 0; i <= c->rs_filter_l; i++) sum += f[i] = rs_blackman(i - offset, fcn, c->rs_filter_l);
        for (i = 0; i <= c->rs_filter_l; i++) f[i] /= sum;
    }
    return 0;
}

static void finish_device_tables(LgDevCfg *c);

/* presets.c:294: apply_abr_preset multiplies the handle's scale by its row's factor - every time it runs, so lame_set_preset(kbps)
 * followed by lame_init_params applies it twice */
extern "C" float lg_abr_preset_scale(int kbps)
{
    static const float sc[17] = { 0.95f, 0.95f, 0.95f, 0.95f, 0.95f, 0.95f, 0.95f, 0.95f, 0.95f, 0.95f, 0.95f, 0.95f, 0.95f, 0.97f, 0.98f, 1.00f, 1.00f };
    return sc[nearest_full_index(kbps)];
}
/* tables.c bitrate_table / samplerate_table rows (0 = MPEG-2, 1 = MPEG-1, 2 = MPEG-2.5) */
extern "C" int lg_table_bitrate(int version, int index) { return LGT_BITRATE[16 * version + index]; }
extern "C" int lg_table_samplerate(int version, int index) { return LGT_SAMPLERATE[4 * version + index]; }

extern "C" void lg_setup_opt_defaults(LgSetupOpt *o)
{
    memset(o, 0, sizeof *o);
    o->scale = o->scale_left = o->scale_right = 1.f;
    o->lowpasswidth = o->highpasswidth = -1;
    o->ath_type = -1; o->ath_curve = -1.f; o->athaa_type = -1; o->msfix = -1.f; o->interch = -1.f;
    o->short_blocks = -1; o->strict_iso = 2; o->use_temporal = -1;
}
extern "C" int lg_setup(LgDevCfg *c, int samplerate_in, int samplerate_out, int channels, int brate, int mode, int quality, int vbr, float vbr_q_frac)
{
    LgSetupOpt o;
    lg_setup_opt_defaults(&o);
    return lg_setup_ex(c, samplerate_in, samplerate_out, channels, brate, mode, quality, vbr, vbr_q_frac, &o);
}
extern "C" int lg_setup_ex(LgDevCfg *c, int samplerate_in, int samplerate_out, int channels, int brate, int mode, int quality, int vbr, float vbr_q_frac,
                           const LgSetupOpt *opt)
{
    /* presets.c:241 abr_switch_map, the columns the CBR path reads */
    static const struct { int kbps, safejoint; float nsmsfix, st_lrm, st_s, scale, masking_adj, ath_lower, ath_curve, interch; int sfscale; }
    pm[17] = {
        {8, 0, 0, 6.60, 145, 0.95, 0, -30.0, 11, 0.0012, 1}, {16, 0, 0, 6.60, 145, 0.95, 0, -25.0, 11, 0.0010, 1},
        {24, 0, 0, 6.60, 145, 0.95, 0, -20.0, 11, 0.0010, 1}, {32, 0, 0, 6.60, 145, 0.95, 0, -15.0, 11, 0.0010, 1},
        {40, 0, 0, 6.60, 145, 0.95, 0, -10.0, 11, 0.0009, 1}, {48, 0, 0, 6.60, 145, 0.95, 0, -10.0, 11, 0.0009, 1},
        {56, 0, 0, 6.60, 145, 0.95, 0, -6.0, 11, 0.0008, 1}, {64, 0, 0, 6.60, 145, 0.95, 0, -2.0, 11, 0.0008, 1},
        {80, 0, 0, 6.60, 145, 0.95, 0, .0, 8, 0.0007, 1}, {96, 0, 2.50, 6.60, 145, 0.95, 0, 1.0, 5.5, 0.0006, 1},
        {112, 0, 2.25, 6.60, 145, 0.95, 0, 2.0, 4.5, 0.0005, 1}, {128, 0, 1.95, 6.40, 140, 0.95, 0, 3.0, 4, 0.0002, 1},
        {160, 1, 1.79, 6.00, 135, 0.95, -2, 5.0, 3.5, 0, 1}, {192, 1, 1.49, 5.60, 125, 0.97, -4, 7.0, 3, 0, 0},
        {224, 1, 1.25, 5.20, 125, 0.98, -6, 9.0, 2, 0, 0}, {256, 1, 0.97, 5.20, 125, 1.00, -8, 10.0, 1, 0, 0},
        {320, 1, 0.90, 5.20, 125, 1.00, -10, 12.0, 0, 0, 0} };
    static const int lowpass_map[17] = { 2000, 3700, 3900, 5500, 7000, 7500, 10000, 11000, 13500, 15100,
        15600, 17000, 17500, 18600, 19400, 19700, 20500 };
    /* presets.c:99 vbr_mt_psy_switch_map: st_lrm, st_s, masking_adj (long, short), ath_lower, ath_curve, ath_sensitivity,
     * interch, safejoint, sfb21mod, msfix, minval, ath_fixpoint; expY = (q >= 3) */
    static const struct { float st_lrm, st_s, madj, madj_s, ath_lower, ath_curve, ath_sens, interch; int safejoint, sfb21mod; float msfix, minval, ath_fixpoint; }
    vm[11] = {
        {4.20, 25.0, -6.8, -6.8, 7.1, 1, 0, 0, 2, 31, 1.000, 5, 100}, {4.20, 25.0, -4.8, -4.8, 5.4, 1.4, -1, 0, 2, 27, 1.122, 5, 98},
        {4.20, 25.0, -2.6, -2.6, 3.7, 2.0, -3, 0, 2, 23, 1.288, 5, 97}, {4.20, 25.0, -1.6, -1.6, 2.0, 2.0, -5, 0, 2, 18, 1.479, 5, 96},
        {4.20, 25.0, -0.0, -0.0, 0.0, 2.0, -8, 0, 2, 12, 1.698, 5, 95}, {4.20, 25.0, 1.3, 1.3, -6, 3.5, -11, 0, 2, 8, 1.950, 5, 94.2},
        {4.50, 100.0, 2.2, 2.3, -12.0, 6.0, -14, 0, 2, 4, 2.239, 3, 93.9}, {4.80, 200.0, 2.7, 2.7, -18.0, 9.0, -17, 0, 2, 0, 2.570, 1, 93.6},
        {5.30, 300.0, 2.8, 2.8, -21.0, 10.0, -23, 0.0002, 0, 0, 2.951, 0, 93.3}, {6.60, 300.0, 2.8, 2.8, -23.0, 11.0, -25, 0.0006, 0, 0, 3.388, 0, 93.3},
        {25.00, 300.0, 2.8, 2.8, -25.0, 12.0, -27, 0.0025, 0, 0, 3.500, 0, 93.3} },                 /* level 10: what level 9 interpolates towards (presets.c:124) */
    /* presets.c:88 vbr_old_switch_map (vbr_rh), same columns, levels 0..10 */
    vo[11] = {
        {5.20, 125.0, -4.2, -6.3, 4.8, 1, 0, 0, 2, 21, 0.97, 5, 100}, {5.30, 125.0, -3.6, -5.6, 4.5, 1.5, 0, 0, 2, 21, 1.35, 5, 100},
        {5.60, 125.0, -2.2, -3.5, 2.8, 2, 0, 0, 2, 21, 1.49, 5, 100}, {5.80, 130.0, -1.8, -2.8, 2.6, 3, -4, 0, 2, 20, 1.64, 5, 100},
        {6.00, 135.0, -0.7, -1.1, 1.1, 3.5, -8, 0, 2, 0, 1.79, 5, 100}, {6.40, 140.0, 0.5, 0.4, -7.5, 4, -12, 0.0002, 0, 0, 1.95, 5, 100},
        {6.60, 145.0, 0.67, 0.65, -14.7, 6.5, -19, 0.0004, 0, 0, 2.30, 5, 100}, {6.60, 145.0, 0.8, 0.75, -19.7, 8, -22, 0.0006, 0, 0, 2.70, 5, 100},
        {6.60, 145.0, 1.2, 1.15, -27.5, 10, -23, 0.0007, 0, 0, 0, 5, 100}, {6.60, 145.0, 1.6, 1.6, -36, 11, -25, 0.0008, 0, 0, 0, 5, 100},
        {6.60, 145.0, 2.0, 2.0, -36, 12, -25, 0.0008, 0, 0, 0, 5, 100} };
    static const int vbr_old_lowpass[11] = { 19500, 19000, 18600, 18000, 17500, 16000, 15600, 14900, 12500, 10000, 3950 };
    static const int vbr_lowpass[11] = { 24000, 19500, 18500, 18000, 17500, 17000, 16500, 15600, 15200, 7230, 3950 };
    int i, j, r, exp_nspsytune = 0, version = 1, best, sr_index, vbr_q = 0, brow = 1;
    int vbr_no_lowpass = 0;
    int samplerate = samplerate_out;                                   /* 0 = chosen below the way lame_init_params does */
    float athaa_sensitivity = 0;
    float maskingadjust, maskingadjust_short, ath_lower_db, attackthre, attackthre_s;
    double lowpass;

    memset(c, 0, sizeof *c);
    if (channels != 1 && channels != 2) return -1;
    if (samplerate_in < 1) return -1;
    c->channels = channels;
    if (channels == 1) mode = LG_MONO;                                 /* lame.c:597 */
    if (mode == LG_MONO) c->channels = 1;
    c->force_ms = (mode == LG_MONO) ? 0 : (opt->force_ms != 0);        /* lame.c:603 */
    float scale = opt->scale;
    if (vbr != 0 && vbr != 2 && vbr != 3 && vbr != 4) return -1;                  /* vbr_off, vbr_abr, vbr_mtrh (lame.h:94) */
    if (vbr == 2) {                                                    /* vbr_rh: `brate` carries VBR_q, no mapping to other rates */
        vbr_q = brate;
        if (vbr_q < 0 || vbr_q > 9 || !(vbr_q_frac >= 0.f && vbr_q_frac < 1.f)) return -1;
        brate = 128;
    }
    if (vbr == 4) {
        /* `brate` carries VBR_q.  Levels 7..9 make lame_init_params pick a lower output rate at these input rates
         * (lame.c:661-698), which needs the resampler; at 32 kHz it rescales VBR_q to a fractional level (:679-686) */
        vbr_q = brate;
        if (vbr_q < 0 || vbr_q > 9 || !(vbr_q_frac >= 0.f && vbr_q_frac < 1.f)) return -1;
        if (samplerate == 0) {
            /* lame.c:661-698: the -V scale is mapped to the internal quality scale of the output rate it implies; e.g. -V7 at
             * 44.1 kHz becomes quality 5.63 at 32 kHz, and at 32 kHz input the levels below 6.5 shrink to 0..5.2 */
            static const struct { int sr_a; float qa, qb, ta, tb; } qm[9] = {
                {48000, 0.0, 6.5, 0.0, 6.5}, {44100, 0.0, 6.5, 0.0, 6.5}, {32000, 6.5, 8.0, 5.2, 6.5}, {24000, 8.0, 8.5, 5.2, 6.0},
                {22050, 8.5, 9.01, 5.2, 6.5}, {16000, 9.01, 9.4, 4.9, 6.5}, {12000, 9.4, 9.6, 4.5, 6.0}, {11025, 9.6, 9.9, 5.1, 6.5},
                {8000, 9.9, 10., 4.9, 6.5} };
            float const qval = vbr_q + vbr_q_frac;
            for (i = 2; i < 9; ++i) {
                if (samplerate_in == qm[i].sr_a && qval < qm[i].qa) {
                    double d = qval / qm[i].qa;
                    d = d * qm[i].ta;
                    vbr_q = (int) d;
                    vbr_q_frac = d - vbr_q;
                }
                if (samplerate_in >= qm[i].sr_a && qm[i].qa <= qval && qval < qm[i].qb) {
                    float const q_ = qm[i].qb - qm[i].qa;
                    float const t_ = qm[i].tb - qm[i].ta;
                    double d = qm[i].ta + t_ * (qval - qm[i].qa) / q_;
                    vbr_q = (int) d;
                    vbr_q_frac = d - vbr_q;
                    samplerate = qm[i].sr_a;
                    vbr_no_lowpass = 1;                                /* gfp->lowpassfreq = -1 */
                    break;
                }
            }
        }
        brate = 128;                                                   /* gfp->brate stays unused; keeps the arithmetic below defined */
    }
    if (vbr == 4 || vbr == 2) { }
    else if (vbr == 3) {
        /* ABR keeps the requested mean bitrate as it is (lame_set_VBR_mean_bitrate_kbps, default 128), clamped for
         * MPEG-1 rates (lame.c:655-658 and :1088-1093 with the default index range 1..14) */
        if (brate == 0) brate = 128;
        if (samplerate) {                                              /* lame.c:645-658: only with an explicit output rate */
            int const lo = samplerate < 32000 ? 8 : 32, hi = samplerate < 16000 ? 64 : (samplerate < 32000 ? 160 : 320);
            if (brate < lo) brate = lo;
            if (brate > hi) brate = hi;
        }
    }
    else {
        if (brate == 0 || opt->compression_ratio > 0) {               /* lame.c:623-644 */
            float const ratio = opt->compression_ratio > 0 ? opt->compression_ratio : 11.025f;
            if (samplerate == 0) samplerate = map_to_mp3_frequency((int) (0.97 * samplerate_in));
            brate = samplerate * 16 * c->channels / (1.e3 * ratio);
        }
    }
    /* lame.c:704-762: low-pass from the bitrate */
    lowpass = lowpass_map[nearest_full_index(brate)];
    if (vbr == 4) {                                                    /* lame.c:730-741 */
        double const a = vbr_lowpass[vbr_q], b = vbr_lowpass[vbr_q + 1], m = vbr_q_frac;
        lowpass = a + m * (b - a);
        if (vbr_no_lowpass) lowpass = -1;
    }
    else if (vbr == 2) {                                               /* lame.c:717-728 */
        double const a = vbr_old_lowpass[vbr_q], b = vbr_old_lowpass[vbr_q + 1], m = vbr_q_frac;
        lowpass = a + m * (b - a);
    }
    else if (mode == LG_MONO) lowpass *= 1.5;
    if (opt->lowpassfreq != 0) lowpass = opt->lowpassfreq;             /* lame.c:704: the automatic choice only when none was asked for */
    c->lowpassfreq = lowpass;
    if (samplerate == 0) {                                             /* lame.c:764-769 */
        if (2 * c->lowpassfreq > samplerate_in) c->lowpassfreq = samplerate_in / 2;
        samplerate = optimum_samplefreq(c->lowpassfreq, samplerate_in);
    }
    /* util.c:433 SmpFrqIndex: MPEG-1 (32/44.1/48 kHz), MPEG-2 (16/22.05/24 kHz) and MPEG-2.5 (8/11.025/12 kHz, also version 0) */
    sr_index = -1;
    for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) if (samplerate == LGT_SAMPLERATE[4 * i + j]) { sr_index = j; version = (i == 1) ? 1 : 0; }
    if (sr_index < 0) return -1;
    brow = (version == 1) ? 1 : (samplerate < 16000 ? 2 : 0);          /* row of bitrate_table: MPEG-2, MPEG-1, MPEG-2.5 */
    if (vbr == 0) {
        /* util.c:320 FindNearestBitrate on that row (lame.c:905-915) */
        best = LGT_BITRATE[16 * brow + 1];
        for (i = 1; i <= 14; i++)
            if (LGT_BITRATE[16 * brow + i] > 0 && abs(LGT_BITRATE[16 * brow + i] - brate) < abs(best - brate)) best = LGT_BITRATE[16 * brow + i];
        brate = best;
    }
    int const preset_brate = brate;                                    /* apply_preset (lame.c:1038) sees the mean bitrate before it is clamped into the index range (:1088-1093) */
    c->samplerate = samplerate;
    c->samplerate_in = samplerate_in;
    if (setup_resampler(c) != 0) return -1;
    if (vbr == 4) c->lowpassfreq = c->lowpassfreq < 24000 ? c->lowpassfreq : 24000;     /* lame.c:770-775 */
    else c->lowpassfreq = c->lowpassfreq < 20500 ? c->lowpassfreq : 20500;
    c->lowpassfreq = samplerate / 2 < c->lowpassfreq ? samplerate / 2 : c->lowpassfreq;
    c->mode_gr = samplerate <= 24000 ? 1 : 2;                          /* lame.c:797 */
    if (mode == LG_MODE_NOT_SET || mode < 0) mode = LG_JOINT;
    /* dual channel (mode 2): two independent channels - no M/S, block types not coupled, its own header code */
    c->mode = mode;
    c->highpass1 = c->highpass2 = 0;
    if (opt->highpassfreq > 0) {                                       /* lame.c:860-875 */
        c->highpass1 = 2. * opt->highpassfreq;
        if (opt->highpasswidth >= 0) c->highpass2 = 2. * (opt->highpassfreq + opt->highpasswidth);
        else c->highpass2 = (1 + 0.00) * 2. * opt->highpassfreq;
        c->highpass1 /= samplerate;
        c->highpass2 /= samplerate;
    }
    c->lowpass1 = c->lowpass2 = 0;
    if (c->lowpassfreq > 0 && c->lowpassfreq < (samplerate / 2)) {
        c->lowpass2 = 2. * c->lowpassfreq;
        if (opt->lowpasswidth >= 0) {                                  /* lame.c:880-884 */
            c->lowpass1 = 2. * (c->lowpassfreq - opt->lowpasswidth);
            if (c->lowpass1 < 0) c->lowpass1 = 0;
        }
        else c->lowpass1 = (1 - 0.00) * 2. * c->lowpassfreq;
        c->lowpass1 /= samplerate;
        c->lowpass2 /= samplerate;
    }
    setup_polyphase_gains(c);
    c->samplerate_index = sr_index;
    c->version = version;
    c->brate = brate;
    c->vbr = vbr;
    c->vbr_mean_kbps = brate;
    c->vbr_min_bitrate_index = 1;                                      /* lame.c:1067-1068 */
    c->vbr_max_bitrate_index = samplerate < 16000 ? 8 : 14;            /* 64 kbps with MPEG-2.5 */
    c->enforce_min_bitrate = opt->vbr_hard_min != 0;                   /* lame.c:557 */
    if (vbr != 0) {                                                    /* lame.c:1071-1093: the user's bitrate range, nearest table entries */
        for (int pass = 0; pass < 2; pass++) {
            int const want = pass ? opt->vbr_max_kbps : opt->vbr_min_kbps;
            if (!want) continue;
            int bestk = LGT_BITRATE[16 * brow + 1], idx = -1;          /* util.c:320 FindNearestBitrate, :360 BitrateIndex */
            for (i = 1; i <= 14; i++)
                if (LGT_BITRATE[16 * brow + i] > 0 && abs(LGT_BITRATE[16 * brow + i] - want) < abs(bestk - want)) bestk = LGT_BITRATE[16 * brow + i];
            for (i = 0; i <= 14; i++) if (LGT_BITRATE[16 * brow + i] > 0 && LGT_BITRATE[16 * brow + i] == bestk) { idx = i; break; }
            if (idx < 0) return -1;
            if (pass) c->vbr_max_bitrate_index = idx; else c->vbr_min_bitrate_index = idx;
        }
        if (vbr == 3) {
            int const hi = LGT_BITRATE[16 * version + c->vbr_max_bitrate_index], lo = LGT_BITRATE[16 * version + c->vbr_min_bitrate_index];
            if (brate > hi) brate = hi;
            if (brate < lo) brate = lo;
            c->brate = brate; c->vbr_mean_kbps = brate;
        }
    }
    c->compression_ratio = samplerate * 16 * c->channels / (1.e3 * (vbr == 3 ? preset_brate : brate));   /* lame.c:839-846, before the clamp of :1085 */
    for (i = 0; i < 16; i++) c->bitrate_kbps[i] = LGT_BITRATE[16 * version + i];
    c->vbr_q = vbr_q;
    c->vbr_q_frac = (vbr == 4 || vbr == 2) ? vbr_q_frac : 0.f;
    if (vbr != 0) c->bitrate_index = 1;                                /* lame.c:921 */
    else {
        c->bitrate_index = -1;
        for (i = 0; i <= 14; i++) if (LGT_BITRATE[16 * brow + i] == brate) { c->bitrate_index = i; break; }
        if (c->bitrate_index <= 0) return -1;
    }
    j = sr_index + 3 * version + 6 * (samplerate < 16000);             /* lame.c:927 */
    for (i = 0; i < 23; i++) c->sfb_l[i] = LGT_SFB_LONG[j * 23 + i];
    for (i = 0; i < 7; i++) c->psfb21[i] = c->sfb_l[21] + i * ((c->sfb_l[22] - c->sfb_l[21]) / 6);
    c->psfb21[6] = 576;
    for (i = 0; i < 14; i++) c->sfb_s[i] = LGT_SFB_SHORT[j * 14 + i];
    for (i = 0; i < 7; i++) c->psfb12[i] = c->sfb_s[12] + i * ((c->sfb_s[13] - c->sfb_s[12]) / 6);
    c->psfb12[6] = 192;
    if (version == 1) c->sideinfo_len = (c->channels == 1) ? 4 + 17 : 4 + 32;     /* lame.c:1228-1235 */
    else c->sideinfo_len = (c->channels == 1) ? 4 + 9 : 4 + 17;
    c->original = 1;

    if (vbr == 4 || vbr == 2) {
        /* lame.c:975-1029 + presets.c:143 apply_vbr_preset(VBR_q) with every option still at its default */
        c->noise_shaping = 0;
        c->quant_comp = 9;
        c->quant_comp_short = 9;
        {   /* presets.c:143-161: every float column moves towards the next level by VBR_q_frac, sfb21mod does so as an int */
            float const x = vbr_q_frac;
            int const n = vbr_q + 1;
#define VP(i) (vbr == 4 ? vm[i] : vo[i])
            int sfb21mod = VP(vbr_q).sfb21mod;
#define VLERP(f) (VP(vbr_q).f + x * (VP(n).f - VP(vbr_q).f))
            attackthre = VLERP(st_lrm);
            attackthre_s = VLERP(st_s);
            maskingadjust = VLERP(madj);
            maskingadjust_short = VLERP(madj_s);
            /* presets.c:34 SET_OPTION: the preset's value unless the option was set (differs from its "unset" value) */
            ath_lower_db = (fabs(opt->ath_lower_db - 0) > 0) ? opt->ath_lower_db : VLERP(ath_lower);
            c->athcurve = (fabs(opt->ath_curve - -1) > 0) ? opt->ath_curve : VLERP(ath_curve);
            athaa_sensitivity = (fabs(opt->athaa_sensitivity - 0) > 0) ? opt->athaa_sensitivity : VLERP(ath_sens);
            c->interch = opt->interch;
            if (VLERP(interch) > 0 && !(fabs(opt->interch - -1) > 0)) c->interch = VLERP(interch);      /* presets.c:185-187 */
            if (c->interch < 0) c->interch = 0;                        /* lame.c:1158 */
            sfb21mod = sfb21mod + x * (VP(n).sfb21mod - sfb21mod);
            c->msfix = (fabs(opt->msfix - -1) > 0) ? opt->msfix : VLERP(msfix);
            if (c->msfix < 0) c->msfix = 0;                            /* lame.c:1143 */
            c->minval = VLERP(minval);
            {   /* presets.c:208-212: the fixpoint follows the input scaling */
                double const ax = fabs(opt->scale);
                double const y = (ax > 0.f) ? (10.f * log10(ax)) : 0.f;
                c->athfixpoint = VLERP(ath_fixpoint) - y;
            }
#undef VLERP
            if (VP(vbr_q).safejoint > 0) exp_nspsytune |= 2;
#undef VP
            if (sfb21mod > 0) exp_nspsytune |= sfb21mod << 20;
        }
        if (vbr == 2) {                                                 /* lame.c:1006-1029: at least level 6 */
            if (quality > 6) quality = 6;
            if (quality < 0) quality = 3;
        }
        else {
        if (quality < 0) quality = 3;
        if (quality < 5) quality = 0;
        if (quality > 7) quality = 7;
        }
        c->sfb21_extra = (vbr_q >= 3) ? 0 : (samplerate > 44000);     /* experimentalY from the preset, lame.c:996-999 */
        goto presets_done;
    }
    /* presets.c:216 apply_abr_preset(brate) with every option still at its default */
    r = nearest_full_index(preset_brate);
    if (pm[r].safejoint > 0) exp_nspsytune |= 2;
    c->noise_shaping = pm[r].sfscale > 0 ? 2 : 0;
    c->quant_comp = 9;
    c->quant_comp_short = 9;
    c->msfix = (fabs(opt->msfix - -1) > 0) ? opt->msfix : pm[r].nsmsfix;
    if (c->msfix < 0) c->msfix = 0;
    attackthre = pm[r].st_lrm;
    attackthre_s = pm[r].st_s;
    scale = scale * pm[r].scale;                                       /* presets.c:294: on top of the user's scale */
    maskingadjust = pm[r].masking_adj;
    if (pm[r].masking_adj > 0) maskingadjust_short = pm[r].masking_adj * .9;
    else maskingadjust_short = pm[r].masking_adj * 1.1;
    ath_lower_db = (fabs(opt->ath_lower_db - 0) > 0) ? opt->ath_lower_db : pm[r].ath_lower;
    c->athcurve = (fabs(opt->ath_curve - -1) > 0) ? opt->ath_curve : pm[r].ath_curve;
    c->interch = (fabs(opt->interch - -1) > 0) ? opt->interch : pm[r].interch;
    if (c->interch < 0) c->interch = 0;
    athaa_sensitivity = opt->athaa_sensitivity;
    c->minval = 5. * (pm[r].kbps / 320.);

    c->sfb21_extra = 0;
presets_done:
    c->preset = (vbr == 4 || vbr == 2) ? 500 - 10 * vbr_q : preset_brate;     /* lame.c:983, :1038 */
    c->mask_adjust = maskingadjust;
    c->mask_adjust_short = maskingadjust_short;
    c->substep_shaping = 0;
    c->subblock_gain = -1;
    c->use_best_huffman = 0;
    if (quality < 0) quality = 3;
    if (quality > 9) quality = 9;
    if (quality == 8) quality = 7;
    c->quality = quality;
    if (setup_quality(c, quality) < 0) return -1;
    c->ath_use_adjust = opt->athaa_type < 0 ? 3 : opt->athaa_type;      /* lame.c:1104-1107 */
    c->ath_aa_sensitivity_p = pow(10.0, athaa_sensitivity / -10.0);
    {   /* lame.c:1115-1133: not set -> allowed; allowed becomes coupled in the two stereo modes */
        int sb = opt->short_blocks;
        if (sb < 0) sb = 0;
        if (sb == 0 && (c->mode == LG_JOINT || c->mode == LG_STEREO)) sb = 1;
        c->short_blocks = sb;
    }
    c->athtype = (vbr == 4) ? 5 : (opt->ath_type < 0 ? 4 : opt->ath_type);   /* presets.c:180 forces 5; lame.c:1149 */
    c->use_temporal = opt->use_temporal >= 0 ? (opt->use_temporal != 0) : ((vbr == 4) ? 0 : 1);      /* lame.c:979-981, :1161 */
    c->no_ath = opt->no_ath != 0;
    c->ath_only = opt->ath_only != 0 || opt->ath_short != 0;
    c->disable_reservoir = opt->disable_reservoir != 0;
    c->ath_offset_db = 0 - ath_lower_db;
    c->ath_offset_factor = powf(10.f, c->ath_offset_db * 0.1f);
    c->use_safe_joint_stereo = exp_nspsytune & 2;
    c->adjust_bass_db = c->adjust_alto_db = c->adjust_treble_db = 0;
    c->adjust_sfb21_db = (exp_nspsytune >> 20) & 63;                    /* lame.c:1186-1190 */
    if (c->adjust_sfb21_db >= 32.f) c->adjust_sfb21_db -= 64.f;
    c->adjust_sfb21_db *= 0.25f;
    c->adjust_sfb21_db += c->adjust_treble_db;
    {
        float m[2][2] = { {1.0f, 0.0f}, {0.0f, 1.0f} };
        m[0][0] *= scale; m[0][1] *= scale; m[1][0] *= scale; m[1][1] *= scale;
        m[0][0] *= opt->scale_left; m[0][1] *= opt->scale_left;
        m[1][0] *= opt->scale_right; m[1][1] *= opt->scale_right;
        if (channels == 2 && c->channels == 1) {
            m[0][0] = 0.5f * (m[0][0] + m[1][0]);
            m[0][1] = 0.5f * (m[0][1] + m[1][1]);
            m[1][0] = 0; m[1][1] = 0;
        }
        memcpy(c->pcm_transform, m, sizeof m);
    }
    c->frac_spf = (vbr == 0) ? ((version + 1) * 72000L * brate) % samplerate : 0;      /* lame.c:1245 */
    setup_quantizer_tables(c);
    setup_psy(c, attackthre, attackthre_s, (vbr == 4 || vbr == 2) ? vbr_q : 4, (vbr == 4 || vbr == 2) ? vbr_q_frac : 0.f);
    /* bitstream.c:91 get_max_frame_buffer_size_by_constraint (the handle's default is MDB_MAXIMUM, lame.c:2341) */
    if (opt->strict_iso == 2) c->buffer_constraint = 7680 * (version + 1);
    else if (opt->strict_iso == 1) c->buffer_constraint = 8 * ((version + 1) * 72000 * LGT_BITRATE[16 * version + (samplerate < 16000 ? 8 : 14)] / samplerate);
    else c->buffer_constraint = 8 * 1440;
    finish_device_tables(c);
    return 0;
}

/* Derived layouts the kernels want: prefix sums of partition widths, the FHT twiddle recurrence
 * unrolled into a table (same float operations as fft.c:105-144, so the values are identical), the
 * two per-frame masking_lower constants (quantize.c:2019-2029) and the static tables. */
static void line_starts(LgBands *b)
{
    int i, j = 0, m = 0;
    for (i = 0; i < b->npart; i++) { b->linestart[i] = j; j += b->numlines[i]; if (b->numlines[i] > m) m = b->numlines[i]; }
    for (; i <= LG_CBANDS; i++) b->linestart[i] = j;
    b->maxlines = m;
}
static void finish_device_tables(LgDevCfg *c)
{
    static const float costab[8] = {
        9.238795325112867e-01, 3.826834323650898e-01, 9.951847266721969e-01, 9.801714032956060e-02,
        9.996988186962042e-01, 2.454122852291229e-02, 9.999811752826011e-01, 6.135884649154475e-03 };
    int s, i, t, n;
    line_starts(&c->l); line_starts(&c->s); line_starts(&c->l2s);
    for (s = 0; s < 4; s++) {
        int const kx = 2 << (2 * s);               /* 2, 8, 32, 128 */
        float c1 = costab[2 * s], s1 = costab[2 * s + 1];
        for (i = 1; i < kx; i++) {
            float c2, s2;
            c2 = 1 - (2 * s1) * s1;
            s2 = (2 * s1) * c1;
            c->fht_tw[s][i][0] = c1; c->fht_tw[s][i][1] = s1; c->fht_tw[s][i][2] = c2; c->fht_tw[s][i][3] = s2;
            c2 = c1;
            c1 = c2 * costab[2 * s] - s1 * costab[2 * s + 1];
            s1 = c2 * costab[2 * s + 1] + s1 * costab[2 * s];
        }
    }
    c->masking_lower_long = pow(10.0, (c->mask_adjust - 0) * 0.1);
    c->masking_lower_short = pow(10.0, (c->mask_adjust_short - 0) * 0.1);
    memcpy(c->enwindow, LGT_ENWINDOW_bits, sizeof LGT_ENWINDOW_bits);
    memcpy(c->mdctwin, LGT_MDCTWIN_bits, sizeof LGT_MDCTWIN_bits);
    /* Huffman length books: 16..23 share one length table, 24..31 another (tables.c:425-441) */
    n = 0;
    for (t = 0; t < 34; t++) {
        int cnt;
        if (t > 16 && t < 24) { c->huff_off[t] = c->huff_off[16]; }
        else if (t > 24 && t < 32) { c->huff_off[t] = c->huff_off[24]; }
        else {
            cnt = (t < 33 ? LGT_HUFF_OFF[t + 1] : (int) sizeof LGT_HUFF_LEN) - LGT_HUFF_OFF[t];
            c->huff_off[t] = n;
            memcpy(c->huff_len + n, LGT_HUFF_LEN + LGT_HUFF_OFF[t], cnt);
            for (int k = 0; k < cnt; k++) c->huff_code[n + k] = LGT_HUFF_CODE[LGT_HUFF_OFF[t] + k];
            n += cnt;
        }
        c->huff_xlen[t] = LGT_HUFF_XLEN[t];
        c->huff_linmax[t] = LGT_HUFF_LINMAX[t];
    }
    memcpy(c->largetbl, LGT_LARGETBL, sizeof LGT_LARGETBL);
    memcpy(c->table23, LGT_TABLE23, sizeof LGT_TABLE23);
    memcpy(c->table56, LGT_TABLE56, sizeof LGT_TABLE56);
    /* packed per-class books (see lg_types.h): candidates {1} {2,3} {5,6} {7,8,9} {10,11,12} {13,14,15} {16..23, 24..31} */
    {
        static const int cand[6][3] = { { 1, 1, 1 }, { 2, 3, 3 }, { 5, 6, 6 }, { 7, 8, 9 }, { 10, 11, 12 }, { 13, 14, 15 } };
        static const int lim[6] = { 2, 3, 4, 6, 8, 16 };
        memset(c->huff_pk, 0, sizeof c->huff_pk);
        for (int k = 0; k < 6; k++) {
            int const xlen = c->huff_xlen[cand[k][0]];
            for (int x = 0; x < lim[k]; x++)
                for (int y = 0; y < lim[k]; y++) {
                    uint32_t e = 0;
                    for (int f = 0; f < 3; f++) e |= (uint32_t) c->huff_len[c->huff_off[cand[k][f]] + x * xlen + y] << (10 * f);
                    c->huff_pk[k * 256 + x * 16 + y] = e;
                }
        }
        for (int b = 0; b < 22; b++) for (int i = c->sfb_l[b]; i < c->sfb_l[b + 1]; i++) c->line_sfb_l[i] = (uint8_t) b;
        for (int sfb = 0; sfb < 13; sfb++) {
            int const ws = c->sfb_s[sfb + 1] - c->sfb_s[sfb];
            for (int wn = 0; wn < 3; wn++)
                for (int i = 0; i < ws; i++) c->line_sfb_s[3 * c->sfb_s[sfb] + wn * ws + i] = (uint8_t) (3 * sfb + wn);
        }
        for (int i = 0; i < 256; i++) {
            uint32_t const a = c->largetbl[i] >> 16, b = c->largetbl[i] & 0xffffu;
            c->huff_pk[6 * 256 + i] = a | (b << 10) | (b << 20);
        }
    }
}
