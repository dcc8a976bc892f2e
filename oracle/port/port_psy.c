/* oracle/port/port_psy.c - TEST INFRASTRUCTURE (see lame_port.h).
 * Restates fft.c (windowed split-radix FHT) and the psychoacoustic model L3psycho_anal_vbr
 * (psymodel.c:1397) with its helpers. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include "lame_port.h"

#define SQRT2_D 1.41421356237309504880
#define LOG2_D 0.69314718055994530942
#define LOG10_D 2.30258509299404568402

/* util.c:977 fast_log2: 512-entry table + linear interpolation */
float lp_fast_log2(const lp_config *c, float x)
{
    float log2val, partial;
    union { float f; int i; } fi;
    int mantisse;
    fi.f = x;
    mantisse = fi.i & 0x7fffff;
    log2val = ((fi.i >> 23) & 0xFF) - 0x7f;
    partial = (mantisse & ((1 << (23 - 9)) - 1));
    partial *= 1.0f / ((1 << (23 - 9)));
    mantisse >>= (23 - 9);
    log2val += c->log_table[mantisse] * (1.0f - partial) + c->log_table[mantisse + 1] * partial;
    return log2val;
}
/* util.h:96 FAST_LOG10(x) - a double-valued expression */
#define FAST_LOG10_D(c, x) (lp_fast_log2(c, x) * (LOG2_D / LOG10_D))
#define FAST_LOG10_X_D(c, x, y) (lp_fast_log2(c, x) * (LOG2_D / LOG10_D * (y)))

/* fft.c:64 fht: in-place fast Hartley transform of n2*2 points.  The twiddle recurrence runs in
 * float; the sqrt(2) products run in double and are rounded to float on store. */
static void fht(float *fz, int n2)
{
    static const float costab[8] = {
        9.238795325112867e-01, 3.826834323650898e-01, 9.951847266721969e-01, 9.801714032956060e-02,
        9.996988186962042e-01, 2.454122852291229e-02, 9.999811752826011e-01, 6.135884649154475e-03 };
    const float *tri = costab;
    int const n = n2 << 1;
    float *const fn = fz + n;
    int k4 = 4;
    do {
        float s1, c1;
        int i, k1, k2, k3, kx;
        float *fi, *gi;
        kx = k4 >> 1; k1 = k4; k2 = k4 << 1; k3 = k2 + k1; k4 = k2 << 1;
        fi = fz; gi = fi + kx;
        do {
            float f0, f1, f2, f3;
            f1 = fi[0] - fi[k1]; f0 = fi[0] + fi[k1];
            f3 = fi[k2] - fi[k3]; f2 = fi[k2] + fi[k3];
            fi[k2] = f0 - f2; fi[0] = f0 + f2; fi[k3] = f1 - f3; fi[k1] = f1 + f3;
            f1 = gi[0] - gi[k1]; f0 = gi[0] + gi[k1];
            f3 = SQRT2_D * gi[k3]; f2 = SQRT2_D * gi[k2];
            gi[k2] = f0 - f2; gi[0] = f0 + f2; gi[k3] = f1 - f3; gi[k1] = f1 + f3;
            gi += k4; fi += k4;
        } while (fi < fn);
        c1 = tri[0]; s1 = tri[1];
        for (i = 1; i < kx; i++) {
            float c2, s2;
            c2 = 1 - (2 * s1) * s1;
            s2 = (2 * s1) * c1;
            fi = fz + i; gi = fz + k1 - i;
            do {
                float a, b, g0, f0, f1, g1, f2, g2, f3, g3;
                b = s2 * fi[k1] - c2 * gi[k1]; a = c2 * fi[k1] + s2 * gi[k1];
                f1 = fi[0] - a; f0 = fi[0] + a; g1 = gi[0] - b; g0 = gi[0] + b;
                b = s2 * fi[k3] - c2 * gi[k3]; a = c2 * fi[k3] + s2 * gi[k3];
                f3 = fi[k2] - a; f2 = fi[k2] + a; g3 = gi[k2] - b; g2 = gi[k2] + b;
                b = s1 * f2 - c1 * g3; a = c1 * f2 + s1 * g3;
                fi[k2] = f0 - a; fi[0] = f0 + a; gi[k3] = g1 - b; gi[k1] = g1 + b;
                b = c1 * g2 - s1 * f3; a = s1 * g2 + c1 * f3;
                gi[k2] = g0 - a; gi[0] = g0 + a; fi[k3] = f1 - b; fi[k1] = f1 + b;
                gi += k4; fi += k4;
            } while (fi < fn);
            c2 = c1;
            c1 = c2 * tri[0] - s1 * tri[1];
            s1 = c2 * tri[1] + s1 * tri[0];
        }
        tri += 2;
    } while (k4 < n);
}

static int bitrev8(int v)
{
    int r = 0, b;
    for (b = 0; b < 8; b++) if (v & (1 << b)) r |= 0x80 >> b;
    return r;        /* fft.c:151 rv_tbl[v>>1] == bitrev8(v>>1 << 1)... see callers */
}

/* fft.c:246 fft_long: Blackman window, first radix-4 pass in bit-reversed order, then fht */
void lp_fft_long(const lp_config *c, float x[LP_BLK], const float *buf)
{
    const float *w = c->window;
    int jj = LP_BLK / 8 - 1;
    x += LP_BLK / 2;
    do {
        float f0, f1, f2, f3, v;
        int i = bitrev8(jj) & 0xfe;       /* rv_tbl[jj] = 7-bit reversal of jj, shifted left by one */
        f0 = w[i] * buf[i]; v = w[i + 0x200] * buf[i + 0x200]; f1 = f0 - v; f0 = f0 + v;
        f2 = w[i + 0x100] * buf[i + 0x100]; v = w[i + 0x300] * buf[i + 0x300]; f3 = f2 - v; f2 = f2 + v;
        x -= 4;
        x[0] = f0 + f2; x[2] = f0 - f2; x[1] = f1 + f3; x[3] = f1 - f3;
        f0 = w[i + 1] * buf[i + 1]; v = w[i + 0x201] * buf[i + 0x201]; f1 = f0 - v; f0 = f0 + v;
        f2 = w[i + 0x101] * buf[i + 0x101]; v = w[i + 0x301] * buf[i + 0x301]; f3 = f2 - v; f2 = f2 + v;
        x[LP_BLK / 2 + 0] = f0 + f2; x[LP_BLK / 2 + 2] = f0 - f2;
        x[LP_BLK / 2 + 1] = f1 + f3; x[LP_BLK / 2 + 3] = f1 - f3;
    } while (--jj >= 0);
    fht(x, LP_BLK / 2);
}

/* fft.c:194 fft_short: three 256-point transforms at offsets 192*(b+1), symmetric Hann half-window */
void lp_fft_short(const lp_config *c, float xs[3][LP_BLK_S], const float *buf)
{
    const float *w = c->window_s;
    int b;
    for (b = 0; b < 3; b++) {
        float *x = &xs[b][LP_BLK_S / 2];
        int const k = (576 / 3) * (b + 1);
        int j = LP_BLK_S / 8 - 1;
        do {
            float f0, f1, f2, f3, v;
            int i = bitrev8(j << 2) & 0xfe;
            f0 = w[i] * buf[i + k]; v = w[0x7f - i] * buf[i + k + 0x80]; f1 = f0 - v; f0 = f0 + v;
            f2 = w[i + 0x40] * buf[i + k + 0x40]; v = w[0x3f - i] * buf[i + k + 0xc0]; f3 = f2 - v; f2 = f2 + v;
            x -= 4;
            x[0] = f0 + f2; x[2] = f0 - f2; x[1] = f1 + f3; x[3] = f1 - f3;
            f0 = w[i + 1] * buf[i + k + 1]; v = w[0x7e - i] * buf[i + k + 0x81]; f1 = f0 - v; f0 = f0 + v;
            f2 = w[i + 0x41] * buf[i + k + 0x41]; v = w[0x3e - i] * buf[i + k + 0xc1]; f3 = f2 - v; f2 = f2 + v;
            x[LP_BLK_S / 2 + 0] = f0 + f2; x[LP_BLK_S / 2 + 2] = f0 - f2;
            x[LP_BLK_S / 2 + 1] = f1 + f3; x[LP_BLK_S / 2 + 3] = f1 - f3;
        } while (--j >= 0);
        fht(x, LP_BLK_S / 2);
    }
}

/* psymodel.c:258 tab[], :270 tab_mask_add_delta[] */
static const float tonal_tab[9] = { 1.0, 0.79433, 0.63096, 0.63096, 0.63096, 0.63096, 0.63096, 0.25119, 0.11749 };
static const int mask_add_delta_tab[9] = { 2, 2, 2, 1, 1, 1, 0, 0, -1 };

/* psymodel.c:294 vbrpsy_mask_add */
static float mask_add(const lp_config *c, float m1, float m2, int b, int delta)
{
    static const float table2[10] = { 1.33352 * 1.33352, 1.35879 * 1.35879, 1.38454 * 1.38454, 1.39497 * 1.39497,
        1.40548 * 1.40548, 1.3537 * 1.3537, 1.30382 * 1.30382, 1.22321 * 1.22321, 1.14758 * 1.14758, 1 };
    float ratio;
    if (m1 < 0) m1 = 0;
    if (m2 < 0) m2 = 0;
    if (m1 <= 0) return m2;
    if (m2 <= 0) return m1;
    if (m2 > m1) ratio = m2 / m1; else ratio = m1 / m2;
    if (abs(b) <= delta) {
        if (ratio >= c->ma_max_i1) return m1 + m2;
        else {
            int i = (int) (FAST_LOG10_X_D(c, ratio, 16.0f));
            return (m1 + m2) * table2[i];
        }
    }
    if (ratio < c->ma_max_i2) return m1 + m2;
    if (m1 < m2) m1 = m2;
    return m1;
}

/* psymodel.c:350 convert_partition2scalefac */
static void partition2sfb(const lp_bands *gd, const float *eb, const float *thr, float enn_out[], float thm_out[])
{
    float enn, thmm;
    int sb, b, n = gd->n_sb;
    enn = thmm = 0.0f;
    for (sb = b = 0; sb < n; ++b, ++sb) {
        int const bo_sb = gd->bo[sb];
        int const npart = gd->npart;
        int const b_lim = bo_sb < npart ? bo_sb : npart;
        while (b < b_lim) { enn += eb[b]; thmm += thr[b]; b++; }
        if (b >= npart) { enn_out[sb] = enn; thm_out[sb] = thmm; ++sb; break; }
        {
            float const w_curr = gd->bo_weight[sb];
            float const w_next = 1.0f - w_curr;
            enn += w_curr * eb[b];
            thmm += w_curr * thr[b];
            enn_out[sb] = enn;
            thm_out[sb] = thmm;
            enn = w_next * eb[b];
            thmm = w_next * thr[b];
        }
    }
    for (; sb < n; ++sb) { enn_out[sb] = 0; thm_out[sb] = 0; }
}

/* psymodel.c:443 NS_INTERP */
static float ns_interp(float x, float y, float r)
{
    if (r >= 1.0f) return x;
    if (r <= 0.0f) return y;
    if (y > 0.0f) return powf(x / y, r) * y;
    return 0.0f;
}

/* psymodel.c:458 pecalc_s / :503 pecalc_l */
static float pecalc_s(const lp_config *c, const lp_ratio *mr, float masking_lower)
{
    static const float regcoef_s[12] = { 11.8, 13.6, 17.2, 32, 46.5, 51.3, 57.5, 67.1, 71.5, 84.6, 97.6, 130 };
    float pe_s = 1236.28f / 4;
    unsigned sb, sblock;
    for (sb = 0; sb < LP_SBMAX_S - 1; sb++)
        for (sblock = 0; sblock < 3; sblock++) {
            float const thm = mr->thm.s[sb][sblock];
            if (thm > 0.0f) {
                float const x = thm * masking_lower;
                float const en = mr->en.s[sb][sblock];
                if (en > x) {
                    if (en > x * 1e10f) pe_s += regcoef_s[sb] * (10.0f * LOG10_D);
                    else pe_s += regcoef_s[sb] * FAST_LOG10_D(c, en / x);
                }
            }
        }
    return pe_s;
}
static float pecalc_l(const lp_config *c, const lp_ratio *mr, float masking_lower)
{
    static const float regcoef_l[21] = { 6.8, 5.8, 5.8, 6.4, 6.5, 9.9, 12.1, 14.4, 15, 18.9, 21.6, 26.9, 34.2, 40.2,
        46.8, 56.5, 60.7, 73.9, 85.7, 93.4, 126.1 };
    float pe_l = 1124.23f / 4;
    unsigned sb;
    for (sb = 0; sb < LP_SBMAX_L - 1; sb++) {
        float const thm = mr->thm.l[sb];
        if (thm > 0.0f) {
            float const x = thm * masking_lower;
            float const en = mr->en.l[sb];
            if (en > x) {
                if (en > x * 1e10f) pe_l += regcoef_l[sb] * (10.0f * LOG10_D);
                else pe_l += regcoef_l[sb] * FAST_LOG10_D(c, en / x);
            }
        }
    }
    return pe_l;
}

/* psymodel.c:583 calc_mask_index_l and :958 vbrpsy_calc_mask_index_s (same arithmetic) */
static void mask_index(const lp_bands *gd, const float *max, const float *avg, unsigned char *mask_idx)
{
    float m, a;
    int b, k;
    int const last_tab_entry = 8;
    b = 0;
    a = avg[b] + avg[b + 1];
    if (a > 0.0f) {
        m = max[b];
        if (m < max[b + 1]) m = max[b + 1];
        a = 20.0f * (m * 2.0f - a) / (a * (gd->numlines[b] + gd->numlines[b + 1] - 1));
        k = (int) a;
        if (k > last_tab_entry) k = last_tab_entry;
        mask_idx[b] = k;
    }
    else mask_idx[b] = 0;
    for (b = 1; b < gd->npart - 1; b++) {
        a = avg[b - 1] + avg[b] + avg[b + 1];
        if (a > 0.0f) {
            m = max[b - 1];
            if (m < max[b]) m = max[b];
            if (m < max[b + 1]) m = max[b + 1];
            a = 20.0f * (m * 3.0f - a) / (a * (gd->numlines[b - 1] + gd->numlines[b] + gd->numlines[b + 1] - 1));
            k = (int) a;
            if (k > last_tab_entry) k = last_tab_entry;
            mask_idx[b] = k;
        }
        else mask_idx[b] = 0;
    }
    a = avg[b - 1] + avg[b];
    if (a > 0.0f) {
        m = max[b - 1];
        if (m < max[b]) m = max[b];
        a = 20.0f * (m * 2.0f - a) / (a * (gd->numlines[b - 1] + gd->numlines[b] - 1));
        k = (int) a;
        if (k > last_tab_entry) k = last_tab_entry;
        mask_idx[b] = k;
    }
    else mask_idx[b] = 0;
}

/* psymodel.c:655 vbrpsy_compute_fft_l */
static void compute_fft_l(lp_encoder *e, const float *const buffer[2], int chn, float fftenergy[LP_HBLK],
                          float (*wsamp_l)[LP_BLK])
{
    int j;
    if (chn < 2) lp_fft_long(&e->cfg, *wsamp_l, buffer[chn]);
    else if (chn == 2) {
        float const sqrt2_half = SQRT2_D * 0.5f;
        for (j = LP_BLK - 1; j >= 0; --j) {
            float const l = wsamp_l[0][j], r = wsamp_l[1][j];
            wsamp_l[0][j] = (l + r) * sqrt2_half;
            wsamp_l[1][j] = (l - r) * sqrt2_half;
        }
    }
    fftenergy[0] = wsamp_l[0][0];
    fftenergy[0] *= fftenergy[0];
    for (j = LP_BLK / 2 - 1; j >= 0; --j) {
        float const re = (*wsamp_l)[LP_BLK / 2 - j];
        float const im = (*wsamp_l)[LP_BLK / 2 + j];
        fftenergy[LP_BLK / 2 - j] = (re * re + im * im) * 0.5f;
    }
    {
        float totalenergy = 0.0f;
        for (j = 11; j < LP_HBLK; j++) totalenergy += fftenergy[j];
        e->psy.tot_ener[chn] = totalenergy;
    }
}

/* psymodel.c:707 vbrpsy_compute_fft_s */
static void compute_fft_s(lp_encoder *e, const float *const buffer[2], int chn, int sblock,
                          float (*fftenergy_s)[LP_HBLK_S], float (*wsamp_s)[3][LP_BLK_S])
{
    int j;
    if (sblock == 0 && chn < 2) lp_fft_short(&e->cfg, *wsamp_s, buffer[chn]);
    if (chn == 2) {
        float const sqrt2_half = SQRT2_D * 0.5f;
        for (j = LP_BLK_S - 1; j >= 0; --j) {
            float const l = wsamp_s[0][sblock][j], r = wsamp_s[1][sblock][j];
            wsamp_s[0][sblock][j] = (l + r) * sqrt2_half;
            wsamp_s[1][sblock][j] = (l - r) * sqrt2_half;
        }
    }
    fftenergy_s[sblock][0] = (*wsamp_s)[sblock][0];
    fftenergy_s[sblock][0] *= fftenergy_s[sblock][0];
    for (j = LP_BLK_S / 2 - 1; j >= 0; --j) {
        float const re = (*wsamp_s)[sblock][LP_BLK_S / 2 - j];
        float const im = (*wsamp_s)[sblock][LP_BLK_S / 2 + j];
        fftenergy_s[sblock][LP_BLK_S / 2 - j] = (re * re + im * im) * 0.5f;
    }
}

/* psymodel.c:759 vbrpsy_attack_detection */
static void attack_detection(lp_encoder *e, const float *const buffer[2], int gr_out, lp_ratio masking_ratio[2][2],
                             lp_ratio masking_ms[2][2], float energy[4], float sub_short_factor[4][3],
                             int ns_attacks[4][4], int uselongblock[2])
{
    static const float fircoef[10] = { -8.65163e-18 * 2, -0.00851586 * 2, -6.74764e-18 * 2, 0.0209036 * 2,
        -3.36639e-17 * 2, -0.0438162 * 2, -1.54175e-17 * 2, 0.0931738 * 2, -5.52212e-17 * 2, -0.313819 * 2 };
    float ns_hpfsmpl[2][576];
    const lp_config *cfg = &e->cfg;
    lp_psy_state *psv = &e->psy;
    int const n_chn_out = cfg->channels;
    int const n_chn_psy = (cfg->mode == LP_JOINT) ? 4 : n_chn_out;
    int chn, i, j;
    memset(ns_hpfsmpl, 0, sizeof ns_hpfsmpl);
    for (chn = 0; chn < n_chn_out; chn++) {
        const float *const firbuf = &buffer[chn][576 - 350 - 21 + 192];
        for (i = 0; i < 576; i++) {
            float sum1, sum2;
            sum1 = firbuf[i + 10];
            sum2 = 0.0;
            for (j = 0; j < ((21 - 1) / 2) - 1; j += 2) {
                sum1 += fircoef[j] * (firbuf[i + j] + firbuf[i + 21 - j]);
                sum2 += fircoef[j + 1] * (firbuf[i + j + 1] + firbuf[i + 21 - j - 1]);
            }
            ns_hpfsmpl[chn][i] = sum1 + sum2;
        }
        masking_ratio[gr_out][chn].en = psv->en[chn];
        masking_ratio[gr_out][chn].thm = psv->thm[chn];
        if (n_chn_psy > 2) {
            masking_ms[gr_out][chn].en = psv->en[chn + 2];
            masking_ms[gr_out][chn].thm = psv->thm[chn + 2];
        }
    }
    for (chn = 0; chn < n_chn_psy; chn++) {
        float attack_intensity[12], en_subshort[12], en_short[4] = { 0, 0, 0, 0 };
        float const *pf = ns_hpfsmpl[chn & 1];
        int ns_uselongblock = 1;
        if (chn == 2) {
            for (i = 0, j = 576; j > 0; ++i, --j) {
                float const l = ns_hpfsmpl[0][i], r = ns_hpfsmpl[1][i];
                ns_hpfsmpl[0][i] = l + r;
                ns_hpfsmpl[1][i] = l - r;
            }
        }
        for (i = 0; i < 3; i++) {
            en_subshort[i] = psv->last_en_subshort[chn][i + 6];
            attack_intensity[i] = en_subshort[i] / psv->last_en_subshort[chn][i + 4];
            en_short[0] += en_subshort[i];
        }
        for (i = 0; i < 9; i++) {
            float const *const pfe = pf + 576 / 9;
            float p = 1.;
            for (; pf < pfe; pf++)
                if (p < fabs(*pf)) p = fabs(*pf);
            psv->last_en_subshort[chn][i] = en_subshort[i + 3] = p;
            en_short[1 + i / 3] += p;
            if (p > en_subshort[i + 3 - 2]) p = p / en_subshort[i + 3 - 2];
            else if (en_subshort[i + 3 - 2] > p * 10.0f) p = en_subshort[i + 3 - 2] / (p * 10.0f);
            else p = 0.0;
            attack_intensity[i + 3] = p;
        }
        for (i = 0; i < 3; ++i) {
            float const enn = en_subshort[i * 3 + 3] + en_subshort[i * 3 + 4] + en_subshort[i * 3 + 5];
            float factor = 1.f;
            if (en_subshort[i * 3 + 5] * 6 < enn) {
                factor *= 0.5f;
                if (en_subshort[i * 3 + 4] * 6 < enn) factor *= 0.5f;
            }
            sub_short_factor[chn][i] = factor;
        }
        {
            float x = cfg->attack_threshold[chn];
            for (i = 0; i < 12; i++)
                if (ns_attacks[chn][i / 3] == 0 && attack_intensity[i] > x) ns_attacks[chn][i / 3] = (i % 3) + 1;
        }
        for (i = 1; i < 4; i++) {
            float const u = en_short[i - 1], v = en_short[i];
            float const m = u > v ? u : v;
            if (m < 40000) {
                if (u < 1.7f * v && v < 1.7f * u) {
                    if (i == 1 && ns_attacks[chn][0] <= ns_attacks[chn][i]) ns_attacks[chn][0] = 0;
                    ns_attacks[chn][i] = 0;
                }
            }
        }
        if (ns_attacks[chn][0] <= psv->last_attacks[chn]) ns_attacks[chn][0] = 0;
        if (psv->last_attacks[chn] == 3 || ns_attacks[chn][0] + ns_attacks[chn][1] + ns_attacks[chn][2] + ns_attacks[chn][3]) {
            ns_uselongblock = 0;
            if (ns_attacks[chn][1] && ns_attacks[chn][0]) ns_attacks[chn][1] = 0;
            if (ns_attacks[chn][2] && ns_attacks[chn][1]) ns_attacks[chn][2] = 0;
            if (ns_attacks[chn][3] && ns_attacks[chn][2]) ns_attacks[chn][3] = 0;
        }
        if (chn < 2) uselongblock[chn] = ns_uselongblock;
        else if (ns_uselongblock == 0) uselongblock[0] = uselongblock[1] = 0;
        energy[chn] = psv->tot_ener[chn];
    }
}

/* psymodel.c:1031 vbrpsy_compute_masking_s */
static void compute_masking_s(lp_encoder *e, const float (*fftenergy_s)[LP_HBLK_S], float *eb, float *thr, int chn, int sblock)
{
    const lp_bands *gds = &e->cfg.s;
    float max[LP_CBANDS], avg[LP_CBANDS];
    int i, j, b;
    unsigned char mask_idx_s[LP_CBANDS];
    (void) chn;
    memset(max, 0, sizeof max);
    memset(avg, 0, sizeof avg);
    for (b = j = 0; b < gds->npart; ++b) {
        float ebb = 0, m = 0;
        int const n = gds->numlines[b];
        for (i = 0; i < n; ++i, ++j) {
            float const el = fftenergy_s[sblock][j];
            ebb += el;
            if (m < el) m = el;
        }
        eb[b] = ebb;
        max[b] = m;
        avg[b] = ebb * gds->rnumlines[b];
    }
    mask_index(gds, max, avg, mask_idx_s);
    for (j = b = 0; b < gds->npart; b++) {
        int kk = gds->s3ind[b][0];
        int const last = gds->s3ind[b][1];
        int const delta = mask_add_delta_tab[mask_idx_s[b]];
        int dd, dd_n;
        float x, ecb, avg_mask;
        float const masking_lower = gds->masking_lower[b] * e->masking_lower;
        dd = mask_idx_s[kk];
        dd_n = 1;
        ecb = gds->s3[j] * eb[kk] * tonal_tab[mask_idx_s[kk]];
        ++j, ++kk;
        while (kk <= last) {
            dd += mask_idx_s[kk];
            dd_n += 1;
            x = gds->s3[j] * eb[kk] * tonal_tab[mask_idx_s[kk]];
            ecb = mask_add(&e->cfg, ecb, x, kk - b, delta);
            ++j, ++kk;
        }
        dd = (1 + 2 * dd) / (2 * dd_n);
        avg_mask = tonal_tab[dd] * 0.5f;
        ecb *= avg_mask;
        thr[b] = ecb;
        /* nb_s1/nb_s2 of the reference are write-only state: not kept */
        x = max[b];
        x *= gds->minval[b];
        x *= avg_mask;
        if (thr[b] > x) thr[b] = x;
        if (masking_lower > 1) thr[b] *= masking_lower;
        if (thr[b] > eb[b]) thr[b] = eb[b];
        if (masking_lower < 1) thr[b] *= masking_lower;
    }
    for (; b < LP_CBANDS; ++b) { eb[b] = 0; thr[b] = 0; }
}

/* psymodel.c:1134 vbrpsy_compute_masking_l */
static void compute_masking_l(lp_encoder *e, const float fftenergy[LP_HBLK], float eb_l[LP_CBANDS], float thr[LP_CBANDS], int chn)
{
    lp_psy_state *psv = &e->psy;
    const lp_bands *gdl = &e->cfg.l;
    float max[LP_CBANDS], avg[LP_CBANDS];
    unsigned char mask_idx_l[LP_CBANDS + 2];
    int k, b, j, i;
    /* psymodel.c:556 calc_energy */
    for (b = j = 0; b < gdl->npart; ++b) {
        float ebb = 0, m = 0;
        for (i = 0; i < gdl->numlines[b]; ++i, ++j) {
            float const el = fftenergy[j];
            ebb += el;
            if (m < el) m = el;
        }
        eb_l[b] = ebb;
        max[b] = m;
        avg[b] = ebb * gdl->rnumlines[b];
    }
    mask_index(gdl, max, avg, mask_idx_l);
    k = 0;
    for (b = 0; b < gdl->npart; b++) {
        float x, ecb, avg_mask, t;
        float const masking_lower = gdl->masking_lower[b] * e->masking_lower;
        int kk = gdl->s3ind[b][0];
        int const last = gdl->s3ind[b][1];
        int const delta = mask_add_delta_tab[mask_idx_l[b]];
        int dd = 0, dd_n = 0;
        dd = mask_idx_l[kk];
        dd_n += 1;
        ecb = gdl->s3[k] * eb_l[kk] * tonal_tab[mask_idx_l[kk]];
        ++k, ++kk;
        while (kk <= last) {
            dd += mask_idx_l[kk];
            dd_n += 1;
            x = gdl->s3[k] * eb_l[kk] * tonal_tab[mask_idx_l[kk]];
            t = mask_add(&e->cfg, ecb, x, kk - b, delta);
            ecb = t;
            ++k, ++kk;
        }
        dd = (1 + 2 * dd) / (2 * dd_n);
        avg_mask = tonal_tab[dd] * 0.5f;
        ecb *= avg_mask;
        /* long block pre-echo control (psymodel.c:1187-1234) */
        if (psv->blocktype_old[chn & 0x01] == LP_SHORT) {
            float const ecb_limit = 2 * psv->nb_l1[chn][b];
            if (ecb_limit > 0) thr[b] = ecb < ecb_limit ? ecb : ecb_limit;
            else {
                /* Min(ecb, eb_l[b] * NS_PREECHO_ATT2) with NS_PREECHO_ATT2 = 0.3 (double) */
                thr[b] = (ecb < eb_l[b] * 0.3) ? ecb : eb_l[b] * 0.3;
            }
        }
        else {
            float ecb_limit_2 = 16 * psv->nb_l2[chn][b];
            float ecb_limit_1 = 2 * psv->nb_l1[chn][b];
            float ecb_limit;
            if (ecb_limit_2 <= 0) ecb_limit_2 = ecb;
            if (ecb_limit_1 <= 0) ecb_limit_1 = ecb;
            if (psv->blocktype_old[chn & 0x01] == LP_NORM) ecb_limit = ecb_limit_1 < ecb_limit_2 ? ecb_limit_1 : ecb_limit_2;
            else ecb_limit = ecb_limit_1;
            thr[b] = ecb < ecb_limit ? ecb : ecb_limit;
        }
        psv->nb_l2[chn][b] = psv->nb_l1[chn][b];
        psv->nb_l1[chn][b] = ecb;
        x = max[b];
        x *= gdl->minval[b];
        x *= avg_mask;
        if (thr[b] > x) thr[b] = x;
        if (masking_lower > 1) thr[b] *= masking_lower;
        if (thr[b] > eb_l[b]) thr[b] = eb_l[b];
        if (masking_lower < 1) thr[b] *= masking_lower;
    }
    for (; b < LP_CBANDS; ++b) { eb_l[b] = 0; thr[b] = 0; }
}

/* psymodel.c:1326 vbrpsy_compute_MS_thresholds */
static void ms_thresholds(const float eb[4][LP_CBANDS], float thr[4][LP_CBANDS], const float cb_mld[LP_CBANDS],
                          const float ath_cb[LP_CBANDS], float athlower, float msfix, int n)
{
    float const msfix2 = msfix * 2.f;
    float rside, rmid;
    int b;
    for (b = 0; b < n; ++b) {
        float const ebM = eb[2][b], ebS = eb[3][b], thmL = thr[0][b], thmR = thr[1][b];
        float thmM = thr[2][b], thmS = thr[3][b];
        if (thmL <= 1.58f * thmR && thmR <= 1.58f * thmL) {
            float const mld_m = cb_mld[b] * ebS, mld_s = cb_mld[b] * ebM;
            float const tmp_m = thmS < mld_m ? thmS : mld_m;
            float const tmp_s = thmM < mld_s ? thmM : mld_s;
            rmid = thmM > tmp_m ? thmM : tmp_m;
            rside = thmS > tmp_s ? thmS : tmp_s;
        }
        else { rmid = thmM; rside = thmS; }
        if (msfix > 0.f) {
            float thmLR, thmMS;
            float const ath = ath_cb[b] * athlower;
            float const tmp_l = thmL > ath ? thmL : ath;
            float const tmp_r = thmR > ath ? thmR : ath;
            thmLR = tmp_l < tmp_r ? tmp_l : tmp_r;
            thmM = rmid > ath ? rmid : ath;
            thmS = rside > ath ? rside : ath;
            thmMS = thmM + thmS;
            if (thmMS > 0.f && (thmLR * msfix2) < thmMS) {
                float const f = thmLR * msfix2 / thmMS;
                thmM *= f;
                thmS *= f;
            }
            rmid = thmM < rmid ? thmM : rmid;
            rside = thmS < rside ? thmS : rside;
        }
        if (rmid > ebM) rmid = ebM;
        if (rside > ebS) rside = ebS;
        thr[2][b] = rmid;
        thr[3][b] = rside;
    }
}

/* test hook: the en/thm state right after each of the two calls of a frame */
lp_xmin lp_dbg_en_after[2][4], lp_dbg_thm_after[2][4];

/* psymodel.c:1397 L3psycho_anal_vbr */
int lp_psycho(lp_encoder *e, const float *const buffer[2], int gr_out, lp_ratio masking_ratio[2][2],
              lp_ratio masking_ms[2][2], float percep_entropy[2], float percep_ms_entropy[2], float energy[4],
              int blocktype_d[2])
{
    const lp_config *cfg = &e->cfg;
    lp_psy_state *psv = &e->psy;
    const lp_bands *gdl = &cfg->l, *gds = &cfg->s;
    lp_xmin last_thm[4];
    float (*wsamp_l)[LP_BLK];
    float (*wsamp_s)[3][LP_BLK_S];
    float fftenergy[LP_HBLK], fftenergy_s[3][LP_HBLK_S];
    static float wsamp_L[2][LP_BLK], wsamp_S[2][3][LP_BLK_S];
    float eb[4][LP_CBANDS], thr[4][LP_CBANDS];
    float sub_short_factor[4][3];
    float thmm;
    float const pcfact = 0.6f;
    float const ath_factor = (cfg->msfix > 0.f) ? (cfg->ath_offset_factor * e->ath_adjust_factor) : 1.f;
    int ns_attacks[4][4] = { {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0} };
    int uselongblock[2];
    int chn, sb, sblock;
    int const n_chn_psy = (cfg->mode == LP_JOINT) ? 4 : cfg->channels;

    memcpy(&last_thm[0], &psv->thm[0], sizeof last_thm);
    attack_detection(e, buffer, gr_out, masking_ratio, masking_ms, energy, sub_short_factor, ns_attacks, uselongblock);
    /* psymodel.c:1265 vbrpsy_compute_block_type */
    if (cfg->short_blocks == 1 && !(uselongblock[0] && uselongblock[1])) uselongblock[0] = uselongblock[1] = 0;
    for (chn = 0; chn < cfg->channels; chn++) {
        if (cfg->short_blocks == 2) uselongblock[chn] = 1;
        if (cfg->short_blocks == 3) uselongblock[chn] = 0;
    }
    for (chn = 0; chn < n_chn_psy; chn++) {
        int const ch01 = chn & 0x01;
        wsamp_l = wsamp_L + ch01;
        compute_fft_l(e, buffer, chn, fftenergy, wsamp_l);
        if (chn < 2) {      /* psymodel.c:743 loudness approximation, psycho_loudness_approx :213 */
            float loudness_power = 0.0;
            int i;
            e->loudness_sq[gr_out][chn] = psv->loudness_sq_save[chn];
            for (i = 0; i < LP_BLK / 2; ++i) loudness_power += fftenergy[i] * cfg->eql_w[i];
            loudness_power *= (1. / (14752 * 14752) / (LP_BLK / 2));
            psv->loudness_sq_save[chn] = loudness_power;
        }
        compute_masking_l(e, fftenergy, eb[chn], thr[chn], chn);
    }
    if (cfg->mode == LP_JOINT && (uselongblock[0] + uselongblock[1]) == 2)
        ms_thresholds((const float (*)[LP_CBANDS]) eb, thr, gdl->mld_cb, cfg->ath_cb_l, ath_factor, cfg->msfix, gdl->npart);
    for (chn = 0; chn < n_chn_psy; chn++) {
        float enn[LP_SBMAX_S], thm[LP_SBMAX_S];
        partition2sfb(gdl, eb[chn], thr[chn], &psv->en[chn].l[0], &psv->thm[chn].l[0]);
        /* psymodel.c:421 convert_partition2scalefac_l_to_s */
        partition2sfb(&cfg->l2s, eb[chn], thr[chn], enn, thm);
        for (sb = 0; sb < LP_SBMAX_S; ++sb) {
            float const scale = 1. / 64.f;
            float const tmp_enn = enn[sb];
            float const tmp_thm = thm[sb] * scale;
            for (sblock = 0; sblock < 3; ++sblock) {
                psv->en[chn].s[sb][sblock] = tmp_enn;
                psv->thm[chn].s[sb][sblock] = tmp_thm;
            }
        }
    }
    /* short blocks */
    for (sblock = 0; sblock < 3; sblock++) {
        for (chn = 0; chn < n_chn_psy; ++chn) {
            int const ch01 = chn & 0x01;
            if (!uselongblock[ch01]) {
                wsamp_s = wsamp_S + ch01;
                compute_fft_s(e, buffer, chn, sblock, fftenergy_s, wsamp_s);
                compute_masking_s(e, (const float (*)[LP_HBLK_S]) fftenergy_s, eb[chn], thr[chn], chn, sblock);
            }
        }
        if (cfg->mode == LP_JOINT && (uselongblock[0] + uselongblock[1]) == 0)
            ms_thresholds((const float (*)[LP_CBANDS]) eb, thr, gds->mld_cb, cfg->ath_cb_s, ath_factor, cfg->msfix, gds->npart);
        for (chn = 0; chn < n_chn_psy; ++chn) {
            int const ch01 = chn & 0x01;
            if (!uselongblock[ch01]) {
                float enn[LP_SBMAX_S], thm[LP_SBMAX_S];
                partition2sfb(gds, eb[chn], thr[chn], enn, thm);
                for (sb = 0; sb < LP_SBMAX_S; ++sb) {
                    psv->en[chn].s[sb][sblock] = enn[sb];
                    psv->thm[chn].s[sb][sblock] = thm[sb];
                }
            }
        }
    }
    /* short block pre-echo control (psymodel.c:1502-1554) */
    for (chn = 0; chn < n_chn_psy; chn++) {
        for (sb = 0; sb < LP_SBMAX_S; sb++) {
            float new_thmm[3], prev_thm, t1, t2;
            for (sblock = 0; sblock < 3; sblock++) {
                thmm = psv->thm[chn].s[sb][sblock];
                thmm *= 0.8;                /* NS_PREECHO_ATT0, double product */
                t1 = t2 = thmm;
                if (sblock > 0) prev_thm = new_thmm[sblock - 1];
                else prev_thm = last_thm[chn].s[sb][2];
                if (ns_attacks[chn][sblock] >= 2 || ns_attacks[chn][sblock + 1] == 1)
                    t1 = ns_interp(prev_thm, thmm, 0.6 * pcfact);
                thmm = t1 < thmm ? t1 : thmm;
                if (ns_attacks[chn][sblock] == 1) t2 = ns_interp(prev_thm, thmm, 0.3 * pcfact);
                else if ((sblock == 0 && psv->last_attacks[chn] == 3) || (sblock > 0 && ns_attacks[chn][sblock - 1] == 3)) {
                    switch (sblock) {
                    case 0: prev_thm = last_thm[chn].s[sb][1]; break;
                    case 1: prev_thm = last_thm[chn].s[sb][2]; break;
                    case 2: prev_thm = new_thmm[0]; break;
                    }
                    t2 = ns_interp(prev_thm, thmm, 0.3 * pcfact);
                }
                thmm = t1 < thmm ? t1 : thmm;
                thmm = t2 < thmm ? t2 : thmm;
                thmm *= sub_short_factor[chn][sblock];
                new_thmm[sblock] = thmm;
            }
            for (sblock = 0; sblock < 3; sblock++) psv->thm[chn].s[sb][sblock] = new_thmm[sblock];
        }
    }
    for (chn = 0; chn < n_chn_psy; chn++) psv->last_attacks[chn] = ns_attacks[chn][2];
    memcpy(lp_dbg_en_after[gr_out], psv->en, sizeof psv->en);
    memcpy(lp_dbg_thm_after[gr_out], psv->thm, sizeof psv->thm);

    /* psymodel.c:1289 vbrpsy_apply_block_type */
    for (chn = 0; chn < cfg->channels; chn++) {
        int blocktype = LP_NORM;
        if (uselongblock[chn]) {
            if (psv->blocktype_old[chn] == LP_SHORT) blocktype = LP_STOP;
        }
        else {
            blocktype = LP_SHORT;
            if (psv->blocktype_old[chn] == LP_NORM) psv->blocktype_old[chn] = LP_START;
            if (psv->blocktype_old[chn] == LP_STOP) psv->blocktype_old[chn] = LP_SHORT;
        }
        blocktype_d[chn] = psv->blocktype_old[chn];
        psv->blocktype_old[chn] = blocktype;
    }
    for (chn = 0; chn < n_chn_psy; chn++) {
        float *ppe;
        int type;
        const lp_ratio *mr;
        if (chn > 1) {
            ppe = percep_ms_entropy - 2;
            type = LP_NORM;
            if (blocktype_d[0] == LP_SHORT || blocktype_d[1] == LP_SHORT) type = LP_SHORT;
            mr = &masking_ms[gr_out][chn - 2];
        }
        else {
            ppe = percep_entropy;
            type = blocktype_d[chn];
            mr = &masking_ratio[gr_out][chn];
        }
        if (type == LP_SHORT) ppe[chn] = pecalc_s(cfg, mr, e->masking_lower);
        else ppe[chn] = pecalc_l(cfg, mr, e->masking_lower);
    }
    return 0;
}
