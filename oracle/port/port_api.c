/* oracle/port/port_api.c - TEST INFRASTRUCTURE (see lame_port.h).
 * Restates the per-frame driver lame_encode_mp3_frame (encoder.c:305) with adjust_ATH (:56) and
 * the PCM buffering of lame_encode_buffer_sample_t (lame.c:1671) / lame_encode_flush (lame.c:2042). */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "lame_port.h"

lp_encoder *lp_open(int samplerate, int channels, int brate, int mode, int quality)
{
    return lp_open_ex(samplerate, channels, brate, mode, quality, 0);
}

lp_encoder *lp_open_ex(int samplerate, int channels, int brate, int mode, int quality, int vbr)
{
    return lp_open_rs(samplerate, 0, channels, brate, mode, quality, vbr);
}

lp_encoder *lp_open_rs(int samplerate_in, int samplerate_out, int channels, int brate, int mode, int quality, int vbr)
{
    return lp_open_vq(samplerate_in, samplerate_out, channels, brate, mode, quality, vbr, 0.f);
}

lp_encoder *lp_open_vq(int samplerate_in, int samplerate_out, int channels, int brate, int mode, int quality, int vbr, float vbr_q_frac)
{
    lp_encoder *e = calloc(1, sizeof *e);
    int i, j, sb;
    if (!e) return NULL;
    if (lp_setup(&e->cfg, samplerate_in, samplerate_out, channels, brate, mode, quality, vbr, vbr_q_frac) < 0) { free(e); return NULL; }
    e->bitrate_index = e->cfg.bitrate_index;
    e->buf = calloc(1, LP_BITBUF);
    /* lame.c:2274 lame_init_internal_flags, lame.c:962, psymodel.c:1897-1922/2075 */
    e->old_value[0] = e->old_value[1] = 180;
    e->current_step[0] = e->current_step[1] = 4;
    e->masking_lower = 1;
    e->mf_samples_to_encode = 576 + 1152;
    e->mf_size = 576 - 48;
    for (i = 0; i < 19; i++) e->pefirbuf[i] = 700 * e->cfg.mode_gr * e->cfg.channels;
    e->slot_lag = e->cfg.frac_spf;
    e->buf_byte_idx = -1;
    for (i = 0; i < 4; ++i) {
        for (j = 0; j < LP_CBANDS; ++j) { e->psy.nb_l1[i][j] = 1e20; e->psy.nb_l2[i][j] = 1e20; }
        for (sb = 0; sb < LP_SBMAX_L; sb++) { e->psy.en[i].l[sb] = 1e20; e->psy.thm[i].l[sb] = 1e20; }
        for (j = 0; j < 3; ++j)
            for (sb = 0; sb < LP_SBMAX_S; sb++) { e->psy.en[i].s[sb][j] = 1e20; e->psy.thm[i].s[sb][j] = 1e20; }
        for (j = 0; j < 9; j++) e->psy.last_en_subshort[i][j] = 10.;
    }
    e->ath_adjust_factor = 0.01;
    e->ath_adjust_limit = 1.0;
    return e;
}

/* lame_set_error_protection(1), lame.c:954: a CRC-16 behind every frame header, two more bytes of side info */
void lp_set_error_protection(lp_encoder *e)
{
    if (!e->cfg.error_protection) { e->cfg.error_protection = 1; e->cfg.sideinfo_len += 2; }
}

void lp_close(lp_encoder *e)
{
    if (!e) return;
    free(e->buf);
    free(e);
}

/* encoder.c:56 adjust_ATH */
static void adjust_ath(lp_encoder *e)
{
    const lp_config *cfg = &e->cfg;
    float gr2_max, max_pow;
    if (cfg->ath_use_adjust == 0) { e->ath_adjust_factor = 1.0; return; }
    max_pow = e->loudness_sq[0][0];
    gr2_max = e->loudness_sq[1][0];
    if (cfg->channels == 2) { max_pow += e->loudness_sq[0][1]; gr2_max += e->loudness_sq[1][1]; }
    else { max_pow += max_pow; gr2_max += gr2_max; }
    if (cfg->mode_gr == 2) max_pow = max_pow > gr2_max ? max_pow : gr2_max;
    max_pow *= 0.5;
    max_pow *= cfg->ath_aa_sensitivity_p;
    if (max_pow > 0.03125) {
        if (e->ath_adjust_factor >= 1.0) e->ath_adjust_factor = 1.0;
        else if (e->ath_adjust_factor < e->ath_adjust_limit) e->ath_adjust_factor = e->ath_adjust_limit;
        e->ath_adjust_limit = 1.0;
    }
    else {
        float const adj_lim_new = 31.98 * max_pow + 0.000625;
        if (e->ath_adjust_factor >= adj_lim_new) {
            e->ath_adjust_factor *= adj_lim_new * 0.075 + 0.925;
            if (e->ath_adjust_factor < adj_lim_new) e->ath_adjust_factor = adj_lim_new;
        }
        else {
            if (e->ath_adjust_limit >= adj_lim_new) e->ath_adjust_factor = adj_lim_new;
            else if (e->ath_adjust_factor < e->ath_adjust_limit) e->ath_adjust_factor = e->ath_adjust_limit;
        }
        e->ath_adjust_limit = adj_lim_new;
    }
}

/* encoder.c:305 lame_encode_mp3_frame */
static int encode_frame(lp_encoder *e, const float *inbuf_l, const float *inbuf_r, unsigned char *out, int cap)
{
    static const float fircoef[9] = { -0.0207887 * 5, -0.0378413 * 5, -0.0432472 * 5, -0.031183 * 5,
        7.79609e-18 * 5, 0.0467745 * 5, 0.10091 * 5, 0.151365 * 5, 0.187098 * 5 };
    const lp_config *cfg = &e->cfg;
    lp_ratio masking_LR[2][2], masking_MS[2][2];
    lp_ratio (*masking)[2];
    const float *inbuf[2];
    float tot_ener[2][4], ms_ener_ratio[2] = { .5, .5 };
    float pe[2][2] = { {0., 0.}, {0., 0.} }, pe_MS[2][2] = { {0., 0.}, {0., 0.} };
    float (*pe_use)[2];
    int ch, gr, i, mp3count;
    float f;
    inbuf[0] = inbuf_l;
    inbuf[1] = inbuf_r;
    if (!e->frame_init_done) {
        /* encoder.c:189 lame_encode_frame_init: prime the filterbank with a short-block pass over
         * the zero-prefixed first samples */
        static float prime0[286 + 1152 + 576], prime1[286 + 1152 + 576];
        int j;
        e->frame_init_done = 1;
        memset(prime0, 0, sizeof prime0);
        memset(prime1, 0, sizeof prime1);
        for (i = 0, j = 0; i < 286 + 576 * (1 + cfg->mode_gr); ++i) {
            if (i >= 576 * cfg->mode_gr) {
                prime0[i] = inbuf[0][j];
                if (cfg->channels == 2) prime1[i] = inbuf[1][j];
                ++j;
            }
        }
        for (gr = 0; gr < cfg->mode_gr; gr++)
            for (ch = 0; ch < cfg->channels; ch++) e->tt[gr][ch].block_type = LP_SHORT;
        lp_mdct_sub48(e, prime0, prime1);
    }
    e->padding = 0;
    if ((e->slot_lag -= cfg->frac_spf) < 0) {
        e->slot_lag += cfg->samplerate;
        e->padding = 1;
    }
    {
        const float *bufp[2] = { 0, 0 };
        int blocktype[2];
        for (gr = 0; gr < cfg->mode_gr; gr++) {
            for (ch = 0; ch < cfg->channels; ch++) bufp[ch] = &inbuf[ch][576 + gr * 576 - (224 + 48)];
            lp_psycho(e, bufp, gr, masking_LR, masking_MS, pe[gr], pe_MS[gr], tot_ener[gr], blocktype);
            if (cfg->mode == LP_JOINT) {
                ms_ener_ratio[gr] = tot_ener[gr][2] + tot_ener[gr][3];
                if (ms_ener_ratio[gr] > 0) ms_ener_ratio[gr] = tot_ener[gr][3] / ms_ener_ratio[gr];
            }
            for (ch = 0; ch < cfg->channels; ch++) {
                e->tt[gr][ch].block_type = blocktype[ch];
                e->tt[gr][ch].mixed_block_flag = 0;
            }
        }
    }
    if (getenv("LP_DEBUG")) fprintf(stderr, "port frame %d: pe %.9g %.9g %.9g %.9g peMS %.9g %.9g %.9g %.9g bt %d %d %d %d\n", e->frame_number, pe[0][0], pe[0][1], pe[1][0], pe[1][1], pe_MS[0][0], pe_MS[0][1], pe_MS[1][0], pe_MS[1][1], e->tt[0][0].block_type, e->tt[0][1].block_type, e->tt[1][0].block_type, e->tt[1][1].block_type);
    adjust_ath(e);
    lp_mdct_sub48(e, inbuf[0], inbuf[1]);
    e->mode_ext = 0;
    if (cfg->force_ms) e->mode_ext = 2;
    else if (cfg->mode == LP_JOINT) {
        float sum_pe_MS = 0, sum_pe_LR = 0;
        for (gr = 0; gr < cfg->mode_gr; gr++)
            for (ch = 0; ch < cfg->channels; ch++) { sum_pe_MS += pe_MS[gr][ch]; sum_pe_LR += pe[gr][ch]; }
        if (sum_pe_MS <= 1.00 * sum_pe_LR) {
            const lp_granule *gi0 = &e->tt[0][0], *gi1 = &e->tt[cfg->mode_gr - 1][0];
            if (gi0[0].block_type == gi0[1].block_type && gi1[0].block_type == gi1[1].block_type) e->mode_ext = 2;
        }
    }
    if (e->mode_ext == 2) { masking = masking_MS; pe_use = pe_MS; }
    else { masking = masking_LR; pe_use = pe; }
    for (i = 0; i < 18; i++) e->pefirbuf[i] = e->pefirbuf[i + 1];
    f = 0.0;
    for (gr = 0; gr < cfg->mode_gr; gr++)
        for (ch = 0; ch < cfg->channels; ch++) f += pe_use[gr][ch];
    e->pefirbuf[18] = f;
    f = e->pefirbuf[9];
    for (i = 0; i < 9; i++) f += (e->pefirbuf[i] + e->pefirbuf[18 - i]) * fircoef[i];
    f = (670 * 5 * cfg->mode_gr * cfg->channels) / f;
    for (gr = 0; gr < cfg->mode_gr; gr++)
        for (ch = 0; ch < cfg->channels; ch++) pe_use[gr][ch] *= f;
    if (getenv("LP_DEBUG")) { fprintf(stderr, "port frame %d: pefir", e->frame_number); for (i = 0; i < 19; i++) fprintf(stderr, " %.9g", e->pefirbuf[i]); fprintf(stderr, "\n"); }
    if (getenv("LP_DEBUG")) fprintf(stderr, "port frame %d: pe_use %.9g %.9g %.9g %.9g f %.9g mer %.9g %.9g\n", e->frame_number, pe_use[0][0], pe_use[0][1], pe_use[1][0], pe_use[1][1], f, ms_ener_ratio[0], ms_ener_ratio[1]);
    memcpy(e->last_pe, pe_use, sizeof e->last_pe);
    if (cfg->vbr == 3) lp_abr_iteration_loop(e, pe_use, ms_ener_ratio, masking);      /* encoder.c:520-538 */
    else if (cfg->vbr == 4) lp_vbr_new_iteration_loop(e, pe_use, ms_ener_ratio, masking);
    else if (cfg->vbr == 2) lp_vbr_old_iteration_loop(e, pe_use, ms_ener_ratio, masking);
    else lp_cbr_iteration_loop(e, pe_use, ms_ener_ratio, masking);
    lp_format_bitstream(e);
    mp3count = lp_copy_buffer(e, out, cap);
    ++e->frame_number;
    return mp3count;
}

/* util.c:531 fill_buffer_resample: up to `desired` output samples from `len` input samples of one channel.
 * Output sample k sits at input time k*ratio - itime; the filter for its fractional position is the nearest of
 * the 2*bpc + 1 precomputed ones.  Returns the samples made, *used = the input samples consumed. */
static int resample_chunk(lp_encoder *e, float *outbuf, int desired, const float *inbuf, int len, int *used, int ch)
{
    const lp_config *cfg = &e->cfg;
    int const filter_l = cfg->rs_filter_l, taps = filter_l + 1, bpc = cfg->rs_bpc;
    double const ratio = cfg->rs_ratio;
    float *old = e->rs_old[ch];
    int i, j = 0, k;
    for (k = 0; k < desired; k++) {
        double const time0 = k * ratio;
        float offset, xvalue = 0.;
        const float *f;
        int joff;
        j = floor(time0 - e->rs_itime[ch]);
        if ((filter_l + j - filter_l / 2) >= len) break;
        offset = (time0 - e->rs_itime[ch] - (j + .5 * (filter_l % 2)));
        joff = floor((offset * 2 * bpc) + bpc + .5);
        f = cfg->rs_filt + joff * LP_RS_TAPS;
        for (i = 0; i <= filter_l; ++i) {
            int const j2 = i + j - filter_l / 2;
            float const y = (j2 < 0) ? old[taps + j2] : inbuf[j2];
            xvalue += y * f[i];
        }
        outbuf[k] = xvalue;
    }
    *used = len < filter_l + j - filter_l / 2 ? len : filter_l + j - filter_l / 2;
    e->rs_itime[ch] += *used - k * ratio;
    if (*used >= taps) for (i = 0; i < taps; i++) old[i] = inbuf[*used + i - taps];
    else {
        int const shift = taps - *used;
        for (i = 0; i < shift; ++i) old[i] = old[i + *used];
        for (j = 0; i < taps; ++i, ++j) old[i] = inbuf[j];
    }
    return k;
}

/* lame.c:1786 lame_copy_inbuffer + lame.c:1671 lame_encode_buffer_sample_t + util.c:665 fill_buffer */
int lp_encode(lp_encoder *e, const short *l, const short *r, int nsamples, unsigned char *out, int cap)
{
    const lp_config *cfg = &e->cfg;
    int mp3size = 0, ret, i, ch;
    int const fs = 576 * cfg->mode_gr;                     /* samples per frame: 1152 (MPEG-1) or 576 (MPEG-2/2.5) */
    int const mf_needed = 1024 + fs - (224 + 48);          /* lame.c:1627 calcNeeded: max(BLKSIZE + fs - FFTOFFSET, 512 + fs - 32) */
    float m[2][2];
    float *in[2] = { NULL, NULL };
    const float *inp[2];
    if (nsamples == 0) return 0;
    if (cfg->channels < 2 && r == NULL) r = l;
    m[0][0] = 1.0f * cfg->pcm_transform[0][0]; m[0][1] = 1.0f * cfg->pcm_transform[0][1];
    m[1][0] = 1.0f * cfg->pcm_transform[1][0]; m[1][1] = 1.0f * cfg->pcm_transform[1][1];
    in[0] = malloc(sizeof(float) * nsamples); in[1] = malloc(sizeof(float) * nsamples);
    for (i = 0; i < nsamples; i++) {
        float const xl = l[i], xr = r[i];
        in[0][i] = xl * m[0][0] + xr * m[0][1];
        in[1][i] = xl * m[1][0] + xr * m[1][1];
    }
    inp[0] = in[0]; inp[1] = in[1];
    while (nsamples > 0) {
        int n_in = 0, n_out = 0;
        if (cfg->resample) {
            for (ch = 0; ch < cfg->channels; ch++)
                n_out = resample_chunk(e, &e->mfbuf[ch][e->mf_size], fs, inp[ch], nsamples, &n_in, ch);
        }
        else {
            n_in = n_out = nsamples < fs ? nsamples : fs;
            for (ch = 0; ch < cfg->channels; ch++) memcpy(&e->mfbuf[ch][e->mf_size], inp[ch], n_out * sizeof(float));
        }
        nsamples -= n_in; inp[0] += n_in; inp[1] += n_in;
        e->mf_size += n_out;
        if (e->mf_samples_to_encode < 1) e->mf_samples_to_encode = 576 + 1152;
        e->mf_samples_to_encode += n_out;
        if (e->mf_size >= mf_needed) {
            int buf_size = cap - mp3size;
            if (cap == 0) buf_size = 0;
            ret = encode_frame(e, e->mfbuf[0], e->mfbuf[1], out, buf_size);
            if (ret < 0) { mp3size = ret; break; }
            out += ret;
            mp3size += ret;
            e->mf_size -= fs;
            e->mf_samples_to_encode -= fs;
            for (ch = 0; ch < cfg->channels; ch++)
                for (i = 0; i < e->mf_size; i++) e->mfbuf[ch][i] = e->mfbuf[ch][i + fs];
        }
    }
    free(in[0]); free(in[1]);
    return mp3size;
}

/* lame.c:2042 lame_encode_flush (no resampling, no ID3v1) */
int lp_flush(lp_encoder *e, unsigned char *out, int cap)
{
    short buffer[2][1152];
    int imp3 = 0, mp3count = 0, remaining, end_padding, frames_left, samples_to_encode;
    int const fs = 576 * e->cfg.mode_gr;
    int const mf_needed = 1024 + fs - (224 + 48);
    if (e->mf_samples_to_encode < 1) return 0;
    samples_to_encode = e->mf_samples_to_encode - 1152;
    if (e->cfg.resample) samples_to_encode += 16. / e->cfg.rs_ratio;       /* lame.c:2083-2087 the resampler's delay */
    memset(buffer, 0, sizeof buffer);
    end_padding = fs - (samples_to_encode % fs);
    if (end_padding < 576) end_padding += fs;
    frames_left = (samples_to_encode + end_padding) / fs;
    while (frames_left > 0 && imp3 >= 0) {
        int const frame_num = e->frame_number;
        int bunch = mf_needed - e->mf_size;
        if (e->cfg.resample) bunch *= e->cfg.rs_ratio;                         /* int *= double: truncated */
        if (bunch > 1152) bunch = 1152;
        if (bunch < 1) bunch = 1;
        remaining = cap - mp3count;
        if (cap == 0) remaining = 0;
        imp3 = lp_encode(e, buffer[0], buffer[1], bunch, out, remaining);
        out += imp3;
        mp3count += imp3;
        frames_left -= ((frame_num != e->frame_number) ? 1 : 0);
    }
    e->mf_samples_to_encode = 0;
    if (imp3 < 0) return imp3;
    remaining = cap - mp3count;
    if (cap == 0) remaining = 0;
    lp_flush_bitstream(e);
    imp3 = lp_copy_buffer(e, out, remaining);
    if (imp3 < 0) return imp3;
    mp3count += imp3;
    return mp3count;
}
