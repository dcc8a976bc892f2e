/* refdump - TEST INFRASTRUCTURE.  Links the UNMODIFIED reference objects (oracle/_ref/obj) and
 * exposes (a) a tiny handle API around the reference's own lame_init/lame_encode_buffer/
 * lame_encode_flush and (b) snapshots of the reference's internal state after each frame and of its
 * init-time tables, read through the reference's own internal headers.  Nothing here re-implements
 * the algorithm; it only copies values out.  Used by tests/ to pin oracle/port and the CUDA path
 * stage by stage (the reference ships no per-stage golden vectors).
 *
 * Built only where /root/reference exists (oracle/Makefile target `ref`). */
#ifdef HAVE_CONFIG_H
#include <config.h>
#endif
#include <stdlib.h>
#include <string.h>
#include "lame.h"
#include "machine.h"
#include "encoder.h"
#include "util.h"
#include "lame_global_flags.h"
#include "quantize_pvt.h"
#include "refdump.h"

struct refdump_handle { lame_global_flags *gfp; };

void *refdump_open(int brate, int mode, int quality, int vbrmode, int vbr_q, int samplerate, int nch)
{
    return refdump_open_rs(brate, mode, quality, vbrmode, vbr_q, samplerate, 0, nch);
}

void *refdump_open_rs(int brate, int mode, int quality, int vbrmode, int vbr_q, int samplerate, int out_samplerate, int nch)
{
    return refdump_open_vq(brate, mode, quality, vbrmode, vbr_q, 0.f, samplerate, out_samplerate, nch);
}

void *refdump_open_vq(int brate, int mode, int quality, int vbrmode, int vbr_q, float vbr_q_frac, int samplerate, int out_samplerate, int nch)
{
    struct refdump_handle *h = calloc(1, sizeof *h);
    lame_global_flags *gfp = lame_init();
    h->gfp = gfp;
    lame_set_in_samplerate(gfp, samplerate > 0 ? samplerate : 44100);
    lame_set_num_channels(gfp, nch > 0 ? nch : 2);
    if (out_samplerate > 0) lame_set_out_samplerate(gfp, out_samplerate);
    if (getenv("LP_CRC")) lame_set_error_protection(gfp, 1);
    if (vbrmode == vbr_abr) { lame_set_VBR(gfp, vbr_abr); if (brate > 0) lame_set_VBR_mean_bitrate_kbps(gfp, brate); }
    else if (vbrmode > 0) { lame_set_VBR(gfp, (vbr_mode) vbrmode); if (vbr_q_frac > 0) lame_set_VBR_quality(gfp, vbr_q + vbr_q_frac); else lame_set_VBR_q(gfp, vbr_q); }
    else if (brate > 0) lame_set_brate(gfp, brate);
    if (mode >= 0) lame_set_mode(gfp, (MPEG_mode) mode);
    if (quality >= 0) lame_set_quality(gfp, quality);
    lame_set_bWriteVbrTag(gfp, 0);
    if (lame_init_params(gfp) < 0) { lame_close(gfp); free(h); return NULL; }
    return h;
}

int refdump_encode(void *hv, const short *l, const short *r, int n, unsigned char *out, int cap)
{
    struct refdump_handle *h = hv;
    return lame_encode_buffer(h->gfp, l, r, n, out, cap);
}

int refdump_flush(void *hv, unsigned char *out, int cap)
{
    struct refdump_handle *h = hv;
    return lame_encode_flush(h->gfp, out, cap);
}

void refdump_close(void *hv)
{
    struct refdump_handle *h = hv;
    if (!h) return;
    lame_close(h->gfp);
    free(h);
}

int refdump_frame_number(void *hv)
{
    struct refdump_handle *h = hv;
    return h->gfp->internal_flags->ov_enc.frame_number;
}

void refdump_snapshot(void *hv, refdump_frame *d)
{
    struct refdump_handle *h = hv;
    lame_internal_flags const *gfc = h->gfp->internal_flags;
    int gr, ch, i, j;
    memset(d, 0, sizeof *d);
    d->frame_number = gfc->ov_enc.frame_number;
    d->padding = gfc->ov_enc.padding;
    d->mode_ext = gfc->ov_enc.mode_ext;
    d->main_data_begin = gfc->l3_side.main_data_begin;
    d->resv_size = gfc->sv_enc.ResvSize;
    d->resv_max = gfc->sv_enc.ResvMax;
    d->drain_pre = gfc->l3_side.resvDrain_pre;
    d->drain_post = gfc->l3_side.resvDrain_post;
    d->ath_adjust_factor = gfc->ATH->adjust_factor;
    d->ath_adjust_limit = gfc->ATH->adjust_limit;
    d->masking_lower = gfc->sv_qnt.masking_lower;
    d->slot_lag = gfc->sv_enc.slot_lag;
    for (i = 0; i < 19; i++) d->pefirbuf[i] = gfc->sv_enc.pefirbuf[i];
    for (ch = 0; ch < 2; ch++) {
        d->old_value[ch] = gfc->sv_qnt.OldValue[ch];
        d->current_step[ch] = gfc->sv_qnt.CurrentStep[ch];
        d->blocktype_old[ch] = gfc->sv_psy.blocktype_old[ch];
        d->loudness_sq_save[ch] = gfc->sv_psy.loudness_sq_save[ch];
        for (i = 0; i < 4; i++) d->scfsi[ch][i] = gfc->l3_side.scfsi[ch][i];
    }
    for (gr = 0; gr < 2; gr++)
        for (ch = 0; ch < 2; ch++) {
            gr_info const *gi = &gfc->l3_side.tt[gr][ch];
            refdump_granule *g = &d->gi[gr][ch];
            memcpy(g->xr, gi->xr, sizeof g->xr);
            memcpy(g->l3_enc, gi->l3_enc, sizeof g->l3_enc);
            memcpy(g->scalefac, gi->scalefac, sizeof g->scalefac);
            g->part2_3_length = gi->part2_3_length;
            g->big_values = gi->big_values;
            g->count1 = gi->count1;
            g->global_gain = gi->global_gain;
            g->scalefac_compress = gi->scalefac_compress;
            g->block_type = gi->block_type;
            g->mixed_block_flag = gi->mixed_block_flag;
            for (i = 0; i < 3; i++) { g->table_select[i] = gi->table_select[i]; g->subblock_gain[i] = gi->subblock_gain[i]; }
            g->region0_count = gi->region0_count;
            g->region1_count = gi->region1_count;
            g->preflag = gi->preflag;
            g->scalefac_scale = gi->scalefac_scale;
            g->count1table_select = gi->count1table_select;
            g->part2_length = gi->part2_length;
            g->sfbmax = gi->sfbmax;
            g->sfbdivide = gi->sfbdivide;
            g->psymax = gi->psymax;
            g->max_nonzero_coeff = gi->max_nonzero_coeff;
            g->count1bits = gi->count1bits;
            g->xrpow_max = gi->xrpow_max;
            g->loudness_sq = gfc->ov_psy.loudness_sq[gr][ch];
        }
    for (i = 0; i < 4; i++) {
        memcpy(d->en_l[i], gfc->sv_psy.en[i].l, sizeof d->en_l[i]);
        memcpy(d->thm_l[i], gfc->sv_psy.thm[i].l, sizeof d->thm_l[i]);
        memcpy(d->en_s[i], gfc->sv_psy.en[i].s, sizeof d->en_s[i]);
        memcpy(d->thm_s[i], gfc->sv_psy.thm[i].s, sizeof d->thm_s[i]);
        memcpy(d->nb_l1[i], gfc->sv_psy.nb_l1[i], sizeof d->nb_l1[i]);
        memcpy(d->nb_l2[i], gfc->sv_psy.nb_l2[i], sizeof d->nb_l2[i]);
        d->tot_ener[i] = gfc->sv_psy.tot_ener[i];
        d->last_attacks[i] = gfc->sv_psy.last_attacks[i];
        for (j = 0; j < 9; j++) d->last_en_subshort[i][j] = gfc->sv_psy.last_en_subshort[i][j];
    }
    memcpy(d->sb_sample, gfc->sv_enc.sb_sample, sizeof d->sb_sample);
}

static void copy_cb2sb(refdump_cb2sb *o, PsyConst_CB2SB_t const *p)
{
    int i, n = 0;
    memcpy(o->masking_lower, p->masking_lower, sizeof o->masking_lower);
    memcpy(o->minval, p->minval, sizeof o->minval);
    memcpy(o->rnumlines, p->rnumlines, sizeof o->rnumlines);
    memcpy(o->mld_cb, p->mld_cb, sizeof o->mld_cb);
    memcpy(o->mld, p->mld, sizeof o->mld);
    memcpy(o->bo_weight, p->bo_weight, sizeof o->bo_weight);
    memcpy(o->s3ind, p->s3ind, sizeof o->s3ind);
    memcpy(o->numlines, p->numlines, sizeof o->numlines);
    memcpy(o->bm, p->bm, sizeof o->bm);
    memcpy(o->bo, p->bo, sizeof o->bo);
    o->npart = p->npart;
    o->n_sb = p->n_sb;
    for (i = 0; i < p->npart; i++) n += p->s3ind[i][1] - p->s3ind[i][0] + 1;
    o->n_s3 = n;
    if (n > REFDUMP_MAX_S3) n = REFDUMP_MAX_S3;
    if (p->s3) memcpy(o->s3, p->s3, n * sizeof(float));
}

void refdump_tables(void *hv, refdump_tab *t)
{
    struct refdump_handle *h = hv;
    lame_internal_flags const *gfc = h->gfp->internal_flags;
    SessionConfig_t const *cfg = &gfc->cfg;
    int i;
    memset(t, 0, sizeof *t);
    copy_cb2sb(&t->l, &gfc->cd_psy->l);
    copy_cb2sb(&t->s, &gfc->cd_psy->s);
    copy_cb2sb(&t->l2s, &gfc->cd_psy->l_to_s);
    for (i = 0; i < 4; i++) t->attack_threshold[i] = gfc->cd_psy->attack_threshold[i];
    t->decay = gfc->cd_psy->decay;
    memcpy(t->ath_l, gfc->ATH->l, sizeof t->ath_l);
    memcpy(t->ath_s, gfc->ATH->s, sizeof t->ath_s);
    memcpy(t->ath_psfb21, gfc->ATH->psfb21, sizeof t->ath_psfb21);
    memcpy(t->ath_psfb12, gfc->ATH->psfb12, sizeof t->ath_psfb12);
    memcpy(t->ath_cb_l, gfc->ATH->cb_l, sizeof t->ath_cb_l);
    memcpy(t->ath_cb_s, gfc->ATH->cb_s, sizeof t->ath_cb_s);
    memcpy(t->eql_w, gfc->ATH->eql_w, sizeof t->eql_w);
    t->ath_floor = gfc->ATH->floor;
    t->ath_decay = gfc->ATH->decay;
    t->ath_aa_sensitivity_p = gfc->ATH->aa_sensitivity_p;
    t->ath_use_adjust = gfc->ATH->use_adjust;
    memcpy(t->longfact, gfc->sv_qnt.longfact, sizeof t->longfact);
    memcpy(t->shortfact, gfc->sv_qnt.shortfact, sizeof t->shortfact);
    for (i = 0; i < 576; i++) t->bv_scf[i] = gfc->sv_qnt.bv_scf[i];
    memcpy(t->amp_filter, gfc->sv_enc.amp_filter, sizeof t->amp_filter);
    memcpy(t->sfb_l, gfc->scalefac_band.l, sizeof t->sfb_l);
    memcpy(t->sfb_s, gfc->scalefac_band.s, sizeof t->sfb_s);
    memcpy(t->pow43, pow43, sizeof t->pow43);
    memcpy(t->adj43asm, adj43asm, sizeof t->adj43asm);
    memcpy(t->ipow20, ipow20, sizeof t->ipow20);
    memcpy(t->pow20, pow20, sizeof t->pow20);
    t->mask_adjust = gfc->sv_qnt.mask_adjust;
    t->mask_adjust_short = gfc->sv_qnt.mask_adjust_short;
    t->sfb21_extra = gfc->sv_qnt.sfb21_extra;
    t->substep_shaping = gfc->sv_qnt.substep_shaping;
    t->msfix = cfg->msfix;
    t->ath_offset_factor = cfg->ATH_offset_factor;
    t->ath_offset_db = cfg->ATH_offset_db;
    t->athfixpoint = cfg->ATHfixpoint;
    t->athcurve = cfg->ATHcurve;
    t->minval_cfg = cfg->minval;
    t->pcm_transform[0] = cfg->pcm_transform[0][0];
    t->pcm_transform[1] = cfg->pcm_transform[0][1];
    t->pcm_transform[2] = cfg->pcm_transform[1][0];
    t->pcm_transform[3] = cfg->pcm_transform[1][1];
    t->noise_shaping = cfg->noise_shaping;
    t->noise_shaping_amp = cfg->noise_shaping_amp;
    t->noise_shaping_stop = cfg->noise_shaping_stop;
    t->subblock_gain = cfg->subblock_gain;
    t->use_best_huffman = cfg->use_best_huffman;
    t->full_outer_loop = cfg->full_outer_loop;
    t->quant_comp = cfg->quant_comp;
    t->quant_comp_short = cfg->quant_comp_short;
    t->use_temporal = cfg->use_temporal_masking_effect;
    t->short_blocks = cfg->short_blocks;
    t->mode = cfg->mode;
    t->force_ms = cfg->force_ms;
    t->sideinfo_len = cfg->sideinfo_len;
    t->avg_bitrate = cfg->avg_bitrate;
    t->bitrate_index = gfc->ov_enc.bitrate_index;
    t->samplerate_out = cfg->samplerate_out;
    t->buffer_constraint = cfg->buffer_constraint;
    t->frac_spf = gfc->sv_enc.frac_SpF;
    t->lowpass1 = cfg->lowpass1;
    t->lowpass2 = cfg->lowpass2;
    t->vbr = cfg->vbr;
    t->disable_reservoir = cfg->disable_reservoir;
    t->interch = cfg->interChRatio;
    t->athtype = cfg->ATHtype;
}
