/* Test harness: pins oracle/port against the unmodified reference (oracle/_ref/refdump.so).
 * usage: port_vs_ref <signal> <brate> <mode> <quality> <nframes> [samplerate] [wavfile]
 *   signal: noise | sine | click | silence | wav
 * Compares init tables, then MP3 bytes frame by frame, and on the first mismatch the per-frame state.
 * Exit code 0 = byte-identical. Needs /root/reference-built refdump.so, so CPU container only. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
#include "../../oracle/refdump.h"
#include "../../oracle/port/lame_port.h"
#include "siggen.h"

#define CMPF(name, a, b, n) do { int k_, bad_ = 0; for (k_ = 0; k_ < (n); k_++) if (memcmp(&(a)[k_], &(b)[k_], 4)) { if (bad_ < 3) printf("  MISMATCH %s[%d]: ref %.9g port %.9g\n", name, k_, (double) (a)[k_], (double) (b)[k_]); bad_++; } if (bad_) { printf("  %s: %d/%d differ\n", name, bad_, (int) (n)); nbad++; } } while (0)
#define CMPI(name, a, b, n) do { int k_, bad_ = 0; for (k_ = 0; k_ < (n); k_++) if ((a)[k_] != (b)[k_]) { if (bad_ < 3) printf("  MISMATCH %s[%d]: ref %d port %d\n", name, k_, (int) (a)[k_], (int) (b)[k_]); bad_++; } if (bad_) { printf("  %s: %d/%d differ\n", name, bad_, (int) (n)); nbad++; } } while (0)

static int cmp_bands(const char *tag, const refdump_cb2sb *r, const lp_bands *p)
{
    int nbad = 0; char nm[64];
    printf(" bands %s: npart ref %d port %d, n_s3 ref %d port %d\n", tag, r->npart, p->npart, r->n_s3, p->n_s3);
    if (r->npart != p->npart || r->n_s3 != p->n_s3) nbad++;
#define B(f, n) snprintf(nm, sizeof nm, "%s." #f, tag); CMPF(nm, r->f, p->f, n)
#define BI(f, n) snprintf(nm, sizeof nm, "%s." #f, tag); CMPI(nm, r->f, p->f, n)
    B(masking_lower, 64); B(minval, r->npart); B(rnumlines, r->npart); B(mld_cb, 64); B(bo_weight, r->n_sb);
    BI(numlines, r->npart); BI(bo, r->n_sb); BI(bm, r->n_sb);
    { int i; for (i = 0; i < r->npart; i++) if (r->s3ind[i][0] != p->s3ind[i][0] || r->s3ind[i][1] != p->s3ind[i][1]) { printf("  s3ind[%d] differs\n", i); nbad++; break; } }
    B(s3, r->n_s3);
    return nbad;
}

int main(int argc, char **argv)
{
    const char *sig = argc > 1 ? argv[1] : "noise";
    int brate = argc > 2 ? atoi(argv[2]) : 128, mode = argc > 3 ? atoi(argv[3]) : -1;
    int quality = argc > 4 ? atoi(argv[4]) : -1, nframes = argc > 5 ? atoi(argv[5]) : 50;
    int sr = argc > 6 ? atoi(argv[6]) : 44100;
    const char *wav = argc > 7 ? argv[7] : NULL;
    int n = nframes * 1152, f, nbad = 0, total_ref = 0, total_port = 0, first_bad = -1;
    short *l = malloc(n * 2), *r = malloc(n * 2);
    static unsigned char ob_ref[65536], ob_port[65536];
    static refdump_tab tab; static refdump_frame fr;
    void *h; lp_encoder *e;
    if (siggen(sig, l, r, n, sr, wav) < 0) { printf("bad signal\n"); return 2; }
    int const vbr = getenv("LP_VBR") ? atoi(getenv("LP_VBR")) : 0;          /* 0 = CBR, 3 = ABR with mean bitrate `brate` */
    int const out_sr = getenv("LP_OUT_SR") ? atoi(getenv("LP_OUT_SR")) : 0; /* explicit output rate (lame_set_out_samplerate), 0 = automatic */
    int const chunk = getenv("LP_CHUNK") ? atoi(getenv("LP_CHUNK")) : 1152;  /* samples per encode call: the resampler's state depends on it */
    float const qfrac = getenv("LP_VBRQ_FRAC") ? (float) atof(getenv("LP_VBRQ_FRAC")) : 0.f;  /* VBR quality = brate + this (lame_set_VBR_quality) */
    h = refdump_open_vq(brate, mode, quality, vbr, (vbr == 4 || vbr == 2) ? brate : 0, qfrac, sr, out_sr, 2);   /* vbr 4 (vbr_mtrh): `brate` is VBR_q */
    /* the same split of the float level as lame_set_VBR_quality (set_get.c:1169) */
    e = lp_open_vq(sr, out_sr, 2, brate, mode < 0 ? LP_MODE_NOT_SET : mode, quality, vbr, (vbr == 4 || vbr == 2) ? ((float) brate + qfrac) - (float) brate : 0.f);
    if (e && getenv("LP_CRC")) lp_set_error_protection(e);            /* refdump_open reads the same variable */
    if (!h || !e) { printf("open failed ref=%p port=%p\n", h, (void *) e); return (!h && !e) ? 0 : 2; }
    refdump_tables(h, &tab);
    {
        const lp_config *c = &e->cfg;
        nbad += cmp_bands("l", &tab.l, &c->l); nbad += cmp_bands("s", &tab.s, &c->s); nbad += cmp_bands("l2s", &tab.l2s, &c->l2s);
        CMPF("attack_threshold", tab.attack_threshold, c->attack_threshold, 4);
        CMPF("decay", (&tab.decay), (&c->decay), 1);
        CMPF("ath_l", tab.ath_l, c->ath_l, 22); CMPF("ath_s", tab.ath_s, c->ath_s, 13);
        CMPF("ath_cb_l", tab.ath_cb_l, c->ath_cb_l, tab.l.npart); CMPF("ath_cb_s", tab.ath_cb_s, c->ath_cb_s, tab.s.npart);
        CMPF("eql_w", tab.eql_w, c->eql_w, 512); CMPF("ath_floor", (&tab.ath_floor), (&c->ath_floor), 1);
        CMPF("longfact", tab.longfact, c->longfact, 22); CMPF("shortfact", tab.shortfact, c->shortfact, 13);
        CMPI("bv_scf", tab.bv_scf, c->bv_scf, 576); CMPF("amp_filter", tab.amp_filter, c->amp_filter, 32);
        CMPI("sfb_l", tab.sfb_l, c->sfb_l, 23); CMPI("sfb_s", tab.sfb_s, c->sfb_s, 14);
        CMPF("pow43", tab.pow43, c->pow43, 8208); CMPF("adj43asm", tab.adj43asm, c->adj43asm, 8208);
        CMPF("ipow20", tab.ipow20, c->ipow20, 257); CMPF("pow20", tab.pow20, c->pow20, 374);
        CMPF("mask_adjust", (&tab.mask_adjust), (&c->mask_adjust), 1); CMPF("mask_adjust_short", (&tab.mask_adjust_short), (&c->mask_adjust_short), 1);
        CMPF("msfix", (&tab.msfix), (&c->msfix), 1); CMPF("ath_offset_factor", (&tab.ath_offset_factor), (&c->ath_offset_factor), 1);
        CMPF("athfixpoint", (&tab.athfixpoint), (&c->athfixpoint), 1); CMPF("pcm_transform", tab.pcm_transform, (&c->pcm_transform[0][0]), 4);
        CMPI("noise_shaping", (&tab.noise_shaping), (&c->noise_shaping), 1); CMPI("noise_shaping_amp", (&tab.noise_shaping_amp), (&c->noise_shaping_amp), 1);
        CMPI("subblock_gain", (&tab.subblock_gain), (&c->subblock_gain), 1); CMPI("use_best_huffman", (&tab.use_best_huffman), (&c->use_best_huffman), 1);
        CMPI("full_outer_loop", (&tab.full_outer_loop), (&c->full_outer_loop), 1); CMPI("short_blocks", (&tab.short_blocks), (&c->short_blocks), 1);
        CMPI("mode", (&tab.mode), (&c->mode), 1); CMPI("sideinfo_len", (&tab.sideinfo_len), (&c->sideinfo_len), 1);
        CMPI("bitrate_index", (&tab.bitrate_index), (&c->bitrate_index), 1); CMPI("frac_spf", (&tab.frac_spf), (&c->frac_spf), 1);
        CMPI("buffer_constraint", (&tab.buffer_constraint), (&c->buffer_constraint), 1); CMPI("use_temporal", (&tab.use_temporal), (&c->use_temporal), 1);
        CMPI("sfb21_extra", (&tab.sfb21_extra), (&c->sfb21_extra), 1); CMPI("quant_comp", (&tab.quant_comp), (&c->quant_comp), 1);
        printf("setup tables: %s\n", nbad ? "MISMATCH" : "identical");
    }
    int const ncalls = (n + chunk - 1) / chunk;
    for (f = 0; f <= ncalls; f++) {
        int nr, np, gr, ch;
        if (f < ncalls) {
            int const cn = (f + 1) * chunk <= n ? chunk : n - f * chunk;
            nr = refdump_encode(h, l + f * chunk, r + f * chunk, cn, ob_ref, sizeof ob_ref);
            np = lp_encode(e, l + f * chunk, r + f * chunk, cn, ob_port, sizeof ob_port);
        }
        else {
            nr = refdump_flush(h, ob_ref, sizeof ob_ref);
            np = lp_flush(e, ob_port, sizeof ob_port);
        }
        total_ref += nr > 0 ? nr : 0; total_port += np > 0 ? np : 0;
        if (nr != np || (nr > 0 && memcmp(ob_ref, ob_port, nr))) {
            if (first_bad < 0) {
                first_bad = f;
                printf("frame call %d: bytes ref %d port %d -> MISMATCH\n", f, nr, np);
                refdump_snapshot(h, &fr);
                { int a[3] = { fr.mode_ext, fr.padding, fr.resv_size }, b[3] = { e->mode_ext, e->padding, e->resv_size }; CMPI("mode_ext/padding/resv", a, b, 3); }
                CMPF("ath_adjust_factor", (&fr.ath_adjust_factor), (&e->ath_adjust_factor), 1);
                CMPF("pefirbuf", fr.pefirbuf, e->pefirbuf, 19);
                CMPF("sb_sample", (&fr.sb_sample[0][0][0][0]), (&e->sb_sample[0][0][0][0]), 2 * 2 * 18 * 32);
                { int i; for (i = 0; i < 4; i++) { CMPF("en_l", fr.en_l[i], e->psy.en[i].l, 22); CMPF("thm_l", fr.thm_l[i], e->psy.thm[i].l, 22);
                    CMPF("en_s", (&fr.en_s[i][0][0]), (&e->psy.en[i].s[0][0]), 39); CMPF("thm_s", (&fr.thm_s[i][0][0]), (&e->psy.thm[i].s[0][0]), 39);
                    CMPF("nb_l1", fr.nb_l1[i], e->psy.nb_l1[i], 64); } }
                CMPF("tot_ener", fr.tot_ener, e->psy.tot_ener, 4);
                CMPI("last_attacks", fr.last_attacks, e->psy.last_attacks, 4);
                for (gr = 0; gr < 2; gr++) for (ch = 0; ch < 2; ch++) {
                    refdump_granule *g = &fr.gi[gr][ch]; lp_granule *p = &e->tt[gr][ch];
                    int a[10] = { g->block_type, g->global_gain, g->part2_3_length, g->part2_length, g->big_values, g->count1, g->scalefac_compress, g->scalefac_scale, g->preflag, g->max_nonzero_coeff };
                    int b[10] = { p->block_type, p->global_gain, p->part2_3_length, p->part2_length, p->big_values, p->count1, p->scalefac_compress, p->scalefac_scale, p->preflag, p->max_nonzero_coeff };
                    printf(" gr %d ch %d\n", gr, ch);
                    CMPI("bt/gg/p23/p2/bv/c1/sfc/sfs/pre/mnz", a, b, 10);
                    CMPF("xr", g->xr, p->xr, 576); CMPI("l3_enc", g->l3_enc, p->l3_enc, 576); CMPI("scalefac", g->scalefac, p->scalefac, 39);
                    CMPI("table_select", g->table_select, p->table_select, 3);
                }
            }
            nbad++;
        }
    }
    printf("%s br=%d mode=%d q=%d sr=%d frames=%d: ref %d bytes, port %d bytes, %s\n", sig, brate, mode, quality, sr, nframes,
           total_ref, total_port, nbad ? "MISMATCH" : "IDENTICAL");
    return nbad ? 1 : 0;
}
