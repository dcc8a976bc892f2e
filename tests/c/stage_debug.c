/* Development aid: one stream, F frames in ONE launch, compare every intermediate device buffer with the
 * oracle port's state after the same frames.  usage: stage_debug <signal> <frames> <brate> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/lamegpu.h"
#include "../../oracle/port/lame_port.h"
#include "../../deprecated-lame-mirror_b200/csrc/lg_types.h"
#include "siggen.h"
#define CMPF(name, a, b, n) do { int k_, bad_ = 0; for (k_ = 0; k_ < (n); k_++) if (memcmp(&(a)[k_], &(b)[k_], 4)) { if (bad_ < 3) printf("    MISMATCH %s[%d]: port %.9g dev %.9g\n", name, k_, (double) (a)[k_], (double) (b)[k_]); bad_++; } if (bad_) { printf("    %s: %d/%d differ\n", name, bad_, (int) (n)); nbad++; } } while (0)
#define CMPI(name, a, b) do { if ((int) (a) != (int) (b)) { printf("    MISMATCH %s: port %d dev %d\n", name, (int) (a), (int) (b)); nbad++; } } while (0)
int main(int argc, char **argv)
{
    const char *sig = argc > 1 ? argv[1] : "noise";
    int F = argc > 2 ? atoi(argv[2]) : 3, brate = argc > 3 ? atoi(argv[3]) : 128, mode = argc > 4 ? atoi(argv[4]) : -1;
    int n = F * 1152 + 752 + 1152, f, gr, ch, nbad = 0, i;
    short *l = malloc(n * 2), *r = malloc(n * 2);
    static unsigned char ob[1 << 20];
    const short *pl[1], *pr[1]; int ns[1], oc[1] = { sizeof ob }, obn[1]; unsigned char *po[1] = { ob };
    lamegpu_batch *b = lamegpu_batch_open(44100, 2, brate, mode, -1, 1, F, 0);
    lp_encoder *e = lp_open(44100, 2, brate, mode < 0 ? LP_MODE_NOT_SET : mode, -1);
    float *sb = malloc((2 * F + 1) * 2 * 576 * 4), *xr = malloc(2 * F * 2 * 576 * 4);
    LgPsyOut *psy = malloc(2 * F * sizeof *psy); LgFrameCtl *frm = malloc(F * sizeof *frm);
    LgGranuleOut *go = malloc(2 * F * 2 * sizeof *go); LgFrameOut *fo = malloc(F * sizeof *fo);
    static LgStreamState st;
    siggen(sig, l, r, n, 44100, NULL);
    /* feed exactly enough for F frames: timeline needs 1152*(F-1)+1904 samples = user 1152*(F-1)+1376 */
    pl[0] = l; pr[0] = r; ns[0] = 1152 * (F - 1) + 1376;
    printf("frames encoded by device: %ld\n", lamegpu_batch_encode(b, pl, pr, ns, po, oc, obn));
    lamegpu_batch_debug_copy(b, 0, sb, (2 * F + 1) * 2 * 576 * 4); lamegpu_batch_debug_copy(b, 4, xr, 2 * F * 2 * 576 * 4);
    lamegpu_batch_debug_copy(b, 2, psy, 2 * F * sizeof *psy); lamegpu_batch_debug_copy(b, 3, frm, F * sizeof *frm);
    lamegpu_batch_debug_copy(b, 5, go, 2 * F * 2 * sizeof *go); lamegpu_batch_debug_copy(b, 6, fo, F * sizeof *fo);
    lamegpu_batch_debug_copy(b, 7, &st, sizeof st);
    for (f = 0; f < F; f++) {
        static unsigned char tmp[65536];
        int before = nbad;
        /* port: feed so that exactly one more frame is produced */
        lp_encode(e, l + (f == 0 ? 0 : 1376 + 1152 * (f - 1)), r + (f == 0 ? 0 : 1376 + 1152 * (f - 1)), f == 0 ? 1376 : 1152, tmp, sizeof tmp);
        printf("frame %d (port frame_number %d)\n", f, e->frame_number);
        if (getenv("LP_DEBUG")) printf("dev  frame %d: pe %g %g %g %g peMS %g %g %g %g bt %d %d %d %d\n", f, psy[2*f].pe[0], psy[2*f].pe[1], psy[2*f+1].pe[0], psy[2*f+1].pe[1], psy[2*f].pe[2], psy[2*f].pe[3], psy[2*f+1].pe[2], psy[2*f+1].pe[3], psy[2*f].block_type[0], psy[2*f].block_type[1], psy[2*f+1].block_type[0], psy[2*f+1].block_type[1]);
        { extern lp_xmin lp_dbg_en_after[2][4], lp_dbg_thm_after[2][4]; int c4;
          /* ratios delivered for (f, gr1) are the state after call (f, gr0) */
          for (c4 = 0; c4 < 4; c4++) { CMPF("ratio.en_l(after gr0)", lp_dbg_en_after[0][c4].l, psy[2 * f + 1].en[c4].l, 22); CMPF("ratio.thm_l(after gr0)", lp_dbg_thm_after[0][c4].l, psy[2 * f + 1].thm[c4].l, 22);
            CMPF("ratio.en_s(after gr0)", (&lp_dbg_en_after[0][c4].s[0][0]), (&psy[2 * f + 1].en[c4].s[0][0]), 39); CMPF("ratio.thm_s(after gr0)", (&lp_dbg_thm_after[0][c4].s[0][0]), (&psy[2 * f + 1].thm[c4].s[0][0]), 39); } }
        CMPI("mode_ext", e->mode_ext, frm[f].mode_ext); CMPI("padding", e->padding, frm[f].padding);
        CMPF("ath_adjust", (&e->ath_adjust_factor), (&frm[f].ath_adjust_factor), 1);
        CMPF("pe_use", (&e->last_pe[0][0]), (&frm[f].pe_use[0][0]), 4);
        CMPI("resv_size", e->resv_size, fo[f].resv_size); CMPI("drain_post", e->drain_post, fo[f].drain_post);
        for (gr = 0; gr < 2; gr++) for (ch = 0; ch < 2; ch++) {
            lp_granule *p = &e->tt[gr][ch]; LgGranuleOut *g = &go[(2 * f + gr) * 2 + ch];
            char nm[64];
            printf("  gr %d ch %d\n", gr, ch);
            /* port sb_sample[ch][1-gr] holds granule gr of this frame */
            CMPF("sb", (&e->sb_sample[ch][1 - gr][0][0]), (sb + ((2 * f + gr + 1) * 2 + ch) * 576), 576);
            CMPI("block_type", p->block_type, psy[2 * f + gr].block_type[ch]);
            CMPF("xr", p->xr, (xr + ((2 * f + gr) * 2 + ch) * 576), 576);
            CMPI("global_gain", p->global_gain, g->global_gain); CMPI("part2_3_length", p->part2_3_length, g->part2_3_length);
            CMPI("part2_length", p->part2_length, g->part2_length); CMPI("big_values", p->big_values, g->big_values);
            CMPI("count1", p->count1, g->count1); CMPI("scalefac_compress", p->scalefac_compress, g->scalefac_compress);
            CMPI("scalefac_scale", p->scalefac_scale, g->scalefac_scale); CMPI("preflag", p->preflag, g->preflag);
            CMPI("count1table", p->count1table_select, g->count1table_select);
            CMPI("region0", p->region0_count, g->region0_count); CMPI("region1", p->region1_count, g->region1_count);
            for (i = 0; i < 3; i++) { snprintf(nm, sizeof nm, "table_select[%d]", i); CMPI(nm, p->table_select[i], g->table_select[i]); }
            { int bad = 0; for (i = 0; i < 576; i++) { int v = p->l3_enc[i]; if (p->xr[i] < 0) v = -v; if (v != g->ix[i]) { if (bad < 3) printf("    MISMATCH ix[%d]: port %d dev %d\n", i, v, g->ix[i]); bad++; } } if (bad) { printf("    ix: %d differ\n", bad); nbad++; } }
            { int bad = 0; for (i = 0; i < 39; i++) if (p->scalefac[i] != g->scalefac[i]) { if (bad < 3) printf("    MISMATCH scalefac[%d]: port %d dev %d\n", i, p->scalefac[i], g->scalefac[i]); bad++; } if (bad) nbad++; }
        }
        if (nbad != before) { printf("first mismatching frame: %d\n", f); if (!getenv("STATE")) break; }
    }
    if (getenv("STATE") || !nbad) {
        for (i = 0; i < 4; i++) { CMPF("state.en_s", (&e->psy.en[i].s[0][0]), (&st.en[i].s[0][0]), 39); CMPI("state.last_attacks", e->psy.last_attacks[i], st.last_attacks[i]); CMPF("state.last_en_subshort", e->psy.last_en_subshort[i], st.last_en_subshort[i], 9); CMPF("state.nb_l2", e->psy.nb_l2[i], st.nb_l2[i], 64); CMPF("state.en_l", e->psy.en[i].l, st.en[i].l, 22); CMPF("state.thm_l", e->psy.thm[i].l, st.thm[i].l, 22);
            CMPF("state.thm_s", (&e->psy.thm[i].s[0][0]), (&st.thm[i].s[0][0]), 39); CMPF("state.nb_l1", e->psy.nb_l1[i], st.nb_l1[i], 64); }
    }
    printf("%s\n", nbad ? "STAGE MISMATCH" : "all stages identical");
    return nbad != 0;
}
