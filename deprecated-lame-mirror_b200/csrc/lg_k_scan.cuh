// lg_k_scan.cuh - kernel B: the ordered part of the psycho-acoustic model (SURVEY.md section 7,
// "Psy stage B").  Everything here depends on what earlier granules of the same stream decided:
// attack/block-type state machine, pre-echo limits (nb_l1/nb_l2), the one-granule delay of en/thm,
// ATH auto-adjust, the M/S decision and the 19-frame PE FIR.  It is cheap (a few thousand cycles per
// granule), so one warp walks one stream's granules in order; partitions / scalefactor bands /
// channels are spread over lanes, and every serial sum of the reference stays serial inside one lane.
//
// Reference: L3psycho_anal_vbr psymodel.c:1397 (minus the stateless parts done by kernel A),
// lame_encode_mp3_frame encoder.c:305 stages 1 and 3, adjust_ATH encoder.c:56.
#pragma once
#include <stddef.h>
#include "lg_math.cuh"

struct LgSmemB {
    LgAnalysis A;                     /* first: copied with 8-byte accesses (the dynamic smem base is 16-byte aligned) */
    float eb[4][LG_CBANDS], thr[4][LG_CBANDS];
    float last_thm_s[4][LG_SBMAX_S][3];
    float ssf[4][3];
    int   ns_att[4][4];
    int   nsuse[4];
    float pe[4];
    float frame_pe[2][4];
    float frame_tot[2][4];
    int   frame_bt[2][2];
    float bcast[4];
    /* convert_partition2scalefac's walk over the partitions is fixed by the band tables: where every scalefactor band's sum starts and
     * ends, so that the bands can be summed side by side (lg_p2s_plan / lg_p2s_band); one plan each for long, long->short, short */
    uint8_t p2s_prev[3][LG_SBMAX_L], p2s_first[3][LG_SBMAX_L], p2s_cur[3][LG_SBMAX_L], p2s_kind[3][LG_SBMAX_L];
    LgXmin pe_en[4], pe_thm[4];       /* the delayed ratios handed to the caller, kept for the perceptual entropy */
    float log_table[513];             /* LgDevCfg.log_table (util.c:977 fast_log2) */
    /* the stream's carried state and the current granule's analysis record live in shared memory while the warp walks
     * the stream: the scan is a chain of dependent steps, so every global load would be paid in full latency */
    LgStreamState st;
};

/* psymodel.c:350 convert_partition2scalefac, run serially by one lane */
__device__ __forceinline__ void lg_partition2sfb(const LgBands *__restrict__ gd, const float *eb, const float *thr,
                                                 float *enn_out, float *thm_out, int out_stride)
{
    float enn = 0.0f, thmm = 0.0f;
    int sb, b, n = gd->n_sb;
    for (sb = b = 0; sb < n; ++b, ++sb) {
        int const bo_sb = gd->bo[sb];
        int const npart = gd->npart;
        int const b_lim = bo_sb < npart ? bo_sb : npart;
        while (b < b_lim) { enn += eb[b]; thmm += thr[b]; b++; }
        if (b >= npart) { enn_out[sb * out_stride] = enn; thm_out[sb * out_stride] = thmm; ++sb; break; }
        float const w_curr = gd->bo_weight[sb];
        float const w_next = 1.0f - w_curr;
        enn += w_curr * eb[b];
        thmm += w_curr * thr[b];
        enn_out[sb * out_stride] = enn;
        thm_out[sb * out_stride] = thmm;
        enn = w_next * eb[b];
        thmm = w_next * thr[b];
    }
    for (; sb < n; ++sb) { enn_out[sb * out_stride] = 0; thm_out[sb * out_stride] = 0; }
}

/* The same sums, one scalefactor band at a time.  The serial walk above visits the partitions in a way that depends only on bo[] and
 * npart: band sb starts from the boundary partition of the band before it (weight 1 - bo_weight[sb-1]), adds the whole partitions
 * [first, cur) one by one in ascending order and ends with bo_weight[sb] of partition cur - or, for the band that reaches npart, without
 * that last term; the bands after it are zero.  lg_p2s_plan replays the integer part of the walk once per kernel; lg_p2s_band then
 * does the floating-point part of one band in exactly the order of the serial loop, so the bands can go to different lanes. */
enum { LG_P2S_NORMAL = 0, LG_P2S_LAST = 1, LG_P2S_ZERO = 2 };
__device__ __forceinline__ void lg_p2s_plan(const LgBands *__restrict__ gd, uint8_t *prev, uint8_t *first, uint8_t *cur, uint8_t *kind)
{
    int const n = gd->n_sb, npart = gd->npart;
    int sb, b;
    for (sb = b = 0; sb < n; ++b, ++sb) {
        int const bo_sb = gd->bo[sb];
        int const b_lim = bo_sb < npart ? bo_sb : npart;
        prev[sb] = (uint8_t) (b > 0 ? b - 1 : 0);
        first[sb] = (uint8_t) b;
        if (b < b_lim) b = b_lim;
        cur[sb] = (uint8_t) b;
        if (b >= npart) { kind[sb] = LG_P2S_LAST; ++sb; break; }
        kind[sb] = LG_P2S_NORMAL;
    }
    for (; sb < n; ++sb) { prev[sb] = first[sb] = cur[sb] = 0; kind[sb] = LG_P2S_ZERO; }
}
__device__ __forceinline__ void lg_p2s_band(const LgBands *__restrict__ gd, const uint8_t *prev, const uint8_t *first, const uint8_t *cur,
                                            const uint8_t *kind, const float *eb, const float *thr, int sb, float *enn_out, float *thm_out)
{
    int const k = kind[sb];
    float enn = 0.0f, thmm = 0.0f;
    if (k != LG_P2S_ZERO) {
        if (sb > 0) {
            float const w_next = 1.0f - gd->bo_weight[sb - 1];
            enn = w_next * eb[prev[sb]];
            thmm = w_next * thr[prev[sb]];
        }
        int const c = cur[sb];
        for (int b = first[sb]; b < c; b++) { enn += eb[b]; thmm += thr[b]; }
        if (k == LG_P2S_NORMAL) {
            float const w_curr = gd->bo_weight[sb];
            enn += w_curr * eb[c];
            thmm += w_curr * thr[c];
        }
    }
    *enn_out = enn; *thm_out = thmm;
}

/* psymodel.c:503 pecalc_l / :458 pecalc_s */
__device__ __forceinline__ float lg_pecalc_l(const float *log_table, const LgXmin *en, const LgXmin *thm, float masking_lower)
{
    const float regcoef_l[21] = { 6.8, 5.8, 5.8, 6.4, 6.5, 9.9, 12.1, 14.4, 15, 18.9, 21.6, 26.9, 34.2, 40.2,
        46.8, 56.5, 60.7, 73.9, 85.7, 93.4, 126.1 };
    float pe_l = 1124.23f / 4;
    for (int sb = 0; sb < LG_SBMAX_L - 1; sb++) {
        float const t = thm->l[sb];
        if (t > 0.0f) {
            float const x = t * masking_lower;
            float const e = en->l[sb];
            if (e > x) {
                if (e > x * 1e10f) pe_l = (float) (pe_l + regcoef_l[sb] * (10.0f * LG_LOG10_D));
                else pe_l = (float) (pe_l + regcoef_l[sb] * LG_FAST_LOG10_SMEM_D(log_table, e / x));
            }
        }
    }
    return pe_l;
}
__device__ __forceinline__ float lg_pecalc_s(const float *log_table, const LgXmin *en, const LgXmin *thm, float masking_lower)
{
    const float regcoef_s[12] = { 11.8, 13.6, 17.2, 32, 46.5, 51.3, 57.5, 67.1, 71.5, 84.6, 97.6, 130 };
    float pe_s = 1236.28f / 4;
    for (int sb = 0; sb < LG_SBMAX_S - 1; sb++)
        for (int sblock = 0; sblock < 3; sblock++) {
            float const t = thm->s[sb][sblock];
            if (t > 0.0f) {
                float const x = t * masking_lower;
                float const e = en->s[sb][sblock];
                if (e > x) {
                    if (e > x * 1e10f) pe_s = (float) (pe_s + regcoef_s[sb] * (10.0f * LG_LOG10_D));
                    else pe_s = (float) (pe_s + regcoef_s[sb] * LG_FAST_LOG10_SMEM_D(log_table, e / x));
                }
            }
        }
    return pe_s;
}

/* psymodel.c:1326 vbrpsy_compute_MS_thresholds for partition b */
__device__ __forceinline__ void lg_ms_threshold(float (*eb)[LG_CBANDS], float (*thr)[LG_CBANDS], float mld, float ath_cb,
                                                float athlower, float msfix, int b)
{
    float const msfix2 = msfix * 2.f;
    float rside, rmid;
    float const ebM = eb[2][b], ebS = eb[3][b], thmL = thr[0][b], thmR = thr[1][b];
    float thmM = thr[2][b], thmS = thr[3][b];
    if (thmL <= 1.58f * thmR && thmR <= 1.58f * thmL) {
        float const mld_m = mld * ebS, mld_s = mld * ebM;
        float const tmp_m = thmS < mld_m ? thmS : mld_m;
        float const tmp_s = thmM < mld_s ? thmM : mld_s;
        rmid = thmM > tmp_m ? thmM : tmp_m;
        rside = thmS > tmp_s ? thmS : tmp_s;
    }
    else { rmid = thmM; rside = thmS; }
    if (msfix > 0.f) {
        float const ath = ath_cb * athlower;
        float const tmp_l = thmL > ath ? thmL : ath;
        float const tmp_r = thmR > ath ? thmR : ath;
        float const thmLR = tmp_l < tmp_r ? tmp_l : tmp_r;
        thmM = rmid > ath ? rmid : ath;
        thmS = rside > ath ? rside : ath;
        float const thmMS = thmM + thmS;
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

/* psymodel.c:759 vbrpsy_attack_detection, the per-channel decision part (one lane per channel) */
__device__ __forceinline__ void lg_attack_channel(const LgDevCfg *__restrict__ c, LgStreamState *st, const float *newsub /*9*/,
                                                  int chn, int ns_attacks[4], float ssf[3], int *uselong)
{
    float attack_intensity[12], en_subshort[12], en_short[4] = { 0, 0, 0, 0 };
    int i;
    for (i = 0; i < 4; i++) ns_attacks[i] = 0;
    for (i = 0; i < 3; i++) {
        en_subshort[i] = st->last_en_subshort[chn][i + 6];
        attack_intensity[i] = en_subshort[i] / st->last_en_subshort[chn][i + 4];
        en_short[0] += en_subshort[i];
    }
    for (i = 0; i < 9; i++) {
        float p = newsub[i];
        st->last_en_subshort[chn][i] = en_subshort[i + 3] = p;
        en_short[1 + i / 3] += p;
        if (p > en_subshort[i + 3 - 2]) p = p / en_subshort[i + 3 - 2];
        else if (en_subshort[i + 3 - 2] > p * 10.0f) p = en_subshort[i + 3 - 2] / (p * 10.0f);
        else p = 0.0f;
        attack_intensity[i + 3] = p;
    }
    for (i = 0; i < 3; ++i) {
        float const enn = en_subshort[i * 3 + 3] + en_subshort[i * 3 + 4] + en_subshort[i * 3 + 5];
        float factor = 1.f;
        if (en_subshort[i * 3 + 5] * 6 < enn) {
            factor *= 0.5f;
            if (en_subshort[i * 3 + 4] * 6 < enn) factor *= 0.5f;
        }
        ssf[i] = factor;
    }
    {
        float const x = c->attack_threshold[chn];
        for (i = 0; i < 12; i++)
            if (ns_attacks[i / 3] == 0 && attack_intensity[i] > x) ns_attacks[i / 3] = (i % 3) + 1;
    }
    for (i = 1; i < 4; i++) {
        float const u = en_short[i - 1], v = en_short[i];
        float const m = u > v ? u : v;
        if (m < 40000) {
            if (u < 1.7f * v && v < 1.7f * u) {
                if (i == 1 && ns_attacks[0] <= ns_attacks[i]) ns_attacks[0] = 0;
                ns_attacks[i] = 0;
            }
        }
    }
    if (ns_attacks[0] <= st->last_attacks[chn]) ns_attacks[0] = 0;
    int ul = 1;
    if (st->last_attacks[chn] == 3 || ns_attacks[0] + ns_attacks[1] + ns_attacks[2] + ns_attacks[3]) {
        ul = 0;
        if (ns_attacks[1] && ns_attacks[0]) ns_attacks[1] = 0;
        if (ns_attacks[2] && ns_attacks[1]) ns_attacks[2] = 0;
        if (ns_attacks[3] && ns_attacks[2]) ns_attacks[3] = 0;
    }
    *uselong = ul;
}

/* encoder.c:56 adjust_ATH (lane 0) */
__device__ __forceinline__ void lg_adjust_ath(const LgDevCfg *__restrict__ cfg, LgStreamState *st, float (*loud)[2])
{
    float gr2_max, max_pow;
    if (cfg->ath_use_adjust == 0) { st->ath_adjust_factor = 1.0f; return; }
    max_pow = loud[0][0];
    gr2_max = loud[1][0];
    if (cfg->channels == 2) { max_pow += loud[0][1]; gr2_max += loud[1][1]; }
    else { max_pow += max_pow; gr2_max += gr2_max; }
    if (cfg->mode_gr == 2) max_pow = max_pow > gr2_max ? max_pow : gr2_max;
    max_pow = (float) (max_pow * 0.5);
    max_pow *= cfg->ath_aa_sensitivity_p;
    if (max_pow > 0.03125) {
        if (st->ath_adjust_factor >= 1.0) st->ath_adjust_factor = 1.0f;
        else if (st->ath_adjust_factor < st->ath_adjust_limit) st->ath_adjust_factor = st->ath_adjust_limit;
        st->ath_adjust_limit = 1.0f;
    }
    else {
        float const adj_lim_new = (float) (31.98 * max_pow + 0.000625);
        if (st->ath_adjust_factor >= adj_lim_new) {
            st->ath_adjust_factor = (float) (st->ath_adjust_factor * (adj_lim_new * 0.075 + 0.925));
            if (st->ath_adjust_factor < adj_lim_new) st->ath_adjust_factor = adj_lim_new;
        }
        else {
            if (st->ath_adjust_limit >= adj_lim_new) st->ath_adjust_factor = adj_lim_new;
            else if (st->ath_adjust_factor < st->ath_adjust_limit) st->ath_adjust_factor = st->ath_adjust_limit;
        }
        st->ath_adjust_limit = adj_lim_new;
    }
}

/* corrupted state: stop with an error instead of encoding something else than the reference (the reference asserts in such places) */
__device__ __forceinline__ void lg_scan_inconsistent()
{
#ifdef LG_EMULATE
    abort();
#else
    __trap();
#endif
}

__global__ void __launch_bounds__(32)
lg_kernel_scan(const LgDevCfg *__restrict__ cfg, const LgAnalysis *__restrict__ ana, LgPsyOut *__restrict__ psy,
               LgFrameCtl *__restrict__ frm, LgStreamState *__restrict__ state,
               const int *__restrict__ nfr, int nframes, int f0, int f1 /* this launch: frames f0 .. f1-1 of the batch */)
{
    LG_DYN_SMEM(LgSmemB, sm);
    int const lane = threadIdx.x & 31;
    int const stream = blockIdx.x;
    LgStreamState *st = &sm->st;
    {
        const int *src = reinterpret_cast<const int *>(state + stream);
        int *dst = reinterpret_cast<int *>(st);
        for (int i = lane; i < (int) (sizeof(LgStreamState) / 4); i += 32) dst[i] = src[i];
    }
    for (int i = lane; i < 513; i += 32) sm->log_table[i] = cfg->log_table[i];
    if (lane < 3) {
        const LgBands *g = lane == 0 ? &cfg->l : (lane == 1 ? &cfg->l2s : &cfg->s);
        lg_p2s_plan(g, sm->p2s_prev[lane], sm->p2s_first[lane], sm->p2s_cur[lane], sm->p2s_kind[lane]);
    }
    __syncwarp();
    const LgBands *gdl = &cfg->l, *gds = &cfg->s;
    int const nch = cfg->channels;
    int const n_chn_psy = (cfg->mode == LG_JOINT) ? 4 : nch;
    float const pcfact = 0.6f;

    int const my_frames = min(nfr[stream], f1);
    int const mgr = cfg->mode_gr;                 /* granules per frame: 2 (MPEG-1) or 1 (MPEG-2/2.5) */
    for (int gb = mgr * f0; gb < mgr * my_frames; gb++) {
        const LgAnalysis *A = &sm->A;
        {
            static_assert(sizeof(LgAnalysis) % 8 == 0 && LG_ANALYSIS_LONG_BYTES % 8 == 0, "LgAnalysis is copied in 8-byte words");
            const LgAnalysis *rec = ana + (size_t) stream * 2 * nframes + gb;
            const float2 *src = reinterpret_cast<const float2 *>(rec);
            float2 *dst = reinterpret_cast<float2 *>(&sm->A);
            /* the short-block part (two thirds of the record) only where kernel A found that the granule can switch to short blocks */
            int const words = (int) ((__ldg(&rec->has_short) ? sizeof(LgAnalysis) : LG_ANALYSIS_LONG_BYTES) / 8);
            for (int i = lane; i < words; i += 32) dst[i] = __ldg(src + i);
        }
        __syncwarp();
        LgPsyOut *P = psy + (size_t) stream * 2 * nframes + gb;
        int const gr = gb % mgr;
        float const ath_factor = (cfg->msfix > 0.f) ? (cfg->ath_offset_factor * st->ath_adjust_factor) : 1.f;
        float const qml = st->masking_lower;

        /* (a) last_thm (short part) and (b) the delayed ratios handed to the caller (psymodel.c:796-803, :1441) */
        for (int i = lane; i < 4 * 39; i += 32) (&sm->last_thm_s[0][0][0])[i] = (&st->thm[i / 39].s[0][0])[i % 39];
        for (int i = lane; i < 4 * 61; i += 32) {
            int const chn = i / 61, k = i % 61;
            float const e_old = ((const float *) &st->en[chn])[k], t_old = ((const float *) &st->thm[chn])[k];
            ((float *) &P->en[chn])[k] = e_old;
            ((float *) &P->thm[chn])[k] = t_old;
            ((float *) &sm->pe_en[chn])[k] = e_old;
            ((float *) &sm->pe_thm[chn])[k] = t_old;
        }
        __syncwarp();
        /* (c) attack detection, one lane per channel */
        if (lane < n_chn_psy) {
            int ul;
            lg_attack_channel(cfg, st, A->en_subshort[lane], lane, sm->ns_att[lane], sm->ssf[lane], &ul);
            sm->nsuse[lane] = ul;
            P->tot_ener[lane] = st->tot_ener[lane];            /* psymodel.c:938 energy[chn] (one granule old) */
            st->tot_ener[lane] = A->tot_ener[lane];            /* psymodel.c:695 */
            if (lane < 2) {                                    /* psymodel.c:749-750 */
                P->loudness_sq[lane] = st->loudness_sq_save[lane];
                st->loudness_sq_save[lane] = A->loudness[lane];
            }
        }
        __syncwarp();
        int ul0 = sm->nsuse[0], ul1 = (nch == 2) ? sm->nsuse[1] : 1;
        if (n_chn_psy > 2) { if (!sm->nsuse[2] || !sm->nsuse[3]) ul0 = ul1 = 0; }
        /* psymodel.c:1265 vbrpsy_compute_block_type */
        if (cfg->short_blocks == 1 && !(ul0 && ul1)) ul0 = ul1 = 0;
        if (cfg->short_blocks == 2) ul0 = ul1 = 1;
        if (cfg->short_blocks == 3) ul0 = ul1 = 0;

        /* (d) long-block thresholds: pre-echo control + clamps (psymodel.c:1187-1256) */
        for (int chn = 0; chn < n_chn_psy; chn++) {
            int const bt_old = st->blocktype_old[chn & 1];
            for (int b = lane; b < LG_CBANDS; b += 32) {
                float e = 0, t = 0;
                if (b < gdl->npart) {
                    float const ecb = A->ecb_l[chn][b];
                    float const masking_lower = gdl->masking_lower[b] * qml;
                    e = A->eb_l[chn][b];
                    if (bt_old == LG_SHORT) {
                        float const ecb_limit = 2 * st->nb_l1[chn][b];
                        if (ecb_limit > 0) t = ecb < ecb_limit ? ecb : ecb_limit;
                        else t = (float) ((ecb < e * 0.3) ? ecb : e * 0.3);
                    }
                    else {
                        float ecb_limit_2 = 16 * st->nb_l2[chn][b];
                        float ecb_limit_1 = 2 * st->nb_l1[chn][b];
                        float ecb_limit;
                        if (ecb_limit_2 <= 0) ecb_limit_2 = ecb;
                        if (ecb_limit_1 <= 0) ecb_limit_1 = ecb;
                        if (bt_old == LG_NORM) ecb_limit = ecb_limit_1 < ecb_limit_2 ? ecb_limit_1 : ecb_limit_2;
                        else ecb_limit = ecb_limit_1;
                        t = ecb < ecb_limit ? ecb : ecb_limit;
                    }
                    st->nb_l2[chn][b] = st->nb_l1[chn][b];
                    st->nb_l1[chn][b] = ecb;
                    float const x = A->lim_l[chn][b];
                    if (t > x) t = x;
                    if (masking_lower > 1) t *= masking_lower;
                    if (t > e) t = e;
                    if (masking_lower < 1) t *= masking_lower;
                }
                sm->eb[chn][b] = e;
                sm->thr[chn][b] = t;
            }
        }
        __syncwarp();
        /* (e) psymodel.c:1458-1463 */
        if (cfg->mode == LG_JOINT && (ul0 + ul1) == 2) {
            for (int b = lane; b < gdl->npart; b += 32)
                lg_ms_threshold(sm->eb, sm->thr, gdl->mld_cb[b], cfg->ath_cb_l[b], ath_factor, cfg->msfix, b);
            __syncwarp();
        }
        /* (f) partitions -> scalefactor bands, long and long->short (psymodel.c:411, :421): one job per channel and band, side by side */
        for (int job = lane; job < n_chn_psy * (LG_SBMAX_L + LG_SBMAX_S); job += 32) {
            int const chn = job / (LG_SBMAX_L + LG_SBMAX_S), q = job % (LG_SBMAX_L + LG_SBMAX_S);
            float enn, thmm;
            if (q < LG_SBMAX_L) {
                lg_p2s_band(gdl, sm->p2s_prev[0], sm->p2s_first[0], sm->p2s_cur[0], sm->p2s_kind[0], sm->eb[chn], sm->thr[chn], q, &enn, &thmm);
                st->en[chn].l[q] = enn; st->thm[chn].l[q] = thmm;
            }
            else {
                int const sb = q - LG_SBMAX_L;
                lg_p2s_band(&cfg->l2s, sm->p2s_prev[1], sm->p2s_first[1], sm->p2s_cur[1], sm->p2s_kind[1], sm->eb[chn], sm->thr[chn], sb, &enn, &thmm);
                float const scale = (float) (1. / 64.f);
                float const tmp_thm = thmm * scale;
                for (int k = 0; k < 3; ++k) { st->en[chn].s[sb][k] = enn; st->thm[chn].s[sb][k] = tmp_thm; }
            }
        }
        __syncwarp();
        /* (g) short blocks (psymodel.c:1470-1500); kernel A already produced min(ecb, clamp) */
        if (!(ul0 && ul1)) {
            if (!A->has_short) lg_scan_inconsistent();         /* kernel A's test is a superset of the attack detection above: cannot happen */
            for (int sblock = 0; sblock < 3; sblock++) {
                for (int chn = 0; chn < n_chn_psy; ++chn) {
                    int const ul = (chn & 1) ? ul1 : ul0;
                    if (ul) continue;
                    for (int b = lane; b < LG_CBANDS; b += 32) {
                        float e = 0, t = 0;
                        if (b < gds->npart) {
                            float const masking_lower = gds->masking_lower[b] * qml;
                            e = A->eb_s[sblock][chn][b];
                            t = A->thr_s[sblock][chn][b];
                            if (masking_lower > 1) t *= masking_lower;
                            if (t > e) t = e;
                            if (masking_lower < 1) t *= masking_lower;
                        }
                        sm->eb[chn][b] = e;
                        sm->thr[chn][b] = t;
                    }
                }
                __syncwarp();
                if (cfg->mode == LG_JOINT && (ul0 + ul1) == 0) {
                    for (int b = lane; b < gds->npart; b += 32)
                        lg_ms_threshold(sm->eb, sm->thr, gds->mld_cb[b], cfg->ath_cb_s[b], ath_factor, cfg->msfix, b);
                    __syncwarp();
                }
                for (int job = lane; job < n_chn_psy * LG_SBMAX_S; job += 32) {
                    int const chn = job / LG_SBMAX_S, sb = job % LG_SBMAX_S;
                    int const ul = (chn & 1) ? ul1 : ul0;
                    if (!ul) lg_p2s_band(gds, sm->p2s_prev[2], sm->p2s_first[2], sm->p2s_cur[2], sm->p2s_kind[2], sm->eb[chn], sm->thr[chn], sb,
                                         &st->en[chn].s[sb][sblock], &st->thm[chn].s[sb][sblock]);
                }
                __syncwarp();
            }
        }
        /* (h) short block pre-echo control (psymodel.c:1502-1554), one lane per (chn, sb) */
        for (int task = lane; task < n_chn_psy * LG_SBMAX_S; task += 32) {
            int const chn = task / LG_SBMAX_S, sb = task % LG_SBMAX_S;
            const int *na = sm->ns_att[chn];
            float new_thmm[3], prev_thm, t1, t2, thmm;
            for (int sblock = 0; sblock < 3; sblock++) {
                thmm = st->thm[chn].s[sb][sblock];
                thmm = (float) (thmm * 0.8);
                t1 = t2 = thmm;
                if (sblock > 0) prev_thm = new_thmm[sblock - 1];
                else prev_thm = sm->last_thm_s[chn][sb][2];
                if (na[sblock] >= 2 || na[sblock + 1] == 1) t1 = lg_ns_interp(prev_thm, thmm, (float) (0.6 * pcfact));
                thmm = t1 < thmm ? t1 : thmm;
                if (na[sblock] == 1) t2 = lg_ns_interp(prev_thm, thmm, (float) (0.3 * pcfact));
                else if ((sblock == 0 && st->last_attacks[chn] == 3) || (sblock > 0 && na[sblock - 1] == 3)) {
                    switch (sblock) {
                    case 0: prev_thm = sm->last_thm_s[chn][sb][1]; break;
                    case 1: prev_thm = sm->last_thm_s[chn][sb][2]; break;
                    case 2: prev_thm = new_thmm[0]; break;
                    }
                    t2 = lg_ns_interp(prev_thm, thmm, (float) (0.3 * pcfact));
                }
                thmm = t1 < thmm ? t1 : thmm;
                thmm = t2 < thmm ? t2 : thmm;
                thmm *= sm->ssf[chn][sblock];
                new_thmm[sblock] = thmm;
            }
            for (int sblock = 0; sblock < 3; sblock++) st->thm[chn].s[sb][sblock] = new_thmm[sblock];
        }
        __syncwarp();
        /* (i) psymodel.c:1555-1557 */
        if (lane < n_chn_psy) st->last_attacks[lane] = sm->ns_att[lane][2];
        /* (j) psymodel.c:1289 vbrpsy_apply_block_type (uniform) */
        int btd[2] = { 0, 0 };
        for (int chn = 0; chn < nch; chn++) {
            int const ul = chn ? ul1 : ul0;
            int old = st->blocktype_old[chn];
            int blocktype = LG_NORM;
            if (ul) { if (old == LG_SHORT) blocktype = LG_STOP; }
            else {
                blocktype = LG_SHORT;
                if (old == LG_NORM) old = LG_START;
                if (old == LG_STOP) old = LG_SHORT;
            }
            btd[chn] = old;
            __syncwarp();
            if (lane == 0) { st->blocktype_old[chn] = blocktype; P->block_type[chn] = old; }
        }
        __syncwarp();
        /* (k) perceptual entropy on the delayed ratios (psymodel.c:1568-1595) */
        if (lane < n_chn_psy) {
            int type;
            if (lane > 1) type = (btd[0] == LG_SHORT || btd[1] == LG_SHORT) ? LG_SHORT : LG_NORM;
            else type = btd[lane];
            float const v = (type == LG_SHORT) ? lg_pecalc_s(sm->log_table, &sm->pe_en[lane], &sm->pe_thm[lane], qml)
                                                : lg_pecalc_l(sm->log_table, &sm->pe_en[lane], &sm->pe_thm[lane], qml);
            P->pe[lane] = v;
            sm->frame_pe[gr][lane] = v;
            sm->frame_tot[gr][lane] = P->tot_ener[lane];
        }
        if (lane < 2) sm->frame_bt[gr][lane] = btd[lane];
        __syncwarp();

        /* ---- frame-level decisions once both granules are known (encoder.c:380-518) */
        if (gr == mgr - 1) {
            int const frame = gb / mgr;
            LgFrameCtl *F = frm + (size_t) stream * nframes + frame;
            if (lane == 0) {
                const LgPsyOut *P0 = P - (mgr - 1);
                float loud[2][2] = { { P0->loudness_sq[0], P0->loudness_sq[1] }, { P->loudness_sq[0], P->loudness_sq[1] } };
                float ms_ener_ratio[2] = { .5f, .5f };
                int mode_ext = 0;
                /* padding (encoder.c:348-352) */
                int padding = 0;
                if ((st->slot_lag -= cfg->frac_spf) < 0) { st->slot_lag += cfg->samplerate; padding = 1; }
                if (cfg->mode == LG_JOINT)
                    for (int g = 0; g < mgr; g++) {
                        ms_ener_ratio[g] = sm->frame_tot[g][2] + sm->frame_tot[g][3];
                        if (ms_ener_ratio[g] > 0) ms_ener_ratio[g] = sm->frame_tot[g][3] / ms_ener_ratio[g];
                    }
                lg_adjust_ath(cfg, st, loud);
                if (cfg->force_ms) mode_ext = 2;
                else if (cfg->mode == LG_JOINT) {
                    float sum_pe_MS = 0, sum_pe_LR = 0;
                    for (int g = 0; g < mgr; g++)
                        for (int ch = 0; ch < nch; ch++) { sum_pe_MS += sm->frame_pe[g][2 + ch]; sum_pe_LR += sm->frame_pe[g][ch]; }
                    if (sum_pe_MS <= 1.00 * sum_pe_LR) {
                        if (sm->frame_bt[0][0] == sm->frame_bt[0][1] && sm->frame_bt[mgr - 1][0] == sm->frame_bt[mgr - 1][1]) mode_ext = 2;
                    }
                }
                int const off = (mode_ext == 2) ? 2 : 0;
                const float fircoef[9] = { -0.0207887 * 5, -0.0378413 * 5, -0.0432472 * 5, -0.031183 * 5,
                    7.79609e-18 * 5, 0.0467745 * 5, 0.10091 * 5, 0.151365 * 5, 0.187098 * 5 };
                for (int i = 0; i < 18; i++) st->pefirbuf[i] = st->pefirbuf[i + 1];
                float f = 0.0f;
                for (int g = 0; g < mgr; g++)
                    for (int ch = 0; ch < nch; ch++) f += sm->frame_pe[g][off + ch];
                st->pefirbuf[18] = f;
                f = st->pefirbuf[9];
                for (int i = 0; i < 9; i++) f += (st->pefirbuf[i] + st->pefirbuf[18 - i]) * fircoef[i];
                f = (670 * 5 * cfg->mode_gr * nch) / f;
                for (int g = 0; g < 2; g++)
                    for (int ch = 0; ch < 2; ch++) F->pe_use[g][ch] = (ch < nch && g < mgr) ? sm->frame_pe[g][off + ch] * f : 0.f;
                F->mode_ext = mode_ext;
                F->padding = padding;
                F->ms_ener_ratio[0] = ms_ener_ratio[0];
                F->ms_ener_ratio[1] = ms_ener_ratio[1];
                F->ath_adjust_factor = st->ath_adjust_factor;
                /* quantize.c:2019-2029: masking_lower left behind by the last gr/ch of this frame */
                st->masking_lower = (cfg->vbr == 4 || sm->frame_bt[mgr - 1][nch - 1] != LG_SHORT) ? cfg->masking_lower_long : cfg->masking_lower_short;   /* VBR-new: mask_adjust for every block type, quantize.c:1613 */
                if (cfg->vbr == 2) {
                    /* VBR-old lowers the masking with the perceptual entropy of the last granule.channel (quantize.c:1418-1426); that
                     * value is what the next frame's psycho-acoustics sees.  exp and pow in double, rounded to float as the reference's. */
                    float const pe_last = sm->frame_pe[mgr - 1][off + nch - 1] * f;
                    float adjust, db;
                    if (sm->frame_bt[mgr - 1][nch - 1] != LG_SHORT) { adjust = (float) (1.28 / (1 + lg_exp(3.5 - pe_last / 300.)) - 0.05); db = cfg->mask_adjust - adjust; }
                    else { adjust = (float) (2.56 / (1 + lg_exp(3.5 - pe_last / 300.)) - 0.14); db = cfg->mask_adjust_short - adjust; }
                    st->masking_lower = (float) lg_pow(10.0, db * 0.1);
                }
                F->masking_lower = st->masking_lower;
                st->frames_done++;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    {
        /* write back what this kernel owns: the quantiser's fields of the state (reservoir, step-size memory, ancillary flag) may be
         * written by kernel D of an earlier piece of the batch at the same time */
        int *dst = reinterpret_cast<int *>(state + stream);
        const int *src = reinterpret_cast<const int *>(st);
        int const q0 = (int) (offsetof(LgStreamState, resv_size) / 4), q1 = (int) (offsetof(LgStreamState, frames_done) / 4);
        int const qa = (int) (offsetof(LgStreamState, ancillary_flag) / 4);
        for (int i = lane; i < (int) (sizeof(LgStreamState) / 4); i += 32) if (!((i >= q0 && i < q1) || i == qa)) dst[i] = src[i];
    }
}
