// lg_k_vbrold.cuh - kernel D'' : VBR-old (vbr_rh), quantize.c:1491 VBR_old_iteration_loop.
//
// The reference finds, for every granule.channel of a frame, the smallest bit budget at which outer_loop (the CBR
// noise-shaping search, kernel D's lg_outer_loop) leaves no band distorted - a bisection of 5..7 outer_loop runs
// (VBR_encode_granule :1246) - then picks the smallest frame that holds the sum; if even the largest does not, the allowed
// noise is raised towards the high bands and everything is searched again (bitpressure_strategy :1455).
//
// Mapping: one CTA per stream, warp = channel (as kernel D), frames in order.  The granules of a channel run one after the
// other on the same warp because bin_search_StepSize starts every search from the gain the previous one left behind
// (OldValue[ch]); both granules' work sets stay in shared memory until the frame's size is settled.  outer_loop is
// instantiated with its VBR-old flag: xrpow is kept as of the chosen quantisation between runs (save_xrpow), sfb21_extra
// is switched off near the bit limit.  The masking feedback of this mode (masking_lower from the perceptual entropy, a
// double exp and pow) is applied by the scan kernel, see lg_k_scan.cuh.
#pragma once
#include "lg_k_quant.cuh"

struct LgVOBest { float xrpow[576]; int16_t ix[576]; int sf[40]; float tail[40]; };
struct LgSmemO {
    LgQWarp w[2][2];                 /* [granule][channel] */
    LgVOBest bst[2];                 /* the bisection's best so far, one per channel warp */
    LgQInfo gi[2][2];
    LgQConst qc[2][2];
    int used[2], ath_over[2][2];
    int bt_gr0[2];
};

/* quantize.c:160 psfb21_analogsilence: from the top of the spectrum down, lines of the six sub-bands of sfb21 (sfb12 of each
 * window for short blocks) that are below the ATH become zero, up to the first line that is not */
__device__ __noinline__ void lg_psfb21_analog_silence(const LgDevCfg *__restrict__ c, LgQWarp *w, int block_type, float ath_adjust_factor, int lane)
{
    float *xr = w->xr;
    if (block_type != LG_SHORT) {
        float ath[6];
        for (int g = 0; g < 6; g++) {
            ath[g] = lg_ath_adjust(c, ath_adjust_factor, c->ath_psfb21[g], c->ath_floor, 0.f);
            if (c->longfact[21] > 1e-12f) ath[g] *= c->longfact[21];
        }
        int const start = c->psfb21[0];
        int top = start - 1;                       /* highest line that stays */
        for (int j = start + lane; j < 576; j += 32) {
            int g = 0;
            while (g < 5 && j >= c->psfb21[g + 1]) g++;
            if (!(fabsf(xr[j]) < ath[g])) top = j;
        }
        top = lg_wmax_i(top);
        for (int j = top + 1 + lane; j < 576; j += 32) xr[j] = 0.f;
        __syncwarp();
        return;
    }
    float ath[6];
    for (int g = 0; g < 6; g++) {
        ath[g] = lg_ath_adjust(c, ath_adjust_factor, c->ath_psfb12[g], c->ath_floor, 0.f);
        if (c->shortfact[12] > 1e-12f) ath[g] *= c->shortfact[12];
    }
    int const wd = c->sfb_s[13] - c->sfb_s[12];
    for (int block = 0; block < 3; block++) {
        int const start = c->sfb_s[12] * 3 + wd * block, end = start + (c->psfb12[6] - c->psfb12[0]);
        int top = start - 1;
        for (int j = start + lane; j < end; j += 32) {
            int const rel = j - start + c->psfb12[0];
            int g = 0;
            while (g < 5 && rel >= c->psfb12[g + 1]) g++;
            if (!(fabsf(xr[j]) < ath[g])) top = j;
        }
        top = lg_wmax_i(top);
        for (int j = top + 1 + lane; j < end; j += 32) xr[j] = 0.f;
        __syncwarp();
    }
}

/* best-so-far of the bisection <-> work set (the reference's gr_info struct assignment plus bst_xrpow) */
__device__ __noinline__ void lg_vo_copy(LgQWarp *w, LgVOBest *b, int to_best, int lane)
{
    float2 *dx = reinterpret_cast<float2 *>(to_best ? b->xrpow : w->xrpow);
    const float2 *sx = reinterpret_cast<const float2 *>(to_best ? w->xrpow : b->xrpow);
    unsigned *di = reinterpret_cast<unsigned *>(to_best ? b->ix : w->ixw);
    const unsigned *si = reinterpret_cast<const unsigned *>(to_best ? w->ixw : b->ix);
    for (int i = lane; i < 288; i += 32) { dx[i] = sx[i]; di[i] = si[i]; }
    for (int i = lane; i < 40; i += 32) {
        if (to_best) { b->sf[i] = w->sfw[i]; b->tail[i] = w->tail_max[i]; }
        else { w->sfw[i] = b->sf[i]; w->tail_max[i] = b->tail[i]; }
    }
    __syncwarp();
}

template <int SUB>
__global__ void __launch_bounds__(64)
lg_kernel_vbrold(const LgDevCfg *__restrict__ cfg, const float *__restrict__ xr_in, const LgPsyOut *__restrict__ psy,
                 const LgFrameCtl *__restrict__ frm, LgGranuleOut *__restrict__ gout, LgFrameOut *__restrict__ fout,
                 LgStreamState *__restrict__ state, const int *__restrict__ nfr, int nframes)
{
    LG_DYN_SMEM(LgSmemO, sm);
    int const lane = threadIdx.x & 31, ch = threadIdx.x >> 5;
    int const stream = blockIdx.x;
    const LgDevCfg *c = cfg;
    int const nch = c->channels, mgr = c->mode_gr;
    int const active = ch < nch;
    LgStreamState *st = state + stream;
    LgVOBest *bst = &sm->bst[ch];
    int resv_size = st->resv_size, main_data_begin = st->main_data_begin;
    int old_value = st->old_value[ch], current_step = st->current_step[ch];
    int anc_flag = st->ancillary_flag, pay_off = 0;
    int const max_index = c->vbr_max_bitrate_index;

    int const my_frames = nfr[stream];
    if (my_frames <= 0) return;                        /* nothing of this stream in this step: its state is not ours to write back (another step's kernel may own it) */
    for (int frame = 0; frame < my_frames; frame++) {
        const LgFrameCtl *F = frm + (size_t) stream * nframes + frame;
        int const padding = F->padding, mode_ext = F->mode_ext;
        /* ---- VBR_old_prepare (quantize.c:1370): frame sizes, bit budgets, allowed noise */
        int frameBits[16], avg, resv_max_m, dummy;
        avg = lg_resv_frame_begin(c, max_index, padding, resv_size, &dummy, &resv_max_m) / mgr;
        for (int i = 1; i <= max_index; i++) frameBits[i] = lg_resv_frame_begin(c, i, padding, resv_size, &dummy, &dummy);
        int max_bits[2][2] = { { 0, 0 }, { 0, 0 } }, min_bits[2][2] = { { 0, 0 }, { 0, 0 } }, bits = 0;
        for (int g = 0; g < mgr; g++) {
            float pe[2] = { F->pe_use[g][0], F->pe_use[g][1] };
            int tb[2] = { 0, 0 };
            int const mxb = lg_on_pe(c, resv_size, resv_max_m, pe, tb, avg, 0);
            if (mode_ext == 2) lg_reduce_side(tb, F->ms_ener_ratio[g], avg, mxb);
            for (int k = 0; k < nch; k++) { max_bits[g][k] = tb[k]; min_bits[g][k] = 126; bits += tb[k]; }
        }
        for (int g = 0; g < mgr; g++)
            for (int k = 0; k < nch; k++) {
                if (bits > frameBits[max_index] && bits > 0) { max_bits[g][k] *= frameBits[max_index]; max_bits[g][k] /= bits; }
                if (min_bits[g][k] > max_bits[g][k]) min_bits[g][k] = max_bits[g][k];
            }
        if (active) {
            for (int gr = 0; gr < mgr; gr++) {
                int const gb = mgr * frame + gr;
                const LgPsyOut *P = psy + (size_t) stream * 2 * nframes + gb;
                LgQWarp *w = &sm->w[gr][ch];
                LgQInfo gi;
                LgQConst qc;
                /* quantize.c:226 init_outer_loop (the short-block reorder was done by kernel C) */
                gi.part2_3_length = 0; gi.big_values = 0; gi.count1 = 0; gi.global_gain = 210; gi.scalefac_compress = 0;
                gi.table_select[0] = gi.table_select[1] = gi.table_select[2] = 0; gi.sbg = 0;
                gi.region0_count = 0; gi.region1_count = 0; gi.preflag = 0; gi.scalefac_scale = 0;
                gi.count1table_select = 0; gi.part2_length = 0; gi.count1bits = 0; gi.xrpow_max = 0;
                qc.ath_over = 0; qc.block_type = P->block_type[ch]; qc.max_nonzero_coeff = 575; qc.jn = 9;
                qc.sfb_lmax = LG_SBPSY_L; qc.sfb_smin = LG_SBPSY_S;
                qc.psy_lmax = c->sfb21_extra ? LG_SBMAX_L : LG_SBPSY_L;
                if (c->samplerate <= 8000) { qc.sfb_lmax = 17; qc.sfb_smin = 9; qc.psy_lmax = 17; }
                qc.psymax = qc.psy_lmax; qc.sfbmax = qc.sfb_lmax; qc.sfbdivide = 11;
                if (qc.block_type == LG_SHORT) {
                    qc.sfb_smin = 0; qc.sfb_lmax = 0;
                    qc.psymax = 3 * (c->sfb21_extra ? LG_SBMAX_S : LG_SBPSY_S);
                    qc.sfbmax = 3 * LG_SBPSY_S;
                    if (c->samplerate <= 8000) qc.psymax = qc.sfbmax = 3 * 9;
                    qc.sfbdivide = qc.sfbmax - 18;
                    qc.psy_lmax = 0;
                }
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
                lg_psfb21_analog_silence(c, w, qc.block_type, F->ath_adjust_factor, lane);
                int const rch = (mode_ext == 2) ? ch + 2 : ch;
                {
                    LgQConst qx = qc;
                    lg_calc_xmin(c, w, qx, &P->en[rch], &P->thm[rch], F->ath_adjust_factor, lane);
                    qc.max_nonzero_coeff = qx.max_nonzero_coeff; qc.jn = qx.jn; qc.ath_over = qx.ath_over;
                }
                if (lane == 0) { sm->gi[gr][ch] = gi; sm->qc[gr][ch] = qc; sm->ath_over[gr][ch] = qc.ath_over; }
                __syncwarp();
            }
        }
        else if (lane == 0) for (int gr = 0; gr < 2; gr++) sm->ath_over[gr][ch] = 0;
        __syncthreads();
        int analog_silence = 1;
        for (int g = 0; g < mgr; g++) for (int k = 0; k < nch; k++) if (sm->ath_over[g][k]) analog_silence = 0;
        /* ---- the loop: quantise every granule with as few bits as its noise allows, find the frame that holds them */
        int used_bits, bitrate_index, mean_bits = 0, resv_max = 0;
        for (int pass = 0;; pass++) {
            if (pass > 200) lg_runaway();
            int my_used = 0;
            if (active) {
                for (int gr = 0; gr < mgr; gr++) {
                    LgQWarp *w = &sm->w[gr][ch];
                    LgQInfo gi = sm->gi[gr][ch];
                    LgQConst qc = sm->qc[gr][ch];
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
                    for (int i = lane; i < 40; i += 32) w->tail_max[i] = 0.f;       /* xrpow is zero above max_nonzero_coeff here */
                    gi.xrpow_max = lg_wmax_fpos(mx);
                    amax = lg_wmax_fpos(amax);
                    __syncwarp();
                    int nonzero = amax > (float) 1E-20;
                    if (!nonzero && amax > 0.f) {
                        float sum = 0;
                        for (int i = 0; i <= qc.max_nonzero_coeff; ++i) sum += fabsf(w->xr[i]);
                        nonzero = sum > (float) 1E-20;
                    }
                    if (nonzero) {
                        if (lane == 0) w->ph[0] = w->ph[1] = SUB ? ~0u : 0u;              /* quantize.c:131-137 */
                        __syncwarp();
                    }
                    else {
                        for (int i = lane; i < 288; i += 32) reinterpret_cast<unsigned *>(w->ixw)[i] = 0u;
                        __syncwarp();
                    }
                    if (nonzero && max_bits[gr][ch] != 0) {
                        /* quantize.c:1246 VBR_encode_granule */
                        int lo = min_bits[gr][ch], hi = max_bits[gr][ch];
                        int const Max_bits = hi;
                        int this_bits = (hi + lo) / 2, dbits, found = 0;
                        LgQInfo bgi = gi;
                        do {
                            int const sfb21 = (this_bits > Max_bits - 42) ? 0 : c->sfb21_extra;
                            int const over = lg_outer_loop<SUB | 2>(c, w, gi, qc, this_bits, &old_value, &current_step, sfb21, lane);
                            if (over <= 0) {
                                found = 1;
                                bgi = gi;
                                lg_vo_copy(w, bst, 1, lane);
                                hi = gi.part2_3_length - 32;
                                dbits = hi - lo;
                                this_bits = (hi + lo) / 2;
                            }
                            else {
                                lo = this_bits + 32;
                                dbits = hi - lo;
                                this_bits = (hi + lo) / 2;
                                if (found) {
                                    found = 2;
                                    gi = bgi;
                                    lg_vo_copy(w, bst, 0, lane);
                                }
                            }
                        } while (dbits > 12);
                        my_used += gi.part2_3_length + gi.part2_length;
                    }
                    if (lane == 0) sm->gi[gr][ch] = gi;
                    __syncwarp();
                }
            }
            if (lane == 0) sm->used[ch] = my_used;
            __syncthreads();
            used_bits = sm->used[0] + sm->used[1];
            bitrate_index = (analog_silence && !c->enforce_min_bitrate) ? 1 : c->vbr_min_bitrate_index;
            for (; bitrate_index < max_index; bitrate_index++) if (used_bits <= frameBits[bitrate_index]) break;
            int const full = lg_resv_frame_begin(c, bitrate_index, padding, resv_size, &mean_bits, &resv_max);
            __syncthreads();
            if (used_bits <= full) break;
            /* quantize.c:1455 bitpressure_strategy */
            if (active) {
                for (int gr = 0; gr < mgr; gr++) {
                    LgQWarp *w = &sm->w[gr][ch];
                    LgQConst const qc = sm->qc[gr][ch];
                    for (int k = lane; k < 40; k += 32) {
                        int sfb = -1, n = LG_SBMAX_L;
                        if (k < qc.psy_lmax) sfb = k;
                        else if (qc.block_type == LG_SHORT && k - qc.psy_lmax < 3 * (LG_SBMAX_S - qc.sfb_smin)) { sfb = qc.sfb_smin + (k - qc.psy_lmax) / 3; n = LG_SBMAX_S; }
                        if (sfb >= 0) w->l3_xmin[k] = (float) ((double) w->l3_xmin[k] * (1. + .029 * sfb * sfb / n / n));
                    }
                }
                __syncwarp();
            }
            for (int g = 0; g < mgr; g++)
                for (int k = 0; k < nch; k++) {
                    double const m = 0.9 * max_bits[g][k];
                    max_bits[g][k] = (int) (min_bits[g][k] > m ? (double) min_bits[g][k] : m);
                }
        }
        /* ---- iteration_finish_one for every granule.channel (quantize.c:1213), reservoir, hand-over to the packer */
        uint8_t scfsi[4] = { 0, 0, 0, 0 };
        if (active) {
            for (int gr = 0; gr < mgr; gr++) {
                int const gb = mgr * frame + gr;
                LgQWarp *w = &sm->w[gr][ch];
                LgQInfo gi = sm->gi[gr][ch];
                LgQConst qc = sm->qc[gr][ch];
                {
                    LgQInfo gm = gi;
                    LgQConst qx = qc;
                    lg_best_scalefac_store(c, w, gm, qx, gr, sm->w[0][ch].sfw, sm->bt_gr0[ch], scfsi, lane);
                    if (c->use_best_huffman == 1) lg_best_huffman_divide(c, w, gm, qx, lane);
                    gi = gm;
                }
                if (gr == 0 && lane == 0) sm->bt_gr0[ch] = qc.block_type;
                __syncwarp();
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
                    sm->gi[gr][ch] = gi;
                }
                __syncwarp();
            }
        }
        if (lane == 0) {
            int u = 0;
            if (active) for (int gr = 0; gr < mgr; gr++) u += sm->gi[gr][ch].part2_3_length + sm->gi[gr][ch].part2_length;
            sm->used[ch] = u;
        }
        __syncthreads();
        {
            int const frame_used = sm->used[0] + sm->used[1];
            resv_size -= frame_used;                                          /* ResvAdjust for every granule.channel */
            /* reservoir.c:239 ResvFrameEnd + the main_data_begin recurrence of format_bitstream (bitstream.c:937) */
            int stuffingBits = 0, over_bits, drain_pre = 0, drain_post = 0;
            resv_size += mean_bits * mgr;
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
            int const pay_bits = drain_pre + frame_used + drain_post;
            if (pay_bits & 7) lg_runaway();
            int const anc_pre = anc_flag;
            if (!c->disable_reservoir) anc_flag ^= (lg_drain_tail_bits(drain_pre) + lg_drain_tail_bits(drain_post)) & 1;
            LgFrameOut *fo = fout + (size_t) stream * nframes + frame;
            if (lane == 0) {
                if (ch == 0) {
                    fo->main_data_begin = mdb_header; fo->drain_pre = drain_pre; fo->drain_post = drain_post;
                    fo->padding = padding; fo->mode_ext = mode_ext; fo->resv_size = resv_size;
                    fo->pay_off = pay_off; fo->pay_bytes = pay_bits >> 3;
                    fo->anc_pre = (uint8_t) anc_pre; fo->anc_post = (uint8_t) anc_flag; fo->pad_[0] = fo->pad_[1] = 0; fo->bitrate_index = bitrate_index;
                }
                for (int i = 0; i < 4; i++) fo->scfsi[ch][i] = active ? scfsi[i] : 0;
            }
            pay_off += pay_bits >> 3;
        }
        __syncthreads();
    }
    if (lane == 0) {
        if (ch == 0) { st->resv_size = resv_size; st->main_data_begin = main_data_begin; st->ancillary_flag = anc_flag; }
        st->old_value[ch] = old_value;
        st->current_step[ch] = current_step;
    }
}
