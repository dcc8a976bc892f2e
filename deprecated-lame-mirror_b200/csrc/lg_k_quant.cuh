// lg_k_quant.cuh - kernel D: the CBR quantisation / noise-shaping / Huffman bit-counting loop.
//
// One CTA per stream, one warp per channel ("one warp per granule/channel").  A stream's granules are
// strictly ordered (bit reservoir ResvSize, OldValue/CurrentStep of the step-size search, scfsi between
// granule 0 and 1), so the kernel walks frames and granules in order; the two channels of a granule are
// independent (targ_bits is fixed before the channel loop, quantize.c:2008) and run concurrently.
//
// Inside a warp the 576 lines are owned pairwise (lane p owns pairs p, p+32, ... -> conflict-free 32-bit
// shared-memory accesses of int16 pairs, 64-bit of float pairs); the 22/39 scalefactor bands are owned
// one per lane for everything the reference sums serially per band (calc_xmin, calc_noise), so every
// float sum keeps the reference's order.  All bit counting is integer and uses warp reductions.
//
// Reference: CBR_iteration_loop quantize.c:1988, outer_loop :1010, bin_search_StepSize :367,
// balance_noise :940, amp_scalefac_bands :720, inc_scalefac_scale :808, inc_subblock_gain :847,
// calc_xmin quantize_pvt.c:589, calc_noise :815, on_pe :428, reduce_side :492, count_bits takehiro.c:767,
// quantize_xrpow :281, noquant_count_bits :654, choose_table :618, best_huffman_divide :884,
// best_scalefac_store :1021, scale_bitcount :1318, reservoir.c:83-293.
//
// Algorithmic HBM bytes per gr.ch: read 2304 B MDCT lines + 488 B ratios, write sizeof(LgGranuleOut).
#pragma once
#include "lg_math.cuh"

struct LgQInfo {                 /* the scalar part of the reference's gr_info (l3side.h:47), warp-uniform */
    float xrpow_max;
    int part2_3_length, big_values, count1, global_gain, scalefac_compress;
    int table_select[3];
    int sbg;                     /* subblock_gain[0..2] in 4-bit fields (field 3 stays 0: long-block bands use "window 3") */
    int region0_count, region1_count, preflag, scalefac_scale, count1table_select, part2_length, count1bits;
};
struct LgQConst {                /* per gr.ch constants set by init_outer_loop / calc_xmin */
    int block_type, sfb_lmax, sfb_smin, psy_lmax, sfbmax, psymax, sfbdivide, max_nonzero_coeff;
    int ath_over;                /* calc_xmin: some band's energy exceeds the ATH (ABR's analog-silence test) */
    int jn;                      /* line loops run j < jn: pairs at or above ((max_nonzero_coeff + 2) & ~1) / 2 stay zero */
};
struct LgNoiseRes { float max_noise; int over_count, over_SSD, bits; };
struct LgPrev { int valid, global_gain, sfb_count1; };   /* scalar part of calc_noise_data (quantize_pvt.h:75) */

struct __attribute__((aligned(16))) LgQWarp {
    float xr[576], xrpow[576], save_xrpow[576], sq[576];
    int16_t ixw[576], ixb[576];
    float l3_xmin[40], distort[40], pn_noise[40], pn_noise_log[40];
    int   pn_step[40];
    int   sfw[40], sfbst[40];
    int   width[40], lstart[41];
    int   act[80];
    float tail_max[40];
    float tail_save[40];             /* tail_max as of save_xrpow (VBR-old keeps xrpow across outer_loop calls) */
    int   r01_bits[24], r01_div[24], r0_tbl[24], r1_tbl[24];
    int   comb_bits[128], comb_tbl[128], r0b[16], r0t[16];
    uint8_t line_sfb[576];
    /* bytes, not ints: sizeof(LgSmemD) has to stay below 32 329 B or only six instead of seven CTAs fit an SM (it matters from 889 streams on) */
    uint8_t window[40];              /* window 0..2 of a short-block band, 3 for long-block bands */
    uint8_t eac[40];                 /* calc_xmin: energy_above_cutoff per band (quantize_pvt.c:643), read by the VBR search */
    unsigned ph[2];                  /* QntStateVar_t.pseudohalf as a bit per band: substep shaping (quality 0-2) has the band in its half step */
};
struct LgSmemD {
    LgQWarp w[2];
    int targ_bits[2];
    int used_bits[2];
    int sf_gr0[2][40];           /* granule-0 scalefactors for scfsi (takehiro.c:964) */
    int bt_gr0[2];
};
static_assert(sizeof(LgSmemD) + 1024 <= 233472 / 7, "kernel D: seven CTAs per SM need at most 32 329 B of dynamic shared memory each");

__constant__ uint8_t LG_PRETAB[22] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0 };
/* Per-lane table lookups with a different index in every lane are serialised by the constant cache; the small tables of
 * this kernel are therefore packed into immediates: pretab (2 bits x 22), slen1/slen2 (4 bits x 16), the three
 * scale_bitcount length tables (8 bits x 16). */
__device__ __forceinline__ int lg_pretab(int sfb) { return (int) ((0x2FE95400000ull >> (2 * (sfb < 21 ? sfb : 21))) & 3ull); }
__device__ __forceinline__ int lg_slen1(int k) { return (int) ((0x4433322211130000ull >> (4 * k)) & 15ull); }
__device__ __forceinline__ int lg_slen2(int k) { return (int) ((0x3232132132103210ull >> (4 * k)) & 15ull); }
__device__ __forceinline__ int lg_scale_len(int block_type_short, int k)
{
    /* LG_SCALE_SHORT / LG_SCALE_LONG, entries 0..7 and 8..15 */
    unsigned long long const lo = block_type_short ? 0x4836243636241200ull : 0x291F15211E140A00ull;
    unsigned long long const hi = block_type_short ? 0x7E6C6C5A485A4836ull : 0x4A403F352B342A20ull;
    return (int) (((k < 8 ? lo : hi) >> (8 * (k & 7))) & 255ull);
}
__constant__ int LG_SLEN1_N[16] = { 1, 1, 1, 1, 8, 2, 2, 2, 4, 4, 4, 8, 8, 8, 16, 16 };
__constant__ int LG_SLEN2_N[16] = { 1, 2, 4, 8, 1, 2, 4, 8, 2, 4, 8, 2, 4, 8, 4, 8 };
__constant__ int LG_SLEN1_TAB[16] = { 0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4 };
__constant__ int LG_SLEN2_TAB[16] = { 0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3 };
__constant__ int LG_SCALE_SHORT[16] = { 0, 18, 36, 54, 54, 36, 54, 72, 54, 72, 90, 72, 90, 108, 108, 126 };
__constant__ int LG_SCALE_MIXED[16] = { 0, 18, 36, 54, 51, 35, 53, 71, 52, 70, 88, 69, 87, 105, 104, 122 };
__constant__ int LG_SCALE_LONG[16] = { 0, 10, 20, 30, 33, 21, 31, 41, 32, 42, 52, 43, 53, 63, 64, 74 };
__constant__ int LG_HUF_NOESC[15] = { 1, 2, 5, 7, 7, 10, 10, 13, 13, 13, 13, 13, 13, 13, 13 };
__constant__ int LG_SCFSI_BAND[5] = { 0, 6, 11, 16, 21 };

/* ---------------------------------------------------------------- warp reductions (integer: order-free) */
__device__ __forceinline__ int lg_wmax_i(int v)
{
#if defined(LG_EMULATE)
    for (int d = 16; d > 0; d >>= 1) { int o = __shfl_xor_sync(LG_FULL, v, d); v = v > o ? v : o; }
    return v;
#else
    return __reduce_max_sync(LG_FULL, v);
#endif
}
__device__ __forceinline__ int lg_wmin_i(int v)
{
#if defined(LG_EMULATE)
    for (int d = 16; d > 0; d >>= 1) { int o = __shfl_xor_sync(LG_FULL, v, d); v = v < o ? v : o; }
    return v;
#else
    return __reduce_min_sync(LG_FULL, v);
#endif
}
__device__ __forceinline__ unsigned lg_wsum_u(unsigned v)
{
#if defined(LG_EMULATE)
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(LG_FULL, v, d);
    return v;
#else
    return __reduce_add_sync(LG_FULL, v);
#endif
}
__device__ __forceinline__ unsigned lg_wor_u(unsigned v)
{
#if defined(LG_EMULATE)
    for (int d = 16; d > 0; d >>= 1) v |= __shfl_xor_sync(LG_FULL, v, d);
    return v;
#else
    return __reduce_or_sync(LG_FULL, v);
#endif
}
/* max of NON-NEGATIVE floats: their bit patterns order like integers, one REDUX instead of five shuffles */
__device__ __forceinline__ float lg_wmax_fpos(float v)
{
#if defined(LG_EMULATE)
    for (int d = 16; d > 0; d >>= 1) { float o = __shfl_xor_sync(LG_FULL, v, d); v = v > o ? v : o; }
    return v;
#else
    return __int_as_float(__reduce_max_sync(LG_FULL, __float_as_int(v)));
#endif
}
/* float max is exact and order-free */
__device__ __forceinline__ float lg_wmax_f(float v)
{
#if defined(LG_EMULATE)
    for (int d = 16; d > 0; d >>= 1) { float o = __shfl_xor_sync(LG_FULL, v, d); v = v > o ? v : o; }
    return v;
#else
    /* order-preserving map float -> int (flip the magnitude bits of negative values), then one integer reduction */
    int k = __float_as_int(v);
    k ^= (k >> 31) & 0x7fffffff;
    k = __reduce_max_sync(LG_FULL, k);
    k ^= (k >> 31) & 0x7fffffff;
    return __int_as_float(k);
#endif
}

__device__ __forceinline__ const uint8_t *lg_hlen(const LgDevCfg *__restrict__ c, int t) { return c->huff_len + c->huff_off[t]; }

template <class W>
__device__ __forceinline__ int lg_band_step(const LgQInfo &gi, const W *w, const int *sf, int sfb)
{
    return gi.global_gain - ((sf[sfb] + (gi.preflag ? lg_pretab(sfb) : 0)) << (gi.scalefac_scale + 1))
         - ((gi.sbg >> (4 * w->window[sfb])) & 15) * 8;
}

/* ---------------------------------------------------------------- Huffman table choice for a region whose largest
 * magnitude is mx (takehiro.c:618 choose_table_nonMMX): the magnitude class selects a packed length book
 * (LgDevCfg::huff_pk) and up to three candidate tables; escape classes add linbits per value >= 15. */
struct LgRegion { int base, t0, t1, t2, lb0, lb1; };
__device__ __forceinline__ void lg_region_class(const LgDevCfg *__restrict__ c, int mx, LgRegion &r)
{
    r.lb0 = r.lb1 = 0;
    if (mx > 15) {
        int const m = mx - 15;
        int choice, choice2;
        for (choice2 = 24; choice2 < 32; choice2++) if ((int) c->huff_linmax[choice2] >= m) break;
        for (choice = choice2 - 8; choice < 24; choice++) if ((int) c->huff_linmax[choice] >= m) break;
        r.base = 6 * 256; r.t0 = choice; r.t1 = r.t2 = choice2;
        r.lb0 = c->huff_xlen[choice]; r.lb1 = c->huff_xlen[choice2];
    }
    else {
        int const k = (mx <= 1) ? 0 : ((mx <= 3) ? mx - 1 : (mx <= 5 ? 3 : (mx <= 7 ? 4 : 5)));
        int const t = LG_HUF_NOESC[mx > 0 ? mx - 1 : 0];
        r.base = k * 256; r.t0 = t;
        r.t1 = (mx == 1) ? t : t + 1;
        r.t2 = (mx <= 3) ? r.t1 : t + 2;
    }
}
/* s0/s1/s2 = whole-region bit sums under the three candidates -> chosen table, adds its bits (count_bit_* epilogues) */
__device__ __forceinline__ int lg_region_pick(const LgRegion &r, unsigned s0, unsigned s1, unsigned s2, unsigned n15, int *bits)
{
    s0 += n15 * (unsigned) r.lb0; s1 += n15 * (unsigned) r.lb1; s2 += n15 * (unsigned) r.lb1;
    int t = r.t0;
    if (s0 > s1) { s0 = s1; t = r.t1; }
    if (s0 > s2) { s0 = s2; t = r.t2; }
    *bits += (int) s0;
    return t;
}

/* count1 quadruples [bigv, count1) counted with both count1 books (warp) */
__device__ __noinline__ void lg_count1_bits(const LgDevCfg *__restrict__ c, const int16_t *ix, int bigv, int count1, int lane, int *a1, int *a2)
{
    const uint8_t *t32l = lg_hlen(c, 32), *t33l = lg_hlen(c, 33);
    unsigned s1 = 0, s2 = 0;
    int const nq = (count1 - bigv) >> 2;
    for (int k = lane; k < nq; k += 32) {
        int const i = count1 - 4 * k;
        int const p = ((ix[i - 4] * 2 + ix[i - 3]) * 2 + ix[i - 2]) * 2 + ix[i - 1];
        s1 += __ldg(&t32l[p]); s2 += __ldg(&t33l[p]);
    }
    unsigned const a = lg_wsum_u(s1 | (s2 << 16));
    *a1 = (int) (a & 0xffffu);
    *a2 = (int) (a >> 16);
}

#ifndef LG_UNROLL_Q
#define LG_UNROLL_Q 1
#endif
#ifndef LG_UNROLL_C
#define LG_UNROLL_C 1
#endif
#define LG_PRAGMA_(x) _Pragma(#x)
#define LG_UNROLL(n) LG_PRAGMA_(unroll n)

/* ---------------------------------------------------------------- takehiro.c:767 count_bits = quantize_xrpow (:281) +
 * noquant_count_bits (:654).  pv.valid == 0 is the reference's prev_noise == 0 (step-size search).
 *
 * Inlined at its single call site (lg_outer_loop is a state machine around one count_bits and one calc_noise), so the
 * granule's scalar state lives in registers.  The loops over the lane's line pairs are rolled on purpose (the search
 * loop has to stay inside the instruction cache) and stop at qc.jn: lines above max_nonzero_coeff are zero when the
 * granule starts and quantize_xrpow only ever keeps or clears them (takehiro.c:300-330), so they are never touched. */
__device__ __forceinline__ int lg_noquant_tail(const LgDevCfg *__restrict__ c, LgQWarp *w, LgQInfo &gi, const LgQConst &qc, LgPrev &pv,
                                               int hi_nz, int hi_big, int lane);
/* takehiro.c:781-797, quality 0-2 only: in the bands that are in their half step, values whose xrpow is below
 * 0.6345 / IPOW20(gain) - the ones that only just rounded up to 1 - are dropped.  Runs over the quantised pairs again and
 * recomputes the lane's highest non-zero / big pair, so the main loop stays as it is for the other quality levels. */
__device__ __noinline__ unsigned lg_substep_zero(const LgDevCfg *__restrict__ c, LgQWarp *w, int gain, int sfbmax, int jn, int ilim, int lane)
{
    float const roundfac = (float) (0.634521682242439 / (double) __ldg(&c->ipow20[gain]));
    unsigned const ph0 = w->ph[0], ph1 = w->ph[1];
    int nz = -1, big = -1;
    for (int j = 0; j < jn; j++) {
        int const P = lane + 32 * j, i = 2 * P;
        int const sfb = w->line_sfb[i];
        unsigned nv = *reinterpret_cast<const unsigned *>(&w->ixw[i]);
        if (sfb < sfbmax && (((sfb < 32 ? ph0 : ph1) >> (sfb & 31)) & 1u)) {
            float2 const xp = *reinterpret_cast<const float2 *>(&w->xrpow[i]);
            if (!(xp.x >= roundfac)) nv &= 0xffff0000u;
            if (!(xp.y >= roundfac)) nv &= 0x0000ffffu;
            *reinterpret_cast<unsigned *>(&w->ixw[i]) = nv;
        }
        if (i < ilim) {
            if (nv != 0u) nz = P;
            if ((nv & 0xfffefffeu) != 0u) big = P;
        }
    }
    return (unsigned) (nz + 1) | ((unsigned) (big + 1) << 16);     /* both are < 288 */
}

template <int SUB>
__device__ __forceinline__ int lg_count_bits(const LgDevCfg *__restrict__ c, LgQWarp *w, LgQInfo &gi, const LgQConst &qc, LgPrev &pv, int lane)
{
    float const istep = __ldg(&c->ipow20[gi.global_gain]);
    if (gi.xrpow_max > (LG_IXMAX) / istep) return LG_LARGE_BITS;
    int const nsfb = (qc.block_type == LG_SHORT) ? 39 : 22;
    int const mnz = qc.max_nonzero_coeff;
    int const prev_data_use = pv.valid && (gi.global_gain == pv.global_gain);
    int const pv_count1 = pv.sfb_count1;
    /* per scalefactor band: 0 keep old values, 1 quantise, 2 quantise with the 0/1 shortcut */
    int T = 64;
    for (int r = 0; 32 * r < nsfb; r++) {
        int const sfb = lane + 32 * r;
        int term = 0;
        if (sfb < nsfb) {
            int step = -1;
            if (prev_data_use || qc.block_type == LG_NORM) step = lg_band_step(gi, w, w->sfw, sfb);
            int const pstep = w->pn_step[sfb];
            int const skip = prev_data_use && (pstep == step);
            int const cross = (w->lstart[sfb] + w->width[sfb]) > mnz;
            int const is01 = pv.valid && pv_count1 > 0 && sfb >= pv_count1 && pstep > 0 && step >= pstep;
            w->act[sfb] = skip ? 0 : (is01 ? 2 : 1);
            term = !skip && cross;
        }
        unsigned const m = __ballot_sync(LG_FULL, term);
        if (m && T == 64) T = 32 * r + (__ffs((int) m) - 1);
    }
    __syncwarp();
    float const compareval0 = (1.0f - 0.4054f) / istep;
    const float *adj = c->adj43asm;
    int const ilim = (mnz + 2) & ~1;               /* <= 576 because max_nonzero_coeff <= 575 */
    int hi_nz = -1, hi_big = -1;
    LG_UNROLL(LG_UNROLL_Q)
    for (int j = 0; j < qc.jn; j++) {
        int const P = lane + 32 * j, i = 2 * P;
        int const sfb = w->line_sfb[i];
        int a0, a1;                       /* action for line i and i+1: 0 keep, 1 quantise, 2 shortcut, 3 zero */
        a0 = a1 = w->act[sfb];
        if (T < 64) {
            if (i > mnz) a0 = a1 = 3;
            else {
                if (sfb == T) a0 = 1;
                a1 = (i + 1 == mnz) ? ((sfb == T) ? 1 : 3) : a0;
            }
        }
        unsigned const old = *reinterpret_cast<const unsigned *>(&w->ixw[i]);
        float2 const xp = *reinterpret_cast<const float2 *>(&w->xrpow[i]);
        int v0 = (int) (old & 0xffffu), v1 = (int) (old >> 16);
        if (a0 == 1) {
            float const xs = istep * xp.x;
            double d = (double) xs + 8388608.0;
            int const idx = __float_as_int((float) d) - 0x4b000000;
            v0 = __float_as_int((float) (d + (double) __ldg(&adj[idx]))) - 0x4b000000;
        }
        else if (a0 == 2) v0 = (compareval0 > xp.x) ? 0 : 1;
        else if (a0 == 3) v0 = 0;
        if (a1 == 1) {
            float const xs = istep * xp.y;
            double d = (double) xs + 8388608.0;
            int const idx = __float_as_int((float) d) - 0x4b000000;
            v1 = __float_as_int((float) (d + (double) __ldg(&adj[idx]))) - 0x4b000000;
        }
        else if (a1 == 2) v1 = (compareval0 > xp.y) ? 0 : 1;
        else if (a1 == 3) v1 = 0;
        unsigned const nv = (unsigned) v0 | ((unsigned) v1 << 16);
        *reinterpret_cast<unsigned *>(&w->ixw[i]) = nv;
        if (i < ilim) {
            if (nv != 0u) hi_nz = P;
            if ((nv & 0xfffefffeu) != 0u) hi_big = P;
        }
    }
    if (SUB) {
        unsigned const r = lg_substep_zero(c, w, gi.global_gain + gi.scalefac_scale, qc.sfbmax, qc.jn, ilim, lane);
        hi_nz = (int) (r & 0xffffu) - 1; hi_big = (int) (r >> 16) - 1;
    }
    return lg_noquant_tail(c, w, gi, qc, pv, hi_nz, hi_big, lane);
}

/* takehiro.c:654 noquant_count_bits on the ix in w->ixw; hi_nz / hi_big = the lane's highest pair (below the limit of
 * max_nonzero_coeff) that is non-zero / holds a value > 1 */
__device__ __forceinline__ int lg_noquant_tail(const LgDevCfg *__restrict__ c, LgQWarp *w, LgQInfo &gi, const LgQConst &qc, LgPrev &pv,
                                               int hi_nz, int hi_big, int lane)
{
    /* ---- noquant_count_bits: count1 / big_values split */
    pv.sfb_count1 = 0;
    hi_nz = lg_wmax_i(hi_nz);
    hi_big = lg_wmax_i(hi_big);
    __syncwarp();                                  /* the ix written above are read by other lanes below */
    int const c1p = hi_nz + 1;
    gi.count1 = 2 * c1p;
    int const nquads = (c1p - 1 - hi_big) >> 1;
    int const bigv = gi.count1 - 4 * nquads;
    gi.big_values = bigv;
    /* region split [0,a1) [a1,a2) [a2,bigv); the count1 quadruples and the regions' largest magnitudes in one go */
    int a1 = 0, a2 = 0, has2 = 0;
    if (bigv > 0) {
        if (qc.block_type == LG_SHORT) {
            a1 = 3 * c->sfb_s[3];
            if (a1 > bigv) a1 = bigv;
            a2 = bigv;
        }
        else if (qc.block_type == LG_NORM) {
            a1 = gi.region0_count = c->bv_scf[bigv - 2];
            a2 = gi.region1_count = c->bv_scf[bigv - 1];
            a2 = c->sfb_l[a1 + a2 + 2];
            a1 = c->sfb_l[a1 + 1];
            has2 = a2 < bigv;
        }
        else {
            gi.region0_count = 7;
            gi.region1_count = LG_SBMAX_L - 1 - 7 - 1;
            a1 = c->sfb_l[7 + 1];
            a2 = bigv;
            if (a1 > a2) a1 = a2;
        }
        a1 = a1 < bigv ? a1 : bigv;
        a2 = a2 < bigv ? a2 : bigv;
    }
    unsigned s1 = 0, s2 = 0;
    {
        const uint8_t *t32l = lg_hlen(c, 32), *t33l = lg_hlen(c, 33);
        const int16_t *ix = w->ixw;
        for (int k = lane; k < nquads; k += 32) {
            int const i = gi.count1 - 4 * k;
            int const p = ((ix[i - 4] * 2 + ix[i - 3]) * 2 + ix[i - 2]) * 2 + ix[i - 1];
            s1 += __ldg(&t32l[p]); s2 += __ldg(&t33l[p]);
        }
    }
    int m0 = 0, m1 = 0, m2 = 0;
    LG_UNROLL(LG_UNROLL_C)
    for (int j = 0; j < qc.jn; j++) {
        int const i = 2 * (lane + 32 * j);
        if (i < bigv) {
            unsigned const u = *reinterpret_cast<const unsigned *>(&w->ixw[i]);
            int const v = max((int) (u & 0xffffu), (int) (u >> 16));
            if (i < a1) m0 = max(m0, v);
            else if (i < a2) m1 = max(m1, v);
            else m2 = max(m2, v);
        }
    }
    unsigned const cb = lg_wsum_u(s1 | (s2 << 16));          /* two fields, each total < 2^13 */
    int bits = (int) (cb & 0xffffu);
    gi.count1table_select = 0;
    if (bits > (int) (cb >> 16)) { bits = (int) (cb >> 16); gi.count1table_select = 1; }
    gi.count1bits = bits;
    if (bigv == 0) return bits;
    m0 = lg_wmax_i(m0); m1 = lg_wmax_i(m1); m2 = lg_wmax_i(m2);
    if (max(m0, max(m1, m2)) > LG_IXMAX) return LG_LARGE_BITS;     /* cannot happen behind the xrpow_max guard above */
    /* lanes 0..2 classify one region each (choose_table's dispatch on the largest magnitude) */
    LgRegion R;
    lg_region_class(c, lane == 0 ? m0 : (lane == 1 ? m1 : m2), R);
    int const b0 = __shfl_sync(LG_FULL, R.base, 0), b1 = __shfl_sync(LG_FULL, R.base, 1), b2 = __shfl_sync(LG_FULL, R.base, 2);
    /* bit sums under the candidate tables, three 10-bit fields per region (a lane adds at most 9 x 31) */
    unsigned acc0 = 0, acc1 = 0, acc2 = 0, n15 = 0;
    const uint32_t *pk = c->huff_pk;
    LG_UNROLL(LG_UNROLL_C)
    for (int j = 0; j < qc.jn; j++) {
        int const i = 2 * (lane + 32 * j);
        if (i < bigv) {
            unsigned const u = *reinterpret_cast<const unsigned *>(&w->ixw[i]);
            unsigned x = u & 0xffffu, y = u >> 16;
            unsigned const over = (x >= 15u) + (y >= 15u);
            x = x < 15u ? x : 15u; y = y < 15u ? y : 15u;
            int const reg = (i >= a1) + (i >= a2);
            int const base = reg == 0 ? b0 : (reg == 1 ? b1 : b2);
            unsigned const e = __ldg(&pk[base + (int) ((x << 4) + y)]);
            if (reg == 0) acc0 += e; else if (reg == 1) acc1 += e; else acc2 += e;
            n15 += over << (10 * reg);
        }
    }
    n15 = lg_wsum_u(n15);
    unsigned const r0 = lg_wsum_u((acc0 & 0x3ffu) | (((acc0 >> 10) & 0x3ffu) << 16)), r1 = lg_wsum_u(acc0 >> 20);
    unsigned const r2 = lg_wsum_u((acc1 & 0x3ffu) | (((acc1 >> 10) & 0x3ffu) << 16)), r3 = lg_wsum_u(acc1 >> 20);
    unsigned const r4 = lg_wsum_u((acc2 & 0x3ffu) | (((acc2 >> 10) & 0x3ffu) << 16)), r5 = lg_wsum_u(acc2 >> 20);
    /* lane r picks region r's table (count_bit_* epilogues); a region the reference does not look at contributes nothing */
    {
        unsigned const s01 = lane == 0 ? r0 : (lane == 1 ? r2 : r4), sx = lane == 0 ? r1 : (lane == 1 ? r3 : r5);
        unsigned const nn = (n15 >> (10 * (lane < 3 ? lane : 0))) & 0x3ffu;
        int const mreg = lane == 0 ? m0 : (lane == 1 ? m1 : m2);
        int const used = lane == 0 ? (0 < a1) : (lane == 1 ? (a1 < a2) : (lane == 2 ? has2 : 0));
        int rb = 0, rt = 0;
        if (used && mreg) rt = lg_region_pick(R, s01 & 0xffffu, s01 >> 16, sx, nn, &rb);
        int const t0 = __shfl_sync(LG_FULL, rt, 0), t1 = __shfl_sync(LG_FULL, rt, 1), t2 = __shfl_sync(LG_FULL, rt, 2);
        bits += __shfl_sync(LG_FULL, rb, 0) + __shfl_sync(LG_FULL, rb, 1) + __shfl_sync(LG_FULL, rb, 2);
        if (0 < a1) gi.table_select[0] = t0;
        if (a1 < a2) gi.table_select[1] = t1;
        if (has2) gi.table_select[2] = t2;
    }
    if (qc.block_type == LG_NORM) {
        /* first sfb whose start is >= big_values (sfb_l is increasing and ends at 576) */
        int const below = (lane < 23) && (c->sfb_l[lane] < bigv);
        pv.sfb_count1 = __popc(__ballot_sync(LG_FULL, below));
    }
    return bits;
}

/* same decision by ONE lane over ix[lo..hi) in shared memory (used where many ranges are evaluated at
 * once, one per lane: best_huffman_divide) */
__device__ __noinline__ int lg_choose_table_serial(const LgDevCfg *__restrict__ c, const int16_t *ix, int lo, int hi, int *bits)
{
    unsigned mx = 0;
    for (int i = lo; i < hi; i++) { unsigned const v = (unsigned) ix[i]; if (v > mx) mx = v; }
    if (mx == 0) return 0;
    if (mx > 15) {
        if (mx > LG_IXMAX) { *bits = LG_LARGE_BITS; return -1; }
        int const m = (int) mx - 15;
        int choice, choice2;
        for (choice2 = 24; choice2 < 32; choice2++) if ((int) c->huff_linmax[choice2] >= m) break;
        for (choice = choice2 - 8; choice < 24; choice++) if ((int) c->huff_linmax[choice] >= m) break;
        unsigned const linbits = c->huff_xlen[choice] * 65536u + c->huff_xlen[choice2];
        unsigned sum = 0;
        for (int i = lo; i < hi; i += 2) {
            unsigned x = (unsigned) ix[i], y = (unsigned) ix[i + 1];
            if (x >= 15u) { x = 15u; sum += linbits; }
            if (y >= 15u) { y = 15u; sum += linbits; }
            sum += __ldg(&c->largetbl[(x << 4) + y]);
        }
        unsigned const sum2 = sum & 0xffffu;
        sum >>= 16u;
        if (sum > sum2) { sum = sum2; choice = choice2; }
        *bits += (int) sum;
        return choice;
    }
    int t1 = LG_HUF_NOESC[mx - 1];
    unsigned const xlen = c->huff_xlen[t1];
    if (mx == 1) {
        const uint8_t *h = lg_hlen(c, 1);
        unsigned sum = 0;
        for (int i = lo; i < hi; i += 2) sum += __ldg(&h[2 * ix[i] + ix[i + 1]]);
        *bits += (int) sum;
        return 1;
    }
    if (mx <= 3) {
        const uint32_t *table = (t1 == 2) ? c->table23 : c->table56;
        unsigned sum = 0;
        for (int i = lo; i < hi; i += 2) sum += __ldg(&table[(unsigned) ix[i] * xlen + (unsigned) ix[i + 1]]);
        unsigned const sum2 = sum & 0xffffu;
        sum >>= 16u;
        if (sum > sum2) { sum = sum2; t1++; }
        *bits += (int) sum;
        return t1;
    }
    const uint8_t *h1 = lg_hlen(c, t1), *h2 = lg_hlen(c, t1 + 1), *h3 = lg_hlen(c, t1 + 2);
    unsigned s1 = 0, s2 = 0, s3 = 0;
    for (int i = lo; i < hi; i += 2) {
        unsigned const x = (unsigned) ix[i] * xlen + (unsigned) ix[i + 1];
        s1 += __ldg(&h1[x]); s2 += __ldg(&h2[x]); s3 += __ldg(&h3[x]);
    }
    int t = t1;
    if (s1 > s2) { s1 = s2; t++; }
    if (s1 > s3) { s1 = s3; t = t1 + 2; }
    *bits += (int) s1;
    return t;
}


/* ---------------------------------------------------------------- quantize_pvt.c:815 calc_noise.
 * The squared errors of all lines of the bands that need recomputing are formed line-parallel (every lane its nine
 * pairs) into sq[]; then one lane per band adds them up serially, in the reference's order (calc_noise_core_c :750),
 * so every float sum is the reference's.  Bands whose step did not change reuse the cached noise (prev_noise). */
__device__ __forceinline__ void lg_calc_noise(const LgDevCfg *__restrict__ c, LgQWarp *w, const LgQInfo &gi, const LgQConst &qc,
                                                 LgNoiseRes *res, LgPrev &pv, int lane)
{
    float *bstep = reinterpret_cast<float *>(w->act);      /* per band: step size, or < 0 = cached */
    int *breg = w->act + 40;                               /* per band: 0 beyond count1, 1 count1 region, 2 big values */
    int need = 0;
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        if (sfb < 40) {
            float st = -1.f;
            if (sfb < qc.psymax) {
                int const s = lg_band_step(gi, w, w->sfw, sfb);
                if (!(pv.valid && w->pn_step[sfb] == s)) {
                    st = __ldg(&c->pow20[s + LG_QMAX2]);
                    int const j = w->lstart[sfb];
                    breg[sfb] = (j > gi.count1) ? 0 : ((j > gi.big_values) ? 1 : 2);
                    need = 1;
                }
            }
            bstep[sfb] = st;
        }
    }
    need = __any_sync(LG_FULL, need);
    __syncwarp();
    if (need) {
        /* lines above max_nonzero_coeff never enter a band's sum (see the truncation of l below) */
        LG_UNROLL(LG_UNROLL_C)
        for (int j = 0; j < qc.jn; j++) {
            int const i = 2 * (lane + 32 * j);
            int const sfb = w->line_sfb[i];
            float const step = bstep[sfb];
            if (step >= 0.f) {
                float2 const x = *reinterpret_cast<const float2 *>(&w->xr[i]);
                unsigned const u = *reinterpret_cast<const unsigned *>(&w->ixw[i]);
                int const reg = breg[sfb];
                float t0, t1;
                if (reg == 0) { t0 = x.x; t1 = x.y; }
                else if (reg == 1) {
                    t0 = fabsf(x.x) - ((u & 0xffffu) ? step : 0.f);
                    t1 = fabsf(x.y) - ((u >> 16) ? step : 0.f);
                }
                else {
                    t0 = fabsf(x.x) - __ldg(&c->pow43[u & 0xffffu]) * step;
                    t1 = fabsf(x.y) - __ldg(&c->pow43[u >> 16]) * step;
                }
                { float2 q2; q2.x = t0 * t0; q2.y = t1 * t1; *reinterpret_cast<float2 *>(&w->sq[i]) = q2; }
            }
        }
        __syncwarp();
    }
    int over = 0, ssd = 0;
    float max_noise = -20.0f;
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        if (sfb < qc.psymax) {
            float const r_l3_xmin = 1.f / w->l3_xmin[sfb];
            float distort_, noise;
            if (bstep[sfb] < 0.f) {
                distort_ = r_l3_xmin * w->pn_noise[sfb];
                noise = w->pn_noise_log[sfb];
            }
            else {
                int const width = w->width[sfb];
                int j = w->lstart[sfb];
                int l = width >> 1;
                if ((j + width) > qc.max_nonzero_coeff) {
                    int const usefullsize = qc.max_nonzero_coeff - j + 1;
                    l = usefullsize > 0 ? usefullsize >> 1 : 0;
                }
                noise = 0;
                const float2 *q = reinterpret_cast<const float2 *>(&w->sq[j]);
                for (int k = 0; k < l; k++) { float2 const v = q[k]; noise += v.x; noise += v.y; }
                w->pn_step[sfb] = lg_band_step(gi, w, w->sfw, sfb);
                w->pn_noise[sfb] = noise;
                distort_ = r_l3_xmin * noise;
                noise = (float) LG_FAST_LOG10_D(c->log_table, (distort_ > 1E-20f ? distort_ : 1E-20f));
                w->pn_noise_log[sfb] = noise;
            }
            w->distort[sfb] = distort_;
            if (noise > 0.0) {
                int tmp = (int) (noise * 10 + .5);
                if (tmp < 1) tmp = 1;
                ssd += tmp * tmp;
                over++;
            }
            max_noise = max_noise > noise ? max_noise : noise;
        }
    }
    pv.global_gain = gi.global_gain;
    res->over_count = (int) lg_wsum_u((unsigned) over);
    res->over_SSD = (int) lg_wsum_u((unsigned) ssd);
    res->max_noise = lg_wmax_f(max_noise);
    __syncwarp();
}

/* takehiro.c:1218 mpeg2_scale_bitcount (MPEG-2/2.5): four partitions of bands (nr_of_sfb_block, quantize_pvt.c:51; table 0 without,
 * table 2 with preflag; the row for short blocks counts the three windows), each coded with the bit length of its largest
 * scalefactor.  A partition's maximum out of range = "over": part2_length and scalefac_compress keep their values. */
__device__ __forceinline__ int lg_lsf_partition(int short_block, int preflag, int sfb)
{
    /* first band of partitions 1..3 (bands of a short block counted per window: sfb / 3) */
    int const b = short_block ? sfb / 3 : sfb;
    if (preflag) return short_block ? (b >= 12 ? 4 : (b >= 6)) : (b >= 21 ? 4 : (b >= 11));
    if (short_block) return b >= 12 ? 4 : b / 3;
    return b >= 21 ? 4 : (b < 6 ? 0 : 1 + (b - 6) / 5);
}
__device__ __forceinline__ int lg_lsf_partition_bands(int short_block, int preflag, int part)
{
    if (preflag) return part < 2 ? (short_block ? 18 : (part ? 10 : 11)) : 0;
    return short_block ? 9 : (part ? 5 : 6);
}
template <class W>
__device__ __noinline__ unsigned lg_scale_bitcount_lsf(W *w, int block_type, int preflag, int compress, int part2_length, int lane)
{
    const int *sf = w->sfw;
    int const sh = block_type == LG_SHORT;
    int m[4] = { 0, 0, 0, 0 };
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        int const part = sfb < 39 ? lg_lsf_partition(sh, preflag, sfb) : 4;
        int const v = part < 4 ? sf[sfb] : 0;
        for (int p = 0; p < 4; p++) if (part == p && v > m[p]) m[p] = v;
    }
    int over = 0, bits = 0, slen[4];
    for (int p = 0; p < 4; p++) {
        m[p] = lg_wmax_i(m[p]);
        int const range = preflag ? (p == 0 ? 7 : (p == 1 ? 3 : 0)) : (p < 2 ? 15 : 7);      /* max_range_sfac_tab rows 2 and 0 */
        over += m[p] > range;
        slen[p] = m[p] == 0 ? 0 : 32 - __clz(m[p]);                                     /* log2tab: bits of the maximum (<= 15) */
        bits += slen[p] * lg_lsf_partition_bands(sh, preflag, p);
    }
    if (!over) {
        part2_length = bits;
        compress = preflag ? 500 + slen[0] * 3 + slen[1] : (((slen[0] * 5) + slen[1]) << 4) + (slen[2] << 2) + slen[3];
    }
    return (unsigned) part2_length | ((unsigned) compress << 20) | ((unsigned) preflag << 29) | ((unsigned) (over != 0) << 30);
}

/* ---------------------------------------------------------------- takehiro.c:1135 mpeg1_scale_bitcount (and :1318 the dispatch to the
 * MPEG-2 form).  By value (three call sites, and the caller's scalars stay in registers): returns part2_length (LG_LARGE_BITS
 * = does not fit) | scalefac_compress << 20 | preflag << 29 | does-not-fit << 30. */
template <class W>
__device__ __noinline__ unsigned lg_scale_bitcount(W *w, int block_type, int sfbmax, int sfbdivide, int preflag, int compress, int lsf_part2, int lane)
{
    if (lsf_part2 >= 0) return lg_scale_bitcount_lsf(w, block_type, preflag, compress, lsf_part2, lane);
    int *sf = w->sfw;
    if (block_type != LG_SHORT) {
        if (!preflag) {
            int bad = 0;
            if (lane >= 11 && lane < LG_SBPSY_L) bad = sf[lane] < lg_pretab(lane);
            if (!__any_sync(LG_FULL, bad)) {
                preflag = 1;
                if (lane >= 11 && lane < LG_SBPSY_L) sf[lane] -= lg_pretab(lane);
                __syncwarp();
            }
        }
    }
    int m1 = 0, m2 = 0;
    for (int sfb = lane; sfb < sfbmax; sfb += 32) {
        if (sfb < sfbdivide) m1 = max(m1, sf[sfb]); else m2 = max(m2, sf[sfb]);
    }
    m1 = lg_wmax_i(m1);
    m2 = lg_wmax_i(m2);
    /* the 16 (slen1, slen2) candidates one per lane; smallest length wins, the lowest index on ties (as the reference's scan) */
    int key = 0x7fffffff;
    if (lane < 16 && m1 < (1 << lg_slen1(lane)) && m2 < (1 << lg_slen2(lane))) key = lg_scale_len(block_type == LG_SHORT, lane) * 16 + lane;
    key = lg_wmin_i(key);
    int part2_length = LG_LARGE_BITS;
    if (key != 0x7fffffff) { part2_length = key >> 4; compress = key & 15; }
    return (unsigned) part2_length | ((unsigned) compress << 20) | ((unsigned) preflag << 29) | ((unsigned) (part2_length == LG_LARGE_BITS) << 30);
}
#define LG_APPLY_SCALE_BITCOUNT(gi, r) do { (gi).part2_length = (int) ((r) & 0xfffffu); (gi).scalefac_compress = (int) (((r) >> 20) & 511u); (gi).preflag = (int) (((r) >> 29) & 1u); } while (0)
/* the last argument of lg_scale_bitcount: -1 for MPEG-1, else the current part2_length (kept when the MPEG-2 form does not fit) */
#define LG_LSF_ARG(c, gi) ((c)->mode_gr == 1 ? (gi).part2_length : -1)

/* quantize.c:540 loop_break */
template <class W>
__device__ __forceinline__ int lg_loop_break(const W *w, int sbg, int sfbmax, int lane)
{
    int unamp = 0;
    for (int sfb = lane; sfb < sfbmax; sfb += 32)
        if (w->sfw[sfb] + ((sbg >> (4 * w->window[sfb])) & 15) == 0) unamp = 1;
    return __any_sync(LG_FULL, unamp) ? 0 : 1;
}

/* Multiply the lines of the flagged bands (factor per band in act[] as float, 0 = untouched); returns the new xrpow_max.
 * Only lines up to max_nonzero_coeff are touched: the quantiser never reads xrpow above it.  The reference scales those
 * lines too and they take part in its running maximum, so their per-band maximum is carried in tail_max[] and scaled here
 * (max and a rounded multiply by a positive factor commute): xrpow_max stays exactly the reference's. */
__device__ __noinline__ float lg_scale_bands(LgQWarp *w, float xrpow_max, int jn, int lane)
{
    const float *fac = reinterpret_cast<const float *>(w->act);
    float mx = xrpow_max;
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        if (sfb < 40) {
            float const f = fac[sfb];
            if (f != 0.f) {
                float const tm = w->tail_max[sfb] * f;
                w->tail_max[sfb] = tm;
                if (tm > mx) mx = tm;
            }
        }
    }
    LG_UNROLL(LG_UNROLL_C)
    for (int j = 0; j < jn; j++) {
        int const i = 2 * (lane + 32 * j);
        float const f = fac[w->line_sfb[i]];
        if (f != 0.f) {
            float2 v = *reinterpret_cast<float2 *>(&w->xrpow[i]);
            v.x *= f; v.y *= f;
            *reinterpret_cast<float2 *>(&w->xrpow[i]) = v;
            if (v.x > mx) mx = v.x;
            if (v.y > mx) mx = v.y;
        }
    }
    mx = lg_wmax_fpos(mx);
    __syncwarp();
    return mx;
}

/* quantize.c:720 amp_scalefac_bands with noise_shaping_amp == 2 (quality 0/1): exactly one band, the first whose
 * distortion is the maximum; with substep shaping every other visit of a band is its half step - the band's flag flips
 * off and nothing is amplified (quantize.c:781-785).  Returns the new xrpow_max.  Out of line: the other quality levels
 * never come here and the search loop has to stay small. */
__device__ __noinline__ float lg_amp_one_band(LgQWarp *w, float trigger, float ifqstep34, float xrpow_max, int sfbmax, int jn, int substep, int lane)
{
    int first = 64;
    for (int r = 1; r >= 0; r--) { int const sfb = lane + 32 * r; if (sfb < sfbmax && !(w->distort[sfb] < trigger)) first = sfb; }
    first = lg_wmin_i(first);
    if (first >= sfbmax) return xrpow_max;
    if (substep) {
        unsigned const bit = 1u << (first & 31);
        unsigned const now = w->ph[first >> 5] ^ bit;
        __syncwarp();
        if (lane == 0) w->ph[first >> 5] = now;
        __syncwarp();
        if (!(now & bit)) return xrpow_max;
    }
    for (int r = 0; r < 2; r++) { int const sfb = lane + 32 * r; if (sfb < 40) reinterpret_cast<float *>(w->act)[sfb] = (sfb == first) ? ifqstep34 : 0.f; }
    if (lane == 0) w->sfw[first]++;
    __syncwarp();
    return lg_scale_bands(w, xrpow_max, jn, lane);
}

/* quantize.c:720 amp_scalefac_bands (noise_shaping_amp 0, 1 and 2; 3 is not selected by any quality level) */
template <int SUB>
__device__ __forceinline__ void lg_amp_scalefac_bands(const LgDevCfg *__restrict__ c, LgQWarp *w, LgQInfo &gi, const LgQConst &qc, int lane)
{
    float const ifqstep34 = (gi.scalefac_scale == 0) ? (float) 1.29683955465100964055 : (float) 1.68179283050742922612;
    float trigger = 0;
    for (int sfb = lane; sfb < qc.sfbmax; sfb += 32) if (trigger < w->distort[sfb]) trigger = w->distort[sfb];
    trigger = lg_wmax_fpos(trigger);
    int const substep = SUB;
    if (SUB && c->noise_shaping_amp == 2) { gi.xrpow_max = lg_amp_one_band(w, trigger, ifqstep34, gi.xrpow_max, qc.sfbmax, qc.jn, substep, lane); return; }
    if (c->noise_shaping_amp == 1) {
        if (trigger > 1.0) trigger = (float) sqrt((double) trigger);      /* pow(trigger, .5), see lg_math.cuh */
        else trigger = (float) (trigger * .95);
    }
    else {
        if (trigger > 1.0) trigger = 1.0f;
        else trigger = (float) (trigger * .95);
    }
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        float f = 0.f;
        if (sfb < qc.sfbmax && !(w->distort[sfb] < trigger)) { w->sfw[sfb]++; f = ifqstep34; }
        if (sfb < 40) reinterpret_cast<float *>(w->act)[sfb] = f;
        if (substep) {                             /* every amplified band flips its half-step flag (quantize.c:781) */
            unsigned const m = __ballot_sync(LG_FULL, f != 0.f);
            if (lane == 0) w->ph[r] ^= m;
        }
    }
    __syncwarp();
    gi.xrpow_max = lg_scale_bands(w, gi.xrpow_max, qc.jn, lane);
}

/* quantize.c:808 inc_scalefac_scale */
__device__ __forceinline__ void lg_inc_scalefac_scale(LgQWarp *w, LgQInfo &gi, const LgQConst &qc, int lane)
{
    float const ifqstep34 = (float) 1.29683955465100964055;
    for (int r = 0; r < 2; r++) {
        int const sfb = lane + 32 * r;
        if (sfb < 40) {
            float f = 0.f;
            if (sfb < qc.sfbmax) {
                int s = w->sfw[sfb];
                if (gi.preflag) s += lg_pretab(sfb);
                if (s & 1) { s++; f = ifqstep34; }
                w->sfw[sfb] = s >> 1;
            }
            reinterpret_cast<float *>(w->act)[sfb] = f;
        }
    }
    __syncwarp();
    gi.xrpow_max = lg_scale_bands(w, gi.xrpow_max, qc.jn, lane);
    gi.preflag = 0;
    gi.scalefac_scale = 1;
}

/* quantize.c:847 inc_subblock_gain (short blocks only; sfb_lmax == 0 because mixed blocks are never used) */
__device__ __forceinline__ int lg_inc_subblock_gain(const LgDevCfg *__restrict__ c, LgQWarp *w, LgQInfo &gi, const LgQConst &qc, int lane)
{
    int *scalefac = w->sfw;
    float *fac = reinterpret_cast<float *>(w->act);
    for (int window = 0; window < 3; window++) {
        int s1 = 0, s2 = 0;
        for (int sfb = qc.sfb_lmax + window + 3 * lane; sfb < qc.sfbmax; sfb += 96) {
            if (sfb < qc.sfbdivide) s1 = max(s1, scalefac[sfb]); else s2 = max(s2, scalefac[sfb]);
        }
        s1 = lg_wmax_i(s1);
        s2 = lg_wmax_i(s2);
        if (s1 < 16 && s2 < 8) continue;
        if (((gi.sbg >> (4 * window)) & 15) >= 7) return 1;
        gi.sbg += 1 << (4 * window);
        for (int r = 0; r < 2; r++) { int const k = lane + 32 * r; if (k < 40) fac[k] = 0.f; }
        __syncwarp();
        {
            int const sfb = qc.sfb_lmax + window + 3 * lane;
            if (sfb < qc.sfbmax) {
                int s = scalefac[sfb];
                s = s - (4 >> gi.scalefac_scale);
                if (s >= 0) scalefac[sfb] = s;
                else {
                    scalefac[sfb] = 0;
                    fac[sfb] = __ldg(&c->ipow20[210 + (s << (gi.scalefac_scale + 1))]);
                }
            }
            /* the band after sfbmax in this window (sfb12) always follows the gain: IPOW20(202) */
            if (lane == 0) fac[qc.sfbmax + window] = __ldg(&c->ipow20[202]);
        }
        __syncwarp();
        gi.xrpow_max = lg_scale_bands(w, gi.xrpow_max, qc.jn, lane);
    }
    return 0;
}

/* quantize.c:940 balance_noise (the two scale_bitcount calls share one site) */
template <int SUB>
__device__ __forceinline__ int lg_balance_noise(const LgDevCfg *__restrict__ c, LgQWarp *w, LgQInfo &gi, const LgQConst &qc, int lane)
{
    lg_amp_scalefac_bands<SUB>(c, w, gi, qc, lane);
    if (lg_loop_break(w, gi.sbg, qc.sfbmax, lane)) return 0;
    for (int pass = 0;; pass++) {
        unsigned const r = lg_scale_bitcount(w, qc.block_type, qc.sfbmax, qc.sfbdivide, gi.preflag, gi.scalefac_compress, LG_LSF_ARG(c, gi), lane);
        LG_APPLY_SCALE_BITCOUNT(gi, r);
        int status = (int) (r >> 30) & 1;
        if (!status) return 1;
        if (pass == 1) return 0;
        if (c->noise_shaping > 1) {
            if (SUB) {                                           /* quantize.c:971 */
                if (lane == 0) { w->ph[0] = 0; w->ph[1] = 0; }
                __syncwarp();
            }
            if (!gi.scalefac_scale) { lg_inc_scalefac_scale(w, gi, qc, lane); status = 0; }
            else if (qc.block_type == LG_SHORT && c->subblock_gain > 0)
                status = lg_inc_subblock_gain(c, w, gi, qc, lane) || lg_loop_break(w, gi.sbg, qc.sfbmax, lane);
        }
        if (status) return 0;
    }
}

/* quantize.c:585 quant_compare, mode 9 (the only one the bitrate presets select) */
__device__ __forceinline__ int lg_quant_compare(const LgNoiseRes &best, const LgNoiseRes &calc)
{
    int better;
    if (best.over_count > 0) {
        better = calc.over_SSD <= best.over_SSD;
        if (calc.over_SSD == best.over_SSD) better = calc.bits < best.bits;
    }
    else better = ((calc.max_noise < 0) && ((calc.max_noise * 10 + calc.bits) <= (best.max_noise * 10 + best.bits)));
    if (best.over_count == 0) better = better && calc.bits < best.bits;
    return better;
}

/* copy work -> best or best -> work (the reference's gr_info struct assignment); lines above max_nonzero_coeff are zero in both */
__device__ __noinline__ void lg_copy_ix_sf(LgQWarp *w, int to_best, int jn, int lane)
{
    unsigned *dix = reinterpret_cast<unsigned *>(to_best ? w->ixb : w->ixw);
    const unsigned *six = reinterpret_cast<const unsigned *>(to_best ? w->ixw : w->ixb);
    int *dsf = to_best ? w->sfbst : w->sfw;
    const int *ssf = to_best ? w->sfw : w->sfbst;
    for (int i = lane; i < 32 * jn; i += 32) dix[i] = six[i];
    for (int i = lane; i < 40; i += 32) dsf[i] = ssf[i];
    __syncwarp();
}

/* xrpow (and the maxima of its lines above max_nonzero_coeff) -> save_xrpow or back: quantize.c:1144 / :1188, VBR-old only */
__device__ __noinline__ void lg_copy_xrpow(LgQWarp *w, int to_save, int jn, int lane)
{
    float2 *d = reinterpret_cast<float2 *>(to_save ? w->save_xrpow : w->xrpow);
    const float2 *s = reinterpret_cast<const float2 *>(to_save ? w->xrpow : w->save_xrpow);
    for (int i = lane; i < 32 * jn; i += 32) d[i] = s[i];
    for (int i = lane; i < 40; i += 32) { if (to_save) w->tail_save[i] = w->tail_max[i]; else w->tail_max[i] = w->tail_save[i]; }
    __syncwarp();
}

/* A search loop that runs away means corrupted state (the reference asserts CurrentStep != 0): stop the kernel with
 * an error instead of hanging the GPU. */
__device__ __forceinline__ void lg_runaway()
{
#ifdef LG_EMULATE
    abort();
#else
    __trap();
#endif
}

/* quantize.c:1010 outer_loop with bin_search_StepSize (:367) folded in, written as a state machine around ONE inlined
 * count_bits and ONE inlined calc_noise, so that the granule's scalar state (gi, best, pv, the search variables) stays in
 * registers for the whole search.  phase: 0 step-size search, main loop; 1 its "while too many bits" tail; 2 the first
 * and 3 the second "raise global_gain until it fits" loop of a noise-shaping round (quantize.c:1083-1101).
 * On return the work set (gi, sfw, ixw) holds the chosen quantisation; the value is the number of distorted bands of it.
 * FL bit 0: substep shaping (quality 0-2).  FL bit 1: VBR-old - the caller runs several searches on one gr.ch, so xrpow is kept as of
 * the chosen quantisation (save_xrpow, quantize.c:1144/:1188) and sfb21_extra is the caller's (it switches it off near the bit limit). */
template <int FL>
__device__ __forceinline__ int lg_outer_loop(const LgDevCfg *__restrict__ c, LgQWarp *w, LgQInfo &gi, const LgQConst &qc, int targ_bits,
                                             int *old_value, int *current_step, int sfb21_vbro, int lane)
{
    constexpr int SUB = FL & 1, VBRO = (FL >> 1) & 1;
    int CurrentStep = *current_step, flag_GoneOver = 0, Direction = 0;
    int const start = *old_value;
    gi.global_gain = start;
    int const desired_rate = targ_bits - gi.part2_length;
    LgPrev pv; pv.valid = 0; pv.global_gain = 0; pv.sfb_count1 = 0;
    LgNoiseRes best_noise; best_noise.max_noise = 0.f; best_noise.over_count = 0; best_noise.over_SSD = 0; best_noise.bits = 0;
    LgQInfo best = gi;
    int age = 0, best_part2_3_length = 9999999, maxggain = 255, huff_bits = 0, phase = 0;
    for (int guard = 0;; guard++) {
        if (guard > 30000) lg_runaway();
        int const nBits = lg_count_bits<SUB>(c, w, gi, qc, pv, lane);
        if (phase == 0) {
            if (!(CurrentStep == 1 || nBits == desired_rate)) {
                int step;
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
                gi.global_gain += step;
                if (gi.global_gain < 0) { gi.global_gain = 0; flag_GoneOver = 1; }
                if (gi.global_gain > 255) { gi.global_gain = 255; flag_GoneOver = 1; }
                continue;
            }
            phase = 1;
        }
        if (phase == 1) {
            if (nBits > desired_rate && gi.global_gain < 255) { gi.global_gain++; continue; }
            *current_step = (start - gi.global_gain >= 4) ? 4 : 2;
            *old_value = gi.global_gain;
            gi.part2_3_length = nBits;
            if (!c->noise_shaping) return 100;
            pv.valid = 1; pv.global_gain = 0; pv.sfb_count1 = 0;
            for (int i = lane; i < 40; i += 32) { w->pn_step[i] = 0; w->pn_noise[i] = 0; w->pn_noise_log[i] = 0; }
            __syncwarp();
        }
        else if (phase == 2) {
            gi.part2_3_length = nBits;
            if (nBits > huff_bits && gi.global_gain <= maxggain) { gi.global_gain++; continue; }
            if (gi.global_gain > maxggain) break;
            if (best_noise.over_count == 0) { phase = 3; continue; }
        }
        else {
            gi.part2_3_length = nBits;
            if (nBits > best_part2_3_length && gi.global_gain <= maxggain) { gi.global_gain++; continue; }
            if (gi.global_gain > maxggain) break;
        }
        LgNoiseRes noise_info;
        lg_calc_noise(c, w, gi, qc, &noise_info, pv, lane);
        noise_info.bits = gi.part2_3_length;
        if (phase == 1) {
            best_noise = noise_info;
            best = gi;                            /* cod_info_w = *cod_info: from here gi is the work copy */
            lg_copy_ix_sf(w, 1, qc.jn, lane);
            if (VBRO) lg_copy_xrpow(w, 1, qc.jn, lane);
        }
        else {
            if (lg_quant_compare(best_noise, noise_info)) {
                best_part2_3_length = best.part2_3_length;
                best_noise = noise_info;
                best = gi;
                lg_copy_ix_sf(w, 1, qc.jn, lane);
                if (VBRO) lg_copy_xrpow(w, 1, qc.jn, lane);
                age = 0;
            }
            else if (c->full_outer_loop == 0) {
                if (++age > (SUB ? 20 : 3) && best_noise.over_count == 0) break;     /* quantize.c:1061-1066 */
                /* the noise_shaping_amp == 3 exits and the second (refinement) pass are not selected by any quality level */
            }
            if (!((gi.global_gain + gi.scalefac_scale) < 255)) break;
        }
        /* top of a noise-shaping round */
        if (VBRO ? sfb21_vbro : c->sfb21_extra) {
            if (w->distort[qc.sfbmax] > 1.0) break;
            if (qc.block_type == LG_SHORT && (w->distort[qc.sfbmax + 1] > 1.0 || w->distort[qc.sfbmax + 2] > 1.0)) break;
        }
        if (lg_balance_noise<SUB>(c, w, gi, qc, lane) == 0) break;
        maxggain = gi.scalefac_scale ? 254 : 255;
        huff_bits = targ_bits - gi.part2_length;
        if (huff_bits <= 0) break;
        phase = 2;
    }
    gi = best;
    lg_copy_ix_sf(w, 0, qc.jn, lane);
    if (VBRO) lg_copy_xrpow(w, 0, qc.jn, lane);
    return best_noise.over_count;
}

/* ---------------------------------------------------------------- quantize_pvt.c:589 calc_xmin: one lane per band */
__device__ __noinline__ void lg_calc_xmin(const LgDevCfg *__restrict__ c, LgQWarp *w, LgQConst &qc, const LgXmin *en, const LgXmin *thm,
                                             float ath_adjust_factor, int lane)
{
    float const eps = (float) 2.2204460492503131e-016;   /* DBL_EPSILON stored in a float */
    int over = 0;                                        /* ath_over != 0 (quantize_pvt.c:628, :715) */
    /* long bands */
    if (lane < qc.psy_lmax) {
        int const gsfb = lane;
        float xmin = lg_ath_adjust(c, ath_adjust_factor, c->ath_l[gsfb], c->ath_floor, c->athfixpoint);
        xmin *= c->longfact[gsfb];
        int const width = w->width[gsfb];
        int j = w->lstart[gsfb];
        float const rh1 = xmin / width;
        float rh2 = eps, en0 = 0.0f, rh3;
        for (int l = 0; l < width; ++l) {
            float const xa = w->xr[j++];
            float const x2 = xa * xa;
            en0 += x2;
            rh2 += (x2 < rh1) ? x2 : rh1;
        }
        if (en0 > xmin) over = 1;
        if (en0 < xmin) rh3 = en0;
        else if (rh2 < xmin) rh3 = xmin;
        else rh3 = rh2;
        xmin = rh3;
        float const e = en->l[gsfb];
        if (e > 1e-12f) {
            float x = en0 * thm->l[gsfb] / e;
            x *= c->longfact[gsfb];
            if (xmin < x) xmin = x;
        }
        xmin = ((double) xmin > 2.2204460492503131e-016) ? xmin : eps;
        w->l3_xmin[gsfb] = xmin;
        w->eac[gsfb] = (en0 > xmin + 1e-14f) ? 1 : 0;
    }
    /* highest non-zero line */
    int k = 0;
    LG_UNROLL(LG_UNROLL_C)
    for (int j = 0; j < 9; j++) {
        int const i = 2 * (lane + 32 * j);
        if (fabsf(w->xr[i + 1]) > 1e-12f) k = i + 1;
        else if (i > 0 && fabsf(w->xr[i]) > 1e-12f && k < i) k = i;
    }
    int max_nonzero = lg_wmax_i(k);
    if (qc.block_type != LG_SHORT) max_nonzero |= 1;
    else { max_nonzero /= 6; max_nonzero *= 6; max_nonzero += 5; }
    if (c->sfb21_extra == 0 && c->samplerate < 44000) {
        int const limit = (qc.block_type != LG_SHORT) ? c->sfb_l[c->samplerate <= 8000 ? 17 : 21] - 1 : 3 * c->sfb_s[c->samplerate <= 8000 ? 9 : 12] - 1;
        if (max_nonzero > limit) max_nonzero = limit;
    }
    qc.max_nonzero_coeff = max_nonzero;
    {
        int const ilim = (max_nonzero + 2) & ~1;
        qc.jn = (ilim + 63) >> 6;
        /* per band: largest xrpow above max_nonzero_coeff (lg_scale_bands keeps it up to date instead of the lines) */
        for (int r = 0; r < 2; r++) {
            int const sfb = lane + 32 * r;
            if (sfb < 40) {
                float tm = 0.f;
                int const j1 = w->lstart[sfb] + w->width[sfb];
                for (int j = max(w->lstart[sfb], ilim); j < j1; j++) { float const v = w->xrpow[j]; if (v > tm) tm = v; }
                w->tail_max[sfb] = tm;
            }
        }
    }
    /* short bands: one lane per sfb handles its three windows */
    {
        int const sfb = qc.sfb_smin + lane;
        int const gsfb = qc.psy_lmax + 3 * lane;
        if (gsfb < qc.psymax) {
            float tmpATH = lg_ath_adjust(c, ath_adjust_factor, c->ath_s[sfb], c->ath_floor, c->athfixpoint);
            tmpATH *= c->shortfact[sfb];
            int const width = w->width[gsfb];
            int j = w->lstart[gsfb];
            float xm[3];
            for (int b = 0; b < 3; b++) {
                float en0 = 0.0f, xmin, rh2 = eps, rh3;
                float const rh1 = tmpATH / width;
                for (int l = 0; l < width; ++l) {
                    float const xa = w->xr[j++];
                    float const x2 = xa * xa;
                    en0 += x2;
                    rh2 += (x2 < rh1) ? x2 : rh1;
                }
                if (en0 > tmpATH) over = 1;
                if (en0 < tmpATH) rh3 = en0;
                else if (rh2 < tmpATH) rh3 = tmpATH;
                else rh3 = rh2;
                xmin = rh3;
                float const e = en->s[sfb][b];
                if (e > 1e-12f) {
                    float x = en0 * thm->s[sfb][b] / e;
                    x *= c->shortfact[sfb];
                    if (xmin < x) xmin = x;
                }
                xmin = ((double) xmin > 2.2204460492503131e-016) ? xmin : eps;
                xm[b] = xmin;
                w->eac[gsfb + b] = (en0 > xmin + 1e-14f) ? 1 : 0;
            }
            if (c->use_temporal) {
                if (xm[0] > xm[1]) xm[1] += (xm[0] - xm[1]) * c->decay;
                if (xm[1] > xm[2]) xm[2] += (xm[1] - xm[2]) * c->decay;
            }
            w->l3_xmin[gsfb] = xm[0]; w->l3_xmin[gsfb + 1] = xm[1]; w->l3_xmin[gsfb + 2] = xm[2];
        }
    }
    qc.ath_over = __any_sync(LG_FULL, over);
    __syncwarp();
}

/* ---------------------------------------------------------------- takehiro.c:884 best_huffman_divide */
/* the candidates: region 0 = [0, sfb_l[r0+1]) for r0 < 16 and region 1 = [sfb_l[r0+1], sfb_l[r0+r1+2]) for the 16 x 8 pairs, each with its best table; thread t of nt */
template <class W>
__device__ __noinline__ void lg_recalc_divide_cand(const LgDevCfg *__restrict__ c, W *w, int bigv, int t, int nt)
{
    const int16_t *ix = w->ixw;
    for (int q = t; q < 144; q += nt) {
        if (q >= 128) {
            int const r0 = q - 128;
            int const a1 = c->sfb_l[r0 + 1];
            int bits = 0, tb = 0;
            if (a1 < bigv) tb = lg_choose_table_serial(c, ix, 0, a1, &bits);
            w->r0b[r0] = bits; w->r0t[r0] = tb;
        }
        else {
            int const r0 = q >> 3, r1 = q & 7;
            int const a1 = c->sfb_l[r0 + 1], a2 = c->sfb_l[r0 + r1 + 2];
            int bits = LG_LARGE_BITS, tb = 0;
            if (a1 < bigv && a2 < bigv) { bits = 0; tb = lg_choose_table_serial(c, ix, a1, a2, &bits); }
            w->comb_bits[q] = bits; w->comb_tbl[q] = tb;
        }
    }
}
/* per r0 + r1 (0..22) the best split */
template <class W>
__device__ __noinline__ void lg_recalc_divide_comb(W *w, int lane)
{
    if (lane < 23) {
        int best = LG_LARGE_BITS, div = 0, t0 = 0, t1 = 0;
        for (int r0 = 0; r0 < 16; r0++) {
            int const r1 = lane - r0;
            if (r1 < 0 || r1 > 7) continue;
            if (w->comb_bits[r0 * 8 + r1] == LG_LARGE_BITS) continue;
            int const bits = w->r0b[r0] + w->comb_bits[r0 * 8 + r1];
            if (best > bits) { best = bits; div = r0; t0 = w->r0t[r0]; t1 = w->comb_tbl[r0 * 8 + r1]; }
        }
        w->r01_bits[lane] = best; w->r01_div[lane] = div; w->r0_tbl[lane] = t0; w->r1_tbl[lane] = t1;
    }
    __syncwarp();
}
template <class W>
__device__ __forceinline__ void lg_recalc_divide_init(const LgDevCfg *__restrict__ c, W *w, int bigv, int lane)
{
    lg_recalc_divide_cand(c, w, bigv, lane, 32);
    __syncwarp();
    lg_recalc_divide_comb(w, lane);
}
/* takehiro.c:847 recalc_divide_sub: g2 is the candidate base, gi the current best */
template <class W>
__device__ __noinline__ void lg_recalc_divide_sub(const LgDevCfg *__restrict__ c, W *w, const LgQInfo &g2, LgQInfo &gi, int lane)
{
    int const bigv = g2.big_values;
    /* every lane evaluates one r2 candidate; the sequential scan then replays the reference's order */
    int r2 = lane + 2, bits_l = LG_LARGE_BITS, r2t = 0, a2 = 0;
    if (r2 < LG_SBMAX_L + 1) {
        a2 = c->sfb_l[r2];
        if (a2 < bigv && w->r01_bits[r2 - 2] != LG_LARGE_BITS) {
            bits_l = w->r01_bits[r2 - 2] + g2.count1bits;
            r2t = lg_choose_table_serial(c, w->ixw, a2, bigv, &bits_l);
        }
    }
    for (r2 = 2; r2 < LG_SBMAX_L + 1; r2++) {
        int const a = c->sfb_l[r2];
        if (a >= bigv) break;
        int const base = w->r01_bits[r2 - 2] + g2.count1bits;
        if (gi.part2_3_length <= base) break;
        int const bits = __shfl_sync(LG_FULL, bits_l, r2 - 2);
        int const tbl = __shfl_sync(LG_FULL, r2t, r2 - 2);
        if (gi.part2_3_length <= bits) continue;
        gi = g2;
        gi.part2_3_length = bits;
        gi.region0_count = w->r01_div[r2 - 2];
        gi.region1_count = r2 - 2 - w->r01_div[r2 - 2];
        gi.table_select[0] = w->r0_tbl[r2 - 2];
        gi.table_select[1] = w->r1_tbl[r2 - 2];
        gi.table_select[2] = tbl;
    }
}
template <class W>
__device__ __noinline__ void lg_best_huffman_divide(const LgDevCfg *__restrict__ c, W *w, LgQInfo &gi, const LgQConst &qc, int lane, int have_cand = 0)
{
    const int16_t *ix = w->ixw;
    if (qc.block_type == LG_SHORT && c->mode_gr == 1) return;          /* takehiro.c:899: not for short blocks of MPEG-2 */
    LgQInfo g2 = gi;
    if (qc.block_type == LG_NORM) {
        if (have_cand) lg_recalc_divide_comb(w, lane);       /* the group form evaluates the candidates on all its warps beforehand */
        else lg_recalc_divide_init(c, w, gi.big_values, lane);
        lg_recalc_divide_sub(c, w, g2, gi, lane);
    }
    int i = g2.big_values;
    if (i == 0 || (unsigned) (ix[i - 2] | ix[i - 1]) > 1) return;
    i = gi.count1 + 2;
    if (i > 576) return;
    g2 = gi;
    g2.count1 = i;
    int a1, a2;
    lg_count1_bits(c, ix, g2.big_values - 2, i, lane, &a1, &a2);
    i = g2.big_values - 2;
    g2.big_values = i;
    g2.count1table_select = 0;
    if (a1 > a2) { a1 = a2; g2.count1table_select = 1; }
    g2.count1bits = a1;
    if (qc.block_type == LG_NORM) lg_recalc_divide_sub(c, w, g2, gi, lane);
    else {
        g2.part2_3_length = a1;
        a1 = c->sfb_l[7 + 1];
        if (a1 > i) a1 = i;
        int t0 = g2.table_select[0], t1 = g2.table_select[1], bits = g2.part2_3_length;
        if (lane == 0) {
            if (a1 > 0) t0 = lg_choose_table_serial(c, ix, 0, a1, &bits);
            if (i > a1) t1 = lg_choose_table_serial(c, ix, a1, i, &bits);
        }
        g2.table_select[0] = __shfl_sync(LG_FULL, t0, 0);
        g2.table_select[1] = __shfl_sync(LG_FULL, t1, 0);
        g2.part2_3_length = __shfl_sync(LG_FULL, bits, 0);
        if (gi.part2_3_length > g2.part2_3_length) gi = g2;
    }
}

/* ---------------------------------------------------------------- takehiro.c:1021 best_scalefac_store (+ :964 scfsi_calc) */
template <class W>
__device__ __noinline__ void lg_best_scalefac_store(const LgDevCfg *__restrict__ c, W *w, LgQInfo &gi, const LgQConst &qc,
                                                       int gr, const int *sf_gr0, int bt_gr0, uint8_t scfsi[4], int lane)
{
    int *sf = w->sfw;
    int recalc = 0;
    for (int sfb = lane; sfb < qc.sfbmax; sfb += 32) {
        int const j0 = w->lstart[sfb], j1 = j0 + w->width[sfb];
        int l = j0;
        for (; l < j1; ++l) if (w->ixw[l] != 0) break;
        if (l == j1) { sf[sfb] = -2; recalc = 1; }
    }
    recalc = __any_sync(LG_FULL, recalc) ? -2 : 0;
    __syncwarp();
    if (!gi.scalefac_scale && !gi.preflag) {
        unsigned s = 0;
        for (int sfb = lane; sfb < qc.sfbmax; sfb += 32) if (sf[sfb] > 0) s |= (unsigned) sf[sfb];
        s = lg_wor_u(s);
        if (!(s & 1) && s != 0) {
            for (int sfb = lane; sfb < qc.sfbmax; sfb += 32) if (sf[sfb] > 0) sf[sfb] >>= 1;
            gi.scalefac_scale = recalc = 1;
            __syncwarp();
        }
    }
    if (!gi.preflag && qc.block_type != LG_SHORT && c->mode_gr == 2) {
        int bad = 0;
        if (lane >= 11 && lane < LG_SBPSY_L) bad = (sf[lane] < lg_pretab(lane) && sf[lane] != -2);
        if (!__any_sync(LG_FULL, bad)) {
            if (lane >= 11 && lane < LG_SBPSY_L && sf[lane] > 0) sf[lane] -= lg_pretab(lane);
            gi.preflag = recalc = 1;
            __syncwarp();
        }
    }
    for (int i = 0; i < 4; i++) scfsi[i] = 0;
    if (c->mode_gr == 2 && gr == 1 && bt_gr0 != LG_SHORT && qc.block_type != LG_SHORT) {
        /* scfsi_calc */
        for (int i = 0; i < 4; i++) {
            int diff = 0;
            int const sfb = LG_SCFSI_BAND[i] + lane;
            if (sfb < LG_SCFSI_BAND[i + 1]) diff = (sf_gr0[sfb] != sf[sfb] && sf[sfb] >= 0);
            if (!__any_sync(LG_FULL, diff)) {
                if (sfb < LG_SCFSI_BAND[i + 1]) sf[sfb] = -1;
                scfsi[i] = 1;
            }
        }
        __syncwarp();
        int s1 = 0, c1 = 0, s2 = 0, c2 = 0;
        if (lane < 11) { if (sf[lane] != -1) { c1 = 1; s1 = sf[lane]; } }
        else if (lane < LG_SBPSY_L) { if (sf[lane] != -1) { c2 = 1; s2 = sf[lane]; } }
        /* the reference's max starts at 0, so negative entries (-2) never raise it */
        s1 = lg_wmax_i(s1 > 0 ? s1 : 0); s2 = lg_wmax_i(s2 > 0 ? s2 : 0);
        unsigned const cc = lg_wsum_u((unsigned) c1 | ((unsigned) c2 << 8));
        c1 = (int) (cc & 0xff); c2 = (int) (cc >> 8);
        for (int i = 0; i < 16; i++)
            if (s1 < LG_SLEN1_N[i] && s2 < LG_SLEN2_N[i]) {
                int const cbits = LG_SLEN1_TAB[i] * c1 + LG_SLEN2_TAB[i] * c2;
                if (gi.part2_length > cbits) { gi.part2_length = cbits; gi.scalefac_compress = i; }
            }
        recalc = 0;
    }
    for (int sfb = lane; sfb < qc.sfbmax; sfb += 32) if (sf[sfb] == -2) sf[sfb] = 0;
    __syncwarp();
    if (recalc) {
        unsigned const r = lg_scale_bitcount(w, qc.block_type, qc.sfbmax, qc.sfbdivide, gi.preflag, gi.scalefac_compress, LG_LSF_ARG(c, gi), lane);
        LG_APPLY_SCALE_BITCOUNT(gi, r);
    }
}

/* ---------------------------------------------------------------- bit budget (reservoir.c, quantize_pvt.c:428/:492): lane-uniform scalar code */
/* bitstream.c:65 getframebits for the frame's bitrate index */
__device__ __forceinline__ int lg_frame_bits(const LgDevCfg *__restrict__ c, int bitrate_index, int padding)
{
    return 8 * ((c->version + 1) * 72000 * c->bitrate_kbps[bitrate_index] / c->samplerate + padding);
}
/* reservoir.c:83 ResvFrameBegin: returns fullFrameBits */
__device__ __forceinline__ int lg_resv_frame_begin(const LgDevCfg *__restrict__ c, int bitrate_index, int padding, int resv_size, int *mean_bits, int *resv_max)
{
    int const frameLength = lg_frame_bits(c, bitrate_index, padding);
    int const meanBits = (frameLength - c->sideinfo_len * 8) / c->mode_gr;
    int const resvLimit = (8 * 256) * c->mode_gr - 8;
    int rmax = c->buffer_constraint - frameLength;
    if (rmax > resvLimit) rmax = resvLimit;
    if (rmax < 0 || c->disable_reservoir) rmax = 0;
    int full = meanBits * c->mode_gr + (resv_size < rmax ? resv_size : rmax);
    if (full > c->buffer_constraint) full = c->buffer_constraint;
    *mean_bits = meanBits;
    *resv_max = rmax;
    return full;
}
__device__ __forceinline__ void lg_resv_max_bits(const LgDevCfg *__restrict__ c, int resv_size, int resv_max, int mean_bits, int *targ_bits, int *extra_bits, int cbr)
{
    int add_bits, targBits, extraBits;
    int ResvSize = resv_size;
    if (cbr) ResvSize += mean_bits;
    targBits = mean_bits;
    if (ResvSize * 10 > resv_max * 9) {
        add_bits = ResvSize - (resv_max * 9) / 10;
        targBits += add_bits;
    }
    else {
        add_bits = 0;
        if (!c->disable_reservoir) targBits = (int) (targBits - .1 * mean_bits);
    }
    extraBits = (ResvSize < (resv_max * 6) / 10 ? ResvSize : (resv_max * 6) / 10);
    extraBits -= add_bits;
    if (extraBits < 0) extraBits = 0;
    *targ_bits = targBits;
    *extra_bits = extraBits;
}
__device__ __forceinline__ int lg_on_pe(const LgDevCfg *__restrict__ c, int resv_size, int resv_max, const float pe[2], int targ_bits[2], int mean_bits, int cbr)
{
    int extra_bits = 0, tbits, bits, add_bits[2] = { 0, 0 }, max_bits, ch;
    int const nch = c->channels;
    lg_resv_max_bits(c, resv_size, resv_max, mean_bits, &tbits, &extra_bits, cbr);
    max_bits = tbits + extra_bits;
    if (max_bits > LG_MAX_BITS_PER_GRANULE) max_bits = LG_MAX_BITS_PER_GRANULE;
    for (bits = 0, ch = 0; ch < nch; ++ch) {
        targ_bits[ch] = LG_MAX_BITS_PER_CHANNEL < tbits / nch ? LG_MAX_BITS_PER_CHANNEL : tbits / nch;
        add_bits[ch] = (int) ((double) (targ_bits[ch] * pe[ch]) / 700.0 - targ_bits[ch]);
        if (add_bits[ch] > mean_bits * 3 / 4) add_bits[ch] = mean_bits * 3 / 4;
        if (add_bits[ch] < 0) add_bits[ch] = 0;
        if (add_bits[ch] + targ_bits[ch] > LG_MAX_BITS_PER_CHANNEL)
            add_bits[ch] = 0 > LG_MAX_BITS_PER_CHANNEL - targ_bits[ch] ? 0 : LG_MAX_BITS_PER_CHANNEL - targ_bits[ch];
        bits += add_bits[ch];
    }
    if (bits > extra_bits && bits > 0)
        for (ch = 0; ch < nch; ++ch) add_bits[ch] = extra_bits * add_bits[ch] / bits;
    for (ch = 0; ch < nch; ++ch) { targ_bits[ch] += add_bits[ch]; extra_bits -= add_bits[ch]; }
    for (bits = 0, ch = 0; ch < nch; ++ch) bits += targ_bits[ch];
    if (bits > LG_MAX_BITS_PER_GRANULE)
        for (ch = 0; ch < nch; ++ch) { targ_bits[ch] *= LG_MAX_BITS_PER_GRANULE; targ_bits[ch] /= bits; }
    return max_bits;
}
__device__ __forceinline__ void lg_reduce_side(int targ_bits[2], float ms_ener_ratio, int mean_bits, int max_bits)
{
    int move_bits;
    float fac = (float) (.33 * (.5 - ms_ener_ratio) / .5);
    if (fac < 0) fac = 0;
    if (fac > .5) fac = .5f;
    move_bits = (int) (fac * .5 * (targ_bits[0] + targ_bits[1]));
    if (move_bits > LG_MAX_BITS_PER_CHANNEL - targ_bits[0]) move_bits = LG_MAX_BITS_PER_CHANNEL - targ_bits[0];
    if (move_bits < 0) move_bits = 0;
    if (targ_bits[1] >= 125) {
        if (targ_bits[1] - move_bits > 125) {
            if (targ_bits[0] < mean_bits) targ_bits[0] += move_bits;
            targ_bits[1] -= move_bits;
        }
        else { targ_bits[0] += targ_bits[1] - 125; targ_bits[1] = 125; }
    }
    move_bits = targ_bits[0] + targ_bits[1];
    if (move_bits > max_bits) {
        targ_bits[0] = (max_bits * targ_bits[0]) / move_bits;
        targ_bits[1] = (max_bits * targ_bits[1]) / move_bits;
    }
}

/* bitstream.c:214 drain_into_ancillary: after the bytes "LAME" and the version string, how many bits are written one by one
 * with the alternating ancillary flag */
__device__ __forceinline__ int lg_drain_tail_bits(int remainingBits)
{
    for (int i = 0; i < 4; i++) if (remainingBits >= 8) remainingBits -= 8;
    if (remainingBits >= 32) for (int i = 0; i < 6 && remainingBits >= 8; ++i) remainingBits -= 8;
    return remainingBits;
}

/* quantize.c:1768 calc_target_bits (ABR): the four granule.channel targets of a frame from the perceptual entropies,
 * independent of each other; only the frame's maximum depends on the reservoir */
__device__ __forceinline__ void lg_calc_target_bits(const LgDevCfg *__restrict__ c, int resv_size, int padding, const float pe[2][2], const int bt[2][2],
                                                    const float ms_ener_ratio[2], int mode_ext, int targ_bits[2][2], int *analog_silence_bits)
{
    int const nch = c->channels;
    int mean_bits, rmax;
    int const max_frame_bits = lg_resv_frame_begin(c, c->vbr_max_bitrate_index, padding, resv_size, &mean_bits, &rmax);
    mean_bits = lg_frame_bits(c, 1, padding) - c->sideinfo_len * 8;
    *analog_silence_bits = mean_bits / (c->mode_gr * nch);
    mean_bits = c->vbr_mean_kbps * (576 * c->mode_gr) * 1000;
    if (c->substep_shaping & 1) mean_bits = (int) (mean_bits * 1.09);
    mean_bits /= c->samplerate;
    mean_bits -= c->sideinfo_len * 8;
    mean_bits /= (c->mode_gr * nch);
    float res_factor = (float) (.93 + .07 * (11.0 - c->compression_ratio) / (11.0 - 5.5));
    if (res_factor < .90) res_factor = (float) .90;
    if (res_factor > 1.00) res_factor = (float) 1.00;
    for (int gr = 0; gr < c->mode_gr; gr++) {
        int sum = 0;
        for (int ch = 0; ch < nch; ch++) {
            targ_bits[gr][ch] = (int) (res_factor * mean_bits);
            if (pe[gr][ch] > 700) {
                int add_bits = (int) ((pe[gr][ch] - 700) / 1.4);
                if (bt[gr][ch] == LG_SHORT) {
                    if (add_bits < mean_bits / 2) add_bits = mean_bits / 2;
                }
                if (add_bits > mean_bits * 3 / 2) add_bits = mean_bits * 3 / 2;
                else if (add_bits < 0) add_bits = 0;
                targ_bits[gr][ch] += add_bits;
            }
            if (targ_bits[gr][ch] > LG_MAX_BITS_PER_CHANNEL) targ_bits[gr][ch] = LG_MAX_BITS_PER_CHANNEL;
            sum += targ_bits[gr][ch];
        }
        if (sum > LG_MAX_BITS_PER_GRANULE)
            for (int ch = 0; ch < nch; ++ch) { targ_bits[gr][ch] *= LG_MAX_BITS_PER_GRANULE; targ_bits[gr][ch] /= sum; }
    }
    if (mode_ext == 2)
        for (int gr = 0; gr < c->mode_gr; gr++) lg_reduce_side(targ_bits[gr], ms_ener_ratio[gr], mean_bits * nch, LG_MAX_BITS_PER_GRANULE);
    int totbits = 0;
    for (int gr = 0; gr < c->mode_gr; gr++)
        for (int ch = 0; ch < nch; ch++) {
            if (targ_bits[gr][ch] > LG_MAX_BITS_PER_CHANNEL) targ_bits[gr][ch] = LG_MAX_BITS_PER_CHANNEL;
            totbits += targ_bits[gr][ch];
        }
    if (totbits > max_frame_bits && totbits > 0)
        for (int gr = 0; gr < c->mode_gr; gr++)
            for (int ch = 0; ch < nch; ch++) { targ_bits[gr][ch] *= max_frame_bits; targ_bits[gr][ch] /= totbits; }
}

/* ---------------------------------------------------------------- the kernel */
/* SUB = 1: the build of the kernel with substep shaping and one-band amplification (quality 0-2; cfg->substep_shaping & 2).
 * The other quality levels run SUB = 0, whose search loop does not carry that code. */
#ifdef LG_D_RESTRICT
#define LG_XR __restrict__
#define LG_XLD(p) __ldg(p)
#else
#define LG_XR
#define LG_XLD(p) (*(p))
#endif
/* FL bit 2 ("dense"): for batches with more than six streams per SM.  Left alone the compiler takes 149 registers, which is best for the
 * latency of one warp but allows six CTAs per SM; shared memory allows seven, and with the registers kept within that (120) a 4096-stream
 * batch runs 9 % faster (33.4 -> 30.5 ms for 8 frames) while the 512-stream one would lose 2 % (6.47 -> 6.61 ms). */
template <int FL>
__global__ void __launch_bounds__(64, (FL & 4) ? 7 : 1)
lg_kernel_quant(const LgDevCfg *__restrict__ cfg, const float *LG_XR xr_in, const LgPsyOut *LG_XR psy, const LgFrameCtl *LG_XR frm,
                LgGranuleOut *__restrict__ gout, LgFrameOut *__restrict__ fout,
                LgStreamState *__restrict__ state, const int *__restrict__ nfr, int nframes, int f0, int f1 /* this launch: frames f0 .. f1-1 */)
{
    constexpr int SUB = FL & 1;
    LG_DYN_SMEM(LgSmemD, sm);
    int const lane = threadIdx.x & 31, ch = threadIdx.x >> 5;
    int const stream = blockIdx.x;
    int const nch = cfg->channels;
    LgStreamState *st = state + stream;
    LgQWarp *w = &sm->w[ch];
    int resv_size = st->resv_size, main_data_begin = st->main_data_begin;
    int old_value = st->old_value[ch], current_step = st->current_step[ch];
    int anc_flag = st->ancillary_flag, pay_off = 0;

    int const my_frames = min(nfr[stream], f1);
    if (my_frames <= 0) return;                        /* nothing of this stream in this step: its state is not ours to write back (another step's kernel may own it) */
    int const mgr = cfg->mode_gr;                      /* granules per frame: 2 (MPEG-1) or 1 (MPEG-2/2.5) */
    if (f0 > 0 && f0 < my_frames) {                     /* a later piece of the batch: the payload continues behind the previous frame's */
        const LgFrameOut *pf = fout + (size_t) stream * nframes + (f0 - 1);
        pay_off = pf->pay_off + pf->pay_bytes;
    }
    for (int frame = f0; frame < my_frames; frame++) {
        const LgFrameCtl *F = frm + (size_t) stream * nframes + frame;
        int const padding = F->padding, mode_ext = F->mode_ext;
        int const abr = (cfg->vbr == 3);
        int bitrate_index = cfg->bitrate_index;
        int mean_bits, resv_max, analog_silence_bits = 0;
        int targ_abr[2][2] = { { 0, 0 }, { 0, 0 } };
        if (abr) {
            /* quantize.c:1900 ABR_iteration_loop: all four targets up front, the frame size is chosen afterwards */
            const LgPsyOut *P0 = psy + (size_t) stream * 2 * nframes + mgr * frame;
            float const pe4[2][2] = { { F->pe_use[0][0], F->pe_use[0][1] }, { F->pe_use[1][0], F->pe_use[1][1] } };
            int const bt4[2][2] = { { P0[0].block_type[0], P0[0].block_type[1] }, { P0[mgr - 1].block_type[0], P0[mgr - 1].block_type[1] } };
            float const mer[2] = { F->ms_ener_ratio[0], F->ms_ener_ratio[1] };
            lg_calc_target_bits(cfg, resv_size, padding, pe4, bt4, mer, mode_ext, targ_abr, &analog_silence_bits);
            mean_bits = resv_max = 0;
        }
        else (void) lg_resv_frame_begin(cfg, bitrate_index, padding, resv_size, &mean_bits, &resv_max);   /* reservoir.c:83 */
        uint8_t scfsi[4] = { 0, 0, 0, 0 };
        int frame_used = 0;
        for (int gr = 0; gr < mgr; gr++) {
            int const gb = mgr * frame + gr;
            const LgPsyOut *P = psy + (size_t) stream * 2 * nframes + gb;
            int targ_bits[2];
            if (abr) { targ_bits[0] = targ_abr[gr][0]; targ_bits[1] = targ_abr[gr][1]; }
            else {
                float pe[2] = { F->pe_use[gr][0], F->pe_use[gr][1] };
                int const max_bits = lg_on_pe(cfg, resv_size, resv_max, pe, targ_bits, mean_bits, gr);
                if (mode_ext == 2) lg_reduce_side(targ_bits, F->ms_ener_ratio[gr], mean_bits, max_bits);
            }
            int used = 0;
            if (ch < nch) {
                LgQInfo gi;
                LgQConst qc;
                int const rch = (mode_ext == 2) ? ch + 2 : ch;
                const LgXmin *en = &P->en[rch], *thm = &P->thm[rch];
                /* quantize.c:226 init_outer_loop (the short-block reorder was done by kernel C) */
                gi.part2_3_length = 0; gi.big_values = 0; gi.count1 = 0; gi.global_gain = 210; gi.scalefac_compress = 0;
                gi.table_select[0] = gi.table_select[1] = gi.table_select[2] = 0;
                gi.sbg = 0;
                gi.region0_count = 0; gi.region1_count = 0; gi.preflag = 0; gi.scalefac_scale = 0;
                gi.count1table_select = 0; gi.part2_length = 0; gi.count1bits = 0; gi.xrpow_max = 0;
                qc.block_type = P->block_type[ch];
                qc.sfb_lmax = LG_SBPSY_L; qc.sfb_smin = LG_SBPSY_S;
                qc.psy_lmax = cfg->sfb21_extra ? LG_SBMAX_L : LG_SBPSY_L;
                if (cfg->samplerate <= 8000) { qc.sfb_lmax = 17; qc.sfb_smin = 9; qc.psy_lmax = 17; }      /* quantize.c:252-256 */
                qc.psymax = qc.psy_lmax; qc.sfbmax = qc.sfb_lmax; qc.sfbdivide = 11;
                if (qc.block_type == LG_SHORT) {
                    qc.sfb_smin = 0; qc.sfb_lmax = 0;
                    qc.psymax = 3 * (cfg->sfb21_extra ? LG_SBMAX_S : LG_SBPSY_S);
                    qc.sfbmax = 3 * LG_SBPSY_S;
                    if (cfg->samplerate <= 8000) qc.psymax = qc.sfbmax = 3 * 9;                           /* quantize.c:284-289 */
                    qc.sfbdivide = qc.sfbmax - 18;
                    qc.psy_lmax = 0;
                }
                qc.max_nonzero_coeff = 575; qc.jn = 9;
                for (int r = 0; r < 2; r++) {
                    int const k = lane + 32 * r;
                    if (k <= 40) {
                        int ws = 0, wn = 3, ls = 576;
                        if (qc.block_type == LG_SHORT) {
                            if (k < 39) {
                                int const sfb = k / 3;
                                ws = cfg->sfb_s[sfb + 1] - cfg->sfb_s[sfb];
                                wn = k % 3;
                                ls = 3 * cfg->sfb_s[sfb] + wn * ws;
                            }
                        }
                        else if (k < LG_SBMAX_L) { ws = cfg->sfb_l[k + 1] - cfg->sfb_l[k]; ls = cfg->sfb_l[k]; }
                        if (k < 40) { w->width[k] = ws; w->window[k] = wn; w->sfw[k] = 0; }
                        w->lstart[k] = ls;
                    }
                }
                __syncwarp();
                {   /* line -> band map and the lines themselves */
                    const float *src = xr_in + (((size_t) stream * 2 * nframes + gb) * 2 + ch) * 576;
                    for (int i = lane; i < 144; i += 32) reinterpret_cast<float4 *>(w->xr)[i] = LG_XLD(reinterpret_cast<const float4 *>(src) + i);
                    const unsigned *map = reinterpret_cast<const unsigned *>(qc.block_type == LG_SHORT ? cfg->line_sfb_s : cfg->line_sfb_l);
                    for (int i = lane; i < 144; i += 32) reinterpret_cast<unsigned *>(w->line_sfb)[i] = __ldg(map + i);
                }
                __syncwarp();
                /* quantize.c:110 init_xrpow (upper = 575) */
                float mx = 0.f, amax = 0.f;
    LG_UNROLL(LG_UNROLL_C)
                for (int j = 0; j < 9; j++) {
                    int const i = 2 * (lane + 32 * j);
                    float const t0 = fabsf(w->xr[i]), t1 = fabsf(w->xr[i + 1]);
                    float const p0 = (float) sqrt((double) t0 * sqrt((double) t0));
                    float const p1 = (float) sqrt((double) t1 * sqrt((double) t1));
                    w->xrpow[i] = p0; w->xrpow[i + 1] = p1;
                    if (p0 > mx) mx = p0;
                    if (p1 > mx) mx = p1;
                    if (t0 > amax) amax = t0;
                    if (t1 > amax) amax = t1;
                    *reinterpret_cast<unsigned *>(&w->ixw[i]) = 0u;
                }
                gi.xrpow_max = lg_wmax_fpos(mx);
                amax = lg_wmax_fpos(amax);
                __syncwarp();
                /* sum > 1e-20 ?  A serial non-negative float sum is >= its largest term, so only a spectrum
                 * whose largest magnitude is itself <= 1e-20 needs the exact serial sum of the reference. */
                int nonzero = amax > (float) 1E-20;
                if (!nonzero && amax > 0.f) {
                    float sum = 0;
                    for (int i = 0; i < 576; ++i) sum += fabsf(w->xr[i]);
                    nonzero = sum > (float) 1E-20;
                }
                if (nonzero) {
                    {   /* through a copy: qc and gi must not have their address taken, or they leave the registers */
                        LgQConst qx = qc;
                        lg_calc_xmin(cfg, w, qx, en, thm, F->ath_adjust_factor, lane);
                        qc.max_nonzero_coeff = qx.max_nonzero_coeff; qc.jn = qx.jn;
                        if (abr && qx.ath_over == 0) targ_bits[ch] = analog_silence_bits;     /* quantize.c:1951 analog silence */
                    }
                    if (SUB) {                                    /* quantize.c:131-137 */
                        if (lane == 0) w->ph[0] = w->ph[1] = ~0u;
                        __syncwarp();
                    }
                    (void) lg_outer_loop<SUB>(cfg, w, gi, qc, targ_bits[ch], &old_value, &current_step, 0, lane);
                }
                /* quantize.c:1213 iteration_finish_one */
                {
                    LgQInfo gm = gi;
                    LgQConst qx = qc;
                    lg_best_scalefac_store(cfg, w, gm, qx, gr, sm->sf_gr0[ch], sm->bt_gr0[ch], scfsi, lane);
                    if (cfg->use_best_huffman == 1) lg_best_huffman_divide(cfg, w, gm, qx, lane);
                    gi = gm;
                }
                used = gi.part2_3_length + gi.part2_length;
                if (gr == 0) {
                    for (int i = lane; i < 40; i += 32) sm->sf_gr0[ch][i] = w->sfw[i];
                    if (lane == 0) sm->bt_gr0[ch] = qc.block_type;
                }
                /* hand the granule to the bit packer */
                LgGranuleOut *o = gout + (((size_t) stream * 2 * nframes + gb) * 2 + ch);
    LG_UNROLL(LG_UNROLL_C)
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
                    sm->used_bits[ch] = used;
                }
            }
            else if (lane == 0) sm->used_bits[ch] = 0;
            __syncthreads();
            resv_size -= sm->used_bits[0] + sm->used_bits[1];        /* reservoir.c:226 ResvAdjust */
            frame_used += sm->used_bits[0] + sm->used_bits[1];
            __syncthreads();
        }
        /* reservoir.c:239 ResvFrameEnd + the main_data_begin recurrence of format_bitstream (bitstream.c:937) */
        {
            if (abr) {
                /* the smallest frame that brings the reservoir back to a non-negative size (quantize.c:1962-1967) */
                for (bitrate_index = cfg->vbr_min_bitrate_index; bitrate_index <= cfg->vbr_max_bitrate_index; bitrate_index++)
                    if (lg_resv_frame_begin(cfg, bitrate_index, padding, resv_size, &mean_bits, &resv_max) >= 0) break;
                if (bitrate_index > cfg->vbr_max_bitrate_index) lg_runaway();
            }
            int stuffingBits = 0, over_bits, drain_pre = 0, drain_post = 0;
            resv_size += mean_bits * cfg->mode_gr;
            if ((over_bits = resv_size % 8) != 0) stuffingBits += over_bits;
            over_bits = (resv_size - stuffingBits) - resv_max;
            if (over_bits > 0) stuffingBits += over_bits;
            int const mdb_bytes = (main_data_begin * 8 < stuffingBits ? main_data_begin * 8 : stuffingBits) / 8;
            drain_pre += 8 * mdb_bytes;
            stuffingBits -= 8 * mdb_bytes;
            resv_size -= 8 * mdb_bytes;
            int const mdb_header = main_data_begin - mdb_bytes;   /* value written into this frame's side info */
            drain_post += stuffingBits;
            resv_size -= stuffingBits;
            main_data_begin = resv_size / 8;                       /* bitstream.c:937-951: mdb*8 == ResvSize */
            LgFrameOut *fo = fout + (size_t) stream * nframes + frame;
            /* payload of this frame = ancillary drain + main data + ancillary drain: always whole bytes, because
             * ResvFrameEnd keeps the reservoir a multiple of 8 bits */
            int const pay_bits = drain_pre + frame_used + drain_post;
            if (pay_bits & 7) lg_runaway();
            int const anc_pre = anc_flag;
            if (!cfg->disable_reservoir) anc_flag ^= (lg_drain_tail_bits(drain_pre) + lg_drain_tail_bits(drain_post)) & 1;
            if (lane == 0) {
                if (ch == 0) {
                    fo->main_data_begin = mdb_header; fo->drain_pre = drain_pre; fo->drain_post = drain_post;
                    fo->padding = padding; fo->mode_ext = mode_ext; fo->resv_size = resv_size;
                    fo->pay_off = pay_off; fo->pay_bytes = pay_bits >> 3;
                    fo->anc_pre = (uint8_t) anc_pre; fo->anc_post = (uint8_t) anc_flag; fo->pad_[0] = fo->pad_[1] = 0; fo->bitrate_index = bitrate_index;
                }
                if (ch < nch) for (int i = 0; i < 4; i++) fo->scfsi[ch][i] = scfsi[i];
                else for (int i = 0; i < 4; i++) fo->scfsi[ch][i] = 0;
            }
            pay_off += pay_bits >> 3;
        }
    }
    if (lane == 0) {
        if (ch == 0) { st->resv_size = resv_size; st->main_data_begin = main_data_begin; st->ancillary_flag = anc_flag; }
        st->old_value[ch] = old_value;
        st->current_step[ch] = current_step;
    }
}
