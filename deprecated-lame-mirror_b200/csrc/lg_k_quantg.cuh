// lg_k_quantg.cuh - kernel D in its group form: NW warps per granule.channel, the 576 lines in REGISTERS.
//
// lg_k_quant.cuh keeps a gr.ch on one warp with its lines in shared memory: 1024 warps for 512 streams, each a chain of dependent
// shared-memory loops - 24 % of the issue slots used.  Here a gr.ch is worked on by a GROUP of NW warps (NW = 3: 96 threads):
//
//   * line level: thread g of the group owns the line pairs g, g + 32 NW, ... (NP = 3 pairs for NW = 3) and keeps their xrpow, the
//     work quantisation and the best quantisation so far in registers; the loops over them are unrolled, nothing line-sized is read
//     from shared memory inside the search except xr (calc_noise) and the neighbours' values for the count1 quadruples;
//   * band level (22/39 scalefactor bands: steps, scalefactors, distortions, the cached noise): every warp of the group keeps its
//     OWN replica in shared memory and does the (cheap) band arithmetic redundantly, so no broadcast of band decisions is needed and
//     all warps walk through the same control flow;
//   * the warps of a group meet only where line results are combined: three exchanges per count_bits (largest non-zero pair |
//     count1 bits and region maxima | region bit sums) and one per calc_noise (the squared errors, summed per band in the
//     reference's order by one lane per band).  An exchange = warp reduction (REDUX), one record per warp into a two-deep ring in
//     shared memory, one named barrier (bar.sync 1+ch, 32 NW), every thread reads the NW records.
//
// The arithmetic, the search (outer_loop as a state machine around one count_bits and one calc_noise) and every reference quirk are
// those of lg_k_quant.cuh - the per-pair and per-band expressions are the same code; what changed is who executes them.  Quality 0-2
// (substep shaping) and VBR-old keep the one-warp kernel.  Reference lines: see lg_k_quant.cuh.
#pragma once
#include "lg_k_quant.cuh"

struct __attribute__((aligned(16))) LgGBand {            /* band-level work set: one replica per warp of the group */
    float l3_xmin[40], distort[40], pn_noise[40], pn_noise_log[40];
    int   pn_step[40], sfw[40], sfbst[40], width[40], lstart[41], act[80];
    float tail_max[40];
    uint8_t window[40];
    LgQInfo best;                 /* outer_loop's best quantisation so far (the reference's cod_info copy), scalar part */
};
#ifndef LG_G_LINES_SMEM
#define LG_G_LINES_SMEM 1
#endif
#if LG_G_LINES_SMEM
/* experiment: the own lines stay in shared memory (every thread its own elements, no extra barriers) and the loops over them are
 * rolled - smaller code (the search loop has to fit the 32 KB instruction cache) and fewer registers against less ILP */
#define LG_G_UNROLL _Pragma("unroll 1")
#ifndef LG_G_HOT_UNROLL
#define LG_G_HOT_UNROLL 1
#endif
#define LG_G_PRAGMA_(x) _Pragma(#x)
#define LG_G_PRAGMA(x) LG_G_PRAGMA_(x)
#define LG_G_UNROLL_HOT LG_G_PRAGMA(unroll LG_G_HOT_UNROLL)
template <class T, int NT> struct LgGOwn { T *base; int g; __device__ __forceinline__ T &operator[](int j) const { return base[g + NT * j]; } };
template <int NT> struct LgGOwnSfb { const uint8_t *base; int g; __device__ __forceinline__ int operator[](int j) const { return base[g + NT * j]; } };
#else
#define LG_G_UNROLL _Pragma("unroll")
#define LG_G_UNROLL_HOT _Pragma("unroll")
#endif
template <int NW> struct __attribute__((aligned(16))) LgGChan : LgGBand {      /* the base is warp 0's replica */
    float xr[576], sq[576];
#if LG_G_LINES_SMEM
    float xrpow[576];
    int16_t ixb[576];
    uint8_t pair_sfb[320];
#endif
    int16_t ixw[576];
    int r01_bits[24], r01_div[24], r0_tbl[24], r1_tbl[24];
    int comb_bits[128], comb_tbl[128], r0b[16], r0t[16];
    int red[2][NW][8];
    LgGBand rep[NW > 1 ? NW - 1 : 1];
};
struct LgGFrame {                 /* the stream's bit budget: kept and updated by thread 0 of the CTA only, read by everybody */
    int targ_bits[2], analog_silence_bits;
    int old_value[2], current_step[2];
    int used_bits[2];
    int more;                     /* another granule follows */
};
template <int NW> struct LgSmemG {
    LgGChan<NW> c[2];
    LgGFrame fr;
    int sf_gr0[2][40];
    int bt_gr0[2];
    int sfb_l[24];
    uint8_t bv_scf[576];
};
#if LG_G_LINES_SMEM
#define LG_G_XRPOW0(cs) (cs)->xrpow
#else
#define LG_G_XRPOW0(cs) (cs)->sq
#endif
template <int NW> struct LgGT {
    static constexpr int NT = 32 * NW;
    static constexpr int NP = (288 + NT - 1) / NT;
};
struct LgGId { int lane, wid, g, ch; };

template <int NW> __device__ __forceinline__ void lg_g_bar(int ch)
{
    if constexpr (NW == 1) __syncwarp();
    else LG_NAMED_BARRIER(1 + ch, 32 * NW);
}

/* every warp hands in one record (already reduced over its lanes), every thread gets all NW records back */
template <int NW, int N>
__device__ __forceinline__ void lg_g_exchange(LgGChan<NW> *cs, int &ring, const LgGId &id, const int (&v)[N], int (&o)[NW][N])
{
    static_assert(N <= 8, "record too large");
    if constexpr (NW == 1) {
        for (int n = 0; n < N; n++) o[0][n] = v[n];
        __syncwarp();
    }
    else {
        int (*slot)[8] = cs->red[ring];
        if (id.lane == 0) for (int n = 0; n < N; n++) slot[id.wid][n] = v[n];
        lg_g_bar<NW>(id.ch);
        for (int w = 0; w < NW; w++) for (int n = 0; n < N; n++) o[w][n] = slot[w][n];
        ring ^= 1;          /* the other half is rewritten only behind the next barrier, which every reader of this one has reached */
    }
}

/* ---------------------------------------------------------------- count_bits (takehiro.c:767) for a group */
template <int NW, class XP, class IV, class SF>
__device__ __forceinline__ int lg_g_count_bits(const LgDevCfg *__restrict__ c, LgSmemG<NW> *sm, LgGChan<NW> *cs, LgGBand *w, LgQInfo &gi, const LgQConst &qc,
                                               LgPrev &pv, const XP &xp, IV &iv, const SF &sfbp,
                                               float xm_warp, int &ring, const LgGId &id)
{
    constexpr int NT = LgGT<NW>::NT, NP = LgGT<NW>::NP;
    int const lane = id.lane;
    float const istep = __ldg(&c->ipow20[gi.global_gain]);
    float const lim = (LG_IXMAX) / istep;
    int const nsfb = (qc.block_type == LG_SHORT) ? 39 : 22;
    int const mnz = qc.max_nonzero_coeff;
    int const prev_data_use = pv.valid && (gi.global_gain == pv.global_gain);
    int const pv_count1 = pv.sfb_count1;
    int const plim = 32 * qc.jn;                   /* pairs the quantiser may touch (lg_k_quant.cuh: j < qc.jn) */
    int const ilim = (mnz + 2) & ~1;
    int hi_nz = -1, hi_big = -1;
    /* A warp whose own lines already overflow skips its share (the table index would leave adj43asm): the group's maximum is then
     * over the limit as well and the result is LARGE_BITS, after which every band is quantised afresh (the gain has changed). */
    if (!(xm_warp > lim)) {
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
LG_G_UNROLL_HOT
        for (int j = 0; j < NP; j++) {
            int const P = id.g + NT * j, i = 2 * P;
            if (P < plim) {
                int const sfb = sfbp[j];
                int a0, a1;
                a0 = a1 = w->act[sfb];
                if (T < 64) {
                    if (i > mnz) a0 = a1 = 3;
                    else {
                        if (sfb == T) a0 = 1;
                        a1 = (i + 1 == mnz) ? ((sfb == T) ? 1 : 3) : a0;
                    }
                }
                unsigned const old = iv[j];
                int v0 = (int) (old & 0xffffu), v1 = (int) (old >> 16);
                if (a0 == 1) {
                    float const xs = istep * xp[j].x;
                    double d = (double) xs + 8388608.0;
                    int const idx = __float_as_int((float) d) - 0x4b000000;
                    v0 = __float_as_int((float) (d + (double) __ldg(&adj[idx]))) - 0x4b000000;
                }
                else if (a0 == 2) v0 = (compareval0 > xp[j].x) ? 0 : 1;
                else if (a0 == 3) v0 = 0;
                if (a1 == 1) {
                    float const xs = istep * xp[j].y;
                    double d = (double) xs + 8388608.0;
                    int const idx = __float_as_int((float) d) - 0x4b000000;
                    v1 = __float_as_int((float) (d + (double) __ldg(&adj[idx]))) - 0x4b000000;
                }
                else if (a1 == 2) v1 = (compareval0 > xp[j].y) ? 0 : 1;
                else if (a1 == 3) v1 = 0;
                unsigned const nv = (unsigned) v0 | ((unsigned) v1 << 16);
                iv[j] = nv;
#if !LG_G_LINES_SMEM
                reinterpret_cast<unsigned *>(cs->ixw)[P] = nv;
#endif
                if (i < ilim) {
                    if (nv != 0u) hi_nz = P;
                    if ((nv & 0xfffefffeu) != 0u) hi_big = P;
                }
            }
        }
    }
    hi_nz = lg_wmax_i(hi_nz);
    hi_big = lg_wmax_i(hi_big);
    {   /* exchange 1: the group's xrpow_max, highest non-zero pair, highest pair with a value above 1 */
        int const v[3] = { __float_as_int(xm_warp), hi_nz, hi_big };
        int o[NW][3];
        lg_g_exchange<NW, 3>(cs, ring, id, v, o);
        int xm = o[0][0];
        hi_nz = o[0][1]; hi_big = o[0][2];
        for (int k = 1; k < NW; k++) { xm = max(xm, o[k][0]); hi_nz = max(hi_nz, o[k][1]); hi_big = max(hi_big, o[k][2]); }
        gi.xrpow_max = __int_as_float(xm);             /* non-negative floats order like their bit patterns */
    }
    if (gi.xrpow_max > lim) return LG_LARGE_BITS;
    /* ---- noquant_count_bits (takehiro.c:654) */
    pv.sfb_count1 = 0;
    int const c1p = hi_nz + 1;
    gi.count1 = 2 * c1p;
    int const nquads = (c1p - 1 - hi_big) >> 1;
    int const bigv = gi.count1 - 4 * nquads;
    gi.big_values = bigv;
    int a1 = 0, a2 = 0, has2 = 0;
    if (bigv > 0) {
        if (qc.block_type == LG_SHORT) {
            a1 = 3 * c->sfb_s[3];
            if (a1 > bigv) a1 = bigv;
            a2 = bigv;
        }
        else if (qc.block_type == LG_NORM) {
            a1 = gi.region0_count = sm->bv_scf[bigv - 2];
            a2 = gi.region1_count = sm->bv_scf[bigv - 1];
            a2 = sm->sfb_l[a1 + a2 + 2];
            a1 = sm->sfb_l[a1 + 1];
            has2 = a2 < bigv;
        }
        else {
            gi.region0_count = 7;
            gi.region1_count = LG_SBMAX_L - 1 - 7 - 1;
            a1 = sm->sfb_l[7 + 1];
            a2 = bigv;
            if (a1 > a2) a1 = a2;
        }
        a1 = a1 < bigv ? a1 : bigv;
        a2 = a2 < bigv ? a2 : bigv;
    }
    unsigned s1 = 0, s2 = 0;
    {
        const uint8_t *t32l = lg_hlen(c, 32), *t33l = lg_hlen(c, 33);
        const int16_t *ix = cs->ixw;
        for (int k = id.g; k < nquads; k += NT) {
            int const i = gi.count1 - 4 * k;
            int const p = ((ix[i - 4] * 2 + ix[i - 3]) * 2 + ix[i - 2]) * 2 + ix[i - 1];
            s1 += __ldg(&t32l[p]); s2 += __ldg(&t33l[p]);
        }
    }
    int m0 = 0, m1 = 0, m2 = 0;
LG_G_UNROLL_HOT
    for (int j = 0; j < NP; j++) {
        int const i = 2 * (id.g + NT * j);
        if (i < bigv) {
            unsigned const u = iv[j];
            int const v = max((int) (u & 0xffffu), (int) (u >> 16));
            if (i < a1) m0 = max(m0, v);
            else if (i < a2) m1 = max(m1, v);
            else m2 = max(m2, v);
        }
    }
    int bits;
    {   /* exchange 2: count1 bits under both books, the regions' largest magnitudes */
        int const v[4] = { (int) lg_wsum_u(s1 | (s2 << 16)), lg_wmax_i(m0), lg_wmax_i(m1), lg_wmax_i(m2) };
        int o[NW][4];
        lg_g_exchange<NW, 4>(cs, ring, id, v, o);
        unsigned cb = (unsigned) o[0][0];
        m0 = o[0][1]; m1 = o[0][2]; m2 = o[0][3];
        for (int k = 1; k < NW; k++) { cb += (unsigned) o[k][0]; m0 = max(m0, o[k][1]); m1 = max(m1, o[k][2]); m2 = max(m2, o[k][3]); }
        bits = (int) (cb & 0xffffu);
        gi.count1table_select = 0;
        if (bits > (int) (cb >> 16)) { bits = (int) (cb >> 16); gi.count1table_select = 1; }
        gi.count1bits = bits;
    }
    if (bigv == 0) return bits;
    if (max(m0, max(m1, m2)) > LG_IXMAX) return LG_LARGE_BITS;
    LgRegion R;
    lg_region_class(c, lane == 0 ? m0 : (lane == 1 ? m1 : m2), R);
    int const b0 = __shfl_sync(LG_FULL, R.base, 0), b1 = __shfl_sync(LG_FULL, R.base, 1), b2 = __shfl_sync(LG_FULL, R.base, 2);
    unsigned acc0 = 0, acc1 = 0, acc2 = 0, n15 = 0;
    const uint32_t *pk = c->huff_pk;
LG_G_UNROLL_HOT
    for (int j = 0; j < NP; j++) {
        int const i = 2 * (id.g + NT * j);
        if (i < bigv) {
            unsigned const u = iv[j];
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
    unsigned r0, r1, r2, r3, r4, r5;
    {   /* exchange 3: the regions' bit sums under their candidate tables */
        int const v[7] = { (int) lg_wsum_u((acc0 & 0x3ffu) | (((acc0 >> 10) & 0x3ffu) << 16)), (int) lg_wsum_u(acc0 >> 20),
                           (int) lg_wsum_u((acc1 & 0x3ffu) | (((acc1 >> 10) & 0x3ffu) << 16)), (int) lg_wsum_u(acc1 >> 20),
                           (int) lg_wsum_u((acc2 & 0x3ffu) | (((acc2 >> 10) & 0x3ffu) << 16)), (int) lg_wsum_u(acc2 >> 20), (int) lg_wsum_u(n15) };
        int o[NW][7];
        lg_g_exchange<NW, 7>(cs, ring, id, v, o);
        r0 = (unsigned) o[0][0]; r1 = (unsigned) o[0][1]; r2 = (unsigned) o[0][2]; r3 = (unsigned) o[0][3]; r4 = (unsigned) o[0][4]; r5 = (unsigned) o[0][5];
        n15 = (unsigned) o[0][6];
        for (int k = 1; k < NW; k++) {
            r0 += (unsigned) o[k][0]; r1 += (unsigned) o[k][1]; r2 += (unsigned) o[k][2]; r3 += (unsigned) o[k][3]; r4 += (unsigned) o[k][4]; r5 += (unsigned) o[k][5];
            n15 += (unsigned) o[k][6];
        }
    }
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
        int const below = (lane < 23) && (sm->sfb_l[lane] < bigv);
        pv.sfb_count1 = __popc(__ballot_sync(LG_FULL, below));
    }
    return bits;
}

/* ---------------------------------------------------------------- calc_noise (quantize_pvt.c:815) for a group */
template <int NW, class IV, class SF>
__device__ __forceinline__ void lg_g_calc_noise(const LgDevCfg *__restrict__ c, LgGChan<NW> *cs, LgGBand *w, const LgQInfo &gi, const LgQConst &qc,
                                                LgNoiseRes *res, LgPrev &pv, const IV &iv, const SF &sfbp, const LgGId &id)
{
    constexpr int NT = LgGT<NW>::NT, NP = LgGT<NW>::NP;
    int const lane = id.lane;
    float *bstep = reinterpret_cast<float *>(w->act);
    int *breg = w->act + 40;
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
    need = __any_sync(LG_FULL, need);              /* the same in every warp of the group: the band state is replicated */
    __syncwarp();
    if (need) {
        int const plim = 32 * qc.jn;
LG_G_UNROLL_HOT
        for (int j = 0; j < NP; j++) {
            int const P = id.g + NT * j, i = 2 * P;
            if (P < plim) {
                int const sfb = sfbp[j];
                float const step = bstep[sfb];
                if (step >= 0.f) {
                    float2 const x = *reinterpret_cast<const float2 *>(&cs->xr[i]);
                    unsigned const u = iv[j];
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
                    { float2 q2; q2.x = t0 * t0; q2.y = t1 * t1; *reinterpret_cast<float2 *>(&cs->sq[i]) = q2; }
                }
            }
        }
        lg_g_bar<NW>(id.ch);
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
                const float2 *q = reinterpret_cast<const float2 *>(&cs->sq[j]);
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

/* multiply the own lines of the flagged bands (factor per band in act[] as float, 0 = untouched); returns the warp's new maximum */
template <int NW, class XP, class SF>
__device__ __forceinline__ float lg_g_scale_bands(LgGBand *w, float xm_warp, int jn, XP &xp, const SF &sfbp, const LgGId &id)
{
    constexpr int NT = LgGT<NW>::NT, NP = LgGT<NW>::NP;
    const float *fac = reinterpret_cast<const float *>(w->act);
    float mx = xm_warp;
    for (int r = 0; r < 2; r++) {
        int const sfb = id.lane + 32 * r;
        if (sfb < 40) {
            float const f = fac[sfb];
            if (f != 0.f) {
                float const tm = w->tail_max[sfb] * f;
                w->tail_max[sfb] = tm;
                if (tm > mx) mx = tm;
            }
        }
    }
    int const plim = 32 * jn;
LG_G_UNROLL_HOT
    for (int j = 0; j < NP; j++) {
        if (id.g + NT * j < plim) {
            float const f = fac[sfbp[j]];
            if (f != 0.f) {
                xp[j].x *= f; xp[j].y *= f;
                if (xp[j].x > mx) mx = xp[j].x;
                if (xp[j].y > mx) mx = xp[j].y;
            }
        }
    }
    mx = lg_wmax_fpos(mx);
    __syncwarp();
    return mx;
}

/* quantize.c:720 amp_scalefac_bands, noise_shaping_amp 0 and 1 (2 belongs to quality 0/1: one-warp kernel) */
template <int NW, class XP, class SF>
__device__ __forceinline__ float lg_g_amp_scalefac_bands(const LgDevCfg *__restrict__ c, LgGBand *w, const LgQInfo &gi, const LgQConst &qc, float xm_warp,
                                                         XP &xp, const SF &sfbp, const LgGId &id)
{
    int const lane = id.lane;
    float const ifqstep34 = (gi.scalefac_scale == 0) ? (float) 1.29683955465100964055 : (float) 1.68179283050742922612;
    float trigger = 0;
    for (int sfb = lane; sfb < qc.sfbmax; sfb += 32) if (trigger < w->distort[sfb]) trigger = w->distort[sfb];
    trigger = lg_wmax_fpos(trigger);
    if (c->noise_shaping_amp == 1) {
        if (trigger > 1.0) trigger = (float) sqrt((double) trigger);
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
    }
    __syncwarp();
    return lg_g_scale_bands<NW>(w, xm_warp, qc.jn, xp, sfbp, id);
}

/* quantize.c:808 inc_scalefac_scale */
template <int NW, class XP, class SF>
__device__ __forceinline__ float lg_g_inc_scalefac_scale(LgGBand *w, LgQInfo &gi, const LgQConst &qc, float xm_warp, XP &xp,
                                                         const SF &sfbp, const LgGId &id)
{
    float const ifqstep34 = (float) 1.29683955465100964055;
    for (int r = 0; r < 2; r++) {
        int const sfb = id.lane + 32 * r;
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
    xm_warp = lg_g_scale_bands<NW>(w, xm_warp, qc.jn, xp, sfbp, id);
    gi.preflag = 0;
    gi.scalefac_scale = 1;
    return xm_warp;
}

/* quantize.c:847 inc_subblock_gain; returns 1 when a window's gain is exhausted */
template <int NW, class XP, class SF>
__device__ __forceinline__ int lg_g_inc_subblock_gain(const LgDevCfg *__restrict__ c, LgGBand *w, LgQInfo &gi, const LgQConst &qc, float &xm_warp,
                                                      XP &xp, const SF &sfbp, const LgGId &id)
{
    int const lane = id.lane;
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
            if (lane == 0) fac[qc.sfbmax + window] = __ldg(&c->ipow20[202]);
        }
        __syncwarp();
        xm_warp = lg_g_scale_bands<NW>(w, xm_warp, qc.jn, xp, sfbp, id);
    }
    return 0;
}

/* quantize.c:940 balance_noise */
template <int NW, class XP, class SF>
__device__ __forceinline__ int lg_g_balance_noise(const LgDevCfg *__restrict__ c, LgGBand *w, LgQInfo &gi, const LgQConst &qc, float &xm_warp,
                                                  XP &xp, const SF &sfbp, const LgGId &id)
{
    int const lane = id.lane;
    xm_warp = lg_g_amp_scalefac_bands<NW>(c, w, gi, qc, xm_warp, xp, sfbp, id);
    if (lg_loop_break(w, gi.sbg, qc.sfbmax, lane)) return 0;
    for (int pass = 0;; pass++) {
        unsigned const r = lg_scale_bitcount(w, qc.block_type, qc.sfbmax, qc.sfbdivide, gi.preflag, gi.scalefac_compress, LG_LSF_ARG(c, gi), lane);
        LG_APPLY_SCALE_BITCOUNT(gi, r);
        int status = (int) (r >> 30) & 1;
        if (!status) return 1;
        if (pass == 1) return 0;
        if (c->noise_shaping > 1) {
            if (!gi.scalefac_scale) { xm_warp = lg_g_inc_scalefac_scale<NW>(w, gi, qc, xm_warp, xp, sfbp, id); status = 0; }
            else if (qc.block_type == LG_SHORT && c->subblock_gain > 0)
                status = lg_g_inc_subblock_gain<NW>(c, w, gi, qc, xm_warp, xp, sfbp, id) || lg_loop_break(w, gi.sbg, qc.sfbmax, lane);
        }
        if (status) return 0;
    }
}

/* ---------------------------------------------------------------- outer_loop (quantize.c:1010) + bin_search_StepSize (:367): the state
 * machine of lg_outer_loop, every thread of the group walking through it with the same decisions.  On return gi, ib (the own pairs'
 * quantised values) and the warp's sfbst hold the chosen quantisation. */
template <int NW, class XP, class IB, class SF>
__device__ __forceinline__ void lg_g_outer_loop(const LgDevCfg *__restrict__ c, LgSmemG<NW> *sm, LgGChan<NW> *cs, LgGBand *w, LgQInfo &gi, const LgQConst &qc,
                                                int targ_bits, volatile int *old_value, volatile int *current_step, XP &xp, IB &ib,
                                                const SF &sfbp, float xm_warp, int &ring, const LgGId &id)
{
    constexpr int NP = LgGT<NW>::NP;
    int const lane = id.lane;
#if LG_G_LINES_SMEM
    LgGOwn<unsigned, LgGT<NW>::NT> iv = { reinterpret_cast<unsigned *>(cs->ixw), id.g };       /* zeroed by the caller */
#else
    unsigned iv[NP];
LG_G_UNROLL
    for (int j = 0; j < NP; j++) iv[j] = 0u;
#endif
    int CurrentStep = *current_step, flag_GoneOver = 0, Direction = 0;
    int const start = *old_value;
    gi.global_gain = start;
    int const desired_rate = targ_bits - gi.part2_length;
    LgPrev pv; pv.valid = 0; pv.global_gain = 0; pv.sfb_count1 = 0;
    LgNoiseRes best_noise; best_noise.max_noise = 0.f; best_noise.over_count = 0; best_noise.over_SSD = 0; best_noise.bits = 0;
    if (lane == 0) w->best = gi;
    int best_p23 = gi.part2_3_length;          /* best.part2_3_length */
    int age = 0, best_part2_3_length = 9999999, maxggain = 255, huff_bits = 0, phase = 0;
    for (int guard = 0;; guard++) {
        if (guard > 30000) lg_runaway();
        int const nBits = lg_g_count_bits<NW>(c, sm, cs, w, gi, qc, pv, xp, iv, sfbp, xm_warp, ring, id);
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
            if (id.g == 0) {                           /* read again only by the next granule, many barriers from here */
                *current_step = (start - gi.global_gain >= 4) ? 4 : 2;
                *old_value = gi.global_gain;
            }
            gi.part2_3_length = nBits;
            if (!c->noise_shaping) {                   /* quality 7-9: the step-size search is the whole loop */
                if (lane == 0) w->best = gi;
LG_G_UNROLL
                for (int j = 0; j < NP; j++) if (id.g + LgGT<NW>::NT * j < 288) ib[j] = iv[j];
                for (int i = lane; i < 40; i += 32) w->sfbst[i] = w->sfw[i];
                __syncwarp();
                break;
            }
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
        lg_g_calc_noise<NW>(c, cs, w, gi, qc, &noise_info, pv, iv, sfbp, id);
        noise_info.bits = gi.part2_3_length;
        int take = 0;
        if (phase == 1) take = 1;
        else {
            if (lg_quant_compare(best_noise, noise_info)) {
                best_part2_3_length = best_p23;
                take = 1;
                age = 0;
            }
            else if (c->full_outer_loop == 0) {
                if (++age > 3 && best_noise.over_count == 0) break;
            }
        }
        if (take) {
            best_noise = noise_info;
            if (lane == 0) w->best = gi;
            best_p23 = gi.part2_3_length;
LG_G_UNROLL
            for (int j = 0; j < NP; j++) if (id.g + LgGT<NW>::NT * j < 288) ib[j] = iv[j];
            for (int i = lane; i < 40; i += 32) w->sfbst[i] = w->sfw[i];
            __syncwarp();
        }
        if (phase != 1 && !((gi.global_gain + gi.scalefac_scale) < 255)) break;
        if (c->sfb21_extra) {
            if (w->distort[qc.sfbmax] > 1.0) break;
            if (qc.block_type == LG_SHORT && (w->distort[qc.sfbmax + 1] > 1.0 || w->distort[qc.sfbmax + 2] > 1.0)) break;
        }
        if (lg_g_balance_noise<NW>(c, w, gi, qc, xm_warp, xp, sfbp, id) == 0) break;
        maxggain = gi.scalefac_scale ? 254 : 255;
        huff_bits = targ_bits - gi.part2_length;
        if (huff_bits <= 0) break;
        phase = 2;
    }
    __syncwarp();
    gi = w->best;
}

/* ---------------------------------------------------------------- calc_xmin (quantize_pvt.c:589) for a group: the band sums replicated per
 * warp (one lane per band, the reference's order), the highest non-zero line from the own pairs.  xrpow sits in cs->sq while this runs. */
template <int NW, class XR>
__device__ __forceinline__ void lg_g_calc_xmin(const LgDevCfg *__restrict__ c, LgGChan<NW> *cs, LgGBand *w, LgQConst &qc, const LgXmin *en, const LgXmin *thm,
                                               float ath_adjust_factor, const XR &xr2, int &ring, const LgGId &id)
{
    constexpr int NT = LgGT<NW>::NT, NP = LgGT<NW>::NP;
    int const lane = id.lane;
    float const eps = (float) 2.2204460492503131e-016;
    int over = 0;
    if (lane < qc.psy_lmax) {
        int const gsfb = lane;
        float xmin = lg_ath_adjust(c, ath_adjust_factor, c->ath_l[gsfb], c->ath_floor, c->athfixpoint);
        xmin *= c->longfact[gsfb];
        int const width = w->width[gsfb];
        int j = w->lstart[gsfb];
        float const rh1 = xmin / width;
        float rh2 = eps, en0 = 0.0f, rh3;
        for (int l = 0; l < width; ++l) {
            float const xa = cs->xr[j++];
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
    }
    int k = 0;
LG_G_UNROLL
    for (int j = 0; j < NP; j++) {
        int const P = id.g + NT * j, i = 2 * P;
        if (P < 288) {
            if (fabsf(xr2[j].y) > 1e-12f) k = max(k, i + 1);
            else if (i > 0 && fabsf(xr2[j].x) > 1e-12f) k = max(k, i);
        }
    }
    k = lg_wmax_i(k);
    int max_nonzero;
    {
        int const v[1] = { k };
        int o[NW][1];
        lg_g_exchange<NW, 1>(cs, ring, id, v, o);
        max_nonzero = o[0][0];
        for (int q = 1; q < NW; q++) max_nonzero = max(max_nonzero, o[q][0]);
    }
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
        for (int r = 0; r < 2; r++) {
            int const sfb = lane + 32 * r;
            if (sfb < 40) {
                float tm = 0.f;
                int const j1 = w->lstart[sfb] + w->width[sfb];
                for (int j = max(w->lstart[sfb], ilim); j < j1; j++) { float const v = LG_G_XRPOW0(cs)[j]; if (v > tm) tm = v; }
                w->tail_max[sfb] = tm;
            }
        }
    }
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
                    float const xa = cs->xr[j++];
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

/* ---------------------------------------------------------------- the kernel: one CTA per stream, 2 x NW warps.
 * The bit budget (reservoir.c, on_pe, reduce_side, the ABR targets, ResvFrameEnd and the main_data_begin recurrence) is scalar integer
 * code along the stream: thread 0 runs it between the two CTA barriers of a granule and leaves the targets in shared memory. */
struct LgGBudget {                /* thread 0's registers */
    int resv_size, main_data_begin, anc_flag, pay_off, frame_used;
    int padding, mode_ext, bitrate_index, mean_bits, resv_max;
    int targ_abr[2][2];
};
__device__ __noinline__ void lg_g_frame_begin(const LgDevCfg *__restrict__ cfg, const LgPsyOut *__restrict__ P0, const LgFrameCtl *__restrict__ F, LgGBudget &b, LgGFrame *fr)
{
    int const mgr = cfg->mode_gr;
    b.padding = F->padding; b.mode_ext = F->mode_ext; b.bitrate_index = cfg->bitrate_index; b.frame_used = 0;
    fr->analog_silence_bits = 0;
    if (cfg->vbr == 3) {
        float const pe4[2][2] = { { F->pe_use[0][0], F->pe_use[0][1] }, { F->pe_use[1][0], F->pe_use[1][1] } };
        int const bt4[2][2] = { { P0[0].block_type[0], P0[0].block_type[1] }, { P0[mgr - 1].block_type[0], P0[mgr - 1].block_type[1] } };
        float const mer[2] = { F->ms_ener_ratio[0], F->ms_ener_ratio[1] };
        int asb = 0;
        lg_calc_target_bits(cfg, b.resv_size, b.padding, pe4, bt4, mer, b.mode_ext, b.targ_abr, &asb);
        fr->analog_silence_bits = asb;
        b.mean_bits = b.resv_max = 0;
    }
    else (void) lg_resv_frame_begin(cfg, b.bitrate_index, b.padding, b.resv_size, &b.mean_bits, &b.resv_max);
}
__device__ __noinline__ void lg_g_granule_targets(const LgDevCfg *__restrict__ cfg, const LgFrameCtl *__restrict__ F, int gr, const LgGBudget &b, LgGFrame *fr)
{
    int targ_bits[2];
    if (cfg->vbr == 3) { targ_bits[0] = b.targ_abr[gr][0]; targ_bits[1] = b.targ_abr[gr][1]; }
    else {
        float pe[2] = { F->pe_use[gr][0], F->pe_use[gr][1] };
        int const max_bits = lg_on_pe(cfg, b.resv_size, b.resv_max, pe, targ_bits, b.mean_bits, gr);
        if (b.mode_ext == 2) lg_reduce_side(targ_bits, F->ms_ener_ratio[gr], b.mean_bits, max_bits);
    }
    fr->targ_bits[0] = targ_bits[0]; fr->targ_bits[1] = targ_bits[1];
}
/* reservoir.c:239 ResvFrameEnd + the main_data_begin recurrence of format_bitstream (bitstream.c:937) */
__device__ __noinline__ void lg_g_frame_end(const LgDevCfg *__restrict__ cfg, LgGBudget &b, LgFrameOut *fo)
{
    int bitrate_index = b.bitrate_index, mean_bits = b.mean_bits, resv_max = b.resv_max, resv_size = b.resv_size;
    if (cfg->vbr == 3) {
        for (bitrate_index = cfg->vbr_min_bitrate_index; bitrate_index <= cfg->vbr_max_bitrate_index; bitrate_index++)
            if (lg_resv_frame_begin(cfg, bitrate_index, b.padding, resv_size, &mean_bits, &resv_max) >= 0) break;
        if (bitrate_index > cfg->vbr_max_bitrate_index) lg_runaway();
    }
    int stuffingBits = 0, over_bits, drain_pre = 0, drain_post = 0;
    resv_size += mean_bits * cfg->mode_gr;
    if ((over_bits = resv_size % 8) != 0) stuffingBits += over_bits;
    over_bits = (resv_size - stuffingBits) - resv_max;
    if (over_bits > 0) stuffingBits += over_bits;
    int const mdb_bytes = (b.main_data_begin * 8 < stuffingBits ? b.main_data_begin * 8 : stuffingBits) / 8;
    drain_pre += 8 * mdb_bytes;
    stuffingBits -= 8 * mdb_bytes;
    resv_size -= 8 * mdb_bytes;
    int const mdb_header = b.main_data_begin - mdb_bytes;
    drain_post += stuffingBits;
    resv_size -= stuffingBits;
    b.main_data_begin = resv_size / 8;
    int const pay_bits = drain_pre + b.frame_used + drain_post;
    if (pay_bits & 7) lg_runaway();
    int const anc_pre = b.anc_flag;
    if (!cfg->disable_reservoir) b.anc_flag ^= (lg_drain_tail_bits(drain_pre) + lg_drain_tail_bits(drain_post)) & 1;
    fo->main_data_begin = mdb_header; fo->drain_pre = drain_pre; fo->drain_post = drain_post;
    fo->padding = b.padding; fo->mode_ext = b.mode_ext; fo->resv_size = resv_size;
    fo->pay_off = b.pay_off; fo->pay_bytes = pay_bits >> 3;
    fo->anc_pre = (uint8_t) anc_pre; fo->anc_post = (uint8_t) b.anc_flag; fo->pad_[0] = fo->pad_[1] = 0; fo->bitrate_index = bitrate_index;
    b.pay_off += pay_bits >> 3;
    b.resv_size = resv_size;
}

/* Four CTAs per SM (512 streams on 148 SMs: 3.46).  Asking for five (96 registers, so that the next step's analysis kernels find room on
 * every SM underneath this kernel) was measured and is worse: 8.03 against 7.68 ms per pipelined step - the overlap buys almost nothing
 * (serial sum of the kernels: 7.81 ms), the registers cost this kernel 5 %. */
#ifndef LG_G_MINBLOCKS
#define LG_G_MINBLOCKS(NW) 4
#endif
#ifdef LG_G_MAXNREG
#define LG_G_BOUNDS(NW) __maxnreg__(LG_G_MAXNREG)
#else
#define LG_G_BOUNDS(NW) __launch_bounds__(64 * NW, LG_G_MINBLOCKS(NW))
#endif
template <int NW>
__global__ void LG_G_BOUNDS(NW)
lg_kernel_quantg(const LgDevCfg *__restrict__ cfg, const float *__restrict__ xr_in, const LgPsyOut *__restrict__ psy, const LgFrameCtl *__restrict__ frm,
                 LgGranuleOut *__restrict__ gout, LgFrameOut *__restrict__ fout, LgStreamState *__restrict__ state, const int *__restrict__ nfr, int nframes)
{
    constexpr int NT = LgGT<NW>::NT, NP = LgGT<NW>::NP;
    LG_DYN_SMEM(LgSmemG<NW>, sm);
    LgGId id;
    {   /* cheap to recompute (the compiler rematerialises them all over the search loop to stay within its registers): no division by 96 */
        int const warp = (int) threadIdx.x >> 5;
        id.ch = warp >= NW ? 1 : 0;
        id.wid = warp - NW * id.ch;
        id.g = (int) threadIdx.x - NT * id.ch;
        id.lane = (int) threadIdx.x & 31;
    }
    int const lane = id.lane, ch = id.ch;
    int const stream = blockIdx.x;
    int const nch = cfg->channels;
    int const tid0 = threadIdx.x == 0;
    LgStreamState *st = state + stream;
    LgGChan<NW> *cs = &sm->c[ch];
    LgGBand *w = id.wid == 0 ? static_cast<LgGBand *>(cs) : &cs->rep[id.wid - 1];
    int ring = 0;
    int const my_frames = min(nfr[stream], nframes);
    if (my_frames <= 0) return;                        /* nothing of this stream in this step: its state is not ours to write back (another step's kernel may own it) */
    int const mgr = cfg->mode_gr;
    LgGBudget bud;
    for (int i = threadIdx.x; i < 576; i += 2 * NT) sm->bv_scf[i] = cfg->bv_scf[i];
    for (int i = threadIdx.x; i < 23; i += 2 * NT) sm->sfb_l[i] = cfg->sfb_l[i];
    if (tid0) {
        bud.resv_size = st->resv_size; bud.main_data_begin = st->main_data_begin; bud.anc_flag = st->ancillary_flag; bud.pay_off = 0;
        for (int k = 0; k < 2; k++) { sm->fr.old_value[k] = st->old_value[k]; sm->fr.current_step[k] = st->current_step[k]; }
        if (my_frames > 0) {
            lg_g_frame_begin(cfg, psy + (size_t) stream * 2 * nframes, frm + (size_t) stream * nframes, bud, &sm->fr);
            lg_g_granule_targets(cfg, frm + (size_t) stream * nframes, 0, bud, &sm->fr);
        }
    }
    __syncthreads();

    for (int gb = 0; gb < mgr * my_frames; gb++) {
        int const frame = mgr == 2 ? gb >> 1 : gb, gr = mgr == 2 ? gb & 1 : 0;
        const LgFrameCtl *F = frm + (size_t) stream * nframes + frame;
        const LgPsyOut *P = psy + (size_t) stream * 2 * nframes + gb;
        if (ch < nch) {
            int const mode_ext = F->mode_ext;
            LgQInfo gi;
            LgQConst qc;
            int const rch = (mode_ext == 2) ? ch + 2 : ch;
            const LgXmin *en = &P->en[rch], *thm = &P->thm[rch];
            gi.part2_3_length = 0; gi.big_values = 0; gi.count1 = 0; gi.global_gain = 210; gi.scalefac_compress = 0;
            gi.table_select[0] = gi.table_select[1] = gi.table_select[2] = 0;
            gi.sbg = 0;
            gi.region0_count = 0; gi.region1_count = 0; gi.preflag = 0; gi.scalefac_scale = 0;
            gi.count1table_select = 0; gi.part2_length = 0; gi.count1bits = 0; gi.xrpow_max = 0;
            qc.block_type = P->block_type[ch];
            qc.sfb_lmax = LG_SBPSY_L; qc.sfb_smin = LG_SBPSY_S;
            qc.psy_lmax = cfg->sfb21_extra ? LG_SBMAX_L : LG_SBPSY_L;
            if (cfg->samplerate <= 8000) { qc.sfb_lmax = 17; qc.sfb_smin = 9; qc.psy_lmax = 17; }
            qc.psymax = qc.psy_lmax; qc.sfbmax = qc.sfb_lmax; qc.sfbdivide = 11;
            if (qc.block_type == LG_SHORT) {
                qc.sfb_smin = 0; qc.sfb_lmax = 0;
                qc.psymax = 3 * (cfg->sfb21_extra ? LG_SBMAX_S : LG_SBPSY_S);
                qc.sfbmax = 3 * LG_SBPSY_S;
                if (cfg->samplerate <= 8000) qc.psymax = qc.sfbmax = 3 * 9;
                qc.sfbdivide = qc.sfbmax - 18;
                qc.psy_lmax = 0;
            }
            qc.max_nonzero_coeff = 575; qc.jn = 9; qc.ath_over = 0;
            for (int r = 0; r < 2; r++) {              /* every warp its own copy of the band geometry */
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
                    else if (k < LG_SBMAX_L) { ws = sm->sfb_l[k + 1] - sm->sfb_l[k]; ls = sm->sfb_l[k]; }
                    if (k < 40) { w->width[k] = ws; w->window[k] = (uint8_t) wn; w->sfw[k] = 0; w->sfbst[k] = 0; }
                    w->lstart[k] = ls;
                }
            }
            /* the own line pairs: xr, band index, xrpow (quantize.c:110 init_xrpow); xrpow also into cs->sq for calc_xmin's tail maxima */
#if LG_G_LINES_SMEM
            LgGOwn<float2, NT> xp = { reinterpret_cast<float2 *>(cs->xrpow), id.g };
            LgGOwn<unsigned, NT> ib = { reinterpret_cast<unsigned *>(cs->ixb), id.g };
            LgGOwnSfb<NT> sfbp = { cs->pair_sfb, id.g };
#else
            float2 xp[NP];
            unsigned ib[NP];
            int sfbp[NP];
#endif
            unsigned sgn = 0;                          /* sign bits of the own lines, for the hand-over to the packer */
            int targ = 0;
            {
#if LG_G_LINES_SMEM
                LgGOwn<float2, NT> xr2 = { reinterpret_cast<float2 *>(cs->xr), id.g };
#else
                float2 xr2[NP];
#endif
                float mx = 0.f, amax = 0.f;
                const float2 *src = reinterpret_cast<const float2 *>(xr_in + (((size_t) stream * 2 * nframes + gb) * 2 + ch) * 576);
                const uint8_t *map = qc.block_type == LG_SHORT ? cfg->line_sfb_s : cfg->line_sfb_l;
LG_G_UNROLL
                for (int j = 0; j < NP; j++) {
                    int const Pp = id.g + NT * j;
#if !LG_G_LINES_SMEM
                    xr2[j].x = xr2[j].y = 0.f; xp[j].x = xp[j].y = 0.f; ib[j] = 0u; sfbp[j] = 0;
#endif
                    if (Pp < 288) {
                        xr2[j] = __ldg(src + Pp);
#if LG_G_LINES_SMEM
                        ib[j] = 0u;
                        cs->pair_sfb[Pp] = __ldg(map + 2 * Pp);
#else
                        sfbp[j] = __ldg(map + 2 * Pp);
#endif
                        float const t0 = fabsf(xr2[j].x), t1 = fabsf(xr2[j].y);
                        xp[j].x = (float) sqrt((double) t0 * sqrt((double) t0));
                        xp[j].y = (float) sqrt((double) t1 * sqrt((double) t1));
                        if (xp[j].x > mx) mx = xp[j].x;
                        if (xp[j].y > mx) mx = xp[j].y;
                        if (t0 > amax) amax = t0;
                        if (t1 > amax) amax = t1;
                        if (xr2[j].x < 0.0f) sgn |= 1u << (2 * j);
                        if (xr2[j].y < 0.0f) sgn |= 2u << (2 * j);
#if !LG_G_LINES_SMEM
                        *reinterpret_cast<float2 *>(&cs->xr[2 * Pp]) = xr2[j];
                        *reinterpret_cast<float2 *>(&cs->sq[2 * Pp]) = xp[j];
#endif
                        reinterpret_cast<unsigned *>(cs->ixw)[Pp] = 0u;
                    }
                }
                float xm_warp = lg_wmax_fpos(mx);
                amax = lg_wmax_fpos(amax);
                {
                    int const v[2] = { __float_as_int(xm_warp), __float_as_int(amax) };
                    int o[NW][2];
                    lg_g_exchange<NW, 2>(cs, ring, id, v, o);
                    int a = o[0][0], b = o[0][1];
                    for (int q = 1; q < NW; q++) { a = max(a, o[q][0]); b = max(b, o[q][1]); }
                    gi.xrpow_max = __int_as_float(a);
                    amax = __int_as_float(b);
                }
                int nonzero = amax > (float) 1E-20;
                if (!nonzero && amax > 0.f) {
                    float sum = 0;
                    for (int i = 0; i < 576; ++i) sum += fabsf(cs->xr[i]);
                    nonzero = sum > (float) 1E-20;
                }
                if (nonzero) {
                    {
                        LgQConst qx = qc;
                        lg_g_calc_xmin<NW>(cfg, cs, w, qx, en, thm, F->ath_adjust_factor, xr2, ring, id);
                        qc.max_nonzero_coeff = qx.max_nonzero_coeff; qc.jn = qx.jn;
                        targ = sm->fr.targ_bits[ch];
                        if (cfg->vbr == 3 && qx.ath_over == 0) targ = sm->fr.analog_silence_bits;
                    }
                    lg_g_bar<NW>(ch);                      /* cs->sq is free again once every warp has its tail maxima */
                    lg_g_outer_loop<NW>(cfg, sm, cs, w, gi, qc, targ, &sm->fr.old_value[ch], &sm->fr.current_step[ch], xp, ib, sfbp, xm_warp, ring, id);
                }
            }
            /* the chosen quantisation into shared memory for iteration_finish_one (warp 0) and straight out to the packer */
            {
                LgGranuleOut *o = gout + (((size_t) stream * 2 * nframes + gb) * 2 + ch);
LG_G_UNROLL
                for (int j = 0; j < NP; j++) {
                    int const Pp = id.g + NT * j;
                    if (Pp < 288) {
                        reinterpret_cast<unsigned *>(cs->ixw)[Pp] = ib[j];
                        int v0 = (int) (ib[j] & 0xffffu), v1 = (int) (ib[j] >> 16);
                        if ((sgn >> (2 * j)) & 1u) v0 = -v0;
                        if ((sgn >> (2 * j)) & 2u) v1 = -v1;
                        *reinterpret_cast<unsigned *>(&o->ix[2 * Pp]) = ((unsigned) v0 & 0xffffu) | ((unsigned) v1 << 16);
                    }
                }
                lg_g_bar<NW>(ch);
                /* best_huffman_divide's 16 + 128 region candidates depend on ix only: all warps of the group share them out */
                int const cand = cfg->use_best_huffman == 1 && qc.block_type == LG_NORM;
                if (cand) {
                    lg_recalc_divide_cand(cfg, cs, gi.big_values, id.g, NT);
                    lg_g_bar<NW>(ch);
                }
                if (id.wid == 0) {
                    /* quantize.c:1213 iteration_finish_one on warp 0's replica (sfw = the chosen scalefactors) */
                    for (int i = lane; i < 40; i += 32) cs->sfw[i] = cs->sfbst[i];
                    __syncwarp();
                    uint8_t scfsi[4] = { 0, 0, 0, 0 };
                    LgQInfo gm = gi;
                    LgQConst qx = qc;
                    lg_best_scalefac_store(cfg, cs, gm, qx, gr, sm->sf_gr0[ch], sm->bt_gr0[ch], scfsi, lane);
                    if (cfg->use_best_huffman == 1) lg_best_huffman_divide(cfg, cs, gm, qx, lane, cand);
                    if (gr == 0) {
                        for (int i = lane; i < 40; i += 32) sm->sf_gr0[ch][i] = cs->sfw[i];
                        if (lane == 0) sm->bt_gr0[ch] = qc.block_type;
                    }
                    for (int i = lane; i < 40; i += 32) o->scalefac[i] = (int8_t) (i < 39 ? cs->sfw[i] : 0);
                    if (lane == 0) {
                        o->part2_3_length = (int16_t) gm.part2_3_length; o->part2_length = (int16_t) gm.part2_length;
                        o->big_values = (int16_t) gm.big_values; o->count1 = (int16_t) gm.count1;
                        o->global_gain = (uint8_t) gm.global_gain; o->scalefac_compress = (uint8_t) gm.scalefac_compress;
                        o->scalefac_compress_hi = (uint8_t) (gm.scalefac_compress >> 8);
                        o->block_type = (uint8_t) qc.block_type; o->mixed_block_flag = 0;
                        for (int i = 0; i < 3; i++) { o->table_select[i] = (uint8_t) gm.table_select[i]; o->subblock_gain[i] = (uint8_t) ((gm.sbg >> (4 * i)) & 15); }
                        o->region0_count = (uint8_t) gm.region0_count; o->region1_count = (uint8_t) gm.region1_count;
                        o->preflag = (uint8_t) gm.preflag; o->scalefac_scale = (uint8_t) gm.scalefac_scale;
                        o->count1table_select = (uint8_t) gm.count1table_select;
                        o->sfbmax = (uint8_t) qc.sfbmax; o->sfbdivide = (uint8_t) qc.sfbdivide;
                        sm->fr.used_bits[ch] = gm.part2_3_length + gm.part2_length;
                        /* scfsi belongs to the frame: all zero but for granule 1 of an MPEG-1 frame */
                        LgFrameOut *fo = fout + (size_t) stream * nframes + frame;
                        if (gr == mgr - 1) for (int i = 0; i < 4; i++) fo->scfsi[ch][i] = scfsi[i];
                    }
                }
            }
        }
        else if (id.g == 0) {
            sm->fr.used_bits[ch] = 0;
            if (gr == mgr - 1) { LgFrameOut *fo = fout + (size_t) stream * nframes + frame; for (int i = 0; i < 4; i++) fo->scfsi[ch][i] = 0; }
        }
        __syncthreads();
        if (tid0) {
            int const used = sm->fr.used_bits[0] + sm->fr.used_bits[1];
            bud.resv_size -= used;                               /* reservoir.c:226 ResvAdjust */
            bud.frame_used += used;
            if (gr == mgr - 1) {
                lg_g_frame_end(cfg, bud, fout + (size_t) stream * nframes + frame);
                if (frame + 1 < my_frames) {
                    lg_g_frame_begin(cfg, psy + (size_t) stream * 2 * nframes + mgr * (frame + 1), F + 1, bud, &sm->fr);
                    lg_g_granule_targets(cfg, F + 1, 0, bud, &sm->fr);
                }
            }
            else lg_g_granule_targets(cfg, F, gr + 1, bud, &sm->fr);
        }
        __syncthreads();
    }
    if (tid0) {
        st->resv_size = bud.resv_size; st->main_data_begin = bud.main_data_begin; st->ancillary_flag = bud.anc_flag;
        for (int k = 0; k < 2; k++) { st->old_value[k] = sm->fr.old_value[k]; st->current_step[k] = sm->fr.current_step[k]; }
    }
}
