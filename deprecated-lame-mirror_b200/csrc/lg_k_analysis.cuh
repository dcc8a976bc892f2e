// lg_k_analysis.cuh - kernel A: everything in the psycho-acoustic model and the filterbank that is a
// pure function of the PCM around one granule ("Psy stage A").
//
// One CTA (4 warps) per (stream, granule slot).  Slot 0 is the granule *before* the batch: only its
// subband samples are needed (they are the "previous granule" input of the first MDCT).
//
//   phase 1  all warps : int16 PCM -> float with the 2x2 pcm_transform (lame.c:1786), staged in shared
//                        memory in a 33-word-pitch layout so that the filterbank's stride-32 accesses
//                        (lane = time slot) are bank-conflict free
//   phase 2  warp 0,1  : polyphase filterbank of channel 0/1, lane = one of the 18 time slots, the
//                        32-point transform runs entirely in registers (newmdct.c:430); then the three
//                        256-point short FHTs (fft.c:194)
//            warp 2,3  : 1024-point windowed FHT of channel 0/1 (fft.c:246), butterflies spread over lanes
//   phase 3  all warps : 21-tap high-pass FIR (psymodel.c:778) and the 9 sub-block peaks per channel L R M S
//   phase 4  all warps : FFT line energies for L R M S (psymodel.c:681-688, :730-736)
//   phase 5  warp=chn  : partition energies / max / avg (serial per partition, as the reference sums),
//                        tonality index, spreading with the non-linear mask_add chain, loudness, tot_ener
//
// Bit-exactness: each output value is produced by the same sequence of IEEE operations as in the
// reference; only *which lane* produces it changes.  Sums the reference accumulates serially are
// accumulated serially by one lane.
//
// Algorithmic HBM bytes per granule: read 2 ch x 576 x 2 B new PCM (2304 B; the 752-sample look-ahead is
// shared with neighbouring CTAs through L2), write 4608 B subband samples + sizeof(LgAnalysis).
#pragma once
#include "lg_math.cuh"

#define LG_PADIDX(i) ((i) + ((i) >> 5))
#define LG_SPAN_PAD 1376

struct LgSmemA {
    float pcm[2][LG_SPAN_PAD];
    float wl[2][LG_BLK];
    float ws[2][3][LG_BLK_S];
    float hpf[2][576];
    float hpf_prev[2][384];          /* the same for the last six sub-blocks of the granule before (for the short-block decision) */
    int   may_attack[4];
    float fe[4][520];
    float fes[4][3][132];
    float eb[4][LG_CBANDS], mx[4][LG_CBANDS], av[4][LG_CBANDS];
    int   midx[4][LG_CBANDS];
};

/* ---------------------------------------------------------------- polyphase filterbank, one time slot */
#define BF_A(p, q, c) { float t_ = a[p] - a[q]; a[q] += a[p]; a[p] = (float) (t_ * (c)); }
#define BF_B(p, q, c) { float t_ = a[p] - a[q]; a[p] += a[q]; a[q] = (float) (t_ * (c)); }
#define XADD(p, q) { float t_ = a[p]; a[p] = a[q] - t_; a[q] = a[q] + t_; }
#define XSUB(p, q) { float t_ = a[p]; a[p] += a[q]; a[q] -= t_; }
#define CHN(k) { xr = a[k] - xr; a[k] = xr; }

/* x: padded span of one channel; W = unpadded index of the window centre (286 + 32*slot).
 * newmdct.c:430 window_subband */
__device__ __forceinline__ void lg_window_subband(const float *__restrict__ x, int W, const float *__restrict__ enw, float a[32])
{
#define X1(n, off) x[LG_PADIDX(W - (n) + (off))]
#define X2(n, off) x[LG_PADIDX(W - 62 + (n) + (off))]
#pragma unroll
    for (int n = 0; n < 15; n++) {
        const float *wp = enw + 10 + 18 * n;
        float w, s, t;
        w = __ldg(wp - 10); s = X2(n, -224) * w; t = X1(n, 224) * w;
        w = __ldg(wp - 9); s += X2(n, -160) * w; t += X1(n, 160) * w;
        w = __ldg(wp - 8); s += X2(n, -96) * w; t += X1(n, 96) * w;
        w = __ldg(wp - 7); s += X2(n, -32) * w; t += X1(n, 32) * w;
        w = __ldg(wp - 6); s += X2(n, 32) * w; t += X1(n, -32) * w;
        w = __ldg(wp - 5); s += X2(n, 96) * w; t += X1(n, -96) * w;
        w = __ldg(wp - 4); s += X2(n, 160) * w; t += X1(n, -160) * w;
        w = __ldg(wp - 3); s += X2(n, 224) * w; t += X1(n, -224) * w;
        w = __ldg(wp - 2); s += X1(n, -256) * w; t -= X2(n, 256) * w;
        w = __ldg(wp - 1); s += X1(n, -192) * w; t -= X2(n, 192) * w;
        w = __ldg(wp + 0); s += X1(n, -128) * w; t -= X2(n, 128) * w;
        w = __ldg(wp + 1); s += X1(n, -64) * w; t -= X2(n, 64) * w;
        w = __ldg(wp + 2); s += X1(n, 0) * w; t -= X2(n, 0) * w;
        w = __ldg(wp + 3); s += X1(n, 64) * w; t -= X2(n, -64) * w;
        w = __ldg(wp + 4); s += X1(n, 128) * w; t -= X2(n, -128) * w;
        w = __ldg(wp + 5); s += X1(n, 192) * w; t -= X2(n, -192) * w;
        s *= __ldg(wp + 6);
        w = t - s;
        a[2 * n] = t + s;
        a[2 * n + 1] = __ldg(wp + 7) * w;
    }
    const float *wp = enw + 10 + 18 * 15;
    {
        float s, t, u, v;
        t = X1(15, -16) * __ldg(wp - 10); s = X1(15, -32) * __ldg(wp - 2);
        t += (X1(15, -48) - X1(15, 16)) * __ldg(wp - 9); s += X1(15, -96) * __ldg(wp - 1);
        t += (X1(15, -80) + X1(15, 48)) * __ldg(wp - 8); s += X1(15, -160) * __ldg(wp + 0);
        t += (X1(15, -112) - X1(15, 80)) * __ldg(wp - 7); s += X1(15, -224) * __ldg(wp + 1);
        t += (X1(15, -144) + X1(15, 112)) * __ldg(wp - 6); s -= X1(15, 32) * __ldg(wp + 2);
        t += (X1(15, -176) - X1(15, 144)) * __ldg(wp - 5); s -= X1(15, 96) * __ldg(wp + 3);
        t += (X1(15, -208) + X1(15, 176)) * __ldg(wp - 4); s -= X1(15, 160) * __ldg(wp + 4);
        t += (X1(15, -240) - X1(15, 208)) * __ldg(wp - 3); s -= X1(15, 224);
        u = s - t; v = s + t;
        t = a[14]; s = a[15] - t;
        a[31] = v + t; a[30] = u + s; a[15] = u - s; a[14] = v - t;
    }
#undef X1
#undef X2
    {
        float const w2 = __ldg(wp - 2 * 18 + 7), w4 = __ldg(wp - 4 * 18 + 7), w6 = __ldg(wp - 6 * 18 + 7);
        float const w10 = __ldg(wp - 10 * 18 + 7), w12 = __ldg(wp - 12 * 18 + 7), w14 = __ldg(wp - 14 * 18 + 7);
        float xr;
        BF_A(28, 0, w2) BF_A(29, 1, w2) BF_A(26, 2, w4) BF_A(27, 3, w4) BF_A(24, 4, w6) BF_A(25, 5, w6)
        BF_A(22, 6, LG_SQRT2_D)
        xr = a[23] - a[7]; a[7] += a[23]; a[23] = (float) (xr * LG_SQRT2_D - a[7]);
        a[7] -= a[6]; a[22] -= a[7]; a[23] -= a[22];
        XADD(6, 31) XADD(7, 30) XADD(22, 15) XADD(23, 14)
        BF_A(20, 8, w10) BF_A(21, 9, w10) BF_A(18, 10, w12) BF_A(19, 11, w12) BF_A(16, 12, w14) BF_A(17, 13, w14)
        BF_A(24, 20, w12) BF_A(25, 21, w12) BF_B(4, 8, w12) BF_B(5, 9, w12)
        BF_B(0, 12, w4) BF_B(1, 13, w4) BF_B(16, 28, w4) BF_A(29, 17, w4)
        BF_B(2, 10, LG_SQRT2_D) BF_B(3, 11, LG_SQRT2_D)
        xr = (float) (LG_SQRT2_D * (-a[18] + a[26])); a[18] += a[26]; a[26] = xr - a[18];
        xr = (float) (LG_SQRT2_D * (-a[19] + a[27])); a[19] += a[27]; a[27] = xr - a[19];
        xr = a[2]; a[19] -= a[3]; a[3] -= xr; a[2] = a[31] - xr; a[31] += xr;
        xr = a[3]; a[11] -= a[19]; a[18] -= xr; a[3] = a[30] - xr; a[30] += xr;
        xr = a[18]; a[27] -= a[11]; a[19] -= xr; a[18] = a[15] - xr; a[15] += xr;
        xr = a[19]; a[10] -= xr; a[19] = a[14] - xr; a[14] += xr;
        xr = a[10]; a[11] -= xr; a[10] = a[23] - xr; a[23] += xr;
        xr = a[11]; a[26] -= xr; a[11] = a[22] - xr; a[22] += xr;
        xr = a[26]; a[27] -= xr; a[26] = a[7] - xr; a[7] += xr;
        xr = a[27]; a[27] = a[6] - xr; a[6] += xr;
        BF_B(0, 4, LG_SQRT2_D) BF_B(1, 5, LG_SQRT2_D) BF_B(16, 20, LG_SQRT2_D) BF_B(17, 21, LG_SQRT2_D)
        xr = (float) (-LG_SQRT2_D * (a[8] - a[12])); a[8] += a[12]; a[12] = xr - a[8];
        xr = (float) (-LG_SQRT2_D * (a[9] - a[13])); a[9] += a[13]; a[13] = xr - a[9];
        xr = (float) (-LG_SQRT2_D * (a[25] - a[29])); a[25] += a[29]; a[29] = xr - a[25];
        xr = (float) (-LG_SQRT2_D * (a[24] + a[28])); a[24] -= a[28]; a[28] = xr - a[24];
        xr = a[24] - a[16]; a[24] = xr; CHN(20) CHN(28)
        xr = a[25] - a[17]; a[25] = xr; CHN(21) CHN(29)
        xr = a[17] - a[1]; a[17] = xr; CHN(9) CHN(25) CHN(5) CHN(21) CHN(13) CHN(29)
        xr = a[1] - a[0]; a[1] = xr;
        CHN(16) CHN(17) CHN(8) CHN(9) CHN(24) CHN(25) CHN(4) CHN(5) CHN(20) CHN(21) CHN(12) CHN(13) CHN(28) CHN(29)
        XSUB(0, 31) XSUB(1, 30) XSUB(16, 15) XSUB(17, 14) XSUB(8, 23) XSUB(9, 22) XSUB(24, 7) XSUB(25, 6)
        XSUB(4, 27) XSUB(5, 26) XSUB(20, 11) XSUB(21, 10) XSUB(12, 19) XSUB(13, 18) XSUB(28, 3) XSUB(29, 2)
    }
}
#undef BF_A
#undef BF_B
#undef XADD
#undef XSUB
#undef CHN

/* ---------------------------------------------------------------- fast Hartley transform, one warp
 * fft.c:64 fht on n points held in shared memory; every stage is n/8 independent butterfly groups. */
__device__ __forceinline__ void lg_fht_warp(float *__restrict__ fz, int n, const LgDevCfg *__restrict__ c, int lane)
{
    int stage = 0;
    for (int kx = 2; kx * 8 <= n; kx <<= 2, stage++) {
        int const k1 = kx << 1, k2 = kx << 2, k3 = k2 + k1, k4 = kx << 3;
        for (int q = lane; q < (n >> 3); q += 32) {
            int const i = q % kx, base = (q / kx) * k4;
            if (i == 0) {
                float *fi = fz + base, *gi = fi + kx;
                float f0, f1, f2, f3;
                f1 = fi[0] - fi[k1]; f0 = fi[0] + fi[k1];
                f3 = fi[k2] - fi[k3]; f2 = fi[k2] + fi[k3];
                fi[k2] = f0 - f2; fi[0] = f0 + f2; fi[k3] = f1 - f3; fi[k1] = f1 + f3;
                f1 = gi[0] - gi[k1]; f0 = gi[0] + gi[k1];
                f3 = (float) (LG_SQRT2_D * gi[k3]); f2 = (float) (LG_SQRT2_D * gi[k2]);
                gi[k2] = f0 - f2; gi[0] = f0 + f2; gi[k3] = f1 - f3; gi[k1] = f1 + f3;
            }
            else {
                float const c1 = __ldg(&c->fht_tw[stage][i][0]), s1 = __ldg(&c->fht_tw[stage][i][1]);
                float const c2 = __ldg(&c->fht_tw[stage][i][2]), s2 = __ldg(&c->fht_tw[stage][i][3]);
                float *fi = fz + base + i, *gi = fz + base + k1 - i;
                float a, b, g0, f0, f1, g1, f2, g2, f3, g3;
                b = s2 * fi[k1] - c2 * gi[k1]; a = c2 * fi[k1] + s2 * gi[k1];
                f1 = fi[0] - a; f0 = fi[0] + a; g1 = gi[0] - b; g0 = gi[0] + b;
                b = s2 * fi[k3] - c2 * gi[k3]; a = c2 * fi[k3] + s2 * gi[k3];
                f3 = fi[k2] - a; f2 = fi[k2] + a; g3 = gi[k2] - b; g2 = gi[k2] + b;
                b = s1 * f2 - c1 * g3; a = c1 * f2 + s1 * g3;
                fi[k2] = f0 - a; fi[0] = f0 + a; gi[k3] = g1 - b; gi[k1] = g1 + b;
                b = c1 * g2 - s1 * f3; a = s1 * g2 + c1 * f3;
                gi[k2] = g0 - a; gi[0] = g0 + a; fi[k3] = f1 - b; fi[k1] = f1 + b;
            }
        }
        __syncwarp();
    }
}

/* fft.c:246 fft_long: Blackman window + first radix-4 pass in bit-reversed order, then the FHT.
 * buf = padded span, B = unpadded index of bufp[0]. */
__device__ __forceinline__ void lg_fft_long_warp(float *__restrict__ x, const float *__restrict__ buf, int B,
                                                 const LgDevCfg *__restrict__ c, int lane)
{
    const float *w = c->window;
    for (int jj = lane; jj < LG_BLK / 8; jj += 32) {
        int const i = lg_bitrev8(jj);
        float f0, f1, f2, f3, v;
        float *o = x + 4 * jj;
#define S(k) buf[LG_PADIDX(B + (k))]
        f0 = __ldg(&w[i]) * S(i); v = __ldg(&w[i + 0x200]) * S(i + 0x200); f1 = f0 - v; f0 = f0 + v;
        f2 = __ldg(&w[i + 0x100]) * S(i + 0x100); v = __ldg(&w[i + 0x300]) * S(i + 0x300); f3 = f2 - v; f2 = f2 + v;
        o[0] = f0 + f2; o[2] = f0 - f2; o[1] = f1 + f3; o[3] = f1 - f3;
        f0 = __ldg(&w[i + 1]) * S(i + 1); v = __ldg(&w[i + 0x201]) * S(i + 0x201); f1 = f0 - v; f0 = f0 + v;
        f2 = __ldg(&w[i + 0x101]) * S(i + 0x101); v = __ldg(&w[i + 0x301]) * S(i + 0x301); f3 = f2 - v; f2 = f2 + v;
        o[LG_BLK / 2 + 0] = f0 + f2; o[LG_BLK / 2 + 2] = f0 - f2; o[LG_BLK / 2 + 1] = f1 + f3; o[LG_BLK / 2 + 3] = f1 - f3;
    }
    __syncwarp();
    lg_fht_warp(x, LG_BLK, c, lane);
}

/* fft.c:194 fft_short */
__device__ __forceinline__ void lg_fft_short_warp(float (*__restrict__ xs)[LG_BLK_S], const float *__restrict__ buf, int B,
                                                  const LgDevCfg *__restrict__ c, int lane)
{
    const float *w = c->window_s;
    for (int b = 0; b < 3; b++) {
        int const k = (576 / 3) * (b + 1);
        int const j = lane;
        int const i = lg_bitrev8(j << 2);
        float f0, f1, f2, f3, v;
        float *o = &xs[b][4 * j];
        f0 = __ldg(&w[i]) * S(i + k); v = __ldg(&w[0x7f - i]) * S(i + k + 0x80); f1 = f0 - v; f0 = f0 + v;
        f2 = __ldg(&w[i + 0x40]) * S(i + k + 0x40); v = __ldg(&w[0x3f - i]) * S(i + k + 0xc0); f3 = f2 - v; f2 = f2 + v;
        o[0] = f0 + f2; o[2] = f0 - f2; o[1] = f1 + f3; o[3] = f1 - f3;
        f0 = __ldg(&w[i + 1]) * S(i + k + 1); v = __ldg(&w[0x7e - i]) * S(i + k + 0x81); f1 = f0 - v; f0 = f0 + v;
        f2 = __ldg(&w[i + 0x41]) * S(i + k + 0x41); v = __ldg(&w[0x3e - i]) * S(i + k + 0xc1); f3 = f2 - v; f2 = f2 + v;
        o[LG_BLK_S / 2 + 0] = f0 + f2; o[LG_BLK_S / 2 + 2] = f0 - f2; o[LG_BLK_S / 2 + 1] = f1 + f3; o[LG_BLK_S / 2 + 3] = f1 - f3;
#undef S
        __syncwarp();
        lg_fht_warp(xs[b], LG_BLK_S, c, lane);
    }
}

/* psymodel.c:583 calc_mask_index_l / :958 vbrpsy_calc_mask_index_s for partition b */
__device__ __forceinline__ int lg_mask_index(const LgBands *__restrict__ gd, const float *mx, const float *av, int b)
{
    int const n = gd->npart;
    float a, m;
    int nl;
    if (b == 0) {
        a = av[0] + av[1];
        m = mx[0]; if (m < mx[1]) m = mx[1];
        nl = gd->numlines[0] + gd->numlines[1] - 1;
        if (!(a > 0.0f)) return 0;
        a = 20.0f * (m * 2.0f - a) / (a * nl);
    }
    else if (b == n - 1) {
        a = av[b - 1] + av[b];
        m = mx[b - 1]; if (m < mx[b]) m = mx[b];
        nl = gd->numlines[b - 1] + gd->numlines[b] - 1;
        if (!(a > 0.0f)) return 0;
        a = 20.0f * (m * 2.0f - a) / (a * nl);
    }
    else {
        a = av[b - 1] + av[b] + av[b + 1];
        m = mx[b - 1]; if (m < mx[b]) m = mx[b]; if (m < mx[b + 1]) m = mx[b + 1];
        nl = gd->numlines[b - 1] + gd->numlines[b] + gd->numlines[b + 1] - 1;
        if (!(a > 0.0f)) return 0;
        a = 20.0f * (m * 3.0f - a) / (a * nl);
    }
    int k = (int) a;
    if (k > 8) k = 8;
    return k;
}

/* partition energies + tonality + spreading for one channel's spectrum energies fe[], by one warp.
 * Returns through eb/mx/av/midx (shared) and writes ecb (after avg_mask) and the minval clamp. */
__device__ __forceinline__ void lg_partition_and_spread(const LgDevCfg *__restrict__ c, const LgBands *__restrict__ gd,
                                                        const float *__restrict__ fe, float *eb, float *mx, float *av,
                                                        int *midx, float *ecb_out, float *lim_out,
                                                        int lane)
{
    int const npart = gd->npart;
    /* psymodel.c:556 calc_energy: serial sum per partition */
    for (int b = lane; b < LG_CBANDS; b += 32) {
        float ebb = 0, m = 0;
        if (b < npart) {
            int const j0 = gd->linestart[b], nl = gd->numlines[b];
            for (int i = 0; i < nl; ++i) {
                float const el = fe[j0 + i];
                ebb += el;
                if (m < el) m = el;
            }
            av[b] = ebb * gd->rnumlines[b];
        }
        else av[b] = 0;
        eb[b] = ebb;
        mx[b] = m;
    }
    __syncwarp();
    for (int b = lane; b < npart; b += 32) midx[b] = lg_mask_index(gd, mx, av, b);
    __syncwarp();
    /* psymodel.c:1154-1185 / :1063-1084 spreading */
    for (int b = lane; b < LG_CBANDS; b += 32) {
        float ecb = 0, lim = 0;
        if (b < npart) {
            int kk = gd->s3lo[b];
            int const last = gd->s3hi[b];
            int k = gd->s3off[b];
            int const delta = LG_MASK_ADD_DELTA[midx[b]];
            int dd = midx[kk], dd_n = 1;
            ecb = __ldg(&gd->s3[k]) * eb[kk] * LG_TONAL_TAB[midx[kk]];
            ++k, ++kk;
            while (kk <= last) {
                dd += midx[kk];
                dd_n += 1;
                float const x = __ldg(&gd->s3[k]) * eb[kk] * LG_TONAL_TAB[midx[kk]];
                ecb = lg_mask_add(c, ecb, x, kk - b, delta);
                ++k, ++kk;
            }
            dd = (1 + 2 * dd) / (2 * dd_n);
            float const avg_mask = LG_TONAL_TAB[dd] * 0.5f;
            ecb *= avg_mask;
            lim = mx[b];
            lim *= gd->minval[b];
            lim *= avg_mask;
        }
        ecb_out[b] = ecb;
        lim_out[b] = lim;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(128)
lg_kernel_analysis(const LgDevCfg *__restrict__ cfg, const int16_t *__restrict__ pcm, int pcm_stride /* samples per channel */,
                   const float *__restrict__ pcmf, const void *__restrict__ pcmn, int pcmn_esz /* bytes per element of a row of pcmn */,
                   const LgPcmKind *__restrict__ kinds, float *__restrict__ sb, LgAnalysis *__restrict__ ana,
                   const int *__restrict__ nfr, int nslots /* 2F+1, F = frames per launch */, int slot0, int cnt /* this launch: slots slot0 .. slot0+cnt-1 */)
{
    LG_DYN_SMEM(LgSmemA, sm);
    int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int const stream = blockIdx.x / cnt, slot = slot0 + blockIdx.x % cnt;
    int const nch = cfg->channels;
    if (slot > cfg->mode_gr * nfr[stream]) return;           /* this stream has fewer frames (of mode_gr granules) in this launch */

    /* phase 1: lame.c:1786 lame_copy_inbuffer (u = xl*m00 + xr*m01, v = xl*m10 + xr*m11) */
    {
        float const m00 = cfg->pcm_transform[0][0], m01 = cfg->pcm_transform[0][1];
        float const m10 = cfg->pcm_transform[1][0], m11 = cfg->pcm_transform[1][1];
        if (pcm) {
            const int16_t *p0 = pcm + (size_t) stream * 2 * pcm_stride + 576 * slot;
            const int16_t *p1 = p0 + pcm_stride;
            for (int i = tid; i < LG_GR_SPAN; i += 128) {
                float const xl = (float) p0[i], xr = (float) p1[i];
                float const u = xl * m00 + xr * m01;
                float const v = xl * m10 + xr * m11;
                sm->pcm[0][LG_PADIDX(i)] = u;
                sm->pcm[1][LG_PADIDX(i)] = v;
            }
        }
        else if (pcmf) {      /* the resampler's output (kernel R) */
            const float *p0 = pcmf + (size_t) stream * 2 * pcm_stride + 576 * slot;
            const float *p1 = p0 + pcm_stride;
            for (int i = tid; i < LG_GR_SPAN; i += 128) {
                sm->pcm[0][LG_PADIDX(i)] = p0[i];
                sm->pcm[1][LG_PADIDX(i)] = p1[i];
            }
        }
        else {                /* the caller's own sample type, every stream its own (lame.c:1803-1834): convert to sample_t, then the matrix */
            LgPcmKind const kd = kinds[stream];
            float const s = kd.scale;                                /* lame.c:1797-1800: m = s * pcm_transform, in float */
            float const n00 = s * m00, n01 = s * m01, n10 = s * m10, n11 = s * m11;
            const char *r0 = (const char *) pcmn + (size_t) stream * 2 * pcm_stride * pcmn_esz;
            const char *r1 = r0 + (size_t) pcm_stride * pcmn_esz;
            for (int i = tid; i < LG_GR_SPAN; i += 128) {
                int const q = 576 * slot + i;
                float xl, xr;
                switch (kd.kind) {
                case LG_PCM_S16: xl = (float) ((const int16_t *) r0)[q]; xr = (float) ((const int16_t *) r1)[q]; break;
                case LG_PCM_S32: xl = (float) ((const int32_t *) r0)[q]; xr = (float) ((const int32_t *) r1)[q]; break;
                case LG_PCM_S64: xl = (float) ((const long long *) r0)[q]; xr = (float) ((const long long *) r1)[q]; break;
                case LG_PCM_F64: xl = (float) ((const double *) r0)[q]; xr = (float) ((const double *) r1)[q]; break;
                default:         xl = ((const float *) r0)[q]; xr = ((const float *) r1)[q]; break;
                }
                if (kd.kind == LG_PCM_DONE) { sm->pcm[0][LG_PADIDX(i)] = xl; sm->pcm[1][LG_PADIDX(i)] = xr; }
                else {
                    sm->pcm[0][LG_PADIDX(i)] = xl * n00 + xr * n01;
                    sm->pcm[1][LG_PADIDX(i)] = xl * n10 + xr * n11;
                }
            }
        }
    }
    __syncthreads();

    LgAnalysis *out = ana + (size_t) stream * (nslots - 1) + (slot > 0 ? slot - 1 : 0);
    int const n_chn_psy = (cfg->mode == LG_JOINT) ? 4 : nch;
    int has_short = 0;

    /* phase 2: psymodel.c:778-795 high-pass FIR, then :831-838 sub-block peaks - and from the peaks whether this granule can switch to
     * short blocks at all.  The FIR also runs over the last six sub-blocks of the granule before (the window holds those samples): the
     * reference compares this granule's first peaks with sub-blocks 4..8 of that granule (last_en_subshort), and an attack it found in
     * sub-block 5 of that granule (against sub-block 3) forces short blocks here (last_attacks == ns_attacks[2] == 3, psymodel.c:904). */
    if (slot > 0) {
        float const k0 = -8.65163e-18 * 2, k1 = -0.00851586 * 2, k2 = -6.74764e-18 * 2, k3 = 0.0209036 * 2,
                    k4 = -3.36639e-17 * 2, k5 = -0.0438162 * 2, k6 = -1.54175e-17 * 2, k7 = 0.0931738 * 2,
                    k8 = -5.52212e-17 * 2, k9 = -0.313819 * 2;
        for (int o = tid; o < 2 * (576 + 384); o += 128) {
            int const ch = o / (576 + 384), q = o % (576 + 384);
            int const i = q < 576 ? q : q - 576 + 192 - 576;          /* own samples 0..575, then -384..-1 */
            float r = 0.0f;
            if (ch < nch) {
                const float *x = sm->pcm[ch];
                int const B = 304 + 576 - 350 - 21 + 192 + i;
#define F(j) x[LG_PADIDX(B + (j))]
                float sum1 = F(10), sum2 = 0.0f;
                sum1 += k0 * (F(0) + F(21)); sum2 += k1 * (F(1) + F(20));
                sum1 += k2 * (F(2) + F(19)); sum2 += k3 * (F(3) + F(18));
                sum1 += k4 * (F(4) + F(17)); sum2 += k5 * (F(5) + F(16));
                sum1 += k6 * (F(6) + F(15)); sum2 += k7 * (F(7) + F(14));
                sum1 += k8 * (F(8) + F(13)); sum2 += k9 * (F(9) + F(12));
#undef F
                r = sum1 + sum2;
            }
            if (q < 576) sm->hpf[ch][q] = r; else sm->hpf_prev[ch][q - 576] = r;
        }
        __syncthreads();
        if (warp < n_chn_psy) {
            int const chn = warp;
            float pk[15];                                             /* [0..5]: sub-blocks 3..8 of the granule before, [6..14]: this granule's nine */
            for (int i = 0; i < 15; i++) {
                const float *h0 = i < 6 ? &sm->hpf_prev[0][64 * i] : &sm->hpf[0][64 * (i - 6)];
                const float *h1 = i < 6 ? &sm->hpf_prev[1][64 * i] : &sm->hpf[1][64 * (i - 6)];
                float p = 1.f;
                for (int k = lane; k < 64; k += 32) {
                    float const l = h0[k], r = h1[k];
                    float v = (chn == 0) ? l : (chn == 1) ? r : (chn == 2) ? (l + r) : (l - r);
                    v = fabsf(v);
                    if (p < v) p = v;
                }
                for (int d = 16; d > 0; d >>= 1) {
                    float const q = __shfl_xor_sync(LG_FULL, p, d);
                    if (p < q) p = q;
                }
                pk[i] = p;
                if (lane == 0 && i >= 6) out->en_subshort[chn][i - 6] = p;
            }
            if (lane == 0) {
                /* psymodel.c:822-872: attack_intensity against the threshold, before any of the rules that take attacks away again */
                float const x = cfg->attack_threshold[chn];
                int hit = 0;
                for (int i = 0; i < 3; i++) hit |= (pk[i + 3] / pk[i + 1]) > x;                /* last_en_subshort[i + 6] / [i + 4] */
                for (int i = -4; i < 9; i++) {
                    /* i = -4: sub-block 5 of the granule before against its sub-block 3 - what that granule's ns_attacks[2] == 3 came from */
                    if (i == -3) i = 0;
                    float p = pk[6 + i];
                    float const e = pk[6 + i - 2];
                    if (p > e) p = p / e;
                    else if (e > p * 10.0f) p = e / (p * 10.0f);
                    else p = 0.0f;
                    hit |= p > x;
                }
                sm->may_attack[chn] = hit;
            }
        }
        __syncthreads();
        for (int chn = 0; chn < n_chn_psy; chn++) has_short |= sm->may_attack[chn];
        if (cfg->short_blocks == 3) has_short = 1;                    /* forced */
        if (cfg->short_blocks == 2) has_short = 0;                    /* dispensed: kernel B never looks */
        if (tid == 0) { out->has_short = has_short; out->pad_ = 0; }
    }

    /* phase 3 */
    if (warp < 2) {
        int const ch = warp;
        if (ch < nch) {
            if (lane < 18) {
                float a[32];
                lg_window_subband(sm->pcm[ch], 286 + 32 * lane, cfg->enwindow, a);
                float *o = sb + (((size_t) stream * nslots + slot) * 2 + ch) * 576 + lane * 32;
#pragma unroll
                for (int band = 0; band < 32; band++) {
                    float v = a[band];
                    if ((lane & 1) && (band & 1)) v *= -1;               /* newmdct.c:969 */
                    /* newmdct.c:984-991: column `band` feeds MDCT band order[band] (bits 1..4 reversed) and is
                     * scaled once, when new, by that band's low-pass gain unless the band is dropped */
                    int const mb = (band & 1) | ((band & 2) << 3) | ((band & 4) << 1) | ((band & 8) >> 1) | ((band & 16) >> 3);
                    float const af = __ldg(&cfg->amp_filter[mb]);
                    if (!(af < 1e-12) && af < 1.0) v *= af;
                    o[band] = v;
                }
            }
            __syncwarp();
            if (slot > 0 && has_short) lg_fft_short_warp(sm->ws[ch], sm->pcm[ch], 304, cfg, lane);
        }
    }
    else if (slot > 0) {
        int const ch = warp - 2;
        if (ch < nch) lg_fft_long_warp(sm->wl[ch], sm->pcm[ch], 304, cfg, lane);
    }
    if (slot == 0) return;
    __syncthreads();

    /* phase 4: line energies */
    {
        float const sqrt2_half = (float) (LG_SQRT2_D * 0.5f);
        for (int o = tid; o < 4 * LG_HBLK; o += 128) {
            int const chn = o / LG_HBLK, m = o % LG_HBLK;
            if (chn >= n_chn_psy) continue;
            int const m2 = (m == 0) ? 0 : LG_BLK - m;
            float re, im;
            if (chn < 2) { re = sm->wl[chn][m]; im = sm->wl[chn][m2]; }
            else {
                float const l0 = sm->wl[0][m], r0 = sm->wl[1][m], l1 = sm->wl[0][m2], r1 = sm->wl[1][m2];
                if (chn == 2) { re = (l0 + r0) * sqrt2_half; im = (l1 + r1) * sqrt2_half; }
                else { re = (l0 - r0) * sqrt2_half; im = (l1 - r1) * sqrt2_half; }
            }
            sm->fe[chn][m] = (m == 0) ? re * re : (re * re + im * im) * 0.5f;
        }
        for (int o = tid; has_short && o < 4 * 3 * LG_HBLK_S; o += 128) {
            int const chn = o / (3 * LG_HBLK_S), sbk = (o / LG_HBLK_S) % 3, m = o % LG_HBLK_S;
            if (chn >= n_chn_psy) continue;
            int const m2 = (m == 0) ? 0 : LG_BLK_S - m;
            float re, im;
            if (chn < 2) { re = sm->ws[chn][sbk][m]; im = sm->ws[chn][sbk][m2]; }
            else {
                float const l0 = sm->ws[0][sbk][m], r0 = sm->ws[1][sbk][m], l1 = sm->ws[0][sbk][m2], r1 = sm->ws[1][sbk][m2];
                if (chn == 2) { re = (l0 + r0) * sqrt2_half; im = (l1 + r1) * sqrt2_half; }
                else { re = (l0 - r0) * sqrt2_half; im = (l1 - r1) * sqrt2_half; }
            }
            sm->fes[chn][sbk][m] = (m == 0) ? re * re : (re * re + im * im) * 0.5f;
        }
    }
    __syncthreads();

    /* phase 5: one warp per psycho-acoustic channel */
    if (warp < n_chn_psy) {
        int const chn = warp;
        /* the two 512-term serial sums: loudness (psymodel.c:213) for L/R, tot_ener (psymodel.c:690) for M/S */
        if (lane == 0) {
            if (chn < 2) {
                float loudness_power = 0.0f;
                for (int i = 0; i < LG_BLK / 2; ++i) loudness_power += sm->fe[chn][i] * __ldg(&cfg->eql_w[i]);
                loudness_power = (float) (loudness_power * (1. / (14752 * 14752) / (LG_BLK / 2)));
                out->loudness[chn] = loudness_power;
            }
            float totalenergy = 0.0f;
            for (int j = 11; j < LG_HBLK; j++) totalenergy += sm->fe[chn][j];
            out->tot_ener[chn] = totalenergy;
        }
        __syncwarp();
        lg_partition_and_spread(cfg, &cfg->l, sm->fe[chn], sm->eb[chn], sm->mx[chn], sm->av[chn], sm->midx[chn],
                                out->ecb_l[chn], out->lim_l[chn], lane);
        for (int b = lane; b < LG_CBANDS; b += 32) out->eb_l[chn][b] = sm->eb[chn][b];
        __syncwarp();
        for (int sbk = 0; has_short && sbk < 3; sbk++) {
            float *thr = out->thr_s[sbk][chn];
            lg_partition_and_spread(cfg, &cfg->s, sm->fes[chn][sbk], sm->eb[chn], sm->mx[chn], sm->av[chn], sm->midx[chn],
                                    thr, sm->av[chn] /* reuse as clamp scratch */, lane);
            for (int b = lane; b < LG_CBANDS; b += 32) {
                float t = thr[b];
                float const x = sm->av[chn][b];
                if (t > x) t = x;                                       /* psymodel.c:1108-1113 */
                thr[b] = t;
                out->eb_s[sbk][chn][b] = sm->eb[chn][b];
            }
            __syncwarp();
        }
    }
}
