// lg_math.cuh - scalar device helpers whose results must match the host's bit for bit.
//
// The reference calls libm at run time in a few places on this path (VBR-new adds log10f, VBR-old exp and pow: see lg_log10f, lg_exp,
// lg_pow below); for CBR/ABR in three: powf() in athAdjust
// (quantize_pvt.c:572) and NS_INTERP (psymodel.c:452), and pow(x, .5) in amp_scalefac_bands
// (quantize.c:758).  The GPU cannot call glibc, so lg_powf() restates glibc 2.39's powf
// (sysdeps/ieee754/flt-32/e_powf.c: 16-entry log2 table + degree-5 polynomial, 32-entry exp2 table +
// cubic, all in binary64) for non-negative bases (zero, subnormal, normal, +inf) and finite exponents, the only cases
// this path produces.  tests/test_powf.py checks it against the host's powf over 2e8 arguments.
// pow(x, .5) rounded to float equals (float)sqrt((double)x) for every float x (a 24-bit x cannot have a
// square root within 2^-50 of a 25-bit rounding boundary), so the device uses the IEEE sqrt.
// lg_log10f() likewise restates glibc 2.39's log10f (e_log10f.c over e_logf.c) for calc_scalefac (vbrquantize.c:317).
// Everything here is compiled with -fmad=false: no contraction, same operations as the host build.
#pragma once
#include "lg_compat.h"
#include "lg_types.h"

#define LG_SQRT2_D 1.41421356237309504880
#define LG_LOG2_D 0.69314718055994530942
#define LG_LOG10_D 2.30258509299404568402

__constant__ double LG_POW_LOGTAB[16][2] = {
    { 0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2 }, { 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2 },
    { 0x1.49539f0f010bp+0, -0x1.7418b0a1fb77bp-2 },  { 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2 },
    { 0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2 }, { 0x1.25e227b0b8eap+0, -0x1.97c1d1b3b7afp-3 },
    { 0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3 }, { 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4 },
    { 0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5 }, { 0x1p+0, 0x0p+0 },
    { 0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4 },  { 0x1.ca4b31f026aap-1, 0x1.476a9543891bap-3 },
    { 0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2 },
    { 0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2 },  { 0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2 } };
__constant__ unsigned long long LG_POW_EXPTAB[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull };

__device__ __noinline__ float lg_powf(float x, float y)
{
    unsigned ix = __float_as_uint(x);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        /* e_powf.c special cases that this path can reach: x == +0 (a silent short window in NS_INTERP),
         * x == +inf (ratio overflow) and subnormal x; y is always a small positive constant or finite */
        if ((ix << 1) == 0u) return y > 0.f ? 0.f : __uint_as_float(0x7f800000u);
        if (ix == 0x7f800000u) return y > 0.f ? x : 0.f;
        if (ix & 0x80000000u) return __uint_as_float(0x7fc00000u);
        if (ix > 0x7f800000u) return x + y;
        ix = __float_as_uint(x * 0x1p23f);
        ix &= 0x7fffffffu;
        ix -= 23u << 23;
    }
    unsigned const tmp = ix - 0x3f330000u;
    int const i = (int) ((tmp >> 19) & 15u);
    unsigned const top = tmp & 0xff800000u;
    unsigned const iz = ix - top;
    int const k = (int) top >> 23;
    double const invc = LG_POW_LOGTAB[i][0], logc = LG_POW_LOGTAB[i][1];
    double const z = (double) __uint_as_float(iz);
    double r = z * invc - 1;
    double const y0 = logc + (double) k;
    double r2 = r * r;
    double yy = 0x1.27616c9496e0bp-2 * r + -0x1.71969a075c67ap-2;
    double const p = 0x1.ec70a6ca7baddp-2 * r + -0x1.7154748bef6c8p-1;
    double const r4 = r2 * r2;
    double q = 0x1.71547652ab82bp0 * r + y0;
    q = p * r2 + q;
    yy = yy * r4 + q;
    double const ylogx = y * yy;
    if (ylogx > 0x1.fffffffd1d571p+6) return __uint_as_float(0x7f800000u);     /* overflow  */
    if (ylogx <= -150.0) return 0.f;                                           /* underflow */
    double kd = ylogx + 0x1.8p+52 / 32;
    unsigned long long const ki = (unsigned long long) __double_as_longlong(kd);
    kd -= 0x1.8p+52 / 32;
    r = ylogx - kd;
    unsigned long long t = LG_POW_EXPTAB[ki & 31u];
    t += ki << (52 - 5);
    double const s = __longlong_as_double((long long) t);
    double const zz = 0x1.c6af84b912394p-5 * r + 0x1.ebfce50fac4f3p-3;
    r2 = r * r;
    yy = 0x1.62e42ff0c52d6p-1 * r + 1;
    yy = zz * r2 + yy;
    yy = yy * s;
    return (float) yy;
}

/* glibc 2.39 logf_data (sysdeps/ieee754/flt-32/e_logf_data.c): 16 x { 1/c, log(c) } */
__constant__ double LG_LOGF_TAB[16][2] = {
    { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
    { 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2 },  { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
    { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3 },
    { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
    { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1p+0, 0x0p+0 },
    { 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 },  { 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4 },
    { 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3 },
    { 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 },  { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 } };

/* glibc 2.39 log10f = the fdlibm wrapper of sysdeps/ieee754/flt-32/e_log10f.c (x = 2^k * m, log10 = k*log10(2) in two
 * floats + ivln10 * logf(m), all in binary32) around the table-driven logf of e_logf.c (binary64 inside, one rounding).
 * The one run-time caller on this path is calc_scalefac (vbrquantize.c:317, the quality-7 step guess of VBR-new);
 * its argument l3_xmin / width is a positive float, possibly subnormal, inf when the threshold overflowed.
 * tests/test_powf.py checks it against the host's log10f on every positive float. */
__device__ __noinline__ float lg_log10f(float x)
{
    int hx = __float_as_int(x), k = 0;
    if (hx < 0x00800000) {
        if ((hx & 0x7fffffff) == 0) return __uint_as_float(0xff800000u);          /* log10(+-0) = -inf */
        if (hx < 0) return __uint_as_float(0x7fc00000u);                          /* log10(-#) = NaN   */
        k -= 25;
        x *= 3.3554432000e+07f;                                                   /* subnormal: scale by 2^25 */
        hx = __float_as_int(x);
    }
    if (hx >= 0x7f800000) return x + x;
    k += (hx >> 23) - 127;
    int const i = (int) (((unsigned) k & 0x80000000u) >> 31);
    hx = (hx & 0x007fffff) | ((0x7f - i) << 23);
    float const y = (float) (k + i);
    /* e_logf.c on m = 2^-i * 1.f, which is a normal number in [0.5, 2) */
    float lf;
    unsigned const ix = (unsigned) hx;
    if (ix == 0x3f800000u) lf = 0.f;
    else {
        unsigned const tmp = ix - 0x3f330000u;
        int const ti = (int) ((tmp >> 19) & 15u);
        int const tk = (int) tmp >> 23;
        unsigned const iz = ix - (tmp & (0x1ffu << 23));
        double const invc = LG_LOGF_TAB[ti][0], logc = LG_LOGF_TAB[ti][1];
        double const z = (double) __uint_as_float(iz);
        double const r = z * invc - 1;
        double const y0 = logc + (double) tk * 0x1.62e42fefa39efp-1;
        double const r2 = r * r;
        double yy = 0x1.5575b0be00b6ap-2 * r + -0x1.ffffef20a4123p-2;
        yy = -0x1.00ea348b88334p-2 * r2 + yy;
        yy = yy * r2 + (y0 + r);
        lf = (float) yy;
    }
    float const z = y * 7.9034151668e-07f + 4.3429449201e-01f * lf;
    return z + y * 3.0102920532e-01f;
}

#include "lg_libm_tab.inc"

/* glibc 2.39 exp (sysdeps/ieee754/dbl-64/e_exp.c): x = k ln2/128 + r, 2^(k/128) from a 128-entry table with a tail, degree-5 polynomial.
 * On x86-64 glibc dispatches to the copy of that file built with -mfma (sysdeps/x86_64/fpu/multiarch/e_exp-fma.c, picked on every CPU
 * with AVX2 + FMA), in which the compiler contracted each a*b + c of the source into one fused operation; the fma() calls below are that
 * build's operations in its order (read off its machine code), so the result is the host libm's bit for bit - the reference's own result
 * is that of the libm it runs on.  The one run-time caller on this path is the masking feedback of VBR-old (quantize.c:1419-1426):
 * exp(3.5 - pe/300).  Where glibc's result is 1 + x (|x| < 2^-54), inf, 0 or a subnormal number (x < -708.39) CUDA's exp stands in: the
 * first three are the same values, the last is unreachable for a perceptual entropy below 2e5.  tests/test_powf.py checks both against the host's libm. */
__device__ __forceinline__ double lg_exp_core(double x, double xtail, bool has_tail, bool large, bool *subnormal)
{
    double kd = fma(x, 0x1.71547652b82fep+7, 0x1.8p+52);
    unsigned long long const ki = (unsigned long long) __double_as_longlong(kd);
    kd -= 0x1.8p+52;
    double r = fma(kd, -0x1.cf79abc9e3b3ap-47, fma(kd, -0x1.62e42fefa0000p-8, x));
    if (has_tail) r = xtail + r;
    unsigned const idx = 2u * (unsigned) (ki & 127u);
    unsigned long long const top = ki << (52 - 7);
    double const tail = __longlong_as_double((long long) LG_EXP_TAB[idx]);
    unsigned long long const sbits = LG_EXP_TAB[idx + 1] + top;
    double const r2 = r * r;
    double const lo = fma(r, 0x1.555555555543cp-3, 0x1.ffffffffffdbdp-2);
    double const hi = fma(r, 0x1.1111167a4d017p-7, 0x1.55555cf172b91p-5);
    double const tmp = fma(hi, r2 * r2, fma(lo, r2, r + tail));
    if (large) {
        /* e_exp.c specialcase(), 512 <= |x| < 1024: the scale 2^(k/128) alone would over- or underflow */
        if ((ki & 0x80000000ull) == 0) {
            double const scale = __longlong_as_double((long long) (sbits - (1009ull << 52)));
            return 0x1p1009 * fma(scale, tmp, scale);
        }
        double const scale = __longlong_as_double((long long) (sbits + (1022ull << 52)));
        double const yy = scale + scale * tmp;                                  /* not fused in the host's build */
        if (fabs(yy) < 1.0) { *subnormal = true; return 0.0; }                  /* subnormal result: the caller falls back */
        return 0x1p-1022 * yy;
    }
    double const scale = __longlong_as_double((long long) sbits);
    return fma(tmp, scale, scale);
}
__device__ __noinline__ double lg_exp(double x)
{
    unsigned const abstop = (unsigned) ((unsigned long long) __double_as_longlong(x) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x409u - 0x3c9u) return exp(x);                      /* |x| < 2^-54: 1 + x; |x| >= 1024: inf or 0 */
    bool sub = false;
    double const v = lg_exp_core(x, 0.0, false, abstop == 0x408u, &sub);
    return sub ? exp(x) : v;
}

/* glibc 2.39 pow (sysdeps/ieee754/dbl-64/e_pow.c, its -mfma build as above) for a positive normal base and an exponent whose result is a
 * normal number: log(x) as hi + lo from a 128-entry table and a degree-7 polynomial, y log(x) as ehi + elo, then the exp above with the
 * tail.  Caller: pow(10.0, masking_lower_db * 0.1) of VBR-old (quantize.c:1426), |y| < 4; glibc's special cases (y = 0 among them) go to
 * CUDA's pow, exact there. */
__device__ __noinline__ double lg_pow(double x, double y)
{
    unsigned long long const ix = (unsigned long long) __double_as_longlong(x), iy = (unsigned long long) __double_as_longlong(y);
    unsigned const topx = (unsigned) (ix >> 52), topy = (unsigned) (iy >> 52) & 0x7ffu;
    if (topx - 0x001u >= 0x7ffu - 0x001u || topy - 0x3beu >= 0x43eu - 0x3beu) return pow(x, y);
    unsigned long long const tmp = ix - 0x3fe6955500000000ull;
    int const i = (int) ((tmp >> (52 - 7)) & 127u);
    int const k = (int) ((long long) tmp >> 52);
    unsigned long long const iz = ix - (tmp & (0xfffull << 52));
    double const z = __longlong_as_double((long long) iz), kd = (double) k;
    double const invc = LG_POWLOG_TAB[i][0], logc = LG_POWLOG_TAB[i][1], logctail = LG_POWLOG_TAB[i][2];
    double const t1 = fma(kd, 0x1.62e42fefa3800p-1, logc);
    double const lo1 = fma(kd, 0x1.ef35793c76730p-45, logctail);
    double const r = fma(z, invc, -1.0);
    double const ar = r * -0x1p-1;
    double const q12 = fma(r, 0x1.0000000000006p-1, -0x1.555555555556p-1);
    double const q34 = fma(r, -0x1.555555529a47ap-1, 0x1.999999959554ep-1);
    double const t2 = r + t1;
    double const lo2 = (t1 - t2) + r;
    double const ar2 = r * ar;
    double const ar3 = r * ar2;
    double const lo3 = fma(ar, r, -ar2);
    double const hi = t2 + ar2;
    double const q56 = fma(r, 0x1.0002b8b263fc3p+0, -0x1.2495b9b4845e9p+0);
    double const lo4 = (t2 - hi) + ar2;
    double const q = fma(ar2, fma(q56, ar2, q34), q12);
    double const lo = fma(ar3, q, ((lo1 + lo2) + lo3) + lo4);
    double const lhi = hi + lo;
    double const ltail = (hi - lhi) + lo;
    double const ehi = y * lhi;
    double const elo = fma(y, ltail, fma(lhi, y, -ehi));
    unsigned const abstop = (unsigned) ((unsigned long long) __double_as_longlong(ehi) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x409u - 0x3c9u) return pow(x, y);                    /* result 1, inf or 0 */
    bool sub = false;
    double const v = lg_exp_core(ehi, elo, true, abstop == 0x408u, &sub);
    return sub ? pow(x, y) : v;
}

/* util.c:977 fast_log2 (513-entry table + linear interpolation) */
__device__ __forceinline__ float lg_fast_log2(const float *__restrict__ log_table, float x)
{
    int const fi = __float_as_int(x);
    int mantisse = fi & 0x7fffff;
    float log2val = (float) (((fi >> 23) & 0xFF) - 0x7f);
    float partial = (float) (mantisse & ((1 << (23 - 9)) - 1));
    partial *= 1.0f / ((1 << (23 - 9)));
    mantisse >>= (23 - 9);
    log2val += __ldg(&log_table[mantisse]) * (1.0f - partial) + __ldg(&log_table[mantisse + 1]) * partial;
    return log2val;
}
/* the same with the table in shared memory */
__device__ __forceinline__ float lg_fast_log2_smem(const float *log_table, float x)
{
    int const fi = __float_as_int(x);
    int mantisse = fi & 0x7fffff;
    float log2val = (float) (((fi >> 23) & 0xFF) - 0x7f);
    float partial = (float) (mantisse & ((1 << (23 - 9)) - 1));
    partial *= 1.0f / ((1 << (23 - 9)));
    mantisse >>= (23 - 9);
    log2val += log_table[mantisse] * (1.0f - partial) + log_table[mantisse + 1] * partial;
    return log2val;
}
/* util.h:96 FAST_LOG10 / FAST_LOG10_X are double-valued expressions */
#define LG_FAST_LOG10_SMEM_D(tab, x) ((double) lg_fast_log2_smem(tab, x) * (LG_LOG2_D / LG_LOG10_D))
#define LG_FAST_LOG10_D(tab, x) ((double) lg_fast_log2(tab, x) * (LG_LOG2_D / LG_LOG10_D))
#define LG_FAST_LOG10_X_D(tab, x, y) ((double) lg_fast_log2(tab, x) * (LG_LOG2_D / LG_LOG10_D * (y)))

/* psymodel.c:258 tab[] / :270 tab_mask_add_delta[] / :297 table2[] */
__constant__ float LG_TONAL_TAB[9] = { 1.0, 0.79433, 0.63096, 0.63096, 0.63096, 0.63096, 0.63096, 0.25119, 0.11749 };
__constant__ int LG_MASK_ADD_DELTA[9] = { 2, 2, 2, 1, 1, 1, 0, 0, -1 };
__constant__ float LG_MASK_TABLE2[10] = { 1.33352 * 1.33352, 1.35879 * 1.35879, 1.38454 * 1.38454, 1.39497 * 1.39497,
    1.40548 * 1.40548, 1.3537 * 1.3537, 1.30382 * 1.30382, 1.22321 * 1.22321, 1.14758 * 1.14758, 1 };

/* psymodel.c:294 vbrpsy_mask_add */
__device__ __forceinline__ float lg_mask_add(const LgDevCfg *__restrict__ c, float m1, float m2, int b, int delta)
{
    float ratio;
    if (m1 < 0) m1 = 0;
    if (m2 < 0) m2 = 0;
    if (m1 <= 0) return m2;
    if (m2 <= 0) return m1;
    if (m2 > m1) ratio = m2 / m1; else ratio = m1 / m2;
    if ((b < 0 ? -b : b) <= delta) {
        if (ratio >= c->ma_max_i1) return m1 + m2;
        int const i = (int) (LG_FAST_LOG10_X_D(c->log_table, ratio, 16.0f));
        return (m1 + m2) * LG_MASK_TABLE2[i];
    }
    if (ratio < c->ma_max_i2) return m1 + m2;
    if (m1 < m2) m1 = m2;
    return m1;
}

/* psymodel.c:443 NS_INTERP */
__device__ __forceinline__ float lg_ns_interp(float x, float y, float r)
{
    if (r >= 1.0f) return x;
    if (r <= 0.0f) return y;
    if (y > 0.0f) return lg_powf(x / y, r) * y;
    return 0.0f;
}

/* quantize_pvt.c:554 athAdjust */
__device__ __noinline__ float lg_ath_adjust(const LgDevCfg *__restrict__ c, float a, float x, float athFloor, float fixpoint)
{
    float const o = 90.30873362f;
    float const p = (fixpoint < 1.f) ? 94.82444863f : fixpoint;
    float u = (float) LG_FAST_LOG10_X_D(c->log_table, x, 10.0f);
    float const v = a * a;
    float w = 0.0f;
    u -= athFloor;
    if (v > 1E-20f) w = (float) (1.f + LG_FAST_LOG10_X_D(c->log_table, v, 10.0f / o));
    if (w < 0) w = 0.f;
    u *= w;
    u += athFloor + o - p;
    return lg_powf(10.f, 0.1f * u);
}

__device__ __forceinline__ int lg_bitrev8(int v)
{
#ifdef LG_EMULATE
    int r = 0;
    for (int b = 0; b < 8; b++) if (v & (1 << b)) r |= 0x80 >> b;
    return r;
#else
    return (int) (__brev((unsigned) v) >> 24);
#endif
}
