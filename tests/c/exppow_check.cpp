/* Test harness: lg_exp / lg_pow (csrc/lg_math.cuh, the device restatements of glibc's binary64 exp and pow) compiled for the host
 * through the emulator shims, against the host's libm - bit for bit - over what the masking feedback of VBR-old (quantize.c:1419-1426)
 * can produce:  exp(3.5 - pe / 300.) for float pe,  pow(10.0, db * 0.1) for float db;  plus random doubles of a wider range.
 * usage: exppow_check <millions of random arguments>   exit code 0 = all identical */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include "simt_emu.h"
#include "lg_compat.h"
#include "lg_math.cuh"

static uint64_t rng = 88172645463325252ull;
static uint64_t next() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; }
static double unit() { return (double) (next() >> 11) * (1.0 / 9007199254740992.0); }
static int same(double a, double b) { return !memcmp(&a, &b, 8) || (a != a && b != b); }

int main(int argc, char **argv)
{
    long const n = (argc > 1 ? atol(argv[1]) : 20) * 1000000L;
    long bad_e = 0, bad_p = 0, total = 0;
    for (long k = 0; k < n; k++) {
        double xe, yp, xb = 10.0;
        switch (k & 3) {
        case 0: { float const pe = (float) (unit() * 6000.0); xe = 3.5 - pe / 300.; float const db = (float) (unit() * 24.0 - 12.0); yp = db * 0.1; break; }
        case 1: { float const pe = (float) (unit() * 1.0e5); xe = 3.5 - pe / 300.; float const db = (float) (unit() * 2.0 - 1.0); yp = db * 0.1; break; }
        case 2: xe = unit() * 1416.0 - 707.0; yp = unit() * 8.0 - 4.0; break;       /* exp: the whole range of normal results */
        default: xe = (unit() - 0.5) * 1e-3; yp = unit() * 200.0 - 100.0; xb = 0.001 + unit() * 1000.0; break;   /* |y log x| up to 691 */
        }
        if (!same(lg_exp(xe), exp(xe))) { if (bad_e < 5) printf("exp(%a): %a vs %a\n", xe, lg_exp(xe), exp(xe)); bad_e++; }
        if (!same(lg_pow(xb, yp), pow(xb, yp))) { if (bad_p < 5) printf("pow(%a, %a): %a vs %a\n", xb, yp, lg_pow(xb, yp), pow(xb, yp)); bad_p++; }
        total++;
    }
    printf("%ld arguments each, exp differs on %ld, pow on %ld: %s\n", total, bad_e, bad_p, (bad_e || bad_p) ? "MISMATCH" : "IDENTICAL");
    return (bad_e || bad_p) ? 1 : 0;
}
