/* Test harness: lg_powf (csrc/lg_math.cuh, the device restatement of glibc's powf) compiled for the host through the
 * emulator shims, against the host's powf - bit for bit - over the argument ranges the encoder produces:
 *   athAdjust        powf(10, 0.1 * u)          u in about [-200, 200]
 *   NS_INTERP        powf(x / y, r) * y         ratio > 0 over the float range, r in (0, 1)
 * usage: powf_check <millions of random pairs>   exit code 0 = all identical */
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
static int same(float a, float b) { return !memcmp(&a, &b, 4) || (a != a && b != b); }

int main(int argc, char **argv)
{
    long const n = (argc > 1 ? atol(argv[1]) : 20) * 1000000L;
    long bad = 0, total = 0;
    /* 1. base 10, every exponent 0.1f * u for u on a fine grid */
    for (int i = -4000000; i <= 4000000; i++) {
        float const u = i * 0.0001f, y = 0.1f * u;
        if (!same(lg_powf(10.f, y), powf(10.f, y))) { if (bad < 5) printf("powf(10, %.9g): %.9g vs %.9g\n", y, lg_powf(10.f, y), powf(10.f, y)); bad++; }
        total++;
    }
    /* 2. random positive bases over the whole float range (incl. subnormals), exponents in (0, 1) and a few fixed ones */
    static const float fixed[6] = { 0.3f, 0.5f, 0.75f, 0.25f, 0.9f, 0.1f };
    for (long k = 0; k < n; k++) {
        uint32_t bits = (uint32_t) next() & 0x7fffffffu;
        if (bits > 0x7f800000u) bits = 0x7f800000u;
        float x, y;
        memcpy(&x, &bits, 4);
        y = (k & 1) ? fixed[(k >> 1) % 6] : (float) ((next() >> 11) * (1.0 / 9007199254740992.0));
        if (!same(lg_powf(x, y), powf(x, y))) { if (bad < 5) printf("powf(%.9g, %.9g): %.9g vs %.9g\n", x, y, lg_powf(x, y), powf(x, y)); bad++; }
        total++;
    }
    /* 3. special bases */
    { float const sp[4] = { 0.f, INFINITY, 1.f, 1e-45f };
      for (int a = 0; a < 4; a++) for (int b = 0; b < 6; b++) { if (!same(lg_powf(sp[a], fixed[b]), powf(sp[a], fixed[b]))) bad++; total++; } }
    printf("%ld arguments, %ld differ: %s\n", total, bad, bad ? "MISMATCH" : "IDENTICAL");
    return bad ? 1 : 0;
}
