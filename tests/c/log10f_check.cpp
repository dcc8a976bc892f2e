/* Test harness: lg_log10f (csrc/lg_math.cuh, the device restatement of glibc's log10f) compiled for the host through the
 * emulator shims, against the host's log10f - bit for bit - on EVERY float from +0 to +inf inclusive (2^31 - 2^23 + 1
 * arguments; calc_scalefac, vbrquantize.c:317, only passes positive ones), a NaN and two negative arguments.
 * usage: log10f_check [stride]   stride 1 = exhaustive; exit code 0 = all identical */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include "simt_emu.h"
#include "lg_compat.h"
#include "lg_math.cuh"

static int same(float a, float b) { return !memcmp(&a, &b, 4) || (a != a && b != b); }

int main(int argc, char **argv)
{
    uint32_t const stride = argc > 1 ? (uint32_t) atol(argv[1]) : 1u;
    long bad = 0, total = 0;
    for (uint64_t b = 0; b <= 0x7f800000u; b += stride) {
        uint32_t const bits = (uint32_t) b;
        float x;
        memcpy(&x, &bits, 4);
        float const mine = lg_log10f(x), ref = log10f(x);
        if (!same(mine, ref)) { if (bad < 5) printf("log10f(%a): %a vs %a\n", x, mine, ref); bad++; }
        total++;
    }
    { float const sp[4] = { INFINITY, NAN, -1.f, -0.f };
      for (int a = 0; a < 4; a++) { if (!same(lg_log10f(sp[a]), log10f(sp[a]))) { printf("special %d differs\n", a); bad++; } total++; } }
    printf("%ld arguments, %ld differ: %s\n", total, bad, bad ? "MISMATCH" : "IDENTICAL");
    return bad ? 1 : 0;
}
