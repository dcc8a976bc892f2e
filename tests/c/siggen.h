/* Deterministic test signals shared by the C harnesses.
 * xorshift64 with seed 88172645463325252. */
#ifndef SIGGEN_H
#define SIGGEN_H
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
static uint64_t sg_state = 88172645463325252ULL;
static uint32_t sg_next(void) { sg_state ^= sg_state << 13; sg_state ^= sg_state >> 7; sg_state ^= sg_state << 17; return (uint32_t) (sg_state >> 16); }
static int sg_uniform(int amp) { return (int) (sg_next() % (2u * amp + 1u)) - amp; }
static int siggen(const char *kind, short *l, short *r, int n, int sr, const char *wav)
{
    int i;
    sg_state = 88172645463325252ULL;
    if (!strcmp(kind, "noise")) { for (i = 0; i < n; i++) { l[i] = sg_uniform(12000); r[i] = sg_uniform(12000); } return 0; }
    if (!strcmp(kind, "gap")) { for (i = 0; i < n; i++) { l[i] = sg_uniform(12000); r[i] = sg_uniform(12000); if (i < 6000 || (i > 20000 && i < 23000)) l[i] = r[i] = 0; } return 0; }
    if (!strcmp(kind, "silence")) { memset(l, 0, n * 2); memset(r, 0, n * 2); return 0; }
    if (!strcmp(kind, "sine")) {
        for (i = 0; i < n; i++) {
            double t = (double) i / sr;
            l[i] = (short) lrint(8000 * sin(2 * M_PI * 440 * t) + 4000 * sin(2 * M_PI * 3300 * t) + sg_uniform(1000));
            r[i] = (short) lrint(8000 * sin(2 * M_PI * 554.37 * t) + 3000 * sin(2 * M_PI * 7000 * t) + sg_uniform(1000));
        }
        return 0;
    }
    if (!strcmp(kind, "click")) {        /* quiet noise floor with loud decaying bursts: forces short blocks */
        for (i = 0; i < n; i++) {
            int ph = i % 7919, ph2 = (i + 3000) % 10007;
            double env = ph < 400 ? exp(-ph / 60.0) : 0.0, env2 = ph2 < 300 ? exp(-ph2 / 40.0) : 0.0;
            l[i] = (short) lrint(sg_uniform(60) + env * sg_uniform(24000));
            r[i] = (short) lrint(sg_uniform(60) + 0.8 * env * sg_uniform(24000) + env2 * sg_uniform(16000));
        }
        return 0;
    }
    if (!strcmp(kind, "wav") && wav) {   /* 16-bit stereo PCM after a 44-byte header */
        FILE *f = fopen(wav, "rb"); short s[2]; if (!f) return -1;
        fseek(f, 44, SEEK_SET);
        for (i = 0; i < n; i++) { if (fread(s, 2, 2, f) != 2) { s[0] = s[1] = 0; } l[i] = s[0]; r[i] = s[1]; }
        fclose(f); return 0;
    }
    return -1;
}
#endif
