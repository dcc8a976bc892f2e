/* Test harness: the batch C ABI (liblamegpu.so on a GPU box, or the emulator build liblamegpu_emu.so on
 * the CPU box) against the oracle port, stream by stream, byte for byte.
 * usage: gpu_vs_port <nstreams> <frames_per_stream> <frames_per_launch> <brate> <mode> <quality> <samplerate> <chunk>
 * Stream i uses signal kind (i % 4) of {noise, sine, click, silence-then-noise} with a stream-specific seed. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/lamegpu.h"
#include "../../oracle/port/lame_port.h"
#include "siggen.h"

int main(int argc, char **argv)
{
    int S = argc > 1 ? atoi(argv[1]) : 2, NF = argc > 2 ? atoi(argv[2]) : 10, FPL = argc > 3 ? atoi(argv[3]) : 4;
    int brate = argc > 4 ? atoi(argv[4]) : 128, mode = argc > 5 ? atoi(argv[5]) : -1, quality = argc > 6 ? atoi(argv[6]) : -1;
    int sr = argc > 7 ? atoi(argv[7]) : 44100, chunk = argc > 8 ? atoi(argv[8]) : 1152;
    static const char *kinds[4] = { "noise", "sine", "click", "noise" };
    int n = NF * 1152, s, i, bad = 0, cap = 2 * (n * 5 / 4 + 7200) + 65536;
    short **L = malloc(S * sizeof *L), **R = malloc(S * sizeof *R);
    unsigned char **out = malloc(S * sizeof *out), **ref = malloc(S * sizeof *ref);
    int *olen = calloc(S, sizeof *olen), *rlen = calloc(S, sizeof *rlen), *ocap = malloc(S * sizeof *ocap), *ob = malloc(S * sizeof *ob), *ns = malloc(S * sizeof *ns);
    const short **pl = malloc(S * sizeof *pl), **pr = malloc(S * sizeof *pr);
    unsigned char **po = malloc(S * sizeof *po);
    lamegpu_batch *b;
    long frames = 0;
    for (s = 0; s < S; s++) {
        L[s] = malloc(n * 2); R[s] = malloc(n * 2); out[s] = malloc(cap); ref[s] = malloc(cap);
        siggen(kinds[s % 4], L[s], R[s], n, sr, NULL);
        /* decorrelate streams: rotate by a stream-specific offset; kind 3 starts with silence */
        { int off = (s * 7919) % n; short *t = malloc(n * 2);
          memcpy(t, L[s] + off, (n - off) * 2); memcpy(t + n - off, L[s], off * 2); memcpy(L[s], t, n * 2);
          memcpy(t, R[s] + off, (n - off) * 2); memcpy(t + n - off, R[s], off * 2); memcpy(R[s], t, n * 2); free(t); }
        if (s % 4 == 3) { int z = n / 3; memset(L[s], 0, z * 2); memset(R[s], 0, z * 2); }
    }
    int const vbr = getenv("LP_VBR") ? atoi(getenv("LP_VBR")) : 0;          /* 0 = CBR, 3 = ABR with mean bitrate `brate` */
    int const out_sr = getenv("LP_OUT_SR") ? atoi(getenv("LP_OUT_SR")) : 0; /* explicit output rate, 0 = automatic (resampling when it differs from sr) */
    float const qfrac = getenv("LP_VBRQ_FRAC") ? (float) atof(getenv("LP_VBRQ_FRAC")) : 0.f;  /* VBR quality = brate + this */
    int const device = getenv("LP_DEVICE") ? atoi(getenv("LP_DEVICE")) : 0;  /* -1 = the batch spans all visible devices */
    b = lamegpu_batch_open_vq(sr, out_sr, 2, (float) brate + qfrac, mode, quality, vbr, S, FPL, device);
    if (b && device < 0) printf("batch spans %d device(s)\n", lamegpu_batch_devices(b));
    if (b && getenv("LP_PIPELINED")) lamegpu_batch_set_pipelined(b, 1);      /* bytes lag one step behind; the flush brings everything out */
    if (!b) { printf("batch open failed\n"); return 2; }
    for (i = 0; i < n; i += chunk) {
        int c = n - i < chunk ? n - i : chunk;
        long r;
        for (s = 0; s < S; s++) { pl[s] = L[s] + i; pr[s] = R[s] + i; ns[s] = c; po[s] = out[s] + olen[s]; ocap[s] = cap - olen[s]; }
        r = lamegpu_batch_encode(b, pl, pr, ns, po, ocap, ob);
        if (r < 0) { printf("encode error %ld\n", r); return 2; }
        frames += r;
        for (s = 0; s < S; s++) olen[s] += ob[s];
    }
    for (s = 0; s < S; s++) { po[s] = out[s] + olen[s]; ocap[s] = cap - olen[s]; }
    frames += lamegpu_batch_flush(b, po, ocap, ob);
    for (s = 0; s < S; s++) olen[s] += ob[s];
    { float ms[5]; lamegpu_batch_kernel_ms(b, ms); printf("last launch kernel ms: analysis %.3f scan %.3f mdct %.3f quant %.3f\n", ms[0], ms[1], ms[2], ms[3]); }
    lamegpu_batch_close(b);
    for (s = 0; s < S; s++) {
        lp_encoder *e = lp_open_vq(sr, out_sr, 2, brate, mode < 0 ? LP_MODE_NOT_SET : mode, quality, vbr, (vbr == 4 || vbr == 2) ? ((float) brate + qfrac) - (float) brate : 0.f);
        int k;
        if (!e) { printf("port open failed\n"); return 2; }
        /* same call pattern as above: with resampling the reference's state depends on where the calls end */
        for (i = 0; i < n; i += chunk) rlen[s] += lp_encode(e, L[s] + i, R[s] + i, n - i < chunk ? n - i : chunk, ref[s] + rlen[s], cap - rlen[s]);
        rlen[s] += lp_flush(e, ref[s] + rlen[s], cap - rlen[s]);
        lp_close(e);
        if (olen[s] != rlen[s] || memcmp(out[s], ref[s], rlen[s])) {
            for (k = 0; k < olen[s] && k < rlen[s]; k++) if (out[s][k] != ref[s][k]) break;
            printf("stream %d (%s): MISMATCH gpu %d bytes, port %d bytes, first diff at byte %d\n", s, kinds[s % 4], olen[s], rlen[s], k);
            bad++;
        }
    }
    printf("S=%d frames/stream=%d fpl=%d br=%d mode=%d q=%d sr=%d chunk=%d: %ld frames, %s\n", S, NF, FPL, brate, mode, quality, sr, chunk,
           frames, bad ? "MISMATCH" : "IDENTICAL");
    return bad ? 1 : 0;
}
