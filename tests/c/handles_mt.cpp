/* Test harness: many application threads, each with its OWN lame_t, encoding concurrently through the libmp3lame face of the product
 * library (lame_init / lame_set_* / lame_init_params / lame_encode_buffer in 1152-sample calls / lame_encode_flush / lame_close) - the
 * reference's threading contract (HACKING:67-76: one handle = one thread at a time, different handles are independent).  Behind the
 * handles sits one shared batch engine whose dispatcher gathers the threads' frames into common launches (lg_api.cpp LgShared).
 * Every stream is compared byte for byte with the same calls made to the checker named by argv[5]: the reference library itself
 * (oracle/_ref/libmp3lame_ref.so, loaded with dlopen so that its lame_* symbols do not collide with the product's).
 *
 * usage: handles_mt <threads> <frames_per_stream> <chunk_samples> <brate> <checker.so> [vbr_mode vbr_q [first_stream]]
 * prints "IDENTICAL n/n streams, <frames/s aggregate>" or the first mismatch; exit code 0 only when every stream is identical. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include <vector>
#include <thread>
#include <atomic>
#include <chrono>
#include <random>
#include "../../include/lamegpu.h"

struct RefApi {
    void *(*init)(void);
    int (*set_brate)(void *, int);
    int (*set_vbr)(void *, int);
    int (*set_vbr_q)(void *, int);
    int (*set_tag)(void *, int);
    int (*init_params)(void *);
    int (*encode)(void *, const short *, const short *, int, unsigned char *, int);
    int (*flush)(void *, unsigned char *, int);
    int (*close)(void *);
};

static void make_pcm(int stream, int n, std::vector<short> &l, std::vector<short> &r)
{
    std::mt19937 rng(1000u + (unsigned) stream);
    l.resize(n); r.resize(n);
    int const kind = stream % 3;
    for (int i = 0; i < n; i++) {
        if (kind == 0) { l[i] = (short) ((int) (rng() % 24001u) - 12000); r[i] = (short) ((int) (rng() % 24001u) - 12000); }
        else if (kind == 1) {
            int const ph = (i + 977 * stream) % 7919;
            double const env = ph < 400 ? exp(-ph / 60.0) : 0.0;
            l[i] = (short) ((int) (rng() % 121u) - 60 + env * ((int) (rng() % 48001u) - 24000));
            r[i] = (short) ((int) (rng() % 121u) - 60 + 0.8 * env * ((int) (rng() % 48001u) - 24000));
        }
        else {
            double const t = (double) (i + 131 * stream) / 44100.0;
            l[i] = (short) (8000 * sin(2 * M_PI * 440 * t) + 4000 * sin(2 * M_PI * 3300 * t) + (int) (rng() % 2001u) - 1000);
            r[i] = (short) (8000 * sin(2 * M_PI * 554.37 * t) + 3000 * sin(2 * M_PI * 7000 * t) + (int) (rng() % 2001u) - 1000);
        }
    }
}

int main(int argc, char **argv)
{
    int const T = argc > 1 ? atoi(argv[1]) : 8, NF = argc > 2 ? atoi(argv[2]) : 6, chunk = argc > 3 ? atoi(argv[3]) : 1152;
    int const brate = argc > 4 ? atoi(argv[4]) : 128;
    const char *checker = argc > 5 ? argv[5] : "oracle/_ref/libmp3lame_ref.so";
    int const vbr = argc > 6 ? atoi(argv[6]) : 0, vbr_q = argc > 7 ? atoi(argv[7]) : 4;
    int const first = argc > 8 ? atoi(argv[8]) : 0;         /* thread s encodes the signal of stream first + s */
    int const n = NF * 1152, cap = n * 5 / 4 + 7200 + 65536;
    void *h = dlopen(checker, RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND)   /* DEEPBIND: the checker's own lame_* calls must stay inside it */;
    if (!h) { fprintf(stderr, "cannot load the checker %s: %s\n", checker, dlerror()); return 2; }
    RefApi R;
    R.init = (void *(*)(void)) dlsym(h, "lame_init");
    R.set_brate = (int (*)(void *, int)) dlsym(h, "lame_set_brate");
    R.set_vbr = (int (*)(void *, int)) dlsym(h, "lame_set_VBR");
    R.set_vbr_q = (int (*)(void *, int)) dlsym(h, "lame_set_VBR_q");
    R.set_tag = (int (*)(void *, int)) dlsym(h, "lame_set_bWriteVbrTag");
    R.init_params = (int (*)(void *)) dlsym(h, "lame_init_params");
    R.encode = (int (*)(void *, const short *, const short *, int, unsigned char *, int)) dlsym(h, "lame_encode_buffer");
    R.flush = (int (*)(void *, unsigned char *, int)) dlsym(h, "lame_encode_flush");
    R.close = (int (*)(void *)) dlsym(h, "lame_close");
    if (!R.init || !R.encode || !R.flush) { fprintf(stderr, "the checker lacks the libmp3lame entry points\n"); return 2; }

    std::vector<std::vector<short>> L(T), Rr(T);
    std::vector<std::vector<unsigned char>> got(T), want(T);
    for (int s = 0; s < T; s++) make_pcm(first + s, n, L[s], Rr[s]);
    /* the handles are made before the clock starts (the first one creates the engine), as an application would */
    std::vector<lame_global_flags *> g(T);
    for (int s = 0; s < T; s++) {
        g[s] = lame_init();
        lame_set_bWriteVbrTag(g[s], 0);
        if (vbr) { lame_set_VBR(g[s], (vbr_mode) vbr); lame_set_VBR_q(g[s], vbr_q); if (vbr == 3) lame_set_VBR_mean_bitrate_kbps(g[s], brate); }
        else lame_set_brate(g[s], brate);
        if (lame_init_params(g[s]) != 0) { fprintf(stderr, "lame_init_params failed for handle %d\n", s); return 2; }
    }
    std::atomic<int> failed(0);
    auto const t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int s = 0; s < T; s++)
        th.emplace_back([&, s]() {
            std::vector<unsigned char> buf(cap);
            for (int i = 0; i < n; i += chunk) {
                int const k = std::min(chunk, n - i);
                int const b = lame_encode_buffer(g[s], L[s].data() + i, Rr[s].data() + i, k, buf.data(), cap);
                if (b < 0) { failed++; return; }
                got[s].insert(got[s].end(), buf.begin(), buf.begin() + b);
            }
            int const b = lame_encode_flush(g[s], buf.data(), cap);
            if (b < 0) { failed++; return; }
            got[s].insert(got[s].end(), buf.begin(), buf.begin() + b);
        });
    for (auto &t : th) t.join();
    double const secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long frames = 0;
    for (int s = 0; s < T; s++) { frames += lame_get_frameNum(g[s]); lame_close(g[s]); }
    if (failed.load()) { printf("FAILED: %d threads got an error code\n", failed.load()); return 1; }
    if (getenv("HANDLES_MT_RATE_ONLY")) {           /* a rate measurement between checked runs: the checker costs one CPU core 0.3 ms per frame */
        printf("UNCHECKED %d streams, %ld frames in %.3f s = %.0f frames/s aggregate (%d threads x own lame_t x %d-sample calls)\n", T, frames, secs, frames / secs, T, chunk);
        return 0;
    }
    /* the same calls to the checker, stream by stream */
    int bad = 0;
    for (int s = 0; s < T; s++) {
        void *r = R.init();
        R.set_tag(r, 0);
        if (vbr) { R.set_vbr(r, vbr); R.set_vbr_q(r, vbr_q); if (vbr == 3) ((int (*)(void *, int)) dlsym(h, "lame_set_VBR_mean_bitrate_kbps"))(r, brate); }
        else R.set_brate(r, brate);
        R.init_params(r);
        std::vector<unsigned char> buf(cap);
        for (int i = 0; i < n; i += chunk) {
            int const k = std::min(chunk, n - i);
            int const b = R.encode(r, L[s].data() + i, Rr[s].data() + i, k, buf.data(), cap);
            want[s].insert(want[s].end(), buf.begin(), buf.begin() + b);
        }
        int const b = R.flush(r, buf.data(), cap);
        want[s].insert(want[s].end(), buf.begin(), buf.begin() + b);
        R.close(r);
        if (got[s] != want[s]) {
            if (!bad) printf("stream %d differs: %zu bytes against %zu\n", s, got[s].size(), want[s].size());
            bad++;
        }
    }
    if (bad) { printf("DIFFERENT: %d of %d streams\n", bad, T); return 1; }
    printf("IDENTICAL %d/%d streams, %ld frames in %.3f s = %.0f frames/s aggregate (%d threads x own lame_t x %d-sample calls)\n", T, T, frames, secs,
           frames / secs, T, chunk);
    return 0;
}
