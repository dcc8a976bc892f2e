/* Test harness: the calls of the libmp3lame face that read or end a handle's state while frames of its lane may still be outstanding
 * (lame_encode_buffer returns with up to LAMEGPU_HANDLE_DEPTH - 1 frames not yet back, lg_api.cpp handle_wait_frames): encoding on
 * after lame_encode_flush, lame_encode_flush_nogap + lame_init_bitstream in mid-stream (lame.c:1988, :2006), lame_close without a
 * flush and the lane's next owner, lame_get_mf_samples_to_encode after every call, a flush right behind a call on a resampled stream.  Every sequence is made to the product library
 * (linked) and to the checker named by argv[1] - the reference library, loaded with dlopen - and the bytes are compared.
 * usage: handle_edges <checker.so> [chunk_samples]; exit code 0 only when every scenario is identical. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <dlfcn.h>
#include <vector>
#include <random>
#include "../../include/lamegpu.h"

typedef std::vector<unsigned char> Bytes;
struct Api {
    void *(*init)(void);
    int (*set_brate)(void *, int);
    int (*set_tag)(void *, int);
    int (*set_out_rate)(void *, int);
    int (*init_params)(void *);
    int (*encode)(void *, const short *, const short *, int, unsigned char *, int);
    int (*flush)(void *, unsigned char *, int);
    int (*flush_nogap)(void *, unsigned char *, int);
    int (*init_bitstream)(void *);
    int (*mf_left)(const void *);
    int (*close)(void *);
};
static Api product()
{
    Api a;
    a.init = (void *(*)(void)) lame_init; a.set_brate = (int (*)(void *, int)) lame_set_brate; a.set_tag = (int (*)(void *, int)) lame_set_bWriteVbrTag;
    a.set_out_rate = (int (*)(void *, int)) lame_set_out_samplerate;
    a.init_params = (int (*)(void *)) lame_init_params; a.encode = (int (*)(void *, const short *, const short *, int, unsigned char *, int)) lame_encode_buffer;
    a.flush = (int (*)(void *, unsigned char *, int)) lame_encode_flush; a.flush_nogap = (int (*)(void *, unsigned char *, int)) lame_encode_flush_nogap;
    a.init_bitstream = (int (*)(void *)) lame_init_bitstream; a.mf_left = (int (*)(const void *)) lame_get_mf_samples_to_encode; a.close = (int (*)(void *)) lame_close;
    return a;
}
static Api checker(void *h)
{
    Api a;
    a.init = (void *(*)(void)) dlsym(h, "lame_init"); a.set_brate = (int (*)(void *, int)) dlsym(h, "lame_set_brate");
    a.set_tag = (int (*)(void *, int)) dlsym(h, "lame_set_bWriteVbrTag"); a.set_out_rate = (int (*)(void *, int)) dlsym(h, "lame_set_out_samplerate"); a.init_params = (int (*)(void *)) dlsym(h, "lame_init_params");
    a.encode = (int (*)(void *, const short *, const short *, int, unsigned char *, int)) dlsym(h, "lame_encode_buffer");
    a.flush = (int (*)(void *, unsigned char *, int)) dlsym(h, "lame_encode_flush"); a.flush_nogap = (int (*)(void *, unsigned char *, int)) dlsym(h, "lame_encode_flush_nogap");
    a.init_bitstream = (int (*)(void *)) dlsym(h, "lame_init_bitstream"); a.mf_left = (int (*)(const void *)) dlsym(h, "lame_get_mf_samples_to_encode");
    a.close = (int (*)(void *)) dlsym(h, "lame_close");
    return a;
}
static std::vector<short> L, R;
static int chunk = 1152;
static const int CAP = 1 << 17;

static void *open_handle(const Api &a, int brate, int tag, int out_rate = 0)
{
    void *g = a.init();
    a.set_brate(g, brate); a.set_tag(g, tag);
    if (out_rate) a.set_out_rate(g, out_rate);
    if (a.init_params(g) != 0) { fprintf(stderr, "lame_init_params failed\n"); exit(2); }
    return g;
}
static void feed(const Api &a, void *g, int from, int to, Bytes &out, std::vector<int> *mf = nullptr)
{
    Bytes buf(CAP);
    for (int i = from; i < to; i += chunk) {
        int const k = std::min(chunk, to - i);
        int const b = a.encode(g, L.data() + i, R.data() + i, k, buf.data(), CAP);
        if (b < 0) { fprintf(stderr, "lame_encode_buffer returned %d\n", b); exit(2); }
        out.insert(out.end(), buf.begin(), buf.begin() + b);
        if (mf) mf->push_back(a.mf_left(g));
    }
}
static void take(int b, const Bytes &buf, Bytes &out) { if (b < 0) { fprintf(stderr, "flush returned %d\n", b); exit(2); } out.insert(out.end(), buf.begin(), buf.begin() + b); }

/* the four scenarios, the same calls for either library */
static Bytes encode_after_flush(const Api &a)
{
    Bytes out, buf(CAP);
    void *g = open_handle(a, 128, 0);
    feed(a, g, 0, 9 * 1152 + 300, out);
    take(a.flush(g, buf.data(), CAP), buf, out);
    feed(a, g, 9 * 1152 + 300, 20 * 1152, out);
    take(a.flush(g, buf.data(), CAP), buf, out);
    a.close(g);
    return out;
}
static Bytes nogap_and_new_bitstream(const Api &a)
{
    Bytes out, buf(CAP);
    void *g = open_handle(a, 160, 1);
    feed(a, g, 0, 8 * 1152 + 77, out);
    take(a.flush_nogap(g, buf.data(), CAP), buf, out);
    a.init_bitstream(g);
    feed(a, g, 8 * 1152 + 77, 19 * 1152, out);
    take(a.flush(g, buf.data(), CAP), buf, out);
    a.close(g);
    return out;
}
static Bytes close_without_flush_then_next_owner(const Api &a)
{
    Bytes out, waste, buf(CAP);
    void *g = open_handle(a, 128, 0);
    feed(a, g, 0, 7 * 1152, waste);            /* frames of this handle may still be outstanding when it is closed */
    a.close(g);
    g = open_handle(a, 128, 0);                /* the product hands the same lane out again: it must start like a fresh stream */
    feed(a, g, 3 * 1152, 15 * 1152, out);
    take(a.flush(g, buf.data(), CAP), buf, out);
    a.close(g);
    return out;
}
static Bytes samples_left_after_every_call(const Api &a)
{
    Bytes out, buf(CAP);
    std::vector<int> mf;
    void *g = open_handle(a, 128, 0);
    feed(a, g, 0, 12 * 1152 + 500, out, &mf);
    take(a.flush(g, buf.data(), CAP), buf, out);
    mf.push_back(a.mf_left(g));
    a.close(g);
    Bytes rec((const unsigned char *) mf.data(), (const unsigned char *) (mf.data() + mf.size()));     /* what is compared: the values, not the bytes' timing */
    rec.insert(rec.end(), out.begin(), out.end());
    return rec;
}

/* 44.1 -> 32 kHz: the flush pads in bunches sized from what the reference has encoded by then (lame.c:2077-2117); here the frames of
 * the last calls still wait for their launch when the flush begins (both slots of the engine busy with the calls before them) */
static Bytes flush_behind_resampled_calls(const Api &a)
{
    Bytes out, buf(CAP);
    void *g = open_handle(a, 96, 1, 32000);
    feed(a, g, 0, 14 * 1152 + 123, out);
    take(a.flush(g, buf.data(), CAP), buf, out);
    a.close(g);
    return out;
}

int main(int argc, char **argv)
{
    const char *so = argc > 1 ? argv[1] : "oracle/_ref/libmp3lame_ref.so";
    if (argc > 2) chunk = atoi(argv[2]);
    void *h = dlopen(so, RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
    if (!h) { fprintf(stderr, "cannot load the checker %s: %s\n", so, dlerror()); return 2; }
    Api const ours = product(), ref = checker(h);
    int const n = 20 * 1152;
    std::mt19937 rng(4242u);
    L.resize(n); R.resize(n);
    for (int i = 0; i < n; i++) {
        double const t = i / 44100.0, env = (i % 5000) < 300 ? exp(-(i % 5000) / 50.0) : 0.0;
        L[i] = (short) (6000 * sin(2 * M_PI * 440 * t) + env * ((int) (rng() % 40001u) - 20000) + (int) (rng() % 801u) - 400);
        R[i] = (short) (5000 * sin(2 * M_PI * 1234 * t) + 0.7 * env * ((int) (rng() % 40001u) - 20000) + (int) (rng() % 801u) - 400);
    }
    struct { const char *name; Bytes (*run)(const Api &); } const cases[] = {
        { "encode after flush", encode_after_flush }, { "flush_nogap + init_bitstream", nogap_and_new_bitstream },
        { "close without flush, lane reused", close_without_flush_then_next_owner }, { "mf_samples_to_encode per call", samples_left_after_every_call },
        { "flush behind resampled calls", flush_behind_resampled_calls } };
    int bad = 0;
    for (auto const &c : cases) {
        Bytes const a = c.run(ours), b = c.run(ref);
        bool const same = a == b;
        printf("%s: %s (%zu bytes against %zu)\n", c.name, same ? "identical" : "DIFFERENT", a.size(), b.size());
        bad += !same;
    }
    if (bad) { printf("DIFFERENT: %d of 5 scenarios\n", bad); return 1; }
    printf("IDENTICAL 5/5 scenarios (%d-sample calls)\n", chunk);
    return 0;
}
