// lg_api.cpp - host side above the device engine: per-stream PCM buffering (the reference's mfbuf logic,
// lame.c:1671-1775), the batch scheduler that turns "frames that became complete" into GPU launches, the
// multi-threaded bit packing of the results, and the two C-ABI faces declared in include/lamegpu.h.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <thread>
#include <atomic>
#include <algorithm>
#include <functional>
#include <new>
#include <chrono>
#include <math.h>
#include <mutex>
#include <condition_variable>
#include <memory>
#include <string>
#include "../../include/lamegpu.h"
#include <type_traits>
#include "lg_engine.h"
#include "lg_bitstream.h"

namespace {

enum { LG_RS_HIST = 48 };      /* input samples kept before a chunk's first one: the filter reaches filter_l/2 + 1 <= 17 back */

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static const bool g_timing = getenv("LAMEGPU_TIMING") != nullptr;

/* One stream = one lame_t of the reference.  The PCM timeline is the reference's zero-prefixed stream:
 * 528 zeros (ENCDELAY - MDCTDELAY, lame.c:2302) then the user's samples; frame k is encodable once the
 * timeline holds 1152*k + 1904 samples (calcNeeded, lame.c:1627). */
struct Stream {
    std::vector<int16_t> pcm16[2];     /* timeline samples from index tbase on */
    std::vector<float> pcmf[2];        /* same, already through pcm_transform: a stream whose calls mixed sample types (float_mode) */
    bool float_mode = false;
    /* same in the caller's own sample type (lame_encode_buffer_int / _long / _float / _ieee_double ...): raw elements of nat_esz bytes; the
     * device converts them and applies nat_scale * pcm_transform (lame.c:1797-1834) */
    std::vector<unsigned char> nat[2];
    int  nat_kind = 0, nat_esz = 0;    /* LG_PCM_S32 / F32 / S64 / F64, 0 = not in this mode */
    float nat_scale = 1.0f;
    bool fed = false;                  /* samples have arrived: the representation is settled */
    int  fs = 1152, need = 1904;       /* samples per frame (576 per granule) and what a frame needs in the buffer (calcNeeded, lame.c:1627) */
    /* input-rate conversion (util.c:531): the stream's input samples (through pcm_transform) from absolute index
     * raw_base on, the reference's per-call bookkeeping replayed as a list of chunks, the input clock, and the end
     * of the timeline the chunks cover.  The samples themselves are made on the device (kernel R). */
    struct Chunk { double itime; long in_base, out_pos; int count; };
    bool rs_mode = false;
    std::vector<unsigned char> raw[2]; /* elements of rs_esz bytes in the caller's sample type rs_kind (LG_PCM_DONE: floats transformed here) */
    int  rs_kind = LG_PCM_S16, rs_esz = 2;
    float rs_scale = 1.0f;
    long rawn() const { return (long) (raw[0].size() / (size_t) rs_esz); }
    long raw_base = -LG_RS_HIST;
    std::vector<Chunk> chunks;
    double rs_itime = 0;
    long rs_tend = 528;
    long tbase = -LG_PCM_HIST;         /* timeline index of element 0 */
    /* int16 input of the call in progress that has not been copied yet: the caller's buffers cover the timeline from the end of
     * pcm16[] on; frames are staged straight from them and only what later frames still need is kept (unborrow) */
    const short *bl = nullptr, *br = nullptr;
    long bn = 0;
    long frames_done = 0;              /* frames handed to the device */
    long frames_out = 0;               /* frames whose bytes have been spliced into `out` */
    long frames_out_base = 0;          /* ... when the current file began (lame_init_bitstream) */
    long mf_samples_to_encode = 576 + 1152;   /* ENCDELAY + POSTDELAY, lame.c:2299 */
    int  last_padding = 0, last_bitrate_index = 0;
    LgBitWriter bw;
    /* encoder.c:156 updateStats: frames per bitrate index x mode extension (column 4 = all), gr.ch per bitrate index x block
     * type (4 = mixed, column 5 = all); row 15 = totals */
    int  hist_mode[16][5], hist_block[16][6];
    std::vector<unsigned char> out;    /* packed bytes not yet handed to the caller */
    /* Info tag bookkeeping of one lame_t (VbrTag.c): the seek table of AddVbrFrame (:196), the byte count and music
     * CRC of copy_buffer (bitstream.c:1079-1090), what lame_get_lametag_frame (:900) needs at the end */
    struct Tag {
        bool on = false;
        int  frame_size = 0;
        long nframes = 0, nbytes = 0, sum = 0;
        int  seen = 0, want = 1, pos = 0;
        int  bag[400];
        unsigned short music_crc = 0;
        long pending = 0;              /* placeholder bytes still in `out`: handed over without CRC (copy_buffer mp3data = 0) */
        int  mode_ext = 0, enc_padding = 0;
    } tag;

    void init()
    {
        for (int c = 0; c < 2; c++) { pcm16[c].assign(LG_PCM_HIST + 528, 0); pcmf[c].clear(); }
        float_mode = false; nat[0].clear(); nat[1].clear(); nat_kind = nat_esz = 0; nat_scale = 1.0f; fed = false;
        tbase = -LG_PCM_HIST; frames_done = 0; frames_out = 0; mf_samples_to_encode = 576 + 1152; last_padding = 0;
        bl = br = nullptr; bn = 0;
        rs_kind = LG_PCM_S16; rs_esz = 2; rs_scale = 1.0f;
        for (int c = 0; c < 2; c++) raw[c].assign((size_t) LG_RS_HIST * rs_esz, 0);
        raw_base = -LG_RS_HIST; chunks.clear(); rs_itime = 0; rs_tend = 528;
        bw.reset(); out.clear();
        memset(hist_mode, 0, sizeof hist_mode); memset(hist_block, 0, sizeof hist_block);
        tag = Tag();
    }
    long held() const { return (long) (float_mode ? pcmf[0].size() : nat_kind ? nat[0].size() / (size_t) nat_esz : pcm16[0].size()); }
    long tend() const { return rs_mode ? rs_tend : tbase + held() + bn; }
    /* timeline samples [from, from + n) of channel c as int16 into dst: from the kept samples, then from the borrowed input */
    void copy16(int c, long from, size_t n, int16_t *dst) const
    {
        long const have = (long) pcm16[c].size(), off = from - tbase;
        size_t k = 0;
        if (off < have) { k = std::min<size_t>(n, (size_t) (have - off)); memcpy(dst, pcm16[c].data() + off, k * sizeof(int16_t)); }
        if (k < n) memcpy(dst + k, (c ? br : bl) + (off + (long) k - have), (n - k) * sizeof(int16_t));
    }
    void unborrow()
    {
        if (!bn) return;
        for (int c = 0; c < 2; c++) pcm16[c].insert(pcm16[c].end(), c ? br : bl, (c ? br : bl) + bn);
        bl = br = nullptr; bn = 0;
    }
    long frames_ready() const
    {
        long const have = tend();
        long const k = frames_done;
        if (have < (long) fs * k + need) return 0;
        return (have - need - (long) fs * k) / fs + 1;
    }
    template <class T> void nat_to_float(const LgDevCfg *cfg)
    {
        size_t const n = nat[0].size() / sizeof(T);
        float const m00 = nat_scale * cfg->pcm_transform[0][0], m01 = nat_scale * cfg->pcm_transform[0][1];
        float const m10 = nat_scale * cfg->pcm_transform[1][0], m11 = nat_scale * cfg->pcm_transform[1][1];
        const T *a = (const T *) nat[0].data(), *b = (const T *) nat[1].data();
        pcmf[0].resize(n); pcmf[1].resize(n);
        for (size_t i = 0; i < n; i++) {
            float const xl = (float) a[i], xr = (float) b[i];
            pcmf[0][i] = xl * m00 + xr * m01;
            pcmf[1][i] = xl * m10 + xr * m11;
        }
    }
    /* a stream whose calls mix sample types (or normalisations) falls back to floats converted on the host, as lame.c:1803-1834 does it */
    void to_float(const LgDevCfg *cfg)
    {
        if (float_mode) return;
        if (nat_kind) {
            switch (nat_kind) {
            case LG_PCM_S32: nat_to_float<int32_t>(cfg); break;
            case LG_PCM_F32: nat_to_float<float>(cfg); break;
            case LG_PCM_S64: nat_to_float<long long>(cfg); break;
            default:         nat_to_float<double>(cfg); break;
            }
            nat[0].clear(); nat[1].clear(); nat_kind = nat_esz = 0;
            float_mode = true;
            return;
        }
        unborrow();
        size_t const n = pcm16[0].size();
        float const m00 = cfg->pcm_transform[0][0], m01 = cfg->pcm_transform[0][1];
        float const m10 = cfg->pcm_transform[1][0], m11 = cfg->pcm_transform[1][1];
        pcmf[0].resize(n); pcmf[1].resize(n);
        for (size_t i = 0; i < n; i++) {
            float const xl = pcm16[0][i], xr = pcm16[1][i];
            pcmf[0][i] = xl * m00 + xr * m01;
            pcmf[1][i] = xl * m10 + xr * m11;
        }
        pcm16[0].clear(); pcm16[1].clear();
        float_mode = true;
    }
    /* a fresh stream (only the zero prefix, which is zero in every type) takes the sample type of its first input */
    void to_native(int kind, int esz, float scale)
    {
        size_t const n = pcm16[0].size();
        for (int c = 0; c < 2; c++) { nat[c].assign(n * (size_t) esz, 0); pcm16[c].clear(); }
        nat_kind = kind; nat_esz = esz; nat_scale = scale;
    }
    void drop_consumed()
    {
        long const keep_from = (long) fs * frames_done - LG_PCM_HIST;
        if (rs_mode) {
            /* chunks that end before the next window, and the input samples only they needed */
            size_t n = 0;
            while (n < chunks.size() && chunks[n].out_pos + chunks[n].count <= keep_from) n++;
            chunks.erase(chunks.begin(), chunks.begin() + n);
            long const from = (chunks.empty() ? raw_base + rawn() : chunks[0].in_base) - LG_RS_HIST;
            if (from > raw_base) {
                for (int c = 0; c < 2; c++) raw[c].erase(raw[c].begin(), raw[c].begin() + (from - raw_base) * rs_esz);
                raw_base = from;
            }
            return;
        }
        long const d = keep_from - tbase;
        if (d <= 0) return;
        if (float_mode) for (int c = 0; c < 2; c++) pcmf[c].erase(pcmf[c].begin(), pcmf[c].begin() + d);
        else if (nat_kind) for (int c = 0; c < 2; c++) nat[c].erase(nat[c].begin(), nat[c].begin() + d * nat_esz);
        else {
            long const have = (long) pcm16[0].size();
            if (d <= have) for (int c = 0; c < 2; c++) pcm16[c].erase(pcm16[c].begin(), pcm16[c].begin() + d);
            else {                               /* the kept samples are used up: the rest of the consumed range lies in the borrowed input */
                long const skip = std::min<long>(d - have, bn);
                for (int c = 0; c < 2; c++) pcm16[c].clear();
                bl += skip; br += skip; bn -= skip;
            }
        }
        tbase = keep_from;
    }
};

/* VbrTag.c:576 CRC_update_lookup: CRC-16, polynomial x^16 + x^15 + x^2 + 1 (reflected 0xA001) */
static unsigned short crc16_update(unsigned short value, unsigned short crc)
{
    static unsigned short table[256];
    static bool ready = false;
    if (!ready) {
        for (int i = 0; i < 256; i++) {
            unsigned short c = (unsigned short) i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? (unsigned short) ((c >> 1) ^ 0xA001) : (unsigned short) (c >> 1);
            table[i] = c;
        }
        ready = true;
    }
    unsigned short const tmp = crc ^ value;
    return (unsigned short) ((crc >> 8) ^ table[tmp & 0xff]);
}

/* VbrTag.c:123 addVbr */
static void tag_add_frame(Stream::Tag &v, int bitrate)
{
    v.nframes++;
    v.sum += bitrate;
    v.seen++;
    if (v.seen < v.want) return;
    if (v.pos < 400) { v.bag[v.pos] = (int) v.sum; v.pos++; v.seen = 0; }
    if (v.pos == 400) {
        for (int i = 1; i < 400; i += 2) v.bag[i / 2] = v.bag[i];
        v.want *= 2;
        v.pos /= 2;
    }
}

/* Persistent worker pool: the per-call host work (PCM staging, bit packing, output hand-over) is a few hundred
 * microseconds per thread, so spawning threads per call would cost as much as the work itself. */
class Pool {
public:
    explicit Pool(int nthreads) : n_(std::max(1, nthreads))
    {
        for (int t = 1; t < n_; t++) th_.emplace_back([this]() { worker(); });
    }
    ~Pool()
    {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return n_; }
    void run(int n, const std::function<void(int)> &fn)
    {
        if (n_ <= 1 || n <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn; total_ = n; next_.store(0); active_ = (int) th_.size(); gen_++;
        }
        cv_.notify_all();
        drain();
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this]() { return active_ == 0; });
        fn_ = nullptr;
    }
private:
    void drain() { for (;;) { int const i = next_.fetch_add(1); if (i >= total_) break; (*fn_)(i); } }
    void worker()
    {
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&]() { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            drain();
            { std::lock_guard<std::mutex> g(m_); if (--active_ == 0) done_.notify_one(); }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    int total_ = 0, active_ = 0;
    unsigned long gen_ = 0;
    bool stop_ = false;
};

} // namespace

struct lamegpu_batch {
    /* device = -1 with several GPUs visible: the batch is a set of parts, one engine per device, each with a contiguous share of the
     * streams; every call fans out to the parts on one host thread per device (streams are independent: no exchange between parts) */
    std::vector<lamegpu_batch *> parts;
    std::vector<int> part_first;
    template <class Fn> long fan_out(const Fn &fn)      /* fn(part index) -> long; returns the sum, or the first negative result */
    {
        std::vector<long> r(parts.size(), 0);
        std::vector<std::thread> th;
        for (size_t p = 1; p < parts.size(); p++) th.emplace_back([&, p]() { r[p] = fn((int) p); });
        r[0] = fn(0);
        for (auto &t : th) t.join();
        long sum = 0;
        for (long v : r) { if (v < 0) return v; sum += v; }
        return sum;
    }
    LgDevCfg cfg;
    lg_engine *eng = nullptr;
    int S = 0, F = 0, nthreads = 1;
    std::unique_ptr<Pool> pool;
    void parallel_for(int n, const std::function<void(int)> &fn)
    {
        if (n <= 1 || nthreads <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
        if (!pool || pool->size() != nthreads) pool.reset(new Pool(nthreads));
        pool->run(n, fn);
    }
    std::vector<Stream> st;
    long frames_total = 0;

    /* ---- the step pipeline.  A step = the frames of all streams that are complete, at most F per stream: staged into one of the
     * engine's two slots, submitted (asynchronous), completed later (wait + splice of the packed frames into the streams' output).
     * Synchronous mode completes a step right away.  Pipelined mode (lamegpu_batch_set_pipelined) leaves the newest step in flight
     * when a call returns: its bytes come out of the next call (or the flush), and the host work of a call - staging the next step,
     * splicing the previous one - runs while the device works. */
    int next_slot = 0;
    bool pipelined = false;
    /* the handles' engine (LgShared): a launch that gathers crowd_lanes lanes or more takes at most crowd_cap frames of each.  A
     * launch lasts as long as its longest lane (kernels B and D walk a stream's frames one after the other), so lanes with one frame
     * next to lanes with four leave the device three quarters idle; the frames held back go into the next launch, which is staged
     * while this one runs.  0 = off (the batch calls: every stream brings the same number of frames). */
    int crowd_cap = 0, crowd_lanes = 0;
    double acc_ms[5] = { 0, 0, 0, 0, 0 };   /* per-kernel device times summed over the steps of the last lamegpu_batch_run_device_steps */
    int acc_n = 0;
    bool in_flight[2] = { false, false };
    std::vector<int> flight_nfr[2];

    /* wait for slot k's step and splice its frames into the streams */
    int complete(int k)
    {
        if (!in_flight[k]) return 0;
        double const t0 = now_ms();
        if (lg_engine_wait(eng, k) != 0) return -2;
        double const t1 = now_ms();
        in_flight[k] = false;
        const int *nfr = flight_nfr[k].data();
        const LgFrameOut *fo = lg_engine_host_fout(eng, k);
        const unsigned char *pay = lg_engine_host_pay(eng, k), *hdr = lg_engine_host_hdr(eng, k);
        size_t const pay_stride = lg_engine_pay_stride(eng);
        parallel_for(S, [&](int s) {
            if (!nfr[s]) return;
            Stream &x = st[s];
            for (int f = 0; f < nfr[s]; f++) {
                const LgFrameOut *fr = fo + (size_t) s * F + f;
                lg_merge_frame(&x.bw, &cfg, fr, hdr + ((size_t) s * F + f) * LG_HDR_STRIDE, pay + (size_t) s * pay_stride + fr->pay_off);
                x.last_padding = fr->padding;
                x.last_bitrate_index = fr->bitrate_index;
                {
                    int const bi = fr->bitrate_index & 15;
                    const unsigned char *bt = hdr + ((size_t) s * F + f) * LG_HDR_STRIDE + 40;
                    x.hist_mode[bi][4]++; x.hist_mode[15][4]++;
                    if (cfg.channels == 2) { x.hist_mode[bi][fr->mode_ext & 3]++; x.hist_mode[15][fr->mode_ext & 3]++; }
                    for (int q = 0; q < 4; q++)
                        if (bt[q] < 5) { x.hist_block[bi][bt[q]]++; x.hist_block[bi][5]++; x.hist_block[15][bt[q]]++; x.hist_block[15][5]++; }
                }
                if (x.tag.on) { tag_add_frame(x.tag, cfg.bitrate_kbps[fr->bitrate_index]); x.tag.mode_ext = fr->mode_ext; }
            }
            x.out.insert(x.out.end(), x.bw.buf.begin(), x.bw.buf.end());
            x.bw.buf.clear();
            x.frames_out += nfr[s];
        });
        if (g_timing) fprintf(stderr, "lamegpu: slot %d: waited %.2f ms for the device, splice %.2f ms\n", k, t1 - t0, now_ms() - t1);
        return 0;
    }
    /* complete whatever is in flight, oldest first */
    int drain()
    {
        if (complete(next_slot) != 0) return -2;
        return complete(next_slot ^ 1);
    }

    /* One step: stage every complete frame of every stream (at most F per stream) into the next slot and submit it.  Returns the slot,
     * -1 when no stream has a complete frame, -2 on error; *frames gets the number of frames submitted.  The slot's previous step
     * (two steps back) is completed first if it still is in flight. */
    int stage_and_submit(long *frames)
    {
        long submitted = 0;
        int const k = next_slot;
        int maxf = 0, any_float = 0, lanes_ready = 0;
        for (int s = 0; s < S; s++) {
            long const r = st[s].frames_ready();
            if (r > 0) { lanes_ready++; maxf = std::max<int>(maxf, (int) std::min<long>(r, F)); if (st[s].float_mode || st[s].nat_kind) any_float = 1; }
        }
        if (maxf == 0) return -1;
        int const cap = (crowd_cap > 0 && lanes_ready >= crowd_lanes) ? std::min(crowd_cap, F) : F;
        maxf = std::min(maxf, cap);
        if (complete(k) != 0) return -2;                 /* the slot's previous step (two steps back) */
        int *nfr = lg_engine_host_nfr(eng, k);
        for (int s = 0; s < S; s++) nfr[s] = (int) std::max<long>(0, std::min<long>(st[s].frames_ready(), cap));
        size_t const stride = lg_engine_pcm_stride(eng);
        double const t0 = now_ms();
        if (cfg.resample) {
            /* kernel R makes the window: stage the input samples and the chunks that overlap it */
            int const reach = cfg.rs_filter_l - cfg.rs_filter_l / 2;
            std::atomic<int> most(0), bad(0);
            auto window = [&](int s, long &t0, long &t1, size_t &nck) {
                const Stream &x = st[s];
                t0 = (long) x.fs * x.frames_done - LG_PCM_HIST; t1 = t0 + (long) nfr[s] * x.fs + LG_PCM_HALO;
                nck = 0;
                while (nck < x.chunks.size() && x.chunks[nck].out_pos < t1) nck++;
            };
            for (int s = 0; s < S; s++) {
                if (!nfr[s]) continue;
                long t0, t1; size_t nck;
                window(s, t0, t1, nck);
                if ((int) nck > most.load()) most.store((int) nck);
            }
            if (lg_engine_reserve_chunks(eng, k, most.load()) != 0) return -2;
            size_t const raw_stride = lg_engine_raw_stride(eng);
            int const cap = lg_engine_chunk_cap(eng, k);
            int esz = 4;
            for (int s = 0; s < S; s++) if (nfr[s] && st[s].rs_esz > esz) esz = st[s].rs_esz;
            if (lg_engine_need_raw(eng, esz) != 0) return -2;
            esz = lg_engine_raw_esz(eng);
            char *hr = (char *) lg_engine_host_raw(eng, k);
            LgPcmKind *hk = lg_engine_host_kinds(eng, k);
            LgRsChunk *hc = lg_engine_host_chunks(eng, k);
            int *hn = lg_engine_host_rs_counts(eng, k);
            parallel_for(S, [&](int s) {
                hn[2 * s] = hn[2 * s + 1] = 0;
                hk[s].kind = LG_PCM_S16; hk[s].scale = 1.0f;
                if (!nfr[s]) return;
                const Stream &x = st[s];
                hk[s].kind = x.rs_kind; hk[s].scale = x.rs_scale;
                long t0, t1; size_t nck;
                window(s, t0, t1, nck);
                long hi = 0;
                for (size_t i = 0; i < nck; i++) {
                    const Stream::Chunk &c = x.chunks[i];
                    LgRsChunk &o = hc[(size_t) s * cap + i];
                    o.itime = c.itime; o.in_base = c.in_base - x.raw_base; o.out_pos = (int32_t) (c.out_pos - t0); o.count = c.count;
                    long const klast = std::min<long>(c.count, t1 - c.out_pos) - 1;
                    hi = std::max(hi, c.in_base + (long) floor((double) klast * cfg.rs_ratio - c.itime) + reach + 1 - x.raw_base);
                }
                if (hi > (long) raw_stride || hi > x.rawn()) { bad.store(1); return; }
                for (int c = 0; c < 2; c++) memcpy(hr + ((size_t) s * 2 + c) * raw_stride * (size_t) esz, x.raw[c].data(), (size_t) hi * x.rs_esz);
                hn[2 * s] = (int) nck; hn[2 * s + 1] = (int) (t1 - t0);
            });
            if (bad.load()) { fprintf(stderr, "lamegpu: resampler staging overflow\n"); return -2; }
            any_float = 1;
        }
        else if (any_float) {
            /* some stream is not int16: every stream's row goes out in the stream's own sample type, kernel A converts */
            int esz = 4;
            for (int s = 0; s < S; s++) if (nfr[s] && st[s].nat_esz > esz) esz = st[s].nat_esz;
            if (lg_engine_need_native_pcm(eng, esz) != 0) return -2;
            esz = lg_engine_native_esz(eng);
            char *hp = (char *) lg_engine_host_pcmn(eng, k);
            LgPcmKind *hk = lg_engine_host_kinds(eng, k);
            parallel_for(S, [&](int s) {
                hk[s].kind = LG_PCM_S16; hk[s].scale = 1.0f;
                if (!nfr[s]) return;
                const Stream &x = st[s];
                size_t const n = (size_t) nfr[s] * x.fs + LG_PCM_HALO;
                for (int c = 0; c < 2; c++) {
                    char *row = hp + ((size_t) s * 2 + c) * stride * (size_t) esz;
                    if (x.float_mode) memcpy(row, x.pcmf[c].data(), n * sizeof(float));
                    else if (x.nat_kind) memcpy(row, x.nat[c].data(), n * (size_t) x.nat_esz);
                    else x.copy16(c, x.tbase, n, (int16_t *) row);
                }
                if (x.float_mode) hk[s].kind = LG_PCM_DONE;
                else if (x.nat_kind) { hk[s].kind = x.nat_kind; hk[s].scale = x.nat_scale; }
            });
            any_float = 2;
        }
        else {
            int16_t *hp = lg_engine_host_pcm16(eng, k);
            parallel_for(S, [&](int s) {
                if (!nfr[s]) return;
                const Stream &x = st[s];
                size_t const n = (size_t) nfr[s] * x.fs + LG_PCM_HALO;
                for (int c = 0; c < 2; c++) x.copy16(c, x.tbase, n, hp + ((size_t) s * 2 + c) * stride);
            });
        }
        double const t1 = now_ms();
        /* a step that shares no stream with the step still in flight (the lanes of lame_t handles: a lane is not gathered again before its
         * frames are back) need not queue its quantiser behind that step's */
        bool independent = in_flight[k ^ 1];
        if (independent) {
            const std::vector<int> &other = flight_nfr[k ^ 1];
            for (int s = 0; s < S && independent; s++) if (nfr[s] && other[s]) independent = false;
        }
        if ((independent ? lg_engine_submit_independent(eng, k, maxf, any_float) : lg_engine_submit(eng, k, maxf, any_float)) != 0) return -2;
        in_flight[k] = true;
        flight_nfr[k].assign(nfr, nfr + S);
        next_slot = k ^ 1;
        /* the streams move on at submission: the next step can be staged while this one runs */
        parallel_for(S, [&](int s) {
            if (!nfr[s]) return;
            Stream &x = st[s];
            x.frames_done += nfr[s];
            x.mf_samples_to_encode -= (long) x.fs * nfr[s];
            x.drop_consumed();
        });
        for (int s = 0; s < S; s++) submitted += nfr[s];
        if (g_timing) fprintf(stderr, "lamegpu: stage %.2f ms, submit + bookkeeping %.2f ms\n", t1 - t0, now_ms() - t1);
        *frames = submitted;
        return k;
    }

    /* stage and submit every complete frame of every stream; returns frames submitted (their bytes are in the streams' `out` on return
     * unless the batch is pipelined) */
    long pump()
    {
        long done = 0;
        for (;;) {
            long n = 0;
            int const k = stage_and_submit(&n);
            if (k == -1) break;
            if (k < 0) return -2;
            done += n;
            if (complete(k ^ 1) != 0) return -2;             /* the step before this one, while this one runs */
            if (!pipelined && complete(k) != 0) return -2;
        }
        frames_total += done;
        return done;
    }

    /* util.c:531 fill_buffer_resample as lame_encode_buffer_sample_t (lame.c:1671) drives it, bookkeeping only: the `n`
     * input samples of one call are consumed in chunks of at most 1152 output samples.  Output sample k of a chunk
     * sits at input time k*ratio - itime and needs the inputs up to floor(that) + filter_l - filter_l/2, so a chunk
     * ends at the first k whose window passes the end of the call's data (found by bisection: that index grows with k). */
    void rs_schedule(Stream &x, int n)
    {
        double const ratio = cfg.rs_ratio;
        int const filter_l = cfg.rs_filter_l, reach = filter_l - filter_l / 2;
        long in_ptr = x.raw_base + x.rawn() - n;
        int remaining = n;
        while (remaining > 0) {
            double const itime = x.rs_itime;
            auto jof = [&](int k) { return (int) floor((double) k * ratio - itime); };
            int const fs = x.fs;                           /* a chunk is at most one frame of output */
            int lo = 0, hi = fs;                           /* first k in [0, fs) with reach + j(k) >= remaining, else fs */
            while (lo < hi) {
                int const mid = (lo + hi) >> 1;
                if (reach + jof(mid) >= remaining) hi = mid; else lo = mid + 1;
            }
            int const count = lo;
            int const j = jof(count < fs ? count : fs - 1);
            int const used = std::min(remaining, reach + j);
            if (count > 0) x.chunks.push_back(Stream::Chunk{ itime, in_ptr, x.rs_tend, count });
            x.rs_itime += (double) used - (double) count * ratio;
            in_ptr += used; remaining -= used;
            x.rs_tend += count;
            if (x.mf_samples_to_encode < 1) x.mf_samples_to_encode = 576 + 1152;     /* lame.c:1735 */
            x.mf_samples_to_encode += count;
        }
    }
    /* the resampler's input keeps the caller's sample type too (kernel R converts); a stream that mixes types falls back to floats
     * transformed here */
    template <class T> static void raw_to_done(Stream &x, const LgDevCfg &cfg)
    {
        size_t const n = (size_t) x.rawn();
        float const m00 = x.rs_scale * cfg.pcm_transform[0][0], m01 = x.rs_scale * cfg.pcm_transform[0][1];
        float const m10 = x.rs_scale * cfg.pcm_transform[1][0], m11 = x.rs_scale * cfg.pcm_transform[1][1];
        std::vector<unsigned char> o[2];
        o[0].resize(n * sizeof(float)); o[1].resize(n * sizeof(float));
        const T *a = (const T *) x.raw[0].data(), *b = (const T *) x.raw[1].data();
        for (size_t i = 0; i < n; i++) {
            float const xl = (float) a[i], xr = (float) b[i];
            ((float *) o[0].data())[i] = xl * m00 + xr * m01;
            ((float *) o[1].data())[i] = xl * m10 + xr * m11;
        }
        x.raw[0].swap(o[0]); x.raw[1].swap(o[1]);
        x.rs_kind = LG_PCM_DONE; x.rs_esz = 4; x.rs_scale = 1.0f;
    }
    template <class T> void feed_rs(Stream &x, const T *l, const T *r, int n, int jump, float scale)
    {
        int const kind = std::is_same<T, float>::value ? LG_PCM_F32 : std::is_same<T, double>::value ? LG_PCM_F64
                       : sizeof(T) == 8 ? LG_PCM_S64 : sizeof(T) == 4 ? LG_PCM_S32 : LG_PCM_S16;
        if (!x.fed) {                                      /* only zeros so far: they are zero in every type */
            size_t const have = (size_t) x.rawn();
            x.rs_kind = kind; x.rs_esz = (int) sizeof(T); x.rs_scale = scale;
            for (int c = 0; c < 2; c++) x.raw[c].assign(have * sizeof(T), 0);
            x.fed = true;
        }
        if (x.rs_kind == kind && x.rs_scale == scale) {
            size_t const at = x.raw[0].size();
            for (int c = 0; c < 2; c++) {
                x.raw[c].resize(at + (size_t) n * sizeof(T));
                T *d = (T *) (x.raw[c].data() + at);
                const T *src = c ? r : l;
                if (jump == 1) memcpy(d, src, (size_t) n * sizeof(T));
                else for (int i = 0; i < n; i++) d[i] = src[(size_t) i * jump];
            }
        }
        else {
            switch (x.rs_kind) {
            case LG_PCM_S16: raw_to_done<int16_t>(x, cfg); break;
            case LG_PCM_S32: raw_to_done<int32_t>(x, cfg); break;
            case LG_PCM_F32: raw_to_done<float>(x, cfg); break;
            case LG_PCM_S64: raw_to_done<long long>(x, cfg); break;
            case LG_PCM_F64: raw_to_done<double>(x, cfg); break;
            default: break;
            }
            float const m00 = scale * cfg.pcm_transform[0][0], m01 = scale * cfg.pcm_transform[0][1];
            float const m10 = scale * cfg.pcm_transform[1][0], m11 = scale * cfg.pcm_transform[1][1];
            size_t const at = x.raw[0].size();
            x.raw[0].resize(at + (size_t) n * sizeof(float)); x.raw[1].resize(at + (size_t) n * sizeof(float));
            float *d0 = (float *) (x.raw[0].data() + at), *d1 = (float *) (x.raw[1].data() + at);
            for (int i = 0; i < n; i++) {
                float const xl = (float) l[(size_t) i * jump], xr = (float) r[(size_t) i * jump];
                d0[i] = xl * m00 + xr * m01;
                d1[i] = xl * m10 + xr * m11;
            }
        }
        rs_schedule(x, n);
    }
    void feed16(int s, const short *l, const short *r, int n, bool borrow = true)
    {
        Stream &x = st[s];
        if (n <= 0) return;
        if (!r) r = l;
        if (x.rs_mode) { feed_rs<short>(x, l, r, n, 1, 1.0f); return; }
        if (x.nat_kind) x.to_float(&cfg);                                      /* int16 after another sample type: mixed */
        x.fed = true;
        if (x.float_mode) {
            float const m00 = cfg.pcm_transform[0][0], m01 = cfg.pcm_transform[0][1];
            float const m10 = cfg.pcm_transform[1][0], m11 = cfg.pcm_transform[1][1];
            for (int i = 0; i < n; i++) {
                float const xl = l[i], xr = r[i];
                x.pcmf[0].push_back(xl * m00 + xr * m01);
                x.pcmf[1].push_back(xl * m10 + xr * m11);
            }
        }
        else {
            /* not copied here: pump() stages the frames this input completes straight from the caller's buffers, end_call() keeps the rest */
            x.unborrow();
            x.bl = l; x.br = r; x.bn = n;
            if (!borrow) x.unborrow();
        }
        if (x.mf_samples_to_encode < 1) x.mf_samples_to_encode = 576 + 1152;     /* lame.c:1735 */
        x.mf_samples_to_encode += n;
    }
    /* before an API call returns: the caller's buffers are theirs again */
    void end_call() { parallel_for(S, [&](int s) { st[s].unborrow(); }); }
    /* lame.c:1786 lame_copy_inbuffer for the sample types other than int16: T = element type, jump = 1 or 2 (interleaved),
     * scale = the entry point's normalisation factor.  Same operation order as COPY_AND_TRANSFORM: the sample is converted
     * to float first, the factor is folded into the 2x2 matrix in float. */
    template <class T> void feedT(int s, const T *l, const T *r, int n, int jump, float scale)
    {
        Stream &x = st[s];
        if (n <= 0) return;
        if (!r) r = l;
        if (x.rs_mode) { feed_rs<T>(x, l, r, n, jump, scale); return; }
        /* the stream keeps the caller's sample type and the device converts (kernel A, phase 1) - as long as all of a stream's calls bring
         * the same type and normalisation; otherwise it falls back to floats converted here */
        int const kind = std::is_same<T, float>::value ? LG_PCM_F32 : std::is_same<T, double>::value ? LG_PCM_F64 : sizeof(T) == 8 ? LG_PCM_S64 : LG_PCM_S32;
        if (!x.fed && !x.float_mode && !x.nat_kind) { x.unborrow(); x.to_native(kind, (int) sizeof(T), scale); }
        x.fed = true;
        if (x.nat_kind == kind && x.nat_scale == scale && x.nat_esz == (int) sizeof(T)) {
            size_t const at = x.nat[0].size();
            for (int c = 0; c < 2; c++) {
                x.nat[c].resize(at + (size_t) n * sizeof(T));
                T *d = (T *) (x.nat[c].data() + at);
                const T *src = c ? r : l;
                if (jump == 1) memcpy(d, src, (size_t) n * sizeof(T));
                else for (int i = 0; i < n; i++) d[i] = src[(size_t) i * jump];
            }
        }
        else {
            x.to_float(&cfg);
            float const m00 = scale * cfg.pcm_transform[0][0], m01 = scale * cfg.pcm_transform[0][1];
            float const m10 = scale * cfg.pcm_transform[1][0], m11 = scale * cfg.pcm_transform[1][1];
            size_t const at = x.pcmf[0].size();
            x.pcmf[0].resize(at + n); x.pcmf[1].resize(at + n);
            for (int i = 0; i < n; i++) {
                float const xl = (float) l[(size_t) i * jump], xr = (float) r[(size_t) i * jump];
                x.pcmf[0][at + i] = xl * m00 + xr * m01;
                x.pcmf[1][at + i] = xl * m10 + xr * m11;
            }
        }
        if (x.mf_samples_to_encode < 1) x.mf_samples_to_encode = 576 + 1152;
        x.mf_samples_to_encode += n;
    }
    void feedf(int s, const float *l, const float *r, int n, float scale) { feedT<float>(s, l, r, n, 1, scale); }
    /* lame.c:2042 lame_encode_flush: how many zero samples stream s still needs */
    void pad_for_flush(int s)
    {
        Stream &x = st[s];
        if (x.mf_samples_to_encode < 1) return;
        if (x.rs_mode) {
            /* lame.c:2077-2117 with resampling: zero samples go in, in bunches sized from the fill of the frame buffer,
             * until the frames counted at the start have come out.  The reference has encoded every complete frame by now; here
             * some may still wait for their launch (a handle's lane with calls outstanding): they count as done, the buffer fill
             * is what the timeline holds beyond them. */
            int const fs = x.fs;
            long const waiting = x.frames_ready();
            int samples_to_encode = (int) (x.mf_samples_to_encode - (long) fs * waiting - 1152);
            samples_to_encode += 16. / cfg.rs_ratio;
            int end_padding = fs - (samples_to_encode % fs);
            if (end_padding < 576) end_padding += fs;
            x.tag.enc_padding = end_padding;
            int frames_left = (samples_to_encode + end_padding) / fs;
            long virt_done = x.frames_done + waiting;       /* frames the reference would have encoded so far */
            while (frames_left > 0) {
                int bunch = (int) (x.need - (x.rs_tend - (long) fs * virt_done));
                bunch *= cfg.rs_ratio;
                if (bunch > 1152) bunch = 1152;
                if (bunch < 1) bunch = 1;
                for (int c = 0; c < 2; c++) x.raw[c].insert(x.raw[c].end(), (size_t) bunch * x.rs_esz, (unsigned char) 0);
                rs_schedule(x, bunch);
                long const ready = x.rs_tend >= (long) fs * virt_done + x.need ? (x.rs_tend - x.need - (long) fs * virt_done) / fs + 1 : 0;
                if (ready > 0) frames_left -= 1;
                virt_done += ready;
            }
            return;
        }
        long const samples_to_encode = x.mf_samples_to_encode - 1152;
        long const fs = x.fs;
        long end_padding = fs - (samples_to_encode % fs);
        if (end_padding < 576) end_padding += fs;
        x.tag.enc_padding = (int) end_padding;                /* lame.c:2091 */
        long const frames_left = (samples_to_encode + end_padding) / fs;
        long const last = x.frames_done + frames_left - 1;
        long const need = fs * last + x.need - x.tend();
        if (need > 0) {
            if (x.float_mode) for (int c = 0; c < 2; c++) x.pcmf[c].insert(x.pcmf[c].end(), (size_t) need, 0.f);
            else if (x.nat_kind) for (int c = 0; c < 2; c++) x.nat[c].insert(x.nat[c].end(), (size_t) need * x.nat_esz, (unsigned char) 0);
            else for (int c = 0; c < 2; c++) x.pcm16[c].insert(x.pcm16[c].end(), (size_t) need, (int16_t) 0);
        }
    }
    /* lame.c:2134 flush_bitstream for the streams marked in live[] (all their frames have been spliced): the rest of the last frame
     * goes out as ancillary data, and the bit reservoir ends there - ResvSize = 0, main_data_begin = 0 (bitstream.c:886-889), on the
     * device too, so that a stream that is fed again after a flush continues as the reference does */
    int finish_streams(const char *live)
    {
        std::vector<int> idx, anc;
        for (int s = 0; s < S; s++) {
            if (!live[s]) continue;
            Stream &x = st[s];
            x.mf_samples_to_encode = 0;
            if (lg_pack_flush(&x.bw, &cfg, x.last_bitrate_index, x.last_padding)) { idx.push_back(s); anc.push_back(x.bw.ancillary_flag); }
            x.out.insert(x.out.end(), x.bw.buf.begin(), x.bw.buf.end());
            x.bw.buf.clear();
        }
        if (idx.empty()) return 0;
        return lg_engine_end_reservoir(eng, idx.data(), anc.data(), (int) idx.size());
    }
    /* the same for one lane of a shared engine */
    int finish_lane(int s, bool end_of_input = true)
    {
        Stream &x = st[s];
        if (end_of_input) x.mf_samples_to_encode = 0;
        int const drained = lg_pack_flush(&x.bw, &cfg, x.last_bitrate_index, x.last_padding);
        x.out.insert(x.out.end(), x.bw.buf.begin(), x.bw.buf.end());
        x.bw.buf.clear();
        if (!drained) return 0;
        int const anc = x.bw.ancillary_flag;
        return lg_engine_end_reservoir(eng, &s, &anc, 1);
    }
    int take(int s, unsigned char *out, int cap)
    {
        Stream &x = st[s];
        int const n = (int) std::min<size_t>(x.out.size(), cap < 0 ? 0 : (size_t) cap);
        if (n > 0) { memcpy(out, x.out.data(), n); x.out.erase(x.out.begin(), x.out.begin() + n); }
        return n;
    }
};

static thread_local const int *g_header_bits = nullptr;
static thread_local const LgSetupOpt *g_setup_opt = nullptr;  /* likewise: the handle's options that change constants of the configuration */
/* the device new engines of the lame_t face are made on: LAMEGPU_DEVICE, else the calling thread's current CUDA device */
static int lg_default_device()
{
    if (const char *e = getenv("LAMEGPU_DEVICE")) return atoi(e);
    return 0;
}     /* set by lame_init_params around its lamegpu_batch_open_vq call */

extern "C" {

lamegpu_batch *lamegpu_batch_open_ex(int samplerate, int channels, int brate, int mode, int quality, int vbr, int nstreams, int frames_per_launch, int device)
{
    return lamegpu_batch_open_rs(samplerate, 0, channels, brate, mode, quality, vbr, nstreams, frames_per_launch, device);
}

lamegpu_batch *lamegpu_batch_open_rs(int samplerate_in, int samplerate_out, int channels, int brate, int mode, int quality, int vbr, int nstreams,
                                     int frames_per_launch, int device)
{
    return lamegpu_batch_open_vq(samplerate_in, samplerate_out, channels, (float) brate, mode, quality, vbr, nstreams, frames_per_launch, device);
}

lamegpu_batch *lamegpu_batch_open_vq(int samplerate_in, int samplerate_out, int channels, float rate, int mode, int quality, int vbr, int nstreams,
                                     int frames_per_launch, int device)
{
    int const brate = (int) rate;
    float const vbr_q_frac = (vbr == 4 || vbr == 2) ? rate - (float) brate : 0.f;
    if (device < 0) {
        /* all visible GPUs (LAMEGPU_DEVICES limits their number): stream s goes to part s * n / nstreams */
        int n = lg_device_count();
        if (const char *e = getenv("LAMEGPU_DEVICES")) n = std::min(n, std::max(1, atoi(e)));
        if (n > nstreams) n = nstreams;
        if (n <= 1) return lamegpu_batch_open_vq(samplerate_in, samplerate_out, channels, rate, mode, quality, vbr, nstreams, frames_per_launch, 0);
        lamegpu_batch *c = new (std::nothrow) lamegpu_batch;
        if (!c) return NULL;
        for (int d = 0; d < n; d++) {
            int const lo = (int) ((long) nstreams * d / n), hi = (int) ((long) nstreams * (d + 1) / n);
            lamegpu_batch *p = lamegpu_batch_open_vq(samplerate_in, samplerate_out, channels, rate, mode, quality, vbr, hi - lo, frames_per_launch, d);
            if (!p) { for (auto *q : c->parts) lamegpu_batch_close(q); delete c; return NULL; }
            c->parts.push_back(p); c->part_first.push_back(lo);
        }
        c->cfg = c->parts[0]->cfg; c->S = nstreams; c->F = frames_per_launch; c->nthreads = c->parts[0]->nthreads;
        return c;
    }
    lamegpu_batch *b = new (std::nothrow) lamegpu_batch;
    if (!b) return NULL;
    LgSetupOpt defaults;
    lg_setup_opt_defaults(&defaults);
    if (lg_setup_ex(&b->cfg, samplerate_in, samplerate_out, channels, brate, mode < 0 ? LG_MODE_NOT_SET : mode, quality, vbr, vbr_q_frac,
                    g_setup_opt ? g_setup_opt : &defaults) != 0) {
        fprintf(stderr, "lamegpu: unsupported configuration (samplerate %d -> %d, channels %d, brate %d, mode %d, quality %d, vbr %d)\n",
                samplerate_in, samplerate_out, channels, brate, mode, quality, vbr);
        delete b;
        return NULL;
    }
    if (g_header_bits) {             /* lame_init_params: copyright / original / emphasis / extension / error_protection of the handle */
        b->cfg.copyright = g_header_bits[0]; b->cfg.original = g_header_bits[1]; b->cfg.emphasis = g_header_bits[2]; b->cfg.extension = g_header_bits[3];
        if (g_header_bits[4]) { b->cfg.error_protection = 1; b->cfg.sideinfo_len += 2; }       /* lame.c:954 */
    }
    if (getenv("LAMEGPU_DEBUG_CFG")) {
        unsigned long h = 1469598103934665603ul;
        const unsigned char *q = reinterpret_cast<const unsigned char *>(&b->cfg);
        for (size_t i = 0; i < sizeof b->cfg; i++) h = (h ^ q[i]) * 1099511628211ul;
        fprintf(stderr, "lamegpu: configuration %016lx (S=%d F=%d)\n", h, nstreams, frames_per_launch);
    }
    b->eng = lg_engine_create(&b->cfg, nstreams, frames_per_launch, device);
    if (!b->eng) { delete b; return NULL; }
    b->S = nstreams; b->F = frames_per_launch;
    /* host worker pool for staging and splice: memcpy-class work that overlaps the device when the batch is pipelined - a few threads do */
    unsigned const hw = std::thread::hardware_concurrency();
    b->nthreads = (int) std::max(1u, std::min(hw ? hw : 1u, 8u));
    if (const char *e = getenv("LAMEGPU_THREADS")) b->nthreads = std::max(1, atoi(e));
    b->st.resize(nstreams);
    for (auto &s : b->st) {
        s.rs_mode = b->cfg.resample != 0; s.fs = 576 * b->cfg.mode_gr; s.need = 1024 + s.fs - 272;
        s.init(); s.last_bitrate_index = b->cfg.bitrate_index;
    }
    return b;
}

lamegpu_batch *lamegpu_batch_open(int samplerate, int channels, int brate, int mode, int quality, int nstreams, int frames_per_launch, int device)
{
    return lamegpu_batch_open_ex(samplerate, channels, brate, mode, quality, 0 /* vbr_off */, nstreams, frames_per_launch, device);
}

void lamegpu_batch_close(lamegpu_batch *b)
{
    if (!b) return;
    if (!b->parts.empty()) { for (auto *p : b->parts) lamegpu_batch_close(p); delete b; return; }
    lg_engine_destroy(b->eng);
    delete b;
}

int lamegpu_batch_set_threads(lamegpu_batch *b, int n)
{
    if (!b || n < 1) return -1;
    for (auto *p : b->parts) p->nthreads = n;
    b->nthreads = n;
    return 0;
}

long lamegpu_batch_encode(lamegpu_batch *b, const short *const *pcm_l, const short *const *pcm_r, const int *nsamples,
                          unsigned char *const *out, const int *out_cap, int *out_bytes)
{
    if (!b) return -3;
    if (!b->parts.empty())
        return b->fan_out([&](int p) {
            int const f = b->part_first[p];
            return lamegpu_batch_encode(b->parts[p], pcm_l + f, pcm_r ? pcm_r + f : NULL, nsamples + f, out ? out + f : NULL, out_cap ? out_cap + f : NULL, out_bytes + f);
        });
    b->parallel_for(b->S, [&](int s) { b->feed16(s, pcm_l[s], pcm_r ? pcm_r[s] : NULL, nsamples[s]); });
    long const done = b->pump();
    b->end_call();
    if (done < 0) return done;
    b->parallel_for(b->S, [&](int s) { out_bytes[s] = out ? b->take(s, out[s], out_cap[s]) : 0; });
    return done;
}

long lamegpu_batch_flush(lamegpu_batch *b, unsigned char *const *out, const int *out_cap, int *out_bytes)
{
    if (!b) return -3;
    if (!b->parts.empty())
        return b->fan_out([&](int p) { int const f = b->part_first[p]; return lamegpu_batch_flush(b->parts[p], out ? out + f : NULL, out_cap ? out_cap + f : NULL, out_bytes + f); });
    std::vector<char> live(b->S);
    for (int s = 0; s < b->S; s++) {
        live[s] = b->st[s].mf_samples_to_encode >= 1;         /* lame.c:2067: "was flush already called?" */
        b->pad_for_flush(s);
    }
    long const done = b->pump();
    if (done < 0) return done;
    if (b->drain() != 0) return -2;
    if (b->finish_streams(live.data()) != 0) return -2;
    for (int s = 0; s < b->S; s++) out_bytes[s] = out ? b->take(s, out[s], out_cap[s]) : 0;
    return done;
}

long lamegpu_batch_encode_packed(lamegpu_batch *b, const short *pcm, int nsamples, unsigned char *out, int out_stride, int *out_bytes)
{
    if (!b) return -3;
    if (!b->parts.empty())
        return b->fan_out([&](int p) {
            size_t const f = (size_t) b->part_first[p];
            return lamegpu_batch_encode_packed(b->parts[p], pcm + f * 2 * (size_t) nsamples, nsamples, out + f * (size_t) out_stride, out_stride, out_bytes + f);
        });
    double const t0 = now_ms();
    b->parallel_for(b->S, [&](int s) { b->feed16(s, pcm + ((size_t) s * 2) * nsamples, pcm + ((size_t) s * 2 + 1) * nsamples, nsamples); });
    double const t1 = now_ms();
    long const done = b->pump();
    b->end_call();
    if (done < 0) return done;
    double const t2 = now_ms();
    b->parallel_for(b->S, [&](int s) { out_bytes[s] = b->take(s, out + (size_t) s * out_stride, out_stride); });
    if (g_timing) fprintf(stderr, "lamegpu: feed %.2f ms, pump %.2f ms, take %.2f ms\n", t1 - t0, t2 - t1, now_ms() - t2);
    return done;
}

long lamegpu_batch_flush_packed(lamegpu_batch *b, unsigned char *out, int out_stride, int *out_bytes)
{
    if (!b) return -3;
    if (!b->parts.empty())
        return b->fan_out([&](int p) { size_t const f = (size_t) b->part_first[p]; return lamegpu_batch_flush_packed(b->parts[p], out + f * (size_t) out_stride, out_stride, out_bytes + f); });
    std::vector<unsigned char *> po(b->S);
    std::vector<int> cap(b->S, out_stride);
    for (int s = 0; s < b->S; s++) po[s] = out + (size_t) s * out_stride;
    return lamegpu_batch_flush(b, po.data(), cap.data(), out_bytes);
}

int lamegpu_batch_set_pipelined(lamegpu_batch *b, int on)
{
    if (!b) return -1;
    if (!b->parts.empty()) return (int) b->fan_out([&](int p) { return (long) lamegpu_batch_set_pipelined(b->parts[p], on); });
    if (!on && b->drain() != 0) return -2;
    b->pipelined = on != 0;
    return 0;
}

/* bench hooks */
int lamegpu_batch_stage_packed(lamegpu_batch *b, const short *pcm, int nframes)
{
    /* Lay a RING of 2 x nframes frames per stream into the engine for lamegpu_batch_run_device_steps: pcm is [S][2][2 * nframes * fs] (fs =
     * samples per frame); the window of buffer set k is the ring from sample k * nframes * fs - 576 on, nframes * fs + 1328 samples long,
     * so that alternating the two sets walks a periodic signal without a seam - every device step then sees what a long stream sees
     * (the first step's history is the ring's end).  One full step per buffer set brings the windows into device memory. */
    if (!b || nframes < 1 || nframes > b->F) return -1;
    if (!b->parts.empty()) {
        size_t const per_stream = 2 * (2 * (size_t) nframes * b->parts[0]->st[0].fs);
        return (int) b->fan_out([&](int p) { return (long) lamegpu_batch_stage_packed(b->parts[p], pcm + (size_t) b->part_first[p] * per_stream, nframes); });
    }
    if (b->drain() != 0) return -1;
    size_t const stride = lg_engine_pcm_stride(b->eng);
    long const n = (long) nframes * b->st[0].fs, ring = 2 * n;
    /* four staging steps: set 0 with silence for everything the first granule's analysis treats as "the granule before" (a stream's true
     * start, which is what the fresh state stands for: the psycho-acoustic analysis runs 701 samples ahead of the frame, so that is the
     * history plus the first 722 samples), set 1, then set 0 again with the ring's end for history and set 1 once more - from here on the
     * state and the windows belong to one long stream */
    for (int pass = 0; pass < 2 * lg_engine_slots(b->eng); pass++) {
        int const k = pass % lg_engine_slots(b->eng);
        int16_t *hp = lg_engine_host_pcm16(b->eng, k);
        int *nfr = lg_engine_host_nfr(b->eng, k);
        long const start = ((k & 1) * n - LG_PCM_HIST + ring) % ring;
        for (int s = 0; s < b->S; s++) {
            nfr[s] = nframes;
            for (int c = 0; c < 2; c++) {
                int16_t *d = hp + ((size_t) s * 2 + c) * stride;
                const short *src = pcm + ((size_t) s * 2 + c) * (size_t) ring;
                for (long i = 0, pos = start; i < n + LG_PCM_HALO; ) {
                    long const run = std::min<long>(ring - pos, n + LG_PCM_HALO - i);
                    memcpy(d + i, src + pos, (size_t) run * sizeof(int16_t));
                    i += run; pos = (pos + run) % ring;
                }
                if (pass == 0) memset(d, 0, (size_t) (LG_PCM_HIST + 736) * sizeof(int16_t));
            }
        }
        if (lg_engine_submit(b->eng, k, nframes, 0) != 0 || lg_engine_wait(b->eng, k) != 0) return -1;
    }
    return 0;
}
/* `steps` device-only steps on the staged input, back to back on alternating slots, the streams' state carried from step to step
 * (persistent streams: no reset in between).  Returns the device time per step in ms: first kernel's start to last kernel's end over
 * all steps, so that consecutive steps overlap as they do in production. */
float lamegpu_batch_run_device_steps(lamegpu_batch *b, int nframes, int steps)
{
    if (!b || steps < 1) return -1.f;
    if (!b->parts.empty()) {                                  /* all devices at once; the slowest one's time per step */
        std::vector<float> ms(b->parts.size(), 0.f);
        b->fan_out([&](int p) { ms[p] = lamegpu_batch_run_device_steps(b->parts[p], nframes, steps); return 0L; });
        float worst = 0.f;
        for (float v : ms) { if (v < 0.f) return v; worst = std::max(worst, v); }
        return worst;
    }
    if (b->drain() != 0) return -1.f;
    for (int i = 0; i < 5; i++) b->acc_ms[i] = 0;
    b->acc_n = 0;
    auto collect = [&](int k) {
        if (!lg_engine_in_flight(b->eng, k)) return 0;
        if (lg_engine_wait(b->eng, k) != 0) return -1;
        const float *m = lg_engine_last_kernel_ms(b->eng);
        for (int i = 0; i < 5; i++) b->acc_ms[i] += m[i];
        b->acc_n++;
        return 0;
    };
    int k = 0;
    if (lg_engine_mark(b->eng, 0) != 0) return -1.f;
    for (int i = 0; i < steps; i++, k ^= 1) {
        if (collect(k) != 0) return -1.f;                     /* step i-2 (its events are about to be reused); step i-1 keeps the device busy */
        if (lg_engine_run_device(b->eng, k, nframes, 0) != 0) return -1.f;
    }
    if (lg_engine_mark(b->eng, 1) != 0) return -1.f;
    if (collect(k) != 0 || collect(k ^ 1) != 0) return -1.f;
    float const ms = lg_engine_marked_ms(b->eng);
    return ms < 0.f ? ms : ms / (float) steps;
}
/* one more pass over the staged ring: both buffer sets (so that the streams stay where a long stream would be), one step at a time -
 * lamegpu_batch_kernel_ms then gives each kernel's time with the device to itself */
int lamegpu_batch_rerun_device(lamegpu_batch *b, int nframes)
{
    if (!b) return -1;
    if (!b->parts.empty()) return (int) b->fan_out([&](int p) { return (long) lamegpu_batch_rerun_device(b->parts[p], nframes); });
    if (b->drain() != 0) return -1;
    for (int i = 0; i < 5; i++) b->acc_ms[i] = 0;
    b->acc_n = 0;
    for (int k = 0; k < lg_engine_slots(b->eng); k++) {
        if (lg_engine_run_device(b->eng, k, nframes, 0) != 0 || lg_engine_wait(b->eng, k) != 0) return -1;
        const float *m = lg_engine_last_kernel_ms(b->eng);
        for (int i = 0; i < 5; i++) b->acc_ms[i] += m[i];
        b->acc_n++;
    }
    return 0;
}
int lamegpu_batch_kernel_ms(const lamegpu_batch *b, float ms[5])
{
    if (!b) return -1;
    if (!b->parts.empty()) return lamegpu_batch_kernel_ms(b->parts[0], ms);
    if (b->acc_n > 0) { for (int i = 0; i < 5; i++) ms[i] = (float) (b->acc_ms[i] / b->acc_n); return 0; }   /* mean over the last run of device steps */
    const float *m = lg_engine_last_kernel_ms(b->eng);
    for (int i = 0; i < 5; i++) ms[i] = m[i];
    return 0;
}
/* device time of the last launch from the start of its first kernel to the end of its last (the pieces' kernels overlap) */
float lamegpu_batch_step_ms(const lamegpu_batch *b) { return !b ? 0.f : (b->parts.empty() ? lg_engine_last_kernel_ms(b->eng)[7] : lamegpu_batch_step_ms(b->parts[0])); }
long lamegpu_batch_kernel_launches(const lamegpu_batch *b)
{
    if (!b) return 0;
    if (b->parts.empty()) return lg_engine_launch_count(b->eng);
    long n = 0;
    for (auto *p : b->parts) n += lamegpu_batch_kernel_launches(p);
    return n;
}
long lamegpu_batch_debug_copy(lamegpu_batch *b, int what, void *dst, size_t cap)
{
    if (!b) return -1;
    return lg_engine_debug_copy(b->parts.empty() ? b->eng : b->parts[0]->eng, what, dst, cap);
}
int lamegpu_batch_devices(const lamegpu_batch *b) { return !b ? 0 : (b->parts.empty() ? 1 : (int) b->parts.size()); }
size_t lamegpu_sizeof_granule_out(void) { return sizeof(LgGranuleOut); }
/* bytes one launch copies device -> host: frame records, payload bytes, headers */
long lamegpu_batch_d2h_bytes(const lamegpu_batch *b)
{
    if (!b) return 0;
    if (!b->parts.empty()) { long n = 0; for (auto *p : b->parts) n += lamegpu_batch_d2h_bytes(p); return n; }
    return (long) ((size_t) b->S * b->F * sizeof(LgFrameOut) + (size_t) b->S * lg_engine_pay_stride(b->eng) + (size_t) b->S * b->F * LG_HDR_STRIDE);
}
size_t lamegpu_sizeof_analysis(void) { return sizeof(LgAnalysis); }

/* ------------------------------------------------------------------ libmp3lame-compatible face */
#define LAME_ID 0xFFF88E3B      /* lame_global_flags.h / lame.c class_id check */

/* libmp3lame options this library only carries (include/lamegpu_options.h, generated from the reference's lame.h):
 * name, C type, default before and after lame_init_params, honoured */
enum {
#define LG_OPT(n, t, d, d2, h) LG_OPTI_##n,
#include "lg_api_options.inc"
#undef LG_OPT
    LG_NOPT
};
static const struct { const char *name; double def, def2; int honoured; } g_opts[LG_NOPT] = {
#define LG_OPT(n, t, d, d2, h) { #n, (double) d, (double) d2, h },
#include "lg_api_options.inc"
#undef LG_OPT
};

struct lame_global_struct {
    unsigned class_id;
    double opt[LG_NOPT];
    unsigned char opt_set[LG_NOPT];
    int num_channels, samplerate_in, samplerate_out, brate, quality, write_lame_tag, mean_brate, vbr_q;
    float vbr_q_frac;
    MPEG_mode mode;
    vbr_mode VBR;
    int launch_frames;
    struct LgShared *se;         /* the engine this handle is a lane of (lame_init_params) */
    int lane;
    int initialised;
    lame_report_function report[3];      /* lame_set_errorf / debugf / msgf */
    int preset, preset_foreign, write_id3tag_automatic;
    int short_blocks;            /* lame_global_flags.h:20 short_block_t: -1 not set, 0 allowed, 1 coupled, 2 dispensed, 3 forced */
    float msfix;                 /* -1 = not set */
};

/* ---- One engine for many handles.  A lame_t is one stream and a GPU wants hundreds per launch, so handles of equal configuration
 * are LANES of one process-wide batch engine.  lame_encode_buffer() copies the samples into its lane and, when
 * they complete frames, sleeps until those frames have been encoded - the bytes come back from the call that completes the frame, as
 * with the reference (lame.c:1743) - while a dispatcher thread per engine gathers the frames that are ready in ALL lanes into one
 * launch.  Application threads that call lame_encode_buffer concurrently on their own handles (the reference's threading contract,
 * HACKING:67-76) therefore share launches: 512 threads feeding one frame each are one 512 x 1 step.  lame_encode_flush is the lane's
 * barrier.
 * Locks: every lane has its own mutex (its stream's data: the owner thread feeding and taking, the dispatcher staging and splicing),
 * so the application threads do not serialise on one another; the engine's mutex `m` guards the lane table, the dispatcher's state and
 * the engine-level device operations.  Order: m before lane mutexes; the dispatcher holds neither while the device works. */
struct LgShared {
    struct Lane { std::mutex lm; std::condition_variable cv; };
    lamegpu_batch *b = nullptr;
    std::string key;
    std::mutex m;
    std::condition_variable cv_work;
    std::unique_ptr<Lane[]> lane;
    std::vector<char> used;
    int nused = 0;
    std::atomic<unsigned long> work{0};
    bool stop = false;
    std::atomic<bool> failed{false};
    std::thread th;
    long steps = 0, step_frames = 0, last_lanes = 0;
    double t_gather = 0, t_stage = 0, t_device = 0, t_splice = 0;     /* LAMEGPU_TIMING: where the dispatcher's time goes, ms */

    void lock_lanes() { for (int s = 0; s < b->S; s++) lane[s].lm.lock(); }
    void unlock_lanes() { for (int s = b->S - 1; s >= 0; s--) lane[s].lm.unlock(); }
    void fail() { failed = true; for (int s = 0; s < b->S; s++) { std::lock_guard<std::mutex> g(lane[s].lm); lane[s].cv.notify_all(); } }
    /* a feeder has completed frames in its lane.  Without the engine's mutex (hundreds of feeders would queue behind the dispatcher's
     * critical sections): a wake-up that slips between the submitter's test and its sleep is caught by its 100 us poll. */
    void kick() { work.fetch_add(1, std::memory_order_release); cv_work.notify_one(); }
    /* Two dispatcher threads keep two launches in flight (the engine's two slots): the submitter stages and submits whatever is ready as
     * soon as a slot is free, the completer waits for the oldest launch, splices its frames into the lanes and wakes their threads.
     * Threads that run in step fall into two groups that alternate, one being encoded while the other is fed. */
    std::condition_variable cv_queue;
    std::vector<int> queue;                  /* submitted slots, oldest first */
    std::thread th2;

    void run_submit()
    {
        std::unique_lock<std::mutex> lk(m);
        unsigned long seen = 0;
        for (;;) {
            while (!(stop || (work.load(std::memory_order_acquire) != seen && !b->in_flight[b->next_slot]))) cv_work.wait_for(lk, std::chrono::microseconds(100));
            if (stop) break;
            double const tg = now_ms();
            /* Application threads that run in step (each feeds a frame, waits for its bytes, feeds the next) come back within a short time
             * of one another: wait until as many lanes have reported frames as the last launch carried, at most 200 us - a launch per
             * straggler would cost each of them a full device round trip.  A single handle never waits here. */
            if (last_lanes > 1) {
                auto const deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(200);
                cv_work.wait_until(lk, deadline, [&]() { return stop || (long) (work.load() - seen) >= last_lanes; });
                if (stop) break;
            }
            seen = work.load();
            double const t0 = now_ms();
            long n = 0;
            lock_lanes();
            int const k = b->stage_and_submit(&n);
            unlock_lanes();
            if (k == -1) continue;
            if (k < 0) { fail(); break; }
            steps++; step_frames += n;
            { int lanes = 0; const int *q = lg_engine_host_nfr(b->eng, k); for (int i = 0; i < b->S; i++) lanes += q[i] > 0; last_lanes = lanes; }
            queue.push_back(k);
            cv_queue.notify_one();
            seen = work.load() - 1;          /* look again: lanes may hold more than one launch's worth */
            t_gather += t0 - tg; t_stage += now_ms() - t0;
        }
    }
    void run_complete()
    {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv_queue.wait(lk, [&]() { return stop || !queue.empty(); });
            if (queue.empty()) break;        /* stop: what is in flight belongs to lanes that were closed */
            int const k = queue.front();
            double const t1 = now_ms();
            lk.unlock();
            int const rc = lg_engine_wait(b->eng, k);          /* the device works: the lanes can be fed, the next launch staged meanwhile */
            lk.lock();
            double const t2 = now_ms();
            lock_lanes();
            int const rc2 = rc != 0 ? rc : b->complete(k);
            unlock_lanes();
            queue.erase(queue.begin());
            if (rc2 != 0) { fail(); break; }
            /* wake the lanes' threads - a system call each - on the worker pool: several hundred in a row cost this thread more than
             * the splice did */
            const std::vector<int> &nfr = b->flight_nfr[k];
            b->parallel_for(b->S, [&](int s) { if (nfr[s]) lane[s].cv.notify_all(); });
            cv_work.notify_all();            /* a slot is free again */
            t_device += t2 - t1; t_splice += now_ms() - t2;
        }
    }
};
static std::mutex g_shared_m;
static std::vector<LgShared *> g_shared;

static LgShared *shared_acquire(const std::string &key, const std::function<lamegpu_batch *()> &make, int *lane)
{
    std::lock_guard<std::mutex> g(g_shared_m);
    for (LgShared *se : g_shared) {
        if (se->key != key || se->failed) continue;
        std::lock_guard<std::mutex> lk(se->m);               /* the dispatcher is between steps or waiting for the device */
        for (int s = 0; s < se->b->S; s++)
            if (!se->used[s]) {
                /* a lane that was used before starts over: fresh stream on the host, initial state on the device (ordered behind
                 * whatever the engine has in flight - no new step can be submitted while m is held) */
                std::lock_guard<std::mutex> ll(se->lane[s].lm);
                if (lg_engine_reset_streams(se->b->eng, s, 1) != 0) return nullptr;
                se->b->st[s].init();
                se->b->st[s].last_bitrate_index = se->b->cfg.bitrate_index;
                se->used[s] = 1; se->nused++;
                *lane = s;
                return se;
            }
    }
    lamegpu_batch *b = make();
    if (!b) return nullptr;
    LgShared *se = new LgShared;
    se->b = b; se->key = key;
    {
        /* measured with 512 threads (profiles/r2_handles.txt): no cap 2.65e5 frames/s, 3 -> 3.6e5, 2 -> 3.8e5, 1 -> 3.3e5 */
        const char *e = getenv("LAMEGPU_HANDLE_CROWD_CAP"), *e2 = getenv("LAMEGPU_HANDLE_CROWD_LANES");
        b->crowd_cap = e ? std::max(0, atoi(e)) : 2;
        b->crowd_lanes = e2 ? std::max(1, atoi(e2)) : 32;        /* 64 threads: 3.9e4 frames/s without the cap, 6.6e4 with it */
    }
    se->lane.reset(new LgShared::Lane[b->S]);
    se->used.assign(b->S, 0);
    se->used[0] = 1; se->nused = 1;
    *lane = 0;
    se->th = std::thread([se]() { se->run_submit(); });
    se->th2 = std::thread([se]() { se->run_complete(); });
    g_shared.push_back(se);
    return se;
}
static void shared_release(LgShared *se, int lane)
{
    std::lock_guard<std::mutex> g(g_shared_m);
    bool last;
    {
        std::lock_guard<std::mutex> lk(se->m);
        se->used[lane] = 0; se->nused--;
        last = se->nused == 0;
        if (last) se->stop = true;
    }
    if (!last) return;
    /* the last handle of an engine takes it along: nothing of this library stays behind in a process that has closed all its handles */
    se->cv_work.notify_all(); se->cv_queue.notify_all();
    if (se->th.joinable()) se->th.join();
    if (se->th2.joinable()) se->th2.join();
    g_shared.erase(std::remove(g_shared.begin(), g_shared.end(), se), g_shared.end());
    if (g_timing && se->steps)
        fprintf(stderr, "lamegpu: shared engine closed after %ld launches, %.1f frames per launch; per launch: gather %.3f ms, stage + submit %.3f, device %.3f, splice + wake %.3f\n",
                se->steps, (double) se->step_frames / se->steps, se->t_gather / se->steps, se->t_stage / se->steps, se->t_device / se->steps, se->t_splice / se->steps);
    lamegpu_batch_close(se->b);
    delete se;
}

/* VbrTag.c:255 setLameTagFrameHeader (MPEG-1 branch): the tag frame looks like a frame of the stream itself, without padding */
static int tag_xing_kbps(const LgDevCfg *c) { return c->version == 1 ? 128 : (c->samplerate < 16000 ? 32 : 64); }

static void tag_frame_header(const LgDevCfg *c, int mode_ext, unsigned char *buffer)
{
    buffer[0] = 0xff;
    /* sync (MPEG-2.5 clears its last bit), version bit, layer III, protection (VbrTag.c:262-267, :308-322) */
    buffer[1] = (unsigned char) ((c->samplerate < 16000 ? 0xe0 : 0xf0) | (c->version == 1 ? 0x0a : 0x02) | (c->error_protection ? 0 : 1));
    /* CBR: the stream's own bitrate; otherwise XING_BITRATE1 / 2 / 25 = 128 / 64 / 32 kbps (VbrTag.c:285-306) */
    int bidx = c->bitrate_index;
    int const xing_kbps = tag_xing_kbps(c);
    if (c->vbr != 0) for (bidx = 1; bidx < 15 && c->bitrate_kbps[bidx] != xing_kbps; bidx++) { }
    buffer[2] = (unsigned char) ((16 * bidx) | ((c->samplerate_index << 2) & 0x0c) | (c->extension & 1));
    buffer[3] = (unsigned char) ((c->mode << 6) | ((mode_ext & 3) << 4) | ((c->copyright & 1) << 3) | ((c->original & 1) << 2) | (c->emphasis & 3));
}

#define LG_VBRHEADERSIZE (100 + 4 + 4 + 4 + 4 + 4)
#define LG_LAMEHEADERSIZE (LG_VBRHEADERSIZE + 9 + 1 + 1 + 8 + 1 + 1 + 3 + 1 + 1 + 2 + 4 + 2 + 2)

/* VbrTag.c:492 InitVbrTag for CBR: an all-zero frame (but for its header) goes into the stream ahead of the audio */
static void tag_init(lamegpu_batch *b, int lane)
{
    Stream &x = b->st[lane];
    const LgDevCfg *c = &b->cfg;
    int const kbps_header = (c->vbr == 0) ? c->brate : tag_xing_kbps(c);         /* VbrTag.c:517-529 */
    int const total = ((c->version + 1) * 72000 * kbps_header) / c->samplerate;
    if (total < c->sideinfo_len + LG_LAMEHEADERSIZE || total > 2880) return;      /* "disable tag, it wont fit" */
    x.tag.on = true;
    x.tag.frame_size = total;
    std::vector<unsigned char> frame((size_t) total, 0);
    tag_frame_header(c, 0, frame.data());
    x.out.insert(x.out.end(), frame.begin(), frame.end());
    x.tag.pending = total;
    /* add_dummy_byte (bitstream.c:893): the bytes pass through the bit writer and every header slot moves with them */
    x.bw.totbit += 8L * total;
    for (int i = 0; i < LG_MAX_HEADER_BUF; ++i) x.bw.header[i].write_timing += 8L * total;
}

static void put_be32(unsigned char *p, unsigned long v) { p[0] = (v >> 24) & 0xff; p[1] = (v >> 16) & 0xff; p[2] = (v >> 8) & 0xff; p[3] = v & 0xff; }
static void put_be16(unsigned char *p, unsigned v) { p[0] = (v >> 8) & 0xff; p[1] = v & 0xff; }

static int handle_quiesce(lame_global_flags *g);
static inline lamegpu_batch *HB(const lame_global_flags *g) { return g->se ? g->se->b : nullptr; }
static inline Stream &HS(const lame_global_flags *g) { return g->se->b->st[g->lane]; }

lame_global_flags *lame_init(void)
{
    lame_global_flags *g = (lame_global_flags *) calloc(1, sizeof *g);
    if (!g) return NULL;
    g->class_id = LAME_ID;
    for (int i = 0; i < LG_NOPT; i++) g->opt[i] = g_opts[i].def;
    g->num_channels = 2; g->samplerate_in = 44100; g->samplerate_out = 0; g->brate = 0; g->quality = -1;
    g->write_lame_tag = 1; g->mode = NOT_SET; g->VBR = vbr_off; g->launch_frames = 8; g->mean_brate = 128; g->vbr_q = 4; g->vbr_q_frac = 0;
    if (const char *e = getenv("LAMEGPU_HANDLE_FRAMES")) g->launch_frames = std::max(1, atoi(e));
    g->short_blocks = -1; g->msfix = -1.f; g->write_id3tag_automatic = 1;
    return g;
}
static int ok(const lame_global_flags *g) { return g && g->class_id == LAME_ID; }
int lame_set_in_samplerate(lame_global_flags *g, int v) { if (!ok(g)) return -1; g->samplerate_in = v; return 0; }
int lame_get_in_samplerate(const lame_global_flags *g) { return ok(g) ? g->samplerate_in : 0; }
int lame_set_num_channels(lame_global_flags *g, int v) { if (!ok(g) || v < 1 || v > 2) return -1; g->num_channels = v; return 0; }
int lame_get_num_channels(const lame_global_flags *g) { return ok(g) ? g->num_channels : 0; }
int lame_set_out_samplerate(lame_global_flags *g, int v) { if (!ok(g)) return -1; g->samplerate_out = v; return 0; }
int lame_get_out_samplerate(const lame_global_flags *g) { return ok(g) ? g->samplerate_out : 0; }
int lame_set_brate(lame_global_flags *g, int v) { if (!ok(g)) return -1; g->brate = v; if (v > 320) return -1; return 0; }
int lame_get_brate(const lame_global_flags *g) { return ok(g) ? g->brate : 0; }
int lame_set_quality(lame_global_flags *g, int v) { if (!ok(g)) return -1; g->quality = v < 0 ? 0 : v > 9 ? 9 : v; return 0; }
int lame_get_quality(const lame_global_flags *g) { return ok(g) ? g->quality : 0; }
int lame_set_mode(lame_global_flags *g, MPEG_mode m) { if (!ok(g) || (int) m < 0 || m >= MAX_INDICATOR) return -1; g->mode = m; return 0; }
MPEG_mode lame_get_mode(const lame_global_flags *g) { return ok(g) ? g->mode : NOT_SET; }
int lame_set_VBR(lame_global_flags *g, vbr_mode m) { if (!ok(g) || (int) m < 0 || m >= vbr_max_indicator) return -1; g->VBR = m; return 0; }
vbr_mode lame_get_VBR(const lame_global_flags *g) { return ok(g) ? g->VBR : vbr_off; }
int lame_set_VBR_mean_bitrate_kbps(lame_global_flags *g, int v) { if (!ok(g)) return -1; g->mean_brate = v; return 0; }     /* set_get.c:1241 */
int lame_get_VBR_mean_bitrate_kbps(const lame_global_flags *g) { return ok(g) ? g->mean_brate : 0; }
int lame_set_VBR_q(lame_global_flags *g, int v)                     /* set_get.c:1112: clamps to 0..9 and reports -1 */
{
    if (!ok(g)) return -1;
    int ret = 0;
    if (v < 0) { ret = -1; v = 0; }
    if (v > 9) { ret = -1; v = 9; }
    g->vbr_q = v;
    g->vbr_q_frac = 0;
    return ret;
}
int lame_set_VBR_quality(lame_global_flags *g, float v)            /* set_get.c:1152: 0 .. 9.999, integer level + fraction */
{
    if (!ok(g)) return -1;
    int ret = 0;
    if (0 > v) { ret = -1; v = 0; }
    if (9.999 < v) { ret = -1; v = 9.999; }
    g->vbr_q = (int) v;
    g->vbr_q_frac = v - g->vbr_q;
    return ret;
}
float lame_get_VBR_quality(const lame_global_flags *g) { return ok(g) ? g->vbr_q + g->vbr_q_frac : 0; }
int lame_get_VBR_q(const lame_global_flags *g) { return ok(g) ? g->vbr_q : 0; }
int lame_set_bWriteVbrTag(lame_global_flags *g, int v) { if (!ok(g)) return -1; g->write_lame_tag = v; return 0; }
int lame_get_bWriteVbrTag(const lame_global_flags *g) { return ok(g) ? g->write_lame_tag : 0; }
const char *get_lame_short_version(void) { return "3.99.5"; }

int lame_init_params(lame_global_flags *g)
{
    if (!ok(g)) return -1;
    /* an option this library does not act on is accepted only while it has the value of a fresh handle - anything else would be
     * silently different from libmp3lame */
    for (int i = 0; i < LG_NOPT; i++)
        if (g->opt_set[i] && !g_opts[i].honoured && g->opt[i] != g_opts[i].def) {
            fprintf(stderr, "lamegpu: lame_set_%s(%g) is not supported (only %g, the value of a fresh handle)\n", g_opts[i].name, g->opt[i], g_opts[i].def);
            return -1;
        }
    LgSetupOpt so;
    lg_setup_opt_defaults(&so);
    so.scale = (float) g->opt[LG_OPTI_scale]; so.scale_left = (float) g->opt[LG_OPTI_scale_left]; so.scale_right = (float) g->opt[LG_OPTI_scale_right];
    so.compression_ratio = (float) g->opt[LG_OPTI_compression_ratio];
    so.lowpassfreq = (int) g->opt[LG_OPTI_lowpassfreq]; so.lowpasswidth = (int) g->opt[LG_OPTI_lowpasswidth];
    so.highpassfreq = (int) g->opt[LG_OPTI_highpassfreq]; so.highpasswidth = (int) g->opt[LG_OPTI_highpasswidth];
    so.no_ath = (int) g->opt[LG_OPTI_noATH]; so.ath_only = (int) g->opt[LG_OPTI_ATHonly]; so.ath_short = (int) g->opt[LG_OPTI_ATHshort];
    so.ath_type = (int) g->opt[LG_OPTI_ATHtype]; so.ath_lower_db = (float) g->opt[LG_OPTI_ATHlower];
    so.athaa_type = (int) g->opt[LG_OPTI_athaa_type]; so.athaa_sensitivity = (float) g->opt[LG_OPTI_athaa_sensitivity];
    so.interch = (float) g->opt[LG_OPTI_interChRatio]; so.msfix = g->msfix;
    so.short_blocks = g->short_blocks;
    so.force_ms = (int) g->opt[LG_OPTI_force_ms]; so.disable_reservoir = (int) g->opt[LG_OPTI_disable_reservoir];
    so.strict_iso = (int) g->opt[LG_OPTI_strict_ISO]; so.use_temporal = (int) g->opt[LG_OPTI_useTemporal];
    so.vbr_min_kbps = (int) g->opt[LG_OPTI_VBR_min_bitrate_kbps]; so.vbr_max_kbps = (int) g->opt[LG_OPTI_VBR_max_bitrate_kbps];
    so.vbr_hard_min = (int) g->opt[LG_OPTI_VBR_hard_min];
    if (g->se) { (void) handle_quiesce(g); shared_release(g->se, g->lane); g->se = nullptr; }
    if (g->preset_foreign && g->VBR == vbr_off) {
        fprintf(stderr, "lamegpu: lame_set_preset(V%d) without a VBR mode is not supported\n", g->vbr_q);
        return -1;
    }
    int const is_vbr = (g->VBR == vbr_mt || g->VBR == vbr_mtrh || g->VBR == vbr_rh);   /* vbr_mt and vbr_mtrh both select VBR_new_iteration_loop, encoder.c:531 */
    int const header_bits[5] = { (int) g->opt[LG_OPTI_copyright] != 0, (int) g->opt[LG_OPTI_original] != 0, (int) g->opt[LG_OPTI_emphasis] & 3,
                                 (int) g->opt[LG_OPTI_extension] != 0, (int) g->opt[LG_OPTI_error_protection] != 0 };
    float const rate = is_vbr ? g->vbr_q + g->vbr_q_frac : (float) (g->VBR == vbr_abr ? g->mean_brate : g->brate);
    int const mode = g->mode == NOT_SET ? -1 : (int) g->mode, vbr = is_vbr ? (g->VBR == vbr_rh ? 2 : 4) : (g->VBR == vbr_abr ? 3 : 0);
    /* lanes of one engine: LAMEGPU_LANES handles of this configuration share a launch, LAMEGPU_HANDLE_FRAMES frames of each per launch */
    int lanes = 512;
    if (const char *e = getenv("LAMEGPU_LANES")) lanes = std::max(1, atoi(e));
    std::string key;
    {
        char buf[256];
        snprintf(buf, sizeof buf, "%d/%d/%d/%.6f/%d/%d/%d/%d%d%d%d%d/%d/%d/", g->samplerate_in, g->samplerate_out, g->num_channels, (double) rate, mode, g->quality, vbr,
                 header_bits[0], header_bits[1], header_bits[2], header_bits[3], header_bits[4], lanes, g->launch_frames);
        key = buf;
        key.append(reinterpret_cast<const char *>(&so), sizeof so);       /* the option block, byte for byte (memset by lg_setup_opt_defaults) */
    }
    g->se = shared_acquire(key, [&]() {
        g_header_bits = header_bits;
        g_setup_opt = &so;
        lamegpu_batch *b = lamegpu_batch_open_vq(g->samplerate_in, g->samplerate_out, g->num_channels, rate, mode, g->quality, vbr, lanes, g->launch_frames, lg_default_device());
        g_header_bits = nullptr;
        g_setup_opt = nullptr;
        return b;
    }, &g->lane);
    if (!g->se) return -1;
    const LgDevCfg &c = HB(g)->cfg;
    g->samplerate_out = c.samplerate;
    g->brate = c.brate;
    g->mean_brate = c.vbr_mean_kbps;
    g->quality = c.quality;
    if (is_vbr) { g->vbr_q = c.vbr_q; g->vbr_q_frac = c.vbr_q_frac; }      /* presets.c:203-206 */
    g->mode = (MPEG_mode) c.mode;
    g->initialised = 1;
    if (g->write_lame_tag) {                                  /* lame.c:1249 lame_init_bitstream -> InitVbrTag */
        std::lock_guard<std::mutex> lk(g->se->lane[g->lane].lm);
        tag_init(HB(g), g->lane);
    }
    return 0;
}
int lame_get_framesize(const lame_global_flags *g) { return ok(g) && g->initialised ? 576 * HB(g)->cfg.mode_gr : 0; }
int lame_get_frameNum(const lame_global_flags *g) { return ok(g) && HB(g) ? (int) (HS(g).frames_out - HS(g).frames_out_base) : 0; }
int lame_get_encoder_delay(const lame_global_flags *g) { return ok(g) ? 576 : 0; }

static int handle_take(lame_global_flags *g, unsigned char *mp3buf, int mp3buf_size)
{
    Stream &x = HS(g);
    int const have = (int) x.out.size();
    if (mp3buf_size != 0 && have > mp3buf_size) return -1;            /* lame.h:687: mp3buf too small */
    if (have) {
        memcpy(mp3buf, x.out.data(), have);
        if (x.tag.on) {
            /* copy_buffer (bitstream.c:1079): audio bytes feed the music CRC and the byte count, the tag placeholder does not */
            long const skip = std::min<long>(x.tag.pending, have);
            x.tag.pending -= skip;
            for (long i = skip; i < have; i++) x.tag.music_crc = crc16_update(x.out[i], x.tag.music_crc);
            x.tag.nbytes += have - skip;
        }
        x.out.clear();
    }
    return have;
}
/* How many frames of a handle may be outstanding when lame_encode_buffer returns.  1 = the call returns the bytes of the frames its own
 * samples completed (the reference's timing, lame.c:1743) - every call of every thread then is a device round trip with a sleep and a
 * wake-up.  d > 1 = the call returns once all but the newest d - 1 frames are back, with whatever bytes have arrived: the lane's next
 * frame is already staged when the previous one completes, launches carry more than one frame per lane, and the caller seldom sleeps.
 * The stream is the same byte for byte; bytes reach the caller up to d - 1 frames later (libmp3lame's contract allows a call to return
 * 0 bytes, lame.h:687-699), lame_encode_flush collects the rest, and whatever reads a lane's final state waits for it (handle_quiesce).
 * Default 5 (measured, profiles/r2_handles.txt, 512 threads x 1152-sample calls: depth 1 2.07e5 frames/s, 3 -> 2.54e5, and with the
 * crowded launches capped at two frames per lane - lamegpu_batch::crowd_cap - 3 -> 3.1e5, 5 -> 3.8e5; one thread 976 -> 1339);
 * LAMEGPU_HANDLE_DEPTH=1 restores the synchronous call.  Read once. */
static long handle_depth()
{
    static long const d = []() { const char *e = getenv("LAMEGPU_HANDLE_DEPTH"); return e ? std::max(1L, atol(e)) : 5L; }();
    return d;
}
/* with the lane's mutex held: wake the dispatcher and sleep until the frames this lane's samples complete have been spliced - all of
 * them (`all`: flush, close and whatever else reads the lane's final state) or all but the newest handle_depth() - 1 */
static int handle_wait_frames(lame_global_flags *g, std::unique_lock<std::mutex> &lk, bool all = false)
{
    LgShared *se = g->se;
    Stream &x = HS(g);
    long const ready = std::max<long>(0, x.frames_ready());
    long const target = x.frames_done + ready - (all ? 0 : handle_depth() - 1);
    if (ready > 0) {
        lk.unlock();
        se->kick();
        lk.lock();
    }
    if (x.frames_out >= target) return se->failed.load() ? -2 : 0;
    se->lane[g->lane].cv.wait(lk, [&]() { return se->failed.load() || x.frames_out >= target; });
    return se->failed.load() ? -2 : 0;
}
/* nothing of this lane staged or in flight any more (before its state is read as final or the lane given back) */
static int handle_quiesce(lame_global_flags *g)
{
    if (!g->se) return 0;
    std::unique_lock<std::mutex> lk(g->se->lane[g->lane].lm);
    return handle_wait_frames(g, lk, true);
}
int lame_encode_buffer(lame_global_flags *g, const short int l[], const short int r[], const int nsamples, unsigned char *mp3buf, const int mp3buf_size)
{
    if (!ok(g) || !g->initialised) return -3;
    if (nsamples == 0) return 0;
    if (!l || (g->num_channels > 1 && !r)) return 0;                  /* lame.c:1856-1864 */
    std::unique_lock<std::mutex> lk(g->se->lane[g->lane].lm);
    HB(g)->feed16(g->lane, l, g->num_channels > 1 ? r : l, nsamples, false);
    if (handle_wait_frames(g, lk) != 0) return -2;
    return handle_take(g, mp3buf, mp3buf_size);
}
int lame_encode_buffer_interleaved(lame_global_flags *g, short int pcm[], int nsamples, unsigned char *mp3buf, int mp3buf_size)
{
    if (!ok(g) || !g->initialised) return -3;
    if (nsamples == 0) return 0;
    std::vector<short> l(nsamples), r(nsamples);
    for (int i = 0; i < nsamples; i++) { l[i] = pcm[2 * i]; r[i] = pcm[2 * i + 1]; }
    return lame_encode_buffer(g, l.data(), r.data(), nsamples, mp3buf, mp3buf_size);
}
/* lame.c:1839 lame_encode_buffer_template for every sample type but int16 */
extern "C++" {
template <class T> static int handle_encode_T(lame_global_flags *g, const T *l, const T *r, int nsamples, int jump, float norm, unsigned char *mp3buf, int mp3buf_size)
{
    if (!ok(g) || !g->initialised) return -3;
    if (nsamples == 0) return 0;
    if (!l || (g->num_channels > 1 && !r)) return 0;
    std::unique_lock<std::mutex> lk(g->se->lane[g->lane].lm);
    HB(g)->feedT<T>(g->lane, l, g->num_channels > 1 ? r : l, nsamples, jump, norm);
    if (handle_wait_frames(g, lk) != 0) return -2;
    return handle_take(g, mp3buf, mp3buf_size);
}
}
int lame_encode_buffer_float(lame_global_flags *g, const float l[], const float r[], const int n, unsigned char *mp3buf, const int size)
{
    return handle_encode_T<float>(g, l, r, n, 1, 1.0f, mp3buf, size);                      /* +/- 32768 full scale, lame.c:1884 */
}
int lame_encode_buffer_ieee_float(lame_t g, const float l[], const float r[], const int n, unsigned char *mp3buf, const int size)
{
    return handle_encode_T<float>(g, l, r, n, 1, 32767.0f, mp3buf, size);                  /* +/- 1.0 full scale, lame.c:1895 */
}
int lame_encode_buffer_interleaved_ieee_float(lame_t g, const float pcm[], const int n, unsigned char *mp3buf, const int size)
{
    return handle_encode_T<float>(g, pcm, pcm ? pcm + 1 : pcm, n, 2, 32767.0f, mp3buf, size);
}
int lame_encode_buffer_ieee_double(lame_t g, const double l[], const double r[], const int n, unsigned char *mp3buf, const int size)
{
    return handle_encode_T<double>(g, l, r, n, 1, 32767.0f, mp3buf, size);
}
int lame_encode_buffer_interleaved_ieee_double(lame_t g, const double pcm[], const int n, unsigned char *mp3buf, const int size)
{
    return handle_encode_T<double>(g, pcm, pcm ? pcm + 1 : pcm, n, 2, 32767.0f, mp3buf, size);
}
int lame_encode_buffer_int(lame_global_flags *g, const int l[], const int r[], const int n, unsigned char *mp3buf, const int size)
{
    float const norm = (float) (1.0 / (1L << (8 * sizeof(int) - 16)));                     /* +/- MAX_INT full scale, lame.c:1938 */
    return handle_encode_T<int>(g, l, r, n, 1, norm, mp3buf, size);
}
int lame_encode_buffer_long2(lame_global_flags *g, const long l[], const long r[], const int n, unsigned char *mp3buf, const int size)
{
    float const norm = (float) (1.0 / (1L << (8 * sizeof(long) - 16)));                    /* +/- MAX_LONG full scale, lame.c:1949 */
    return handle_encode_T<long>(g, l, r, n, 1, norm, mp3buf, size);
}
int lame_encode_buffer_long(lame_global_flags *g, const long l[], const long r[], const int n, unsigned char *mp3buf, const int size)
{
    return handle_encode_T<long>(g, l, r, n, 1, 1.0f, mp3buf, size);                       /* +/- 32768 full scale, lame.c:1960 */
}
/* lame.c:2042 lame_encode_flush / lame.c:1988 lame_encode_flush_nogap: pad, encode what is left, then flush_bitstream */
static int handle_flush(lame_global_flags *g, unsigned char *mp3buf, int size)
{
    if (!ok(g) || !g->initialised) return -3;
    {
        std::unique_lock<std::mutex> lk(g->se->lane[g->lane].lm);
        Stream &x = HS(g);
        if (x.mf_samples_to_encode < 1) return 0;
        HB(g)->pad_for_flush(g->lane);
        if (handle_wait_frames(g, lk, true) != 0) return -2;
    }
    /* the lane has nothing in flight any more; the engine may (other lanes' step): finish_lane orders the end of the lane's reservoir
     * behind it, and no new step can be submitted while the engine's mutex is held */
    std::lock_guard<std::mutex> le(g->se->m);
    std::lock_guard<std::mutex> lk(g->se->lane[g->lane].lm);
    if (HB(g)->finish_lane(g->lane) != 0) return -2;
    return handle_take(g, mp3buf, size);
}
int lame_encode_flush(lame_global_flags *g, unsigned char *mp3buf, int size) { return handle_flush(g, mp3buf, size); }
int lame_encode_flush_nogap(lame_global_flags *g, unsigned char *mp3buf, int size)            /* lame.c:1988: flush_bitstream + copy_buffer; the buffered samples stay */
{
    if (!ok(g) || !g->initialised) return -3;
    if (handle_quiesce(g) != 0) return -2;
    std::lock_guard<std::mutex> le(g->se->m);
    std::lock_guard<std::mutex> lk(g->se->lane[g->lane].lm);
    if (HB(g)->finish_lane(g->lane, false) != 0) return -2;
    return handle_take(g, mp3buf, size);
}
/* lame.c:2006 lame_init_bitstream: a new file begins - statistics and the Info tag start over, the encoder state continues */
int lame_init_bitstream(lame_global_flags *g)
{
    if (!ok(g) || !g->initialised) return -3;
    std::unique_lock<std::mutex> lk(g->se->lane[g->lane].lm);
    if (handle_wait_frames(g, lk, true) != 0) return -2;
    Stream &x = HS(g);
    x.frames_out_base = x.frames_out;
    memset(x.hist_mode, 0, sizeof x.hist_mode); memset(x.hist_block, 0, sizeof x.hist_block);
    x.tag = Stream::Tag();
    if (g->write_lame_tag) tag_init(HB(g), g->lane);
    return 0;
}
/* the carried options: setter stores, getter returns what was stored (or the reference's default) */
#define LG_OPT(n, t, d, d2, h) \
    int lame_set_##n(lame_global_flags *g, t v) { if (!ok(g)) return -1; g->opt[LG_OPTI_##n] = (double) v; g->opt_set[LG_OPTI_##n] = 1; return 0; } \
    t lame_get_##n(const lame_global_flags *g) { return ok(g) ? (t) g->opt[LG_OPTI_##n] : (t) 0; }
#include "lg_api_options.inc"
#undef LG_OPT

/* version.c:55-260, set_get.c and lame.c read-only queries */
const char *get_lame_version(void) { return "3.99.5"; }
const char *get_lame_very_short_version(void) { return "LAME3.99r5"; }
const char *get_psy_version(void) { return "1.0"; }
const char *get_lame_url(void) { return "http://lame.sf.net"; }
const char *get_lame_os_bitness(void) { return sizeof(void *) == 8 ? "64bits" : "32bits"; }
int lame_get_version(const lame_global_flags *g) { return ok(g) && HB(g) ? HB(g)->cfg.version : 0; }            /* 1 = MPEG-1, 0 = MPEG-2/2.5 */
int lame_get_encoder_padding(const lame_global_flags *g) { return ok(g) && HB(g) ? HS(g).tag.enc_padding : 0; }
int lame_get_mf_samples_to_encode(const lame_global_flags *g)           /* frames that wait for their launch count as encoded, as in the reference */
{
    if (!ok(g) || !HB(g)) return 0;
    std::lock_guard<std::mutex> lk(g->se->lane[g->lane].lm);
    const Stream &x = HS(g);
    return (int) (x.mf_samples_to_encode - (long) x.fs * x.frames_ready());
}
int lame_get_totalframes(const lame_global_flags *g)                                                          /* set_get.c:2121 */
{
    if (!ok(g) || !HB(g)) return 0;
    unsigned long const fs = 576ul * HB(g)->cfg.mode_gr;
    unsigned long n = (unsigned long) g->opt[LG_OPTI_num_samples];
    if (n == (0ul - 1ul) || n == 4294967295ul) return 0;
    if (g->samplerate_in != g->samplerate_out && g->samplerate_in > 0) n *= (double) g->samplerate_out / g->samplerate_in;
    n += 576;
    unsigned long end_padding = fs - (n % fs);
    if (end_padding < 576) end_padding += fs;
    return (int) ((n + end_padding) / fs);
}
void lame_print_config(const lame_global_flags *g)
{
    if (!ok(g) || !HB(g)) return;
    const LgDevCfg *c = &HB(g)->cfg;
    fprintf(stderr, "lamegpu %s: %d Hz -> %d Hz%s, MPEG-%s Layer III, %s, quality %d, lowpass %d Hz\n", get_lame_version(), c->samplerate_in, c->samplerate,
            c->resample ? " (resampled on the device)" : "", c->version == 1 ? "1" : (c->samplerate < 16000 ? "2.5" : "2"),
            c->vbr == 0 ? "CBR" : (c->vbr == 3 ? "ABR" : (c->vbr == 2 ? "VBR (rh)" : "VBR (mtrh)")), c->quality, c->lowpassfreq);
}
void lame_print_internals(const lame_global_flags *g) { lame_print_config(g); }
/* lame.c:2234 lame_mp3_tags_fid: the finished Info tag over the placeholder frame at the start of the file (no ID3v2 to skip) */
void lame_mp3_tags_fid(lame_global_flags *g, FILE *f)
{
    if (!ok(g) || !HB(g) || !f || !HS(g).tag.on) return;
    unsigned char buf[2880];
    size_t const n = lame_get_lametag_frame(g, buf, sizeof buf);
    if (n == 0 || n > sizeof buf) return;
    if (fseek(f, 0, SEEK_SET) != 0) { fprintf(stderr, "lamegpu: could not update LAME tag, file not seekable.\n"); return; }
    if (fwrite(buf, 1, n, f) != n) fprintf(stderr, "lamegpu: could not update LAME tag.\n");
}
/* ---- the rest of the encoder's exports of include/libmp3lame.sym */
int lame_get_size_mp3buffer(const lame_global_flags *g)                 /* set_get.c:2042 over bitstream.c:801 compute_flushbits */
{
    if (!ok(g) || !g->initialised) return 0;
    std::lock_guard<std::mutex> lk(g->se->lane[g->lane].lm);
    const Stream &x = HS(g);
    const LgDevCfg *c = &HB(g)->cfg;
    int last_ptr = x.bw.h_ptr - 1;
    if (last_ptr == -1) last_ptr = LG_MAX_HEADER_BUF - 1;
    long total = x.bw.header[last_ptr].write_timing - x.bw.totbit;
    total += 8 * ((c->version + 1) * 72000 * c->bitrate_kbps[x.last_bitrate_index] / c->samplerate + x.last_padding);
    total = (total % 8) ? 1 + total / 8 : total / 8;
    return (int) (total + (long) x.out.size() + (long) x.bw.buf.size());
}
/* presets.c:320 apply_preset as lame_set_preset (set_get.c:2159) calls it on a handle whose tuning options are untouched: the preset
 * selects the rate mode and level; the tuning values it forces are the ones lame_init_params derives for that mode and level anyway */
int lame_set_preset(lame_global_flags *g, int preset)
{
    if (!ok(g)) return -1;
    g->preset = preset;
    switch (preset) {
    case 1000 /* R3MIX */: preset = 470; g->VBR = vbr_mtrh; break;
    case 1006 /* MEDIUM */: case 1007 /* MEDIUM_FAST */: preset = 460; g->VBR = vbr_mtrh; break;
    case 1001 /* STANDARD */: case 1004 /* STANDARD_FAST */: preset = 480; g->VBR = vbr_mtrh; break;
    case 1002 /* EXTREME */: case 1005 /* EXTREME_FAST */: preset = 500; g->VBR = vbr_mtrh; break;
    case 1003 /* INSANE */: g->preset = 320; g->mean_brate = 320; g->brate = 320; g->VBR = vbr_off; return 320;      /* its row scales by 1.00 */
    default: break;
    }
    g->preset = preset;
    if (preset >= 410 && preset <= 500 && preset % 10 == 0) {            /* V9 = 410 ... V0 = 500 */
        g->vbr_q = (500 - preset) / 10;
        if (g->VBR == vbr_off) g->preset_foreign = 1;                    /* VBR tuning forced onto a CBR handle: not modelled, lame_init_params says so */
        return preset;
    }
    if (8 <= preset && preset <= 320) {                                  /* apply_abr_preset */
        g->VBR = vbr_abr; g->mean_brate = preset; g->brate = preset;
        /* presets.c:294 is not guarded by `enforce`: the row's input scaling is applied here and again by lame_init_params */
        float const sc = (float) g->opt[LG_OPTI_scale] * lg_abr_preset_scale(preset);
        g->opt[LG_OPTI_scale] = sc; g->opt_set[LG_OPTI_scale] = 1;
        return preset;
    }
    g->preset = 0;
    return preset;
}
int lame_set_preset_expopts(lame_global_flags *g, int) { return ok(g) ? 0 : -1; }
/* set_get.c:1650-1850: block switching, one field behind three setters */
int lame_set_allow_diff_short(lame_global_flags *g, int v) { if (!ok(g)) return -1; g->short_blocks = v ? 0 : 1; return 0; }
int lame_get_allow_diff_short(const lame_global_flags *g) { return ok(g) ? (g->short_blocks == 0 ? 1 : 0) : 0; }
int lame_set_no_short_blocks(lame_global_flags *g, int v) { if (!ok(g) || v < 0 || v > 1) return -1; g->short_blocks = v ? 2 : 0; return 0; }
int lame_get_no_short_blocks(const lame_global_flags *g) { return ok(g) ? (g->short_blocks < 0 ? -1 : (g->short_blocks == 2 ? 1 : 0)) : -1; }
int lame_set_force_short_blocks(lame_global_flags *g, int v)
{
    if (!ok(g) || v < 0 || v > 1) return -1;
    if (v == 1) g->short_blocks = 3;
    else if (g->short_blocks == 3) g->short_blocks = 0;
    return 0;
}
int lame_get_force_short_blocks(const lame_global_flags *g) { return ok(g) ? (g->short_blocks < 0 ? -1 : (g->short_blocks == 3 ? 1 : 0)) : -1; }
void lame_set_msfix(lame_global_flags *g, double v) { if (ok(g)) g->msfix = (float) v; }          /* set_get.c:1686 */
float lame_get_msfix(const lame_global_flags *g) { return ok(g) ? g->msfix : 0; }
int lame_set_errorf(lame_global_flags *g, lame_report_function f) { if (!ok(g)) return -1; g->report[0] = f; return 0; }
int lame_set_debugf(lame_global_flags *g, lame_report_function f) { if (!ok(g)) return -1; g->report[1] = f; return 0; }
int lame_set_msgf(lame_global_flags *g, lame_report_function f) { if (!ok(g)) return -1; g->report[2] = f; return 0; }
int lame_set_asm_optimizations(lame_global_flags *g, int optim, int) { return ok(g) ? optim : -1; }
void lame_set_write_id3tag_automatic(lame_global_flags *g, int v) { if (ok(g)) g->write_id3tag_automatic = v; }
int lame_get_write_id3tag_automatic(const lame_global_flags *g) { return ok(g) ? g->write_id3tag_automatic : 1; }
int lame_init_old(lame_global_flags *) { return -1; }
void get_lame_version_numerical(lame_version_t *v)
{
    if (!v) return;
    v->major = 3; v->minor = 99; v->alpha = 0; v->beta = 0;
    v->psy_major = 1; v->psy_minor = 0; v->psy_alpha = 0; v->psy_beta = 0;
    v->features = "";
}
int lame_get_bitrate(int mpeg_version, int table_index)                  /* set_get.c: bitrate_table[version][index] */
{
    if (0 <= mpeg_version && mpeg_version <= 2 && 0 <= table_index && table_index <= 15) return lg_table_bitrate(mpeg_version, table_index);
    return -1;
}
int lame_get_samplerate(int mpeg_version, int table_index)
{
    if (0 <= mpeg_version && mpeg_version <= 2 && 0 <= table_index && table_index <= 3) return lg_table_samplerate(mpeg_version, table_index);
    return -1;
}
int lame_get_RadioGain(const lame_global_flags *) { return 0; }
int lame_get_AudiophileGain(const lame_global_flags *) { return 0; }
float lame_get_PeakSample(const lame_global_flags *) { return 0.f; }
int lame_get_noclipGainChange(const lame_global_flags *) { return 0; }
float lame_get_noclipScale(const lame_global_flags *) { return -1.f; }

/* lame.c:2145 lame_encode_finish = lame_encode_flush + lame_close */
int lame_encode_finish(lame_global_flags *g, unsigned char *mp3buf, int size)
{
    int const ret = lame_encode_flush(g, mp3buf, size);
    (void) lame_close(g);
    return ret;
}

/* lame.c:2462-2606: the statistics of the frames encoded so far (bitrates of this MPEG version, frames per bitrate and
 * stereo mode, gr.ch per bitrate and block type) */
void lame_bitrate_kbps(const lame_global_flags *g, int bitrate_kbps[14])
{
    if (!ok(g) || !HB(g)) return;
    for (int i = 0; i < 14; i++) bitrate_kbps[i] = HB(g)->cfg.bitrate_kbps[i + 1];
}
void lame_bitrate_hist(const lame_global_flags *g, int bitrate_count[14])
{
    if (!ok(g) || !HB(g)) return;
    for (int i = 0; i < 14; i++) bitrate_count[i] = HS(g).hist_mode[i + 1][4];
}
void lame_stereo_mode_hist(const lame_global_flags *g, int stmode_count[4])
{
    if (!ok(g) || !HB(g)) return;
    for (int i = 0; i < 4; i++) stmode_count[i] = HS(g).hist_mode[15][i];
}
void lame_bitrate_stereo_mode_hist(const lame_global_flags *g, int bitrate_stmode_count[14][4])
{
    if (!ok(g) || !HB(g)) return;
    for (int j = 0; j < 14; j++) for (int i = 0; i < 4; i++) bitrate_stmode_count[j][i] = HS(g).hist_mode[j + 1][i];
}
void lame_block_type_hist(const lame_global_flags *g, int btype_count[6])
{
    if (!ok(g) || !HB(g)) return;
    for (int i = 0; i < 6; i++) btype_count[i] = HS(g).hist_block[15][i];
}
void lame_bitrate_block_type_hist(const lame_global_flags *g, int bitrate_btype_count[14][6])
{
    if (!ok(g) || !HB(g)) return;
    for (int j = 0; j < 14; j++) for (int i = 0; i < 6; i++) bitrate_btype_count[j][i] = HS(g).hist_block[j + 1][i];
}
/* VbrTag.c:900 lame_get_lametag_frame (+ :151 Xing_seek_table, :597 PutLameVBR) for the configurations this library
 * encodes: CBR ("Info"), MPEG-1, no CRC, no ReplayGain analysis, no nogap. */
size_t lame_get_lametag_frame(const lame_global_flags *g, unsigned char *buffer, size_t size)
{
    if (!ok(g) || !g->initialised || !HB(g)) return 0;
    const Stream &x = HS(g);
    const LgDevCfg *c = &HB(g)->cfg;
    const Stream::Tag &v = x.tag;
    if (!v.on) return 0;
    if (v.pos <= 0) return 0;
    if (size < (size_t) v.frame_size) return (size_t) v.frame_size;
    if (!buffer) return 0;
    memset(buffer, 0, (size_t) v.frame_size);
    tag_frame_header(c, v.mode_ext, buffer);
    unsigned char toc[100];
    memset(toc, 0, sizeof toc);
    for (int i = 1; i < 100; ++i) {
        float j = i / (float) 100, act, sum;
        int indx = (int) (floor(j * v.pos));
        if (indx > v.pos - 1) indx = v.pos - 1;
        act = (float) v.bag[indx];
        sum = (float) v.sum;
        int seek_point = (int) (256. * act / sum);
        if (seek_point > 255) seek_point = 255;
        toc[i] = (unsigned char) seek_point;
    }
    unsigned n = (unsigned) c->sideinfo_len;
    if (c->error_protection) n -= 2;                        /* VbrTag.c:955-961: the Xing data keeps its offset */
    memcpy(buffer + n, c->vbr == 0 ? "Info" : "Xing", 4); n += 4;
    put_be32(buffer + n, 1 + 2 + 4 + 8); n += 4;            /* FRAMES_FLAG + BYTES_FLAG + TOC_FLAG + VBR_SCALE_FLAG */
    put_be32(buffer + n, (unsigned long) v.nframes); n += 4;
    unsigned long const stream_size = (unsigned long) (v.nbytes + v.frame_size);
    put_be32(buffer + n, stream_size); n += 4;
    memcpy(buffer + n, toc, sizeof toc); n += sizeof toc;
    if (c->error_protection) lg_header_crc(buffer, c->sideinfo_len);     /* VbrTag.c:996 */
    unsigned short crc = 0;
    for (unsigned i = 0; i < n; i++) crc = crc16_update(buffer[i], crc);
    /* PutLameVBR */
    unsigned char *p = buffer + n;
    int k = 0;
    int nQuality = 100 - 10 * g->vbr_q /* default 4, lame.c:2360 */ - g->quality;
    if (nQuality < 0) nQuality = 0;
    double const lp = c->lowpassfreq / 100.0 + .5;
    unsigned char const nLowpass = (unsigned char) (lp > 255 ? 255 : lp);
    unsigned char const nFlags = (unsigned char) (c->athtype + (1 << 4) + ((c->use_safe_joint_stereo != 0) << 5));
    int nStereoMode;
    switch (c->mode) {
    case LG_MONO: nStereoMode = 0; break;
    case LG_STEREO: nStereoMode = 1; break;
    case LG_DUAL: nStereoMode = 2; break;
    case LG_JOINT: nStereoMode = c->force_ms ? 4 : 3; break;
    default: nStereoMode = 7; break;
    }
    int nSourceFreq;
    if (c->samplerate_in <= 32000) nSourceFreq = 0;
    else if (c->samplerate_in == 48000) nSourceFreq = 2;
    else if (c->samplerate_in > 48000) nSourceFreq = 3;
    else nSourceFreq = 1;
    int const bNonOptimal = (c->short_blocks == 2 /* forced */ || c->short_blocks == 3 /* dispensed */ ||
                             (c->disable_reservoir && c->vbr_mean_kbps < 320) || c->athtype == 0 || c->samplerate_in <= 32000);
    unsigned char const nMisc = (unsigned char) (c->noise_shaping + (nStereoMode << 2) + (bNonOptimal << 5) + (nSourceFreq << 6));
    put_be32(p + k, (unsigned long) nQuality); k += 4;
    memcpy(p + k, "LAME3.99r", 9); k += 9;                  /* get_lame_tag_encoder_short_version(), version.c:148 */
    {   /* revision 0 + the method: vbr_mode numbered the Lame tag's way (VbrTag.c:646) */
        static const unsigned char vbr_type_translator[7] = { 1, 5, 3, 2, 4, 0, 3 };
        p[k++] = (unsigned) g->VBR < 7u ? vbr_type_translator[g->VBR] : 0;
    }
    p[k++] = nLowpass;
    put_be32(p + k, 0); k += 4;                             /* peak signal amplitude: no ReplayGain analysis */
    put_be16(p + k, 0); k += 2;
    put_be16(p + k, 0); k += 2;
    p[k++] = nFlags;
    {   /* "if ABR, {store bitrate <= 255} else {store -b}": the VBR modes store the minimal bitrate (VbrTag.c:672-686) */
        int const nABRBitrate = (c->vbr == 4 || c->vbr == 2) ? c->bitrate_kbps[c->vbr_min_bitrate_index] : c->vbr_mean_kbps;
        p[k++] = (unsigned char) (nABRBitrate >= 255 ? 0xFF : nABRBitrate);
    }
    int const enc_delay = 576, enc_padding = v.enc_padding;
    p[k] = (unsigned char) (enc_delay >> 4);
    p[k + 1] = (unsigned char) ((enc_delay << 4) + (enc_padding >> 8));
    p[k + 2] = (unsigned char) enc_padding;
    k += 3;
    p[k++] = nMisc;
    p[k++] = 0;
    put_be16(p + k, (unsigned) c->preset); k += 2;           /* cfg->preset: what apply_preset was called with, presets.c:361 */
    put_be32(p + k, stream_size); k += 4;
    put_be16(p + k, v.music_crc); k += 2;
    for (int i = 0; i < k; i++) crc = crc16_update(p[i], crc);
    put_be16(p + k, crc);
    return (size_t) v.frame_size;
}

int lame_close(lame_global_flags *g)
{
    if (!ok(g)) return -3;
    if (g->se) { (void) handle_quiesce(g); shared_release(g->se, g->lane); }
    g->se = nullptr;
    g->class_id = 0;
    free(g);
    return 0;
}

} // extern "C"
