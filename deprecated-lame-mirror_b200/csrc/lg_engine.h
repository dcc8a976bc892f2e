// lg_engine.h - internal C interface between the device engine (lg_engine.cu) and the host side
// (lg_api.cpp, lg_bitstream.cpp).  Not part of the public ABI (see include/lamegpu.h).
//
// A step (one launch over S streams x up to F frames) runs on one of lg_engine_slots() slots: stage the slot's pinned input buffers,
// lg_engine_submit() (asynchronous), later lg_engine_wait() and read the slot's result buffers.  Steps must be submitted in stream
// order (slot 0, 1, 0, 1, ...: the per-stream state is carried on the device from step to step); a slot's buffers may be restaged
// once its previous step has been waited for.
#pragma once
#include <stddef.h>
#include "lg_types.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct lg_engine lg_engine;
/* libmp3lame options that only change host-computed constants of the configuration (the lame_set_* of the same name), with
 * the values a fresh lame_init() handle has: "not set" is what lame_init_params and the presets test for (presets.c:34 SET_OPTION) */
typedef struct LgSetupOpt {
    float scale, scale_left, scale_right;       /* 1: lame.c:1196-1206 pcm_transform */
    float compression_ratio;                    /* 0 = not set: lame.c:629 */
    int   lowpassfreq, lowpasswidth;            /* Hz; 0 = chosen from the rate, -1 = no filter; width -1 = default (lame.c:704, :878) */
    int   highpassfreq, highpasswidth;          /* Hz; 0 / -1 = none; width -1 = default (lame.c:861) */
    int   no_ath;                               /* quantize_pvt.c:294: ATH at -200 dB */
    int   ath_only, ath_short;                  /* only reach the Info tag's "non-optimal" flag in 3.99.5 (VbrTag.c:789) */
    int   ath_type;                             /* -1 = default: 4, the VBR-new presets force 5 (presets.c:180) */
    float ath_lower_db;                         /* lame_set_ATHlower; 0 = the preset's */
    float ath_curve;                            /* -1 = the preset's */
    int   athaa_type;                           /* -1 = default (3) */
    float athaa_sensitivity;                    /* 0 = the preset's */
    float msfix;                                /* -1 = the preset's */
    float interch;                              /* lame_set_interChRatio; -1 = the preset's */
    int   short_blocks;                         /* -1 not set, 0 allowed, 1 coupled, 2 dispensed (no_short_blocks), 3 forced (lame_global_flags.h:20) */
    int   force_ms, disable_reservoir;
    int   strict_iso;                           /* 0 MDB_DEFAULT, 1 MDB_STRICT_ISO, 2 MDB_MAXIMUM (lame_init: 2); bitstream.c:91 */
    int   use_temporal;                         /* -1 = default: on, off for VBR-new */
    int   vbr_min_kbps, vbr_max_kbps, vbr_hard_min;   /* 0 = not set: lame.c:1067-1093, quantize.c:1550 */
} LgSetupOpt;
void lg_setup_opt_defaults(LgSetupOpt *o);
int lg_setup_ex(LgDevCfg *c, int samplerate_in, int samplerate_out, int channels, int brate, int mode, int quality, int vbr, float vbr_q_frac,
                const LgSetupOpt *opt);
int lg_setup(LgDevCfg *c, int samplerate_in, int samplerate_out /* 0 = as lame_init_params picks it */, int channels, int brate, int mode, int quality,
             int vbr /* 0 vbr_off, 3 vbr_abr, 4 vbr_mtrh (brate = VBR_q) */, float vbr_q_frac /* VBR quality = VBR_q + this */);
float lg_abr_preset_scale(int kbps);
int lg_table_bitrate(int version, int index);
int lg_table_samplerate(int version, int index);
int lg_device_count(void);
lg_engine *lg_engine_create(const LgDevCfg *cfg, int nstreams, int max_frames, int device);
void lg_engine_destroy(lg_engine *e);
int  lg_engine_reset_streams(lg_engine *e, int first, int count);
int  lg_engine_end_reservoir(lg_engine *e, const int *streams, const int *ancillary_flags, int n);
int  lg_engine_need_native_pcm(lg_engine *e, int esz);
int  lg_engine_reserve_chunks(lg_engine *e, int slot, int per_stream);
int  lg_engine_slots(const lg_engine *e);
int  lg_engine_device(const lg_engine *e);
int  lg_engine_submit(lg_engine *e, int slot, int nframes, int mode);      /* mode: 0 int16 window, 1 resampled, 2 native sample types */
int  lg_engine_run_device(lg_engine *e, int slot, int nframes, int mode);
/* the same as lg_engine_submit for a step none of whose streams has frames in the step in flight in the other slot: its quantiser need not
 * wait for that step's */
int  lg_engine_submit_independent(lg_engine *e, int slot, int nframes, int mode);
int  lg_engine_wait(lg_engine *e, int slot);
int  lg_engine_in_flight(const lg_engine *e, int slot);
int  lg_engine_mark(lg_engine *e, int which);
float lg_engine_marked_ms(lg_engine *e);
const LgDevCfg *lg_engine_config(const lg_engine *e);
int  lg_engine_streams(const lg_engine *e);
int  lg_engine_max_frames(const lg_engine *e);
size_t lg_engine_pcm_stride(const lg_engine *e);
size_t lg_engine_raw_stride(const lg_engine *e);
size_t lg_engine_pay_stride(const lg_engine *e);
/* slot buffers (pinned host memory) */
int16_t *lg_engine_host_pcm16(lg_engine *e, int slot);
void *lg_engine_host_pcmn(lg_engine *e, int slot);
LgPcmKind *lg_engine_host_kinds(lg_engine *e, int slot);
int  lg_engine_native_esz(const lg_engine *e);
int *lg_engine_host_nfr(lg_engine *e, int slot);
void *lg_engine_host_raw(lg_engine *e, int slot);
int  lg_engine_raw_esz(const lg_engine *e);
int  lg_engine_need_raw(lg_engine *e, int esz);
LgRsChunk *lg_engine_host_chunks(lg_engine *e, int slot);
int *lg_engine_host_rs_counts(lg_engine *e, int slot);          /* per stream: { nchunks, win_n } */
int  lg_engine_chunk_cap(const lg_engine *e, int slot);
const unsigned char *lg_engine_host_pay(const lg_engine *e, int slot);
const unsigned char *lg_engine_host_hdr(const lg_engine *e, int slot);
const LgFrameOut *lg_engine_host_fout(const lg_engine *e, int slot);
const float *lg_engine_last_kernel_ms(const lg_engine *e);
long lg_engine_launch_count(const lg_engine *e);
long lg_engine_debug_copy(lg_engine *e, int what, void *dst, size_t cap);
#ifdef __cplusplus
}
#endif
