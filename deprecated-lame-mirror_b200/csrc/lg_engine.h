// lg_engine.h - internal C interface between the device engine (lg_engine.cu) and the host side
// (lg_api.cpp, lg_bitstream.cpp).  Not part of the public ABI (see include/lamegpu.h).
#pragma once
#include <stddef.h>
#include "lg_types.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct lg_engine lg_engine;
int lg_setup(LgDevCfg *c, int samplerate_in, int samplerate_out /* 0 = as lame_init_params picks it */, int channels, int brate, int mode, int quality,
             int vbr /* 0 vbr_off, 3 vbr_abr, 4 vbr_mtrh (brate = VBR_q) */, float vbr_q_frac /* VBR quality = VBR_q + this */);
lg_engine *lg_engine_create(const LgDevCfg *cfg, int nstreams, int max_frames, int device);
void lg_engine_destroy(lg_engine *e);
int  lg_engine_reset_streams(lg_engine *e, int first, int count);
int  lg_engine_need_float_pcm(lg_engine *e);
int  lg_engine_reserve_chunks(lg_engine *e, int per_stream);
size_t lg_engine_raw_stride(const lg_engine *e);
float *lg_engine_host_raw(lg_engine *e);
LgRsChunk *lg_engine_host_chunks(lg_engine *e);
int *lg_engine_host_rs_counts(lg_engine *e);          /* per stream: { nchunks, win_n } */
int  lg_engine_chunk_cap(const lg_engine *e);
int  lg_engine_encode(lg_engine *e, int nframes, int use_float);
int  lg_engine_run_device(lg_engine *e, int nframes, int use_float);
int  lg_engine_sync(lg_engine *e);
const LgDevCfg *lg_engine_config(const lg_engine *e);
int  lg_engine_streams(const lg_engine *e);
int  lg_engine_max_frames(const lg_engine *e);
size_t lg_engine_pcm_stride(const lg_engine *e);
int16_t *lg_engine_host_pcm16(lg_engine *e);
float *lg_engine_host_pcmf(lg_engine *e);
int *lg_engine_host_nfr(lg_engine *e);
const unsigned char *lg_engine_host_pay(const lg_engine *e);
const unsigned char *lg_engine_host_hdr(const lg_engine *e);
size_t lg_engine_pay_stride(const lg_engine *e);
const LgFrameOut *lg_engine_host_fout(const lg_engine *e);
const float *lg_engine_last_kernel_ms(const lg_engine *e);
long lg_engine_launch_count(const lg_engine *e);
#ifdef __cplusplus
}
#endif
#ifdef __cplusplus
extern "C" {
#endif
long lg_engine_debug_copy(lg_engine *e, int what, void *dst, size_t cap);
#ifdef __cplusplus
}
#endif
