/* lamegpu.h - C ABI of the B200 batched MP3 encode library (liblamegpu.so).
 *
 * Two faces:
 *
 *  1. The libmp3lame entry points of the hot path, with the reference's names, argument meaning and
 *     error codes, so an application written against include/lame.h of LAME 3.99.5 links unchanged.
 *     Each declaration cites the reference declaration it replaces.  A handle is a lane of a GPU batch
 *     engine: lame_encode_buffer() queues PCM and returns whatever finished bytes exist (the reference
 *     already documents that the return value "can be 0", lame.h:687); lame_encode_flush() drains.
 *
 *  2. An additive batch interface (lamegpu_batch_*) that feeds many independent streams per call - the
 *     form a GPU needs and what bench.py measures.  Plain pointers and sizes only.
 *
 * Scope of this build: MPEG-1 Layer III, 32/44.1/48 kHz, CBR 112..320 kbps (no resampling), stereo /
 * joint stereo / mono, quality 3..9.  Anything else makes lame_init_params()/lamegpu_batch_open() fail
 * with -1 - there is no silent fallback and no CPU path.
 */
#ifndef LAMEGPU_H
#define LAMEGPU_H
#include <stddef.h>
#include <stdio.h>
#include <stdarg.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ 1. libmp3lame-compatible face */
struct lame_global_struct;
typedef struct lame_global_struct lame_global_flags;          /* include/lame.h:146-147 */
typedef lame_global_flags *lame_t;

typedef enum vbr_mode_e { vbr_off = 0, vbr_mt, vbr_rh, vbr_abr, vbr_mtrh, vbr_max_indicator, vbr_default = vbr_mtrh } vbr_mode; /* lame.h:49-57 */
typedef enum MPEG_mode_e { STEREO = 0, JOINT_STEREO, DUAL_CHANNEL, MONO, NOT_SET, MAX_INDICATOR } MPEG_mode;               /* lame.h:61-68 */
typedef enum Padding_type_e { PAD_NO = 0, PAD_ALL, PAD_ADJUST, PAD_MAX_INDICATOR } Padding_type;                             /* lame.h:72-77 */
typedef void (*lame_report_function)(const char *format, va_list ap);                                                        /* lame.h:38 */
typedef struct {                                                                                                             /* lame.h:655-669 */
    int major, minor, alpha, beta;
    int psy_major, psy_minor, psy_alpha, psy_beta;
    const char *features;
} lame_version_t;

lame_global_flags *lame_init(void);                                                  /* lame.h:168  NULL on OOM */
int  lame_set_in_samplerate(lame_global_flags *, int);                               /* lame.h:188 */
int  lame_get_in_samplerate(const lame_global_flags *);                              /* lame.h:189 */
int  lame_set_num_channels(lame_global_flags *, int);                                /* lame.h:192 */
int  lame_get_num_channels(const lame_global_flags *);                               /* lame.h:193 */
int  lame_set_out_samplerate(lame_global_flags *, int);                              /* lame.h:224 */
int  lame_get_out_samplerate(const lame_global_flags *);                             /* lame.h:225 */
int  lame_set_brate(lame_global_flags *, int);                                       /* lame.h:353 */
int  lame_get_brate(const lame_global_flags *);                                      /* lame.h:354 */
int  lame_set_quality(lame_global_flags *, int);                                     /* lame.h:263 */
int  lame_get_quality(const lame_global_flags *);                                    /* lame.h:264 */
int  lame_set_mode(lame_global_flags *, MPEG_mode);                                  /* lame.h:270 */
MPEG_mode lame_get_mode(const lame_global_flags *);                                  /* lame.h:271 */
int  lame_set_VBR(lame_global_flags *, vbr_mode);                                    /* lame.h:432  only vbr_off is accepted */
vbr_mode lame_get_VBR(const lame_global_flags *);
int  lame_set_VBR_mean_bitrate_kbps(lame_global_flags *, int);                       /* lame.h:444  ABR mean bitrate */
int  lame_get_VBR_mean_bitrate_kbps(const lame_global_flags *);                      /* lame.h:445 */
int  lame_set_VBR_q(lame_global_flags *, int);                                       /* lame.h:436  VBR quality 0..9 (vbr_mtrh: 0..6 here) */
int  lame_get_VBR_q(const lame_global_flags *);                                      /* lame.h:437 */
int  lame_set_VBR_quality(lame_global_flags *, float);                               /* lame.h:438  fractional level 0 .. 9.999 */
float lame_get_VBR_quality(const lame_global_flags *);                               /* lame.h:439 */
int  lame_set_VBR_mean_bitrate_kbps(lame_global_flags *, int);                       /* lame.h:447  ABR mean bitrate */
int  lame_get_VBR_mean_bitrate_kbps(const lame_global_flags *);                                    /* lame.h:433 */
int  lame_set_bWriteVbrTag(lame_global_flags *, int);                                /* lame.h:240  the Info tag frame is not produced */
int  lame_get_bWriteVbrTag(const lame_global_flags *);                               /* lame.h:241 */
int  lame_init_params(lame_global_flags *);                                          /* lame.h:636  <0 on error/unsupported */
int  lame_get_framesize(const lame_global_flags *);                                  /* lame.h:582 */
int  lame_get_frameNum(const lame_global_flags *);                                   /* lame.h:597 */
int  lame_get_encoder_delay(const lame_global_flags *);                              /* lame.h:571 */
int  lame_encode_buffer(lame_global_flags *, const short int pcm_l[], const short int pcm_r[], const int nsamples,
                        unsigned char *mp3buf, const int mp3buf_size);               /* lame.h:715 */
int  lame_encode_buffer_interleaved(lame_global_flags *, short int pcm[], int num_samples,
                                    unsigned char *mp3buf, int mp3buf_size);         /* lame.h:730 */
int  lame_encode_buffer_ieee_float(lame_t, const float pcm_l[], const float pcm_r[], const int nsamples,
                                   unsigned char *mp3buf, const int mp3buf_size);    /* lame.h:758  +/-1.0 full scale */
/* the other sample types of the same entry point (lame.c:1839 lame_encode_buffer_template) */
int  lame_encode_buffer_float(lame_global_flags *, const float pcm_l[], const float pcm_r[], const int nsamples,
                              unsigned char *mp3buf, const int mp3buf_size);         /* lame.h:746  +/-32768 full scale */
int  lame_encode_buffer_interleaved_ieee_float(lame_t, const float pcm[], const int nsamples,
                                   unsigned char *mp3buf, const int mp3buf_size);    /* lame.h:765 */
int  lame_encode_buffer_ieee_double(lame_t, const double pcm_l[], const double pcm_r[], const int nsamples,
                                    unsigned char *mp3buf, const int mp3buf_size);   /* lame.h:776 */
int  lame_encode_buffer_interleaved_ieee_double(lame_t, const double pcm[], const int nsamples,
                                    unsigned char *mp3buf, const int mp3buf_size);   /* lame.h:783 */
int  lame_encode_buffer_long(lame_global_flags *, const long pcm_l[], const long pcm_r[], const int nsamples,
                             unsigned char *mp3buf, const int mp3buf_size);          /* lame.h:799  +/-32768 full scale */
int  lame_encode_buffer_long2(lame_global_flags *, const long pcm_l[], const long pcm_r[], const int nsamples,
                              unsigned char *mp3buf, const int mp3buf_size);         /* lame.h:813  +/-MAX_LONG full scale */
int  lame_encode_buffer_int(lame_global_flags *, const int pcm_l[], const int pcm_r[], const int nsamples,
                            unsigned char *mp3buf, const int mp3buf_size);           /* lame.h:831  +/-MAX_INT full scale */
int  lame_encode_flush(lame_global_flags *, unsigned char *mp3buf, int size);        /* lame.h:856 */
int  lame_encode_flush_nogap(lame_global_flags *, unsigned char *mp3buf, int size);  /* lame.h:878  flush_bitstream without the end-of-input padding */
int  lame_init_bitstream(lame_global_flags *);                                       /* lame.h:890  a new file after a nogap flush: statistics and tag start over */
int  lame_get_size_mp3buffer(const lame_global_flags *);                             /* lame.h:594  bytes a nogap flush would hand out now */
int  lame_set_preset(lame_global_flags *, int preset);                               /* lame.h:359  V0..V9, 8..320 (ABR), STANDARD / EXTREME / MEDIUM / INSANE / R3MIX */
int  lame_set_preset_expopts(lame_global_flags *, int);                              /* lame.h:1304 obsolete: does nothing */
int  lame_set_errorf(lame_global_flags *, lame_report_function);                     /* lame.h:346 */
int  lame_set_debugf(lame_global_flags *, lame_report_function);                     /* lame.h:347 */
int  lame_set_msgf(lame_global_flags *, lame_report_function);                       /* lame.h:348 */
int  lame_set_no_short_blocks(lame_global_flags *, int);                             /* lame.h:399  three views of one field, set_get.c:1650-1850 */
int  lame_get_no_short_blocks(const lame_global_flags *);
int  lame_set_force_short_blocks(lame_global_flags *, int);                          /* lame.h:403 */
int  lame_get_force_short_blocks(const lame_global_flags *);
int  lame_set_allow_diff_short(lame_global_flags *, int);                            /* lame.h:392 */
int  lame_get_allow_diff_short(const lame_global_flags *);
void lame_set_msfix(lame_global_flags *, double);                                    /* lame.h:424 */
float lame_get_msfix(const lame_global_flags *);                                     /* lame.h:425 */
int  lame_set_asm_optimizations(lame_global_flags *, int optim, int mode);           /* lame.h:334  no CPU SIMD paths here: accepted, no effect */
void lame_set_write_id3tag_automatic(lame_global_flags *, int);                      /* lame.h:1266 */
int  lame_get_write_id3tag_automatic(const lame_global_flags *);                     /* lame.h:1267 */
int  lame_init_old(lame_global_flags *);                                             /* lame.h:1278 obsolete initialiser of a caller-allocated struct: -1 */
void get_lame_version_numerical(lame_version_t *);                                   /* lame.h:671 */
int  lame_get_bitrate(int mpeg_version, int table_index);                            /* lame.h:1308 */
int  lame_get_samplerate(int mpeg_version, int table_index);                         /* lame.h:1309 */
/* ReplayGain analysis is outside the hot path (off by default in the library, lame.c:2386): the getters
 * report "not computed" as the reference does without the analysis */
int  lame_get_RadioGain(const lame_global_flags *);                                  /* lame.h:610 */
int  lame_get_AudiophileGain(const lame_global_flags *);                             /* lame.h:613 */
float lame_get_PeakSample(const lame_global_flags *);                                /* lame.h:616 */
int  lame_get_noclipGainChange(const lame_global_flags *);                           /* lame.h:620 */
float lame_get_noclipScale(const lame_global_flags *);                               /* lame.h:625 */
/* Info tag (CBR): lame_set_bWriteVbrTag(1) - the reference's default - puts the all-zero placeholder frame ahead of
 * the audio; after lame_encode_flush this returns the finished tag frame to be written at offset 0 (VbrTag.c:900) */
size_t lame_get_lametag_frame(const lame_global_flags *, unsigned char *buffer, size_t size); /* lame.h:970 */
/* statistics of the frames encoded so far (encoder.c:156 updateStats), same layouts as the reference */
void lame_bitrate_kbps(const lame_global_flags *, int bitrate_kbps[14]);                         /* lame.h:912 */
void lame_bitrate_hist(const lame_global_flags *, int bitrate_count[14]);                        /* lame.h:909 */
void lame_stereo_mode_hist(const lame_global_flags *, int stereo_mode_count[4]);                 /* lame.h:915 */
void lame_bitrate_stereo_mode_hist(const lame_global_flags *, int bitrate_stmode_count[14][4]);  /* lame.h:919 */
void lame_block_type_hist(const lame_global_flags *, int btype_count[6]);                        /* lame.h:923 */
void lame_bitrate_block_type_hist(const lame_global_flags *, int bitrate_btype_count[14][6]);    /* lame.h:927 */
int  lame_close(lame_global_flags *);                                                /* lame.h:977 */
const char *get_lame_short_version(void);                                            /* lame.h:645 */
const char *get_lame_version(void);                                                  /* lame.h:644 */
const char *get_lame_very_short_version(void);                                       /* lame.h:646 */
const char *get_psy_version(void);                                                   /* lame.h:647 */
const char *get_lame_url(void);                                                      /* lame.h:648 */
const char *get_lame_os_bitness(void);                                               /* lame.h:649 */
int  lame_get_version(const lame_global_flags *);                                    /* lame.h:568  1 = MPEG-1, 0 = MPEG-2/2.5 */
int  lame_get_encoder_padding(const lame_global_flags *);                            /* lame.h:579 */
int  lame_get_mf_samples_to_encode(const lame_global_flags *);                       /* lame.h:585 */
int  lame_get_totalframes(const lame_global_flags *);                                /* lame.h:603 */
void lame_print_config(const lame_global_flags *);                                   /* lame.h:678 */
void lame_print_internals(const lame_global_flags *);                                /* lame.h:680 */
void lame_mp3_tags_fid(lame_global_flags *, FILE *fid);                              /* lame.h:950 */
int  lame_encode_finish(lame_global_flags *, unsigned char *mp3buf, int size);       /* lame.h:988 */
#include "lamegpu_options.h"                                                         /* the remaining lame_set_X / lame_get_X pairs */

/* ------------------------------------------------------------------ 2. batch face */
typedef struct lamegpu_batch lamegpu_batch;

/* One engine for `nstreams` independent streams that share a configuration.  `frames_per_launch` is how
 * many frames of every stream one GPU launch covers (the reservoir makes a stream's frames sequential,
 * so parallelism comes from the number of streams).  Returns NULL on unsupported configuration, missing
 * GPU or allocation failure (a message goes to stderr). */
lamegpu_batch *lamegpu_batch_open(int samplerate, int channels, int brate, int mode /* MPEG_mode or -1 */,
                                  int quality /* 0..9 or -1 */, int nstreams, int frames_per_launch, int device);
/* same, with the rate mode: vbr = 0 (vbr_off, CBR at `brate`) or 3 (vbr_abr, `brate` is the mean bitrate of
 * lame_set_VBR_mean_bitrate_kbps; quantize.c:1900 ABR_iteration_loop) or 4 (vbr_mtrh, `brate` is VBR_q 0..6;
 * quantize.c:1645 VBR_new_iteration_loop + vbrquantize.c) */
lamegpu_batch *lamegpu_batch_open_ex(int samplerate, int channels, int brate, int mode, int quality, int vbr,
                                     int nstreams, int frames_per_launch, int device);
/* same, with the input rate apart from the MPEG output rate (lame_set_in_samplerate / lame_set_out_samplerate,
 * lame.h:184,196).  samplerate_out = 0 picks the output rate the way lame_init_params does (lame.c:764-769
 * optimum_samplefreq); it must come out as 32000, 44100 or 48000.  When the two differ the streams' samples go
 * through the reference's polyphase resampler (util.c:531 fill_buffer_resample) on the device. */
lamegpu_batch *lamegpu_batch_open_rs(int samplerate_in, int samplerate_out, int channels, int brate, int mode, int quality,
                                     int vbr, int nstreams, int frames_per_launch, int device);
/* same with a fractional rate argument: for vbr = 4 `rate` is the VBR quality 0 .. 9.999 of lame_set_VBR_quality (lame.h:447),
 * otherwise the bitrate as above.  Without an explicit output rate the level is mapped as lame_init_params does it
 * (lame.c:661-698): e.g. 7.0 at 44.1 kHz encodes at 32 kHz with internal quality 5.63. */
lamegpu_batch *lamegpu_batch_open_vq(int samplerate_in, int samplerate_out, int channels, float rate, int mode, int quality,
                                     int vbr, int nstreams, int frames_per_launch, int device);
void lamegpu_batch_close(lamegpu_batch *b);
/* device = -1 in any lamegpu_batch_open*: the batch spans ALL visible GPUs of the process (LAMEGPU_DEVICES limits their number) - stream s
 * lives on device s * n / nstreams, one engine and one host thread per device, no exchange between devices (streams are independent,
 * SURVEY.md section 8e).  Every lamegpu_batch_* call then works on all devices at once.  Returns the number of devices of the batch. */
int  lamegpu_batch_devices(const lamegpu_batch *b);

/* Feed nsamples[i] samples to stream i (pcm_l[i]/pcm_r[i]; pcm_r may be NULL for mono) and encode every
 * frame that became complete.  Bytes produced for stream i are appended at out[i] (capacity out_cap[i]);
 * out_bytes[i] receives the count.  Bytes that do not fit stay queued for the next call.
 * Returns the number of frames encoded over all streams, or <0 on error. */
long lamegpu_batch_encode(lamegpu_batch *b, const short *const *pcm_l, const short *const *pcm_r, const int *nsamples,
                          unsigned char *const *out, const int *out_cap, int *out_bytes);

/* Same contract as lame_encode_flush for every stream: pads, encodes the last frames, drains the bit
 * reservoir into ancillary data. */
long lamegpu_batch_flush(lamegpu_batch *b, unsigned char *const *out, const int *out_cap, int *out_bytes);

/* contiguous convenience form used by language bindings: pcm is [nstreams][2][nsamples] int16, out is
 * [nstreams][out_stride] bytes */
long lamegpu_batch_encode_packed(lamegpu_batch *b, const short *pcm, int nsamples, unsigned char *out, int out_stride, int *out_bytes);
long lamegpu_batch_flush_packed(lamegpu_batch *b, unsigned char *out, int out_stride, int *out_bytes);

/* Pipelined mode (off by default).  A launch ("step") runs on one of two buffer sets; with pipelining on, lamegpu_batch_encode*
 * leaves its newest step in flight when it returns - the bytes of that step come out of the NEXT call (or the flush) - so that
 * the host work of a call (staging the next step, splicing the previous one) and the copies run while the device encodes.
 * The concatenated output per stream is the same either way; like lame_encode_buffer, a call may return 0 bytes.
 * Turning it off completes what is in flight.  0 = ok. */
int  lamegpu_batch_set_pipelined(lamegpu_batch *b, int on);

/* measurement hooks (bench.py): lamegpu_batch_stage_packed lays a ring of 2 x nframes frames per stream (pcm = [S][2][2 * nframes * 1152]) into
 * the engine's two buffer sets so that alternating them walks a periodic signal without a seam (one full step per buffer set, then the
 * streams are reset); lamegpu_batch_run_device_steps runs `steps` device-only steps on them back to back - persistent streams,
 * consecutive steps overlapping as in production - and returns the device time per step in ms (first kernel's start to last kernel's
 * end, CUDA events; < 0 on error); lamegpu_batch_rerun_device = one step on freshly reset streams, after which lamegpu_batch_kernel_ms
 * gives that step's per-kernel times [analysis, scan, mdct, quantise, pack] and lamegpu_batch_step_ms its first-kernel-to-last time */
int  lamegpu_batch_rerun_device(lamegpu_batch *b, int nframes);
float lamegpu_batch_run_device_steps(lamegpu_batch *b, int nframes, int steps);
int  lamegpu_batch_stage_packed(lamegpu_batch *b, const short *pcm, int nframes);
int  lamegpu_batch_kernel_ms(const lamegpu_batch *b, float ms[5]);
float lamegpu_batch_step_ms(const lamegpu_batch *b);
long lamegpu_batch_kernel_launches(const lamegpu_batch *b);
int  lamegpu_batch_set_threads(lamegpu_batch *b, int nthreads);
long lamegpu_batch_debug_copy(lamegpu_batch *b, int what, void *dst, size_t cap);   /* tests: intermediate device buffers */
size_t lamegpu_sizeof_granule_out(void);
long lamegpu_batch_d2h_bytes(const lamegpu_batch *b);              /* device->host bytes per launch */
size_t lamegpu_sizeof_analysis(void);
/* tests: the device restatements of the reference's run-time libm calls (csrc/lg_math.cuh) on n arguments;
 * fn 0 powf(x, y), 1 log10f(x), 2 exp(x), 3 pow(x, y); floats travel widened to double.  0 = ok, -1 = no CUDA device */
int  lamegpu_math_selftest(int fn, const double *x, const double *y, double *out, int n);

#ifdef __cplusplus
}
#endif
#endif
