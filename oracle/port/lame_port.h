/* oracle/port - TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of the LAME 3.99.5 encode hot path (MPEG-1 Layer III,
 * 32/44.1/48 kHz output from any input rate, CBR, ABR and VBR-new (-V0..-V6), stereo / joint stereo / mono), written
 * from the algorithm's description in the reference sources, each function citing the reference
 * file:line it follows.  Its only job is to be the CPU checker for the CUDA path (tests/, smoke(),
 * bench.py cpu_baseline).  Parity of this port is PINNED: tests/test_port_vs_ref.py compares its MP3
 * bytes and per-stage state with the unmodified reference compiled into oracle/_ref/ (see
 * oracle/Makefile), on testcase.wav and on seeded synthetic inputs.
 *
 * Floating point: the reference is compiled -O2 -fno-fast-math -ffp-contract=off, FLOAT=float; this
 * port must be compiled with the same flags.  Every float/double promotion of the reference is
 * reproduced explicitly (comments say "double" where the reference's expression is evaluated in
 * double). */
#ifndef LAME_PORT_H
#define LAME_PORT_H
#include <stdint.h>

#define LP_CBANDS 64
#define LP_SBMAX_L 22
#define LP_SBMAX_S 13
#define LP_SBPSY_L 21
#define LP_SBPSY_S 12
#define LP_SFBMAX 39
#define LP_BLK 1024
#define LP_BLK_S 256
#define LP_HBLK 513
#define LP_HBLK_S 129
#define LP_PRECALC 8208
#define LP_QMAX 257
#define LP_QMAX2 116
#define LP_IXMAX 8206
#define LP_LARGE_BITS 100000
#define LP_MAX_BITS_PER_CHANNEL 4095
#define LP_MAX_BITS_PER_GRANULE 7680
#define LP_RS_BPC 320               /* util.h BPC: at most this many fractional-offset filters each side */
#define LP_RS_TAPS 33               /* filter_l + 1 <= 33 */

enum { LP_NORM = 0, LP_START = 1, LP_SHORT = 2, LP_STOP = 3 };
enum { LP_STEREO = 0, LP_JOINT = 1, LP_DUAL = 2, LP_MONO = 3, LP_MODE_NOT_SET = 4 };

/* partition-band constants, one set for long FFT, short FFT and long->short mapping
 * (reference PsyConst_CB2SB_t, util.h:188) */
typedef struct {
    float masking_lower[LP_CBANDS], minval[LP_CBANDS], rnumlines[LP_CBANDS], mld_cb[LP_CBANDS];
    float mld[LP_SBMAX_L], bo_weight[LP_SBMAX_L];
    int   s3ind[LP_CBANDS][2], numlines[LP_CBANDS], bm[LP_SBMAX_L], bo[LP_SBMAX_L];
    int   npart, n_sb, n_s3;
    float s3[2048];
} lp_bands;

typedef struct {
    /* resolved stream parameters */
    int   samplerate, channels, mode, brate, bitrate_index, samplerate_index, version, mode_gr;
    int   sideinfo_len, frac_spf, buffer_constraint, lowpassfreq, quality;
    int   noise_shaping, noise_shaping_amp, noise_shaping_stop, subblock_gain, use_best_huffman,
          full_outer_loop, quant_comp, quant_comp_short, substep_shaping, sfb21_extra,
          use_temporal, short_blocks, force_ms, use_safe_joint_stereo, disable_reservoir,
          error_protection, copyright, original, extension, emphasis, athtype, ath_use_adjust;
    float msfix, ath_offset_db, ath_offset_factor, athcurve, athfixpoint, minval, interch;
    float mask_adjust, mask_adjust_short, pcm_transform[2][2], lowpass1, lowpass2, highpass1, highpass2;
    float adjust_bass_db, adjust_alto_db, adjust_treble_db, adjust_sfb21_db;
    float ath_aa_sensitivity_p, ath_decay, ath_floor;
    /* vbr: 0 = vbr_off (CBR), 3 = vbr_abr, 4 = vbr_mtrh with quality vbr_q (lame.h:94 vbr_mode); ABR keeps its mean bitrate, the bitrate index range
     * it may choose a frame size from, and the compression ratio calc_target_bits reads (quantize.c:1768) */
    int   vbr, vbr_q, vbr_mean_kbps, vbr_min_bitrate_index, vbr_max_bitrate_index;
    float compression_ratio, vbr_q_frac;
    /* input-rate conversion (util.c:531 fill_buffer_resample): samplerate is the OUTPUT rate */
    int   samplerate_in, resample, rs_filter_l, rs_bpc;
    double rs_ratio;
    /* tables */
    int   sfb_l[23], sfb_s[14], psfb21[7], psfb12[7];
    float amp_filter[32];
    lp_bands l, s, l2s;
    float attack_threshold[4], decay;
    float ath_l[22], ath_s[13], ath_psfb21[6], ath_psfb12[6], ath_cb_l[64], ath_cb_s[64], eql_w[512];
    float longfact[22], shortfact[13];
    int   bv_scf[576];
    float window[LP_BLK], window_s[LP_BLK_S / 2];
    float pow43[LP_PRECALC], adj43asm[LP_PRECALC], ipow20[LP_QMAX], pow20[LP_QMAX + LP_QMAX2 + 1];
    float log_table[513];
    float ma_max_i1, ma_max_i2;
    float rs_filt[(2 * LP_RS_BPC + 1) * LP_RS_TAPS];
} lp_config;

/* one granule/channel of coded data (reference gr_info, l3side.h:47) */
typedef struct {
    float xr[576];
    int   l3_enc[576];
    int   scalefac[LP_SFBMAX];
    float xrpow_max;
    int   part2_3_length, big_values, count1, global_gain, scalefac_compress, block_type,
          mixed_block_flag, table_select[3], subblock_gain[4], region0_count, region1_count, preflag,
          scalefac_scale, count1table_select, part2_length, sfb_lmax, sfb_smin, psy_lmax, sfbmax,
          psymax, sfbdivide, width[LP_SFBMAX], window[LP_SFBMAX], count1bits, max_nonzero_coeff;
    int   slen[4];                      /* MPEG-2/2.5 (LSF) scalefactor coding: bits per partition and the partition sizes (takehiro.c:1218) */
    const int *sfb_partition_table;
    char  energy_above_cutoff[LP_SFBMAX];
} lp_granule;

typedef struct { float l[LP_SBMAX_L]; float s[LP_SBMAX_S][3]; } lp_xmin;
typedef struct { lp_xmin thm, en; } lp_ratio;

typedef struct {
    float nb_l1[4][LP_CBANDS], nb_l2[4][LP_CBANDS];
    lp_xmin thm[4], en[4];
    float loudness_sq_save[2], tot_ener[4], last_en_subshort[4][9];
    int   last_attacks[4], blocktype_old[2];
} lp_psy_state;

#define LP_MFSIZE (3 * 1152 + 576 - 48)
#define LP_MAX_HEADER_BUF 256
#define LP_BITBUF (147456 + 16384)

typedef struct {
    lp_config cfg;
    /* per-stream mutable state */
    lp_psy_state psy;
    float ath_adjust_factor, ath_adjust_limit, masking_lower;
    float loudness_sq[2][2];
    float sb_sample[2][2][18][32];
    float pefirbuf[19];
    int   slot_lag, padding, mode_ext, frame_number, frame_init_done;
    int   bitrate_index;                /* of the frame being coded (constant for CBR, chosen per frame for ABR) */
    int   resv_size, resv_max, main_data_begin, drain_pre, drain_post, scfsi[2][4];
    int   old_value[2], current_step[2];
    lp_granule tt[2][2];
    float mfbuf[2][LP_MFSIZE];
    int   mf_size, mf_samples_to_encode;
    float rs_old[2][LP_RS_TAPS];        /* resampler: the last filter_l + 1 input samples, and the input time of the next chunk */
    double rs_itime[2];
    /* bit writer (reference Bit_stream_struc + header ring, util.h:272) */
    unsigned char *buf;
    int   totbit, buf_byte_idx, buf_bit_idx;
    struct { int write_timing, ptr; unsigned char buf[40]; } header[LP_MAX_HEADER_BUF];   /* sideinfo_len <= 38 */
    int   h_ptr, w_ptr, ancillary_flag;
    /* last frame's psy products kept for tests */
    float last_pe[2][2];
} lp_encoder;

/* API: mirrors lame_init, lame_set_xxx, lame_init_params, lame_encode_buffer, lame_encode_flush, lame_close */
lp_encoder *lp_open(int samplerate, int channels, int brate, int mode, int quality);
lp_encoder *lp_open_ex(int samplerate, int channels, int brate, int mode, int quality, int vbr /* 0 off, 3 abr, 4 mtrh: brate = VBR_q */);
lp_encoder *lp_open_rs(int samplerate_in, int samplerate_out /* 0 = as lame_init_params picks it */, int channels, int brate, int mode, int quality, int vbr);
/* vbr 4 with a fractional level: VBR quality = brate + vbr_q_frac (lame_set_VBR_quality, set_get.c:1152) */
lp_encoder *lp_open_vq(int samplerate_in, int samplerate_out, int channels, int brate, int mode, int quality, int vbr, float vbr_q_frac);
int  lp_encode(lp_encoder *e, const short *l, const short *r, int nsamples, unsigned char *out, int cap);
int  lp_flush(lp_encoder *e, unsigned char *out, int cap);
void lp_close(lp_encoder *e);
void lp_set_error_protection(lp_encoder *e);   /* before the first lp_encode */

/* internals shared between the port's files */
int   lp_setup(lp_config *c, int samplerate_in, int samplerate_out, int channels, int brate, int mode, int quality, int vbr, float vbr_q_frac);
float lp_fast_log2(const lp_config *c, float x);
void  lp_fft_long(const lp_config *c, float x[LP_BLK], const float *buf);
void  lp_fft_short(const lp_config *c, float x[3][LP_BLK_S], const float *buf);
int   lp_psycho(lp_encoder *e, const float *const buffer[2], int gr_out, lp_ratio masking_ratio[2][2],
                lp_ratio masking_ms[2][2], float pe[2], float pe_ms[2], float energy[4], int blocktype_d[2]);
void  lp_mdct_sub48(lp_encoder *e, const float *w0, const float *w1);
void  lp_cbr_iteration_loop(lp_encoder *e, float pe[2][2], const float ms_ener_ratio[2], lp_ratio ratio[2][2]);
void  lp_abr_iteration_loop(lp_encoder *e, float pe[2][2], const float ms_ener_ratio[2], lp_ratio ratio[2][2]);
void  lp_vbr_old_iteration_loop(lp_encoder *e, float pe[2][2], const float ms_ener_ratio[2], lp_ratio ratio[2][2]);
void  lp_vbr_new_iteration_loop(lp_encoder *e, float pe[2][2], const float ms_ener_ratio[2], lp_ratio ratio[2][2]);
int   lp_getframebits(const lp_encoder *e);
void  lp_format_bitstream(lp_encoder *e);
void  lp_flush_bitstream(lp_encoder *e);
int   lp_copy_buffer(lp_encoder *e, unsigned char *out, int cap);

/* static tables (generated numeric data, tools/gen_tables.c) */
const float   *lp_tab_enwindow(void);
const float   *lp_tab_mdctwin(void);
extern const uint8_t lp_pretab[22];

#endif
