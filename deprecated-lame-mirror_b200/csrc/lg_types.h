// lg_types.h - plain-old-data layouts shared by the host side and the kernels.
//
// Vocabulary follows the reference: a *stream* is one lame_t handle, a *frame*
// is 1152 samples x 2 channels, a frame has two *granules* of 576 lines per channel, "gr.ch" is one
// granule of one channel, *sfb* = scalefactor band, *partition* = psychoacoustic partition band.
//
// HBM layout for one batch of S streams x F frames (all arrays stream-major, so a CTA that owns a
// stream or a granule reads contiguous memory):
//   pcm      int16 [S][2][F*1152 + LG_PCM_HALO]   timeline index 0 = sample 1152*k0 - 576 of the
//                                                 zero-prefixed stream (528 leading zeros, lame.c:2302)
//   sb       f32   [S][2F+1][2][18][32]           polyphase subband samples, slot 0 = granule before the batch
//   ana      LgAnalysis [S][2F]                   stateless psycho-acoustic products per granule
//   psy      LgPsyOut   [S][2F]                   ordered-scan products per granule (block types, ratios, pe)
//   frm      LgFrameCtl [S][F]                    per-frame decisions (mode_ext, ATH factor, bit budget inputs)
//   xr       f32   [S][2F][2][576]                MDCT lines, M/S-converted and short-block reordered
//   out      LgGranuleOut [S][2F][2] + LgFrameOut [S][F]   quantised result handed to the bit packer
//   state    LgStreamState [S]                    everything a stream carries from one batch to the next
#pragma once
#include <stdint.h>

#define LG_CBANDS 64
#define LG_SBMAX_L 22
#define LG_SBMAX_S 13
#define LG_SBPSY_L 21
#define LG_SBPSY_S 12
#define LG_SFBMAX 39
#define LG_BLK 1024
#define LG_BLK_S 256
#define LG_HBLK 513
#define LG_HBLK_S 129
#define LG_PRECALC 8208
#define LG_QMAX 257
#define LG_QMAX2 116
#define LG_IXMAX 8206
#define LG_LARGE_BITS 100000
#define LG_MAX_BITS_PER_CHANNEL 4095
#define LG_MAX_BITS_PER_GRANULE 7680
#define LG_PCM_HIST 576                 /* samples kept before the first frame of a batch */
#define LG_PCM_HALO 1328                /* HIST + 752 look-ahead (encoder.c:254-299) */
#define LG_RS_BPC 320                   /* util.h BPC: fractional-offset filters each side of the resampler */
#define LG_RS_TAPS 33                   /* filter_l + 1 <= 33 taps */
#define LG_GR_SPAN 1328                 /* samples one granule's analysis touches: [576g+576, 576g+1904) */

enum { LG_NORM = 0, LG_START = 1, LG_SHORT = 2, LG_STOP = 3 };
/* What a stream's row of the native PCM window holds (lame.c:1803-1834 COPY_AND_TRANSFORM's T) and the entry point's normalisation
 * factor s (lame.c:1884-1960); kernel A converts to sample_t and applies s * pcm_transform.  LG_PCM_DONE: floats the host has already
 * transformed (a stream whose calls mixed sample types). */
enum { LG_PCM_S16 = 0, LG_PCM_DONE = 1, LG_PCM_S32 = 2, LG_PCM_F32 = 3, LG_PCM_S64 = 4, LG_PCM_F64 = 5 };
typedef struct { int kind; float scale; } LgPcmKind;
enum { LG_STEREO = 0, LG_JOINT = 1, LG_DUAL = 2, LG_MONO = 3, LG_MODE_NOT_SET = 4 };

/* partition-band constants (reference PsyConst_CB2SB_t, util.h:188) */
typedef struct {
    float masking_lower[LG_CBANDS], minval[LG_CBANDS], rnumlines[LG_CBANDS], mld_cb[LG_CBANDS];
    float bo_weight[LG_SBMAX_L];
    int   s3lo[LG_CBANDS], s3hi[LG_CBANDS], s3off[LG_CBANDS];  /* first/last masker partition, offset into s3 */
    int   numlines[LG_CBANDS], linestart[LG_CBANDS + 1];
    int   bo[LG_SBMAX_L];
    int   npart, n_sb, n_s3, maxlines;
    float s3[1024];
} LgBands;

/* immutable per-configuration tables, computed on the host with the host's libm exactly as the
 * reference does at lame_init_params time, then uploaded once */
typedef struct {
    int   samplerate, channels, mode, brate, bitrate_index, samplerate_index, version, mode_gr;
    int   sideinfo_len, frac_spf, buffer_constraint, lowpassfreq, quality;
    int   noise_shaping, noise_shaping_amp, noise_shaping_stop, subblock_gain, use_best_huffman,
          full_outer_loop, quant_comp, quant_comp_short, substep_shaping, sfb21_extra,
          use_temporal, short_blocks, force_ms, use_safe_joint_stereo, disable_reservoir,
          error_protection, copyright, original, extension, emphasis, athtype, ath_use_adjust,
          enforce_min_bitrate /* lame_set_VBR_hard_min: silent frames keep the minimum bitrate too */, ath_only, no_ath,
          preset /* what apply_preset was last called with: goes into the LAME tag (VbrTag.c:713) */;
    float msfix, ath_offset_db, ath_offset_factor, athcurve, athfixpoint, minval, interch;
    float mask_adjust, mask_adjust_short, pcm_transform[2][2], lowpass1, lowpass2, highpass1, highpass2;
    float adjust_bass_db, adjust_alto_db, adjust_treble_db, adjust_sfb21_db;
    float ath_aa_sensitivity_p, ath_decay, ath_floor;
    float masking_lower_long, masking_lower_short;  /* (float)pow(10, mask_adjust*0.1), quantize.c:2029 */
    /* vbr: 0 = vbr_off (CBR), 3 = vbr_abr, 4 = vbr_mtrh at quality vbr_q (lame.h:94).  ABR chooses a frame size per frame: mean bitrate, index range,
     * the compression ratio calc_target_bits reads (quantize.c:1768) and the bitrate table row of this MPEG version */
    int   vbr, vbr_q, vbr_mean_kbps, vbr_min_bitrate_index, vbr_max_bitrate_index;
    float compression_ratio, vbr_q_frac;     /* VBR quality = vbr_q + vbr_q_frac, after the mapping of lame.c:661-698 */
    int   bitrate_kbps[16];
    int   sfb_l[23], sfb_s[14], psfb21[7], psfb12[7];
    float amp_filter[32];
    LgBands l, s, l2s;
    float attack_threshold[4], decay;
    float ath_l[22], ath_s[13], ath_psfb21[6], ath_psfb12[6], ath_cb_l[64], ath_cb_s[64], eql_w[512];
    float longfact[22], shortfact[13];
    uint8_t bv_scf[576];
    float window[LG_BLK], window_s[LG_BLK_S / 2];
    float fht_tw[4][128][4];            /* per stage, per i: c1, s1, c2, s2 of the reference's recurrence (fft.c:105-144) */
    float pow43[LG_PRECALC], adj43asm[LG_PRECALC], ipow20[LG_QMAX], pow20[LG_QMAX + LG_QMAX2 + 1];
    float log_table[513];
    float ma_max_i1, ma_max_i2;
    float enwindow[288], mdctwin[4 * 36];
    /* Huffman bit-length books (ISO 11172-3 table B.7) in the packed layout of lg_tables_data.inc */
    uint8_t  huff_len[2048];
    uint16_t huff_off[34];
    uint8_t  huff_xlen[34];
    uint16_t huff_linmax[34];
    uint32_t largetbl[256], table23[9], table56[16];
    /* one packed length book per magnitude class of a region (max 1 | 2 | 3 | 4-5 | 6-7 | 8-15 | escape): entry
     * [class][x*16+y] = bits under the class's three candidate tables in 10-bit fields (two-candidate classes repeat
     * the second), so one code path replaces count_bit_noESC/_from2/_from3/_ESC (takehiro.c:449-573) */
    uint32_t huff_pk[7 * 256];
    /* scalefactor band of every line: long blocks (22 bands) and short blocks after reordering (39 = 13 x 3 windows) */
    uint8_t line_sfb_l[576], line_sfb_s[576];
    uint16_t huff_code[2048];           /* code words, same layout as huff_len (bit packer, tables.c HB tables) */
    /* input-rate conversion (util.c:531 fill_buffer_resample); `samplerate` above is the OUTPUT rate */
    int   samplerate_in, resample, rs_filter_l, rs_bpc;
    double rs_ratio;
    float rs_filt[(2 * LG_RS_BPC + 1) * LG_RS_TAPS];
} LgDevCfg;

/* one call of the reference's resampler = one chunk: `count` output samples starting at timeline index `out_pos`,
 * made from the stream's input samples from absolute index `in_base` on, with the input clock at `itime` */
typedef struct {
    double  itime;
    int64_t in_base;
    int32_t out_pos, count;
} LgRsChunk;

/* ---- stage A (stateless analysis) -> stage B (ordered scan).  The head and the long-block part are always written and read; the
 * short-block part (two thirds of the record) only for granules that can switch to short blocks: kernel A decides that from the
 * sub-block peaks of the granule and of the one before it, a superset of the reference's attack detection (psymodel.c:759-935 only ever
 * REMOVES attacks after comparing the peak ratios with the threshold), as the reference skips its short-block FFTs and masking for
 * long-block granules (psymodel.c:1474 vbrpsy_skip_masking_s) */
typedef struct {
    float en_subshort[4][12];           /* 9 sub-block peaks of the high-passed signal (psymodel.c:831) */
    float tot_ener[4];
    float loudness[2];
    int   has_short, pad_;
    float eb_l[4][LG_CBANDS];           /* partition energies, long FFT, L R M S */
    float ecb_l[4][LG_CBANDS];          /* spread threshold before pre-echo control (psymodel.c:1185) */
    float lim_l[4][LG_CBANDS];          /* max*minval*avg_mask clamp (psymodel.c:1240) */
    float eb_s[3][4][LG_CBANDS];
    float thr_s[3][4][LG_CBANDS];       /* min(ecb, clamp) for the three short windows */
} LgAnalysis;
#define LG_ANALYSIS_LONG_BYTES (sizeof(LgAnalysis) - 2 * 3 * 4 * LG_CBANDS * sizeof(float))

typedef struct { float l[LG_SBMAX_L]; float s[LG_SBMAX_S][3]; } LgXmin;

/* ---- stage B -> stages C/D, one per granule */
typedef struct {
    LgXmin en[4], thm[4];               /* L R M S ratios handed to the quantiser (one granule delay applied) */
    float  pe[4];                        /* perceptual entropy L R M S */
    int    block_type[2];
    float  tot_ener[4];
    float  loudness_sq[2];
} LgPsyOut;

/* ---- stage B -> stages C/D, one per frame */
typedef struct {
    int   mode_ext, padding;
    float ms_ener_ratio[2];
    float pe_use[2][2];                  /* after the 19-frame FIR scaling (encoder.c:489-518) */
    float ath_adjust_factor;
    float masking_lower;
} LgFrameCtl;

/* ---- quantised output of one gr.ch: exactly what encodeSideInfo2/writeMainData read (bitstream.c:321,686) */
typedef struct {
    int16_t ix[576];                     /* quantised magnitudes with the sign of xr applied */
    int8_t  scalefac[LG_SFBMAX + 1];
    int16_t part2_3_length, part2_length, big_values, count1;
    uint8_t global_gain, scalefac_compress, block_type, mixed_block_flag;
    uint8_t table_select[3], subblock_gain[3];
    uint8_t region0_count, region1_count, preflag, scalefac_scale, count1table_select, sfbmax, sfbdivide;
    uint8_t scalefac_compress_hi;        /* MPEG-2/2.5: scalefac_compress has 9 bits (takehiro.c:1281) */
} __attribute__((aligned(16))) LgGranuleOut;

typedef struct {
    int32_t main_data_begin, drain_pre, drain_post, padding, mode_ext, resv_size;
    uint8_t scfsi[2][4];
    /* for the device bit packer (kernel E): where this frame's payload (ancillary drain + main data, a whole number
     * of bytes) sits in the stream's payload buffer of this launch, and the state of the alternating stuffing bit
     * (bitstream.c:246-256) before and after the frame */
    int32_t pay_off, pay_bytes;
    uint8_t anc_pre, anc_post, pad_[2];
    int32_t bitrate_index;               /* of this frame (constant for CBR, chosen per frame by ABR) */
} LgFrameOut;
#define LG_HDR_STRIDE 44                /* bytes per frame record: header + side info (sideinfo_len <= 38 with CRC), block types at 40..43 */
#define LG_PAY_SLACK 1024               /* a launch can drain at most the reservoir (511 bytes) on top of its own frames */

/* ---- per-stream state carried across batches (reference PsyStateVar_t util.h:219, ATH_t :166,
 *      EncStateVar_t :242, QntStateVar_t :318) */
typedef struct {
    float nb_l1[4][LG_CBANDS], nb_l2[4][LG_CBANDS];
    LgXmin thm[4], en[4];
    float loudness_sq_save[2], tot_ener[4], last_en_subshort[4][9];
    int   last_attacks[4], blocktype_old[2];
    float ath_adjust_factor, ath_adjust_limit, masking_lower;
    float pefirbuf[19];
    int   slot_lag;
    int   resv_size, main_data_begin;
    int   old_value[2], current_step[2];
    int   frames_done;
    int   ancillary_flag;
} LgStreamState;
