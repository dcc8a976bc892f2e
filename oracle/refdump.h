/* Plain-C layout of the reference-state snapshots produced by oracle/refdump.c (test infrastructure).
 * All fields are 4-byte so the Python side can mirror it with a numpy structured dtype. */
#ifndef ORACLE_REFDUMP_H
#define ORACLE_REFDUMP_H

#define REFDUMP_MAX_S3 2048

typedef struct {
    float xr[576];
    int   l3_enc[576];
    int   scalefac[39];
    int   part2_3_length, big_values, count1, global_gain, scalefac_compress, block_type,
          mixed_block_flag, table_select[3], subblock_gain[3], region0_count, region1_count,
          preflag, scalefac_scale, count1table_select, part2_length, sfbmax, sfbdivide, psymax,
          max_nonzero_coeff, count1bits;
    float xrpow_max;
    float loudness_sq;
} refdump_granule;

typedef struct {
    int   frame_number, padding, mode_ext, main_data_begin, resv_size, resv_max, drain_pre, drain_post;
    int   slot_lag;
    float ath_adjust_factor, ath_adjust_limit, masking_lower;
    float pefirbuf[19];
    int   old_value[2], current_step[2], blocktype_old[2];
    float loudness_sq_save[2];
    int   scfsi[2][4];
    refdump_granule gi[2][2];
    float en_l[4][22], thm_l[4][22], en_s[4][13][3], thm_s[4][13][3];
    float nb_l1[4][64], nb_l2[4][64];
    float tot_ener[4];
    int   last_attacks[4];
    float last_en_subshort[4][9];
    float sb_sample[2][2][18][32];
} refdump_frame;

typedef struct {
    float masking_lower[64], minval[64], rnumlines[64], mld_cb[64], mld[22], bo_weight[22];
    int   s3ind[64][2], numlines[64], bm[22], bo[22], npart, n_sb, n_s3;
    float s3[REFDUMP_MAX_S3];
} refdump_cb2sb;

typedef struct {
    refdump_cb2sb l, s, l2s;
    float attack_threshold[4], decay;
    float ath_l[22], ath_s[13], ath_psfb21[6], ath_psfb12[6], ath_cb_l[64], ath_cb_s[64], eql_w[512];
    float ath_floor, ath_decay, ath_aa_sensitivity_p;
    int   ath_use_adjust;
    float longfact[22], shortfact[13];
    int   bv_scf[576];
    float amp_filter[32];
    int   sfb_l[23], sfb_s[14];
    float pow43[8208], adj43asm[8208], ipow20[257], pow20[374];
    float mask_adjust, mask_adjust_short;
    int   sfb21_extra, substep_shaping;
    float msfix, ath_offset_factor, ath_offset_db, athfixpoint, athcurve, minval_cfg;
    float pcm_transform[4];
    int   noise_shaping, noise_shaping_amp, noise_shaping_stop, subblock_gain, use_best_huffman,
          full_outer_loop, quant_comp, quant_comp_short, use_temporal, short_blocks, mode, force_ms,
          sideinfo_len, avg_bitrate, bitrate_index, samplerate_out, buffer_constraint, frac_spf;
    float lowpass1, lowpass2;
    int   vbr, disable_reservoir;
    float interch;
    int   athtype;
} refdump_tab;

void *refdump_open(int brate, int mode, int quality, int vbrmode, int vbr_q, int samplerate, int nch);
void *refdump_open_vq(int brate, int mode, int quality, int vbrmode, int vbr_q, float vbr_q_frac, int samplerate, int out_samplerate, int nch);
void *refdump_open_rs(int brate, int mode, int quality, int vbrmode, int vbr_q, int samplerate, int out_samplerate /* 0 = automatic */, int nch);
int   refdump_encode(void *h, const short *l, const short *r, int n, unsigned char *out, int cap);
int   refdump_flush(void *h, unsigned char *out, int cap);
void  refdump_close(void *h);
int   refdump_frame_number(void *h);
void  refdump_snapshot(void *h, refdump_frame *d);
void  refdump_tables(void *h, refdump_tab *t);

#endif
