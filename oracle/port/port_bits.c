/* oracle/port/port_bits.c - TEST INFRASTRUCTURE (see lame_port.h).
 * Restates the serial bit packer of bitstream.c: header/side-info ring (encodeSideInfo2 :321),
 * main data (writeMainData :686, Huffmancode :561, huffman_coder_count1 :482), ancillary drain
 * (drain_into_ancillary :214) and format_bitstream (:918) / flush_bitstream (:863). */
#include <string.h>
#include "lame_port.h"
#include "port_tables.inc"

static const int slen1_tab[16] = { 0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4 };
static const int slen2_tab[16] = { 0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3 };

/* bitstream.c:130 putheader_bits */
static void put_header(lp_encoder *e)
{
    memcpy(&e->buf[e->buf_byte_idx], e->header[e->w_ptr].buf, e->cfg.sideinfo_len);
    e->buf_byte_idx += e->cfg.sideinfo_len;
    e->totbit += e->cfg.sideinfo_len * 8;
    e->w_ptr = (e->w_ptr + 1) & (LP_MAX_HEADER_BUF - 1);
}
/* bitstream.c:150 putbits2 */
static void putbits(lp_encoder *e, int val, int j)
{
    while (j > 0) {
        int k;
        if (e->buf_bit_idx == 0) {
            e->buf_bit_idx = 8;
            e->buf_byte_idx++;
            if (e->header[e->w_ptr].write_timing == e->totbit) put_header(e);
            e->buf[e->buf_byte_idx] = 0;
        }
        k = j < e->buf_bit_idx ? j : e->buf_bit_idx;
        j -= k;
        e->buf_bit_idx -= k;
        e->buf[e->buf_byte_idx] |= ((val >> j) << e->buf_bit_idx);
        e->totbit += k;
    }
}
/* bitstream.c:214 drain_into_ancillary; the version string is get_lame_short_version() = "3.99.5" */
static void drain_ancillary(lp_encoder *e, int remainingBits)
{
    static const char version[] = "3.99.5";
    int i;
    if (remainingBits >= 8) { putbits(e, 0x4c, 8); remainingBits -= 8; }
    if (remainingBits >= 8) { putbits(e, 0x41, 8); remainingBits -= 8; }
    if (remainingBits >= 8) { putbits(e, 0x4d, 8); remainingBits -= 8; }
    if (remainingBits >= 8) { putbits(e, 0x45, 8); remainingBits -= 8; }
    if (remainingBits >= 32)
        for (i = 0; i < (int) strlen(version) && remainingBits >= 8; ++i) {
            remainingBits -= 8;
            putbits(e, version[i], 8);
        }
    for (; remainingBits >= 1; remainingBits -= 1) {
        putbits(e, e->ancillary_flag, 1);
        e->ancillary_flag ^= !e->cfg.disable_reservoir;
    }
}
/* bitstream.c:261 writeheader */
static void writeheader(lp_encoder *e, int val, int j)
{
    int ptr = e->header[e->h_ptr].ptr;
    while (j > 0) {
        int const k = j < 8 - (ptr & 7) ? j : 8 - (ptr & 7);
        j -= k;
        e->header[e->h_ptr].buf[ptr >> 3] |= ((val >> j)) << (8 - (ptr & 7) - k);
        ptr += k;
    }
    e->header[e->h_ptr].ptr = ptr;
}
/* bitstream.c:287 CRC_update + :304 CRC_writeheader: CRC-16 (x^16 + x^15 + x^2 + 1, MSB first, start 0xffff) over header bytes 2-3 and the
 * side info, stored behind the header */
static void crc_writeheader(const lp_config *cfg, unsigned char *header)
{
    int crc = 0xffff, i, k;
    for (i = 2; i < cfg->sideinfo_len; i++) {
        int value;
        if (i == 4 || i == 5) continue;
        value = header[i] << 8;
        for (k = 0; k < 8; k++) {
            value <<= 1;
            crc <<= 1;
            if ((crc ^ value) & 0x10000) crc ^= 0x8005;
        }
    }
    header[4] = crc >> 8;
    header[5] = crc & 255;
}

/* bitstream.c:321 encodeSideInfo2 */
static void encode_side_info(lp_encoder *e, int bitsPerFrame)
{
    const lp_config *cfg = &e->cfg;
    int gr, ch, band, old;
    e->header[e->h_ptr].ptr = 0;
    memset(e->header[e->h_ptr].buf, 0, cfg->sideinfo_len);
    if (cfg->samplerate < 16000) writeheader(e, 0xffe, 12);           /* MPEG-2.5 */
    else writeheader(e, 0xfff, 12);
    writeheader(e, cfg->version, 1);
    writeheader(e, 4 - 3, 2);
    writeheader(e, !cfg->error_protection, 1);
    writeheader(e, e->bitrate_index, 4);
    writeheader(e, cfg->samplerate_index, 2);
    writeheader(e, e->padding, 1);
    writeheader(e, cfg->extension, 1);
    writeheader(e, cfg->mode, 2);
    writeheader(e, e->mode_ext, 2);
    writeheader(e, cfg->copyright, 1);
    writeheader(e, cfg->original, 1);
    writeheader(e, cfg->emphasis, 2);
    if (cfg->error_protection) writeheader(e, 0, 16);                 /* room for the CRC */
    if (cfg->version != 1) {
        /* MPEG-2/2.5: one granule, 8-bit main_data_begin, 9-bit scalefac_compress, no scfsi, no preflag bit */
        writeheader(e, e->main_data_begin, 8);
        writeheader(e, 0, cfg->channels);
        for (ch = 0; ch < cfg->channels; ch++) {
            lp_granule *gi = &e->tt[0][ch];
            writeheader(e, gi->part2_3_length + gi->part2_length, 12);
            writeheader(e, gi->big_values / 2, 9);
            writeheader(e, gi->global_gain, 8);
            writeheader(e, gi->scalefac_compress, 9);
            if (gi->table_select[0] == 14) gi->table_select[0] = 16;
            if (gi->table_select[1] == 14) gi->table_select[1] = 16;
            if (gi->block_type != LP_NORM) {
                writeheader(e, 1, 1);
                writeheader(e, gi->block_type, 2);
                writeheader(e, gi->mixed_block_flag, 1);
                writeheader(e, gi->table_select[0], 5);
                writeheader(e, gi->table_select[1], 5);
                writeheader(e, gi->subblock_gain[0], 3);
                writeheader(e, gi->subblock_gain[1], 3);
                writeheader(e, gi->subblock_gain[2], 3);
            }
            else {
                writeheader(e, 0, 1);
                writeheader(e, gi->table_select[0], 5);
                writeheader(e, gi->table_select[1], 5);
                if (gi->table_select[2] == 14) gi->table_select[2] = 16;
                writeheader(e, gi->table_select[2], 5);
                writeheader(e, gi->region0_count, 4);
                writeheader(e, gi->region1_count, 3);
            }
            writeheader(e, gi->scalefac_scale, 1);
            writeheader(e, gi->count1table_select, 1);
        }
        if (cfg->error_protection) crc_writeheader(cfg, e->header[e->h_ptr].buf);
        old = e->h_ptr;
        e->h_ptr = (old + 1) & (LP_MAX_HEADER_BUF - 1);
        e->header[e->h_ptr].write_timing = e->header[old].write_timing + bitsPerFrame;
        return;
    }
    writeheader(e, e->main_data_begin, 9);
    writeheader(e, 0, cfg->channels == 2 ? 3 : 5);
    for (ch = 0; ch < cfg->channels; ch++)
        for (band = 0; band < 4; band++) writeheader(e, e->scfsi[ch][band], 1);
    for (gr = 0; gr < 2; gr++)
        for (ch = 0; ch < cfg->channels; ch++) {
            lp_granule *gi = &e->tt[gr][ch];
            writeheader(e, gi->part2_3_length + gi->part2_length, 12);
            writeheader(e, gi->big_values / 2, 9);
            writeheader(e, gi->global_gain, 8);
            writeheader(e, gi->scalefac_compress, 4);
            if (gi->block_type != LP_NORM) {
                writeheader(e, 1, 1);
                writeheader(e, gi->block_type, 2);
                writeheader(e, gi->mixed_block_flag, 1);
                if (gi->table_select[0] == 14) gi->table_select[0] = 16;
                writeheader(e, gi->table_select[0], 5);
                if (gi->table_select[1] == 14) gi->table_select[1] = 16;
                writeheader(e, gi->table_select[1], 5);
                writeheader(e, gi->subblock_gain[0], 3);
                writeheader(e, gi->subblock_gain[1], 3);
                writeheader(e, gi->subblock_gain[2], 3);
            }
            else {
                writeheader(e, 0, 1);
                if (gi->table_select[0] == 14) gi->table_select[0] = 16;
                writeheader(e, gi->table_select[0], 5);
                if (gi->table_select[1] == 14) gi->table_select[1] = 16;
                writeheader(e, gi->table_select[1], 5);
                if (gi->table_select[2] == 14) gi->table_select[2] = 16;
                writeheader(e, gi->table_select[2], 5);
                writeheader(e, gi->region0_count, 4);
                writeheader(e, gi->region1_count, 3);
            }
            writeheader(e, gi->preflag, 1);
            writeheader(e, gi->scalefac_scale, 1);
            writeheader(e, gi->count1table_select, 1);
        }
    if (cfg->error_protection) crc_writeheader(cfg, e->header[e->h_ptr].buf);
    old = e->h_ptr;
    e->h_ptr = (old + 1) & (LP_MAX_HEADER_BUF - 1);
    e->header[e->h_ptr].write_timing = e->header[old].write_timing + bitsPerFrame;
}

/* bitstream.c:561 Huffmancode */
static int huffman_code(lp_encoder *e, unsigned tableindex, int start, int end, const lp_granule *gi)
{
    const uint8_t *hlen = LGT_HUFF_LEN + LGT_HUFF_OFF[tableindex];
    const uint16_t *code = LGT_HUFF_CODE + LGT_HUFF_OFF[tableindex];
    unsigned const linbits = LGT_HUFF_XLEN[tableindex];
    int i, bits = 0;
    if (!tableindex) return bits;
    for (i = start; i < end; i += 2) {
        int16_t cbits = 0;
        uint16_t xbits = 0;
        unsigned xlen = LGT_HUFF_XLEN[tableindex], ext = 0;
        unsigned x1 = gi->l3_enc[i], x2 = gi->l3_enc[i + 1];
        if (x1 != 0u) {
            if (gi->xr[i] < 0.0f) ext++;
            cbits--;
        }
        if (tableindex > 15u) {
            if (x1 >= 15u) {
                uint16_t const linbits_x1 = x1 - 15u;
                ext |= linbits_x1 << 1u;
                xbits = linbits;
                x1 = 15u;
            }
            if (x2 >= 15u) {
                uint16_t const linbits_x2 = x2 - 15u;
                ext <<= linbits;
                ext |= linbits_x2;
                xbits += linbits;
                x2 = 15u;
            }
            xlen = 16;
        }
        if (x2 != 0u) {
            ext <<= 1;
            if (gi->xr[i + 1] < 0.0f) ext++;
            cbits--;
        }
        x1 = x1 * xlen + x2;
        xbits -= cbits;
        cbits += hlen[x1];
        putbits(e, code[x1], cbits);
        putbits(e, (int) ext, xbits);
        bits += cbits + xbits;
    }
    return bits;
}
/* bitstream.c:482 huffman_coder_count1 */
static int huffman_count1(lp_encoder *e, const lp_granule *gi)
{
    int const t = gi->count1table_select + 32;
    const uint8_t *hlen = LGT_HUFF_LEN + LGT_HUFF_OFF[t];
    const uint16_t *code = LGT_HUFF_CODE + LGT_HUFF_OFF[t];
    int i, bits = 0;
    const int *ix = &gi->l3_enc[gi->big_values];
    const float *xr = &gi->xr[gi->big_values];
    for (i = (gi->count1 - gi->big_values) / 4; i > 0; --i) {
        int huffbits = 0, p = 0, v;
        v = ix[0]; if (v) { p += 8; if (xr[0] < 0.0f) huffbits++; }
        v = ix[1]; if (v) { p += 4; huffbits *= 2; if (xr[1] < 0.0f) huffbits++; }
        v = ix[2]; if (v) { p += 2; huffbits *= 2; if (xr[2] < 0.0f) huffbits++; }
        v = ix[3]; if (v) { p++; huffbits *= 2; if (xr[3] < 0.0f) huffbits++; }
        ix += 4;
        xr += 4;
        putbits(e, huffbits + code[p], hlen[p]);
        bits += hlen[p];
    }
    return bits;
}
/* bitstream.c:686 writeMainData with Short/LongHuffmancodebits (:633/:650) */
static int write_main_data(lp_encoder *e)
{
    const lp_config *cfg = &e->cfg;
    int gr, ch, sfb, data_bits, tot_bits = 0;
    for (gr = 0; gr < cfg->mode_gr; gr++)
        for (ch = 0; ch < cfg->channels; ch++) {
            const lp_granule *gi = &e->tt[gr][ch];
            data_bits = 0;
            if (cfg->version != 1) {
                /* MPEG-2/2.5: every band of a partition with the partition's width, bands past the last partition not at all */
                int part, i, w, nwin = (gi->block_type == LP_SHORT) ? 3 : 1;
                sfb = 0;
                for (part = 0; part < 4; part++) {
                    int const sfbs = gi->sfb_partition_table[part] / nwin, slen = gi->slen[part];
                    for (i = 0; i < sfbs; i++, sfb++)
                        for (w = 0; w < nwin; w++) {
                            int const v = gi->scalefac[sfb * nwin + w];
                            putbits(e, v > 0 ? v : 0, slen);
                            data_bits += slen;
                        }
                }
            }
            else {
            int const slen1 = slen1_tab[gi->scalefac_compress], slen2 = slen2_tab[gi->scalefac_compress];
            for (sfb = 0; sfb < gi->sfbdivide; sfb++) {
                if (gi->scalefac[sfb] == -1) continue;
                putbits(e, gi->scalefac[sfb], slen1);
                data_bits += slen1;
            }
            for (; sfb < gi->sfbmax; sfb++) {
                if (gi->scalefac[sfb] == -1) continue;
                putbits(e, gi->scalefac[sfb], slen2);
                data_bits += slen2;
            }
            }
            if (gi->block_type == LP_SHORT) {
                int region1Start = 3 * cfg->sfb_s[3];
                if (region1Start > gi->big_values) region1Start = gi->big_values;
                data_bits += huffman_code(e, gi->table_select[0], 0, region1Start, gi);
                data_bits += huffman_code(e, gi->table_select[1], region1Start, gi->big_values, gi);
            }
            else {
                int bigvalues = gi->big_values, region1Start, region2Start;
                unsigned i = gi->region0_count + 1;
                region1Start = cfg->sfb_l[i];
                i += gi->region1_count + 1;
                region2Start = cfg->sfb_l[i];
                if (region1Start > bigvalues) region1Start = bigvalues;
                if (region2Start > bigvalues) region2Start = bigvalues;
                data_bits += huffman_code(e, gi->table_select[0], 0, region1Start, gi);
                data_bits += huffman_code(e, gi->table_select[1], region1Start, region2Start, gi);
                data_bits += huffman_code(e, gi->table_select[2], region2Start, bigvalues, gi);
            }
            data_bits += huffman_count1(e, gi);
            tot_bits += data_bits;
        }
    return tot_bits;
}

/* bitstream.c:918 format_bitstream */
void lp_format_bitstream(lp_encoder *e)
{
    int bits, bitsPerFrame = lp_getframebits(e);
    drain_ancillary(e, e->drain_pre);
    encode_side_info(e, bitsPerFrame);
    bits = 8 * e->cfg.sideinfo_len;
    bits += write_main_data(e);
    drain_ancillary(e, e->drain_post);
    bits += e->drain_post;
    e->main_data_begin += (bitsPerFrame - bits) / 8;
    if (e->totbit > 1000000000) {
        int i;
        for (i = 0; i < LP_MAX_HEADER_BUF; ++i) e->header[i].write_timing -= e->totbit;
        e->totbit = 0;
    }
}

/* bitstream.c:793 compute_flushbits + :863 flush_bitstream */
void lp_flush_bitstream(lp_encoder *e)
{
    int flushbits, remaining_headers, first_ptr = e->w_ptr, last_ptr = e->h_ptr - 1;
    if (last_ptr == -1) last_ptr = LP_MAX_HEADER_BUF - 1;
    flushbits = e->header[last_ptr].write_timing - e->totbit;
    if (flushbits >= 0) {
        remaining_headers = 1 + last_ptr - first_ptr;
        if (last_ptr < first_ptr) remaining_headers = 1 + last_ptr - first_ptr + LP_MAX_HEADER_BUF;
        flushbits -= remaining_headers * 8 * e->cfg.sideinfo_len;
    }
    flushbits += lp_getframebits(e);
    if (flushbits < 0) return;
    drain_ancillary(e, flushbits);
    e->resv_size = 0;
    e->main_data_begin = 0;
}

/* bitstream.c:1059 do_copy_buffer */
int lp_copy_buffer(lp_encoder *e, unsigned char *out, int cap)
{
    int const minimum = e->buf_byte_idx + 1;
    if (minimum <= 0) return 0;
    if (cap != 0 && minimum > cap) return -1;
    memcpy(out, e->buf, minimum);
    e->buf_byte_idx = -1;
    e->buf_bit_idx = 0;
    return minimum;
}
