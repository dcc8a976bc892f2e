// lg_bitstream.cpp - the serial per-stream bit packer that stays on the host:
// side info + scalefactors + Huffman code words + ancillary drain, consuming the quantised granules
// the GPU produced.  Follows bitstream.c: putbits2 :150, drain_into_ancillary :214, writeheader :261,
// encodeSideInfo2 :321, huffman_coder_count1 :482, Huffmancode :561, writeMainData :686,
// format_bitstream :918, compute_flushbits :793, flush_bitstream :863 (MPEG-1 branches).
#include <string.h>
#include <stdlib.h>
#include "lg_bitstream.h"
#include "lg_tables_data.inc"

static const int slen1_tab[16] = { 0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4 };
static const int slen2_tab[16] = { 0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3 };

void LgBitWriter::reset()
{
    buf.clear();
    bit_idx = 0;
    totbit = 0;
    memset(header, 0, sizeof header);
    h_ptr = w_ptr = 0;
    ancillary_flag = 0;
}

static inline void put_header(LgBitWriter *bw, int sideinfo_len)
{
    /* the reference copies the header over the byte it just opened (bitstream.c:137) */
    bw->buf.pop_back();
    bw->buf.insert(bw->buf.end(), bw->header[bw->w_ptr].buf, bw->header[bw->w_ptr].buf + sideinfo_len);
    bw->buf.push_back(0);
    bw->totbit += sideinfo_len * 8;
    bw->w_ptr = (bw->w_ptr + 1) & (LG_MAX_HEADER_BUF - 1);
}

static inline void putbits(LgBitWriter *bw, int sideinfo_len, int val, int j)
{
    while (j > 0) {
        if (bw->bit_idx == 0) {
            bw->bit_idx = 8;
            bw->buf.push_back(0);
            if (bw->header[bw->w_ptr].write_timing == bw->totbit) put_header(bw, sideinfo_len);
        }
        int const k = j < bw->bit_idx ? j : bw->bit_idx;
        j -= k;
        bw->bit_idx -= k;
        bw->buf.back() |= (unsigned char) ((val >> j) << bw->bit_idx);
        bw->totbit += k;
    }
}

static void drain_ancillary(LgBitWriter *bw, const LgDevCfg *cfg, int remainingBits)
{
    static const char version[] = "3.99.5";      /* get_lame_short_version(), version.h:37-40 */
    int const sl = cfg->sideinfo_len;
    if (remainingBits >= 8) { putbits(bw, sl, 0x4c, 8); remainingBits -= 8; }
    if (remainingBits >= 8) { putbits(bw, sl, 0x41, 8); remainingBits -= 8; }
    if (remainingBits >= 8) { putbits(bw, sl, 0x4d, 8); remainingBits -= 8; }
    if (remainingBits >= 8) { putbits(bw, sl, 0x45, 8); remainingBits -= 8; }
    if (remainingBits >= 32)
        for (int i = 0; i < (int) strlen(version) && remainingBits >= 8; ++i) {
            remainingBits -= 8;
            putbits(bw, sl, version[i], 8);
        }
    for (; remainingBits >= 1; remainingBits -= 1) {
        putbits(bw, sl, bw->ancillary_flag, 1);
        bw->ancillary_flag ^= !cfg->disable_reservoir;
    }
}

static inline void writeheader(LgBitWriter *bw, int val, int j)
{
    int ptr = bw->header[bw->h_ptr].ptr;
    while (j > 0) {
        int const k = j < 8 - (ptr & 7) ? j : 8 - (ptr & 7);
        j -= k;
        bw->header[bw->h_ptr].buf[ptr >> 3] |= (unsigned char) (((val >> j)) << (8 - (ptr & 7) - k));
        ptr += k;
    }
    bw->header[bw->h_ptr].ptr = ptr;
}

static int frame_bits(const LgDevCfg *c, int padding)
{
    return 8 * ((c->version + 1) * 72000 * c->brate / c->samplerate + padding);
}

static void encode_side_info(LgBitWriter *bw, const LgDevCfg *cfg, const LgFrameOut *fo, const LgGranuleOut *g, int bitsPerFrame)
{
    bw->header[bw->h_ptr].ptr = 0;
    memset(bw->header[bw->h_ptr].buf, 0, cfg->sideinfo_len);
    writeheader(bw, 0xfff, 12);
    writeheader(bw, cfg->version, 1);
    writeheader(bw, 4 - 3, 2);
    writeheader(bw, !cfg->error_protection, 1);
    writeheader(bw, cfg->bitrate_index, 4);
    writeheader(bw, cfg->samplerate_index, 2);
    writeheader(bw, fo->padding, 1);
    writeheader(bw, cfg->extension, 1);
    writeheader(bw, cfg->mode, 2);
    writeheader(bw, fo->mode_ext, 2);
    writeheader(bw, cfg->copyright, 1);
    writeheader(bw, cfg->original, 1);
    writeheader(bw, cfg->emphasis, 2);
    writeheader(bw, fo->main_data_begin, 9);
    writeheader(bw, 0, cfg->channels == 2 ? 3 : 5);
    for (int ch = 0; ch < cfg->channels; ch++)
        for (int band = 0; band < 4; band++) writeheader(bw, fo->scfsi[ch][band], 1);
    for (int gr = 0; gr < 2; gr++)
        for (int ch = 0; ch < cfg->channels; ch++) {
            const LgGranuleOut *gi = &g[gr * 2 + ch];
            int t0 = gi->table_select[0], t1 = gi->table_select[1], t2 = gi->table_select[2];
            if (t0 == 14) t0 = 16;
            if (t1 == 14) t1 = 16;
            if (t2 == 14) t2 = 16;
            writeheader(bw, gi->part2_3_length + gi->part2_length, 12);
            writeheader(bw, gi->big_values / 2, 9);
            writeheader(bw, gi->global_gain, 8);
            writeheader(bw, gi->scalefac_compress, 4);
            if (gi->block_type != LG_NORM) {
                writeheader(bw, 1, 1);
                writeheader(bw, gi->block_type, 2);
                writeheader(bw, gi->mixed_block_flag, 1);
                writeheader(bw, t0, 5);
                writeheader(bw, t1, 5);
                writeheader(bw, gi->subblock_gain[0], 3);
                writeheader(bw, gi->subblock_gain[1], 3);
                writeheader(bw, gi->subblock_gain[2], 3);
            }
            else {
                writeheader(bw, 0, 1);
                writeheader(bw, t0, 5);
                writeheader(bw, t1, 5);
                writeheader(bw, t2, 5);
                writeheader(bw, gi->region0_count, 4);
                writeheader(bw, gi->region1_count, 3);
            }
            writeheader(bw, gi->preflag, 1);
            writeheader(bw, gi->scalefac_scale, 1);
            writeheader(bw, gi->count1table_select, 1);
        }
    int const old = bw->h_ptr;
    bw->h_ptr = (old + 1) & (LG_MAX_HEADER_BUF - 1);
    bw->header[bw->h_ptr].write_timing = bw->header[old].write_timing + bitsPerFrame;
}

static int huffman_code(LgBitWriter *bw, int sl, unsigned tableindex, int start, int end, const LgGranuleOut *gi)
{
    if (!tableindex) return 0;
    if (tableindex == 14) tableindex = 16;      /* encodeSideInfo2 rewrites 14 -> 16 before the data is written */
    const uint8_t *hlen = LGT_HUFF_LEN + LGT_HUFF_OFF[tableindex];
    const uint16_t *code = LGT_HUFF_CODE + LGT_HUFF_OFF[tableindex];
    unsigned const linbits = LGT_HUFF_XLEN[tableindex];
    int bits = 0;
    for (int i = start; i < end; i += 2) {
        int16_t cbits = 0;
        uint16_t xbits = 0;
        unsigned xlen = LGT_HUFF_XLEN[tableindex], ext = 0;
        int const s1 = gi->ix[i], s2 = gi->ix[i + 1];
        unsigned x1 = (unsigned) abs(s1), x2 = (unsigned) abs(s2);
        if (x1 != 0u) {
            if (s1 < 0) ext++;
            cbits--;
        }
        if (tableindex > 15u) {
            if (x1 >= 15u) {
                uint16_t const linbits_x1 = (uint16_t) (x1 - 15u);
                ext |= (unsigned) linbits_x1 << 1u;
                xbits = (uint16_t) linbits;
                x1 = 15u;
            }
            if (x2 >= 15u) {
                uint16_t const linbits_x2 = (uint16_t) (x2 - 15u);
                ext <<= linbits;
                ext |= linbits_x2;
                xbits = (uint16_t) (xbits + linbits);
                x2 = 15u;
            }
            xlen = 16;
        }
        if (x2 != 0u) {
            ext <<= 1;
            if (s2 < 0) ext++;
            cbits--;
        }
        x1 = x1 * xlen + x2;
        xbits = (uint16_t) (xbits - cbits);
        cbits = (int16_t) (cbits + hlen[x1]);
        putbits(bw, sl, code[x1], cbits);
        putbits(bw, sl, (int) ext, xbits);
        bits += cbits + xbits;
    }
    return bits;
}

static int huffman_count1(LgBitWriter *bw, int sl, const LgGranuleOut *gi)
{
    int const t = gi->count1table_select + 32;
    const uint8_t *hlen = LGT_HUFF_LEN + LGT_HUFF_OFF[t];
    const uint16_t *code = LGT_HUFF_CODE + LGT_HUFF_OFF[t];
    int bits = 0;
    const int16_t *ix = &gi->ix[gi->big_values];
    for (int i = (gi->count1 - gi->big_values) / 4; i > 0; --i) {
        int huffbits = 0, p = 0;
        if (ix[0]) { p += 8; if (ix[0] < 0) huffbits++; }
        if (ix[1]) { p += 4; huffbits *= 2; if (ix[1] < 0) huffbits++; }
        if (ix[2]) { p += 2; huffbits *= 2; if (ix[2] < 0) huffbits++; }
        if (ix[3]) { p++; huffbits *= 2; if (ix[3] < 0) huffbits++; }
        ix += 4;
        putbits(bw, sl, huffbits + code[p], hlen[p]);
        bits += hlen[p];
    }
    return bits;
}

static int write_main_data(LgBitWriter *bw, const LgDevCfg *cfg, const LgGranuleOut *g)
{
    int const sl = cfg->sideinfo_len;
    int tot_bits = 0;
    for (int gr = 0; gr < 2; gr++)
        for (int ch = 0; ch < cfg->channels; ch++) {
            const LgGranuleOut *gi = &g[gr * 2 + ch];
            int const slen1 = slen1_tab[gi->scalefac_compress], slen2 = slen2_tab[gi->scalefac_compress];
            int data_bits = 0, sfb;
            for (sfb = 0; sfb < gi->sfbdivide; sfb++) {
                if (gi->scalefac[sfb] == -1) continue;
                putbits(bw, sl, gi->scalefac[sfb], slen1);
                data_bits += slen1;
            }
            for (; sfb < gi->sfbmax; sfb++) {
                if (gi->scalefac[sfb] == -1) continue;
                putbits(bw, sl, gi->scalefac[sfb], slen2);
                data_bits += slen2;
            }
            if (gi->block_type == LG_SHORT) {
                int region1Start = 3 * cfg->sfb_s[3];
                if (region1Start > gi->big_values) region1Start = gi->big_values;
                data_bits += huffman_code(bw, sl, gi->table_select[0], 0, region1Start, gi);
                data_bits += huffman_code(bw, sl, gi->table_select[1], region1Start, gi->big_values, gi);
            }
            else {
                int const bigvalues = gi->big_values;
                unsigned i = gi->region0_count + 1;
                int region1Start = cfg->sfb_l[i];
                i += gi->region1_count + 1;
                int region2Start = cfg->sfb_l[i];
                if (region1Start > bigvalues) region1Start = bigvalues;
                if (region2Start > bigvalues) region2Start = bigvalues;
                data_bits += huffman_code(bw, sl, gi->table_select[0], 0, region1Start, gi);
                data_bits += huffman_code(bw, sl, gi->table_select[1], region1Start, region2Start, gi);
                data_bits += huffman_code(bw, sl, gi->table_select[2], region2Start, bigvalues, gi);
            }
            data_bits += huffman_count1(bw, sl, gi);
            tot_bits += data_bits;
        }
    return tot_bits;
}

void lg_pack_frame(LgBitWriter *bw, const LgDevCfg *cfg, const LgFrameOut *fo, const LgGranuleOut *g)
{
    int const bitsPerFrame = frame_bits(cfg, fo->padding);
    drain_ancillary(bw, cfg, fo->drain_pre);
    encode_side_info(bw, cfg, fo, g, bitsPerFrame);
    (void) write_main_data(bw, cfg, g);
    drain_ancillary(bw, cfg, fo->drain_post);
    if (bw->totbit > 1000000000) {
        for (int i = 0; i < LG_MAX_HEADER_BUF; ++i) bw->header[i].write_timing -= bw->totbit;
        bw->totbit = 0;
    }
}

/* The product path: kernel E has already formed the frame's bits.  hdr = header + side info (sideinfo_len bytes),
 * pay = the frame's payload (ancillary drain + main data, whole bytes).  What is left is what format_bitstream does
 * with its header ring (bitstream.c:918-960, putheader_bits :130): the payload goes out in order, and each pending
 * header is spliced in when the stream reaches the position where its frame starts. */
void lg_merge_frame(LgBitWriter *bw, const LgDevCfg *cfg, const LgFrameOut *fo, const unsigned char *hdr, const unsigned char *pay)
{
    int const sl = cfg->sideinfo_len;
    memcpy(bw->header[bw->h_ptr].buf, hdr, sl);
    int const old = bw->h_ptr;
    bw->h_ptr = (old + 1) & (LG_MAX_HEADER_BUF - 1);
    bw->header[bw->h_ptr].write_timing = bw->header[old].write_timing + frame_bits(cfg, fo->padding);
    long n = fo->pay_bytes;
    while (n > 0) {
        long const until = bw->header[bw->w_ptr].write_timing - bw->totbit;     /* bits to the next frame start */
        if (until == 0) {
            bw->buf.insert(bw->buf.end(), bw->header[bw->w_ptr].buf, bw->header[bw->w_ptr].buf + sl);
            bw->totbit += 8L * sl;
            bw->w_ptr = (bw->w_ptr + 1) & (LG_MAX_HEADER_BUF - 1);
            continue;
        }
        long chunk = until / 8;
        if (until < 0 || chunk > n) chunk = n;
        bw->buf.insert(bw->buf.end(), pay, pay + chunk);
        pay += chunk; n -= chunk;
        bw->totbit += 8 * chunk;
    }
    bw->ancillary_flag = fo->anc_post;
    if (bw->totbit > 1000000000) {
        for (int i = 0; i < LG_MAX_HEADER_BUF; ++i) bw->header[i].write_timing -= bw->totbit;
        bw->totbit = 0;
    }
}

void lg_pack_flush(LgBitWriter *bw, const LgDevCfg *cfg, int last_padding)
{
    int const first_ptr = bw->w_ptr;
    int last_ptr = bw->h_ptr - 1;
    if (last_ptr == -1) last_ptr = LG_MAX_HEADER_BUF - 1;
    long flushbits = bw->header[last_ptr].write_timing - bw->totbit;
    if (flushbits >= 0) {
        int remaining_headers = 1 + last_ptr - first_ptr;
        if (last_ptr < first_ptr) remaining_headers = 1 + last_ptr - first_ptr + LG_MAX_HEADER_BUF;
        flushbits -= remaining_headers * 8 * cfg->sideinfo_len;
    }
    flushbits += frame_bits(cfg, last_padding);
    if (flushbits < 0) return;
    drain_ancillary(bw, cfg, (int) flushbits);
}
