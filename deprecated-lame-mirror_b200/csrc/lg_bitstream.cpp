// lg_bitstream.cpp - the part of the bit stream that is inherently serial per stream and stays on the host
// (SURVEY.md section 8b): splicing each frame's header + side info into the payload bytes kernel E packed
// (the bit reservoir lets main data start before its own header), and the final flush.  Follows bitstream.c:
// putbits2 :150, putheader_bits :130, drain_into_ancillary :214, format_bitstream :918, compute_flushbits :793,
// flush_bitstream :863.  Side info, scalefactors and Huffman code words are formed on the device (lg_k_pack.cuh).
#include <string.h>
#include <stdlib.h>
#include "lg_bitstream.h"


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

/* bitstream.c:65 getframebits for a frame's bitrate index */
static int frame_bits(const LgDevCfg *c, int bitrate_index, int padding)
{
    return 8 * ((c->version + 1) * 72000 * c->bitrate_kbps[bitrate_index] / c->samplerate + padding);
}

/* The product path: kernel E has already formed the frame's bits.  hdr = header + side info (sideinfo_len bytes),
 * pay = the frame's payload (ancillary drain + main data, whole bytes).  What is left is what format_bitstream does
 * with its header ring (bitstream.c:918-960, putheader_bits :130): the payload goes out in order, and each pending
 * header is spliced in when the stream reaches the position where its frame starts. */
/* bitstream.c:287 CRC_update + :304 CRC_writeheader: CRC-16 (x^16 + x^15 + x^2 + 1, MSB first, start 0xffff) over header bytes 2-3 and the
 * side info, stored in bytes 4-5 (error_protection) */
void lg_header_crc(unsigned char *header, int sideinfo_len)
{
    int crc = 0xffff;
    for (int i = 2; i < sideinfo_len; i++) {
        if (i == 4 || i == 5) continue;
        int value = header[i] << 8;
        for (int k = 0; k < 8; k++) {
            value <<= 1;
            crc <<= 1;
            if ((crc ^ value) & 0x10000) crc ^= 0x8005;
        }
    }
    header[4] = (unsigned char) (crc >> 8);
    header[5] = (unsigned char) (crc & 255);
}

void lg_merge_frame(LgBitWriter *bw, const LgDevCfg *cfg, const LgFrameOut *fo, const unsigned char *hdr, const unsigned char *pay)
{
    int const sl = cfg->sideinfo_len;
    memcpy(bw->header[bw->h_ptr].buf, hdr, sl);
    if (cfg->error_protection) lg_header_crc(bw->header[bw->h_ptr].buf, sl);
    int const old = bw->h_ptr;
    bw->h_ptr = (old + 1) & (LG_MAX_HEADER_BUF - 1);
    bw->header[bw->h_ptr].write_timing = bw->header[old].write_timing + frame_bits(cfg, fo->bitrate_index, fo->padding);
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

/* returns 1 when the stream was drained (the reference then ends the bit reservoir), 0 when there was nothing to flush */
int lg_pack_flush(LgBitWriter *bw, const LgDevCfg *cfg, int last_bitrate_index, int last_padding)
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
    flushbits += frame_bits(cfg, last_bitrate_index, last_padding);
    if (flushbits < 0) return 0;
    drain_ancillary(bw, cfg, (int) flushbits);
    return 1;
}
