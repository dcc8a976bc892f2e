// lg_k_pack.cuh - kernel E: the frame's bits, formed on the device.
//
// The reference's bit packer (bitstream.c) is serial per stream only because every code word starts where
// the previous one ended.  Kernel D already knows every length (part2_length, part2_3_length, the reservoir
// drains), so the start of each frame, of each granule.channel inside it, and - after a prefix sum over code
// lengths - of every code word is known up front, and the words can be written independently:
//
//   one CTA per (stream, frame), one warp per granule.channel
//   payload of a frame = drain_pre ancillary bits | main data of gr0ch0 gr0ch1 gr1ch0 gr1ch1 | drain_post bits
//   (a whole number of bytes, see kernel D); each lane owns a contiguous run of line pairs / count1
//   quadruples, sums their code lengths, a warp scan turns the sums into bit offsets, and the lane ORs its
//   code words into the frame's shared-memory image; the image is stored byte-swapped to HBM.
//   The 4-byte header + side info (encodeSideInfo2, bitstream.c:321) is built the same way into 36 bytes.
//
// What stays on the host is only what is inherently about the byte stream of one lame_t: inserting each
// frame's header where the previous frame ends (the main-data/header interleave of the bit reservoir) and
// handing bytes to the caller - memcpy work (lg_bitstream.cpp lg_merge_frame).
//
// Reference: writeMainData bitstream.c:686, Huffmancode :561, huffman_coder_count1 :482, encodeSideInfo2 :321,
// drain_into_ancillary :214.  Algorithmic HBM bytes per frame: read 4 x sizeof(LgGranuleOut), write <= frame bytes.
#pragma once
#include "lg_compat.h"
#include "lg_types.h"

#define LG_PACK_WORDS 1024              /* frame payload image: 4096 bytes >= 2 x 7680 bits + the largest drain */

struct LgSmemE {
    unsigned img[LG_PACK_WORDS];
    unsigned hdr[LG_HDR_STRIDE / 4];
    int gstart[4];
};

/* OR the low n bits of v (n <= 32) into the big-endian bit image at bit offset o */
__device__ __forceinline__ void lg_put(unsigned *img, int o, unsigned v, int n)
{
    if (n <= 0) return;
    unsigned long long const x = ((unsigned long long) v << (64 - n)) >> (o & 31);
    unsigned const hi = (unsigned) (x >> 32), lo = (unsigned) x;
#ifdef LG_EMULATE
    if (hi) __atomic_fetch_or(&img[o >> 5], hi, __ATOMIC_RELAXED);
    if (lo) __atomic_fetch_or(&img[(o >> 5) + 1], lo, __ATOMIC_RELAXED);
#else
    if (hi) atomicOr(&img[o >> 5], hi);
    if (lo) atomicOr(&img[(o >> 5) + 1], lo);
#endif
}

__device__ __forceinline__ int lg_warp_excl_scan(int v, int lane, int *total)
{
    int s = v;
    for (int d = 1; d < 32; d <<= 1) { int const o = __shfl_up_sync(LG_FULL, s, d); if (lane >= d) s += o; }
    *total = __shfl_sync(LG_FULL, s, 31);
    return s - v;
}

/* bitstream.c:214 drain_into_ancillary: n bits at offset o, whole warp */
__device__ __forceinline__ void lg_put_drain(unsigned *img, int o, int n, int flag, int toggles, int lane)
{
    const char tag[10] = { 0x4c, 0x41, 0x4d, 0x45, '3', '.', '9', '9', '.', '5' };
    int nb = 0;                                      /* leading tag bytes */
    int rem = n;
    for (int i = 0; i < 4; i++) if (rem >= 8) { rem -= 8; nb++; }
    if (rem >= 32) for (int i = 0; i < 6 && rem >= 8; ++i) { rem -= 8; nb++; }
    if (lane < nb) lg_put(img, o + 8 * lane, (unsigned) tag[lane], 8);
    o += 8 * nb;
    /* rem single bits: flag, !flag, flag, ... (or constant when the reservoir is disabled) */
    for (int k = 32 * lane; k < rem; k += 1024) {
        int const m = rem - k < 32 ? rem - k : 32;
        unsigned pat = toggles ? (flag ? 0xaaaaaaaau : 0x55555555u) : (flag ? 0xffffffffu : 0u);
        /* bit k (even offset from the start of the run) carries `flag`: the MSB of pat */
        pat >>= (32 - m);                            /* the first m bits of the pattern */
        lg_put(img, o + k, pat, m);
    }
}

/* one big-value pair: code word + sign/linbits extension (bitstream.c:561 Huffmancode) -> (bits, total length) */
__device__ __forceinline__ int lg_pair_code(const LgDevCfg *__restrict__ c, int t, int s1, int s2, unsigned *code, int *cbits_o, unsigned *ext_o, int *xbits_o)
{
    if (t == 14) t = 16;                             /* encodeSideInfo2 rewrites 14 -> 16 */
    const uint8_t *hlen = c->huff_len + c->huff_off[t];
    const uint16_t *hcode = c->huff_code + c->huff_off[t];
    unsigned const linbits = c->huff_xlen[t];
    unsigned xlen = linbits, ext = 0;
    int cbits = 0, xbits = 0;
    unsigned x1 = (unsigned) (s1 < 0 ? -s1 : s1), x2 = (unsigned) (s2 < 0 ? -s2 : s2);
    if (x1 != 0u) { if (s1 < 0) ext++; cbits--; }
    if (t > 15) {
        if (x1 >= 15u) { ext |= (x1 - 15u) << 1; xbits = (int) linbits; x1 = 15u; }
        if (x2 >= 15u) { ext <<= linbits; ext |= (x2 - 15u); xbits += (int) linbits; x2 = 15u; }
        xlen = 16;
    }
    if (x2 != 0u) { ext <<= 1; if (s2 < 0) ext++; cbits--; }
    x1 = x1 * xlen + x2;
    xbits -= cbits;
    cbits += __ldg(&hlen[x1]);
    *code = __ldg(&hcode[x1]); *cbits_o = cbits; *ext_o = ext; *xbits_o = xbits;
    return cbits + xbits;
}

__global__ void __launch_bounds__(128)
lg_kernel_pack(const LgDevCfg *__restrict__ c, const LgGranuleOut *__restrict__ gout, const LgFrameOut *__restrict__ fout,
               unsigned char *__restrict__ pay, int pay_stride, unsigned char *__restrict__ hdr_out,
               const int *__restrict__ nfr, int nframes, int f0, int cnt /* this launch: frames f0 .. f0+cnt-1 */)
{
    LG_DYN_SMEM(LgSmemE, sm);
    int const stream = blockIdx.x / cnt, frame = f0 + blockIdx.x % cnt;
    if (frame >= nfr[stream]) return;
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int const nch = c->channels;
    const LgFrameOut *fo = fout + (size_t) stream * nframes + frame;
    int const crc = c->error_protection ? 16 : 0;   /* bits left free behind the header for the CRC the host fills in (bitstream.c:348) */
    int const mgr = c->mode_gr;                   /* MPEG-2/2.5: one granule per frame, warps 2 and 3 only help with the drains */
    const LgGranuleOut *g4 = gout + ((size_t) stream * 2 * nframes + mgr * frame) * 2;
    int const pay_bytes = fo->pay_bytes;
    int const nwords = (pay_bytes + 3) >> 2;
    for (int i = threadIdx.x; i <= nwords && i < LG_PACK_WORDS; i += 128) sm->img[i] = 0u;
    if (threadIdx.x < LG_HDR_STRIDE / 4) sm->hdr[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        int o = fo->drain_pre;
        for (int k = 0; k < 4; k++) {
            sm->gstart[k] = o;
            if ((k & 1) < nch && (k >> 1) < mgr) o += g4[k].part2_3_length + g4[k].part2_length;
        }
    }
    __syncthreads();
    int const gr = warp >> 1, ch = warp & 1;
    if (ch < nch && gr < mgr) {
        const LgGranuleOut *gi = &g4[warp];
        int pos = sm->gstart[warp];
        /* ---- scale factors (writeMainData bitstream.c:700-718; MPEG-2/2.5 :747-775: every band of a partition with the partition's
         * width, which scalefac_compress encodes (takehiro.c:1281-1297)) */
        if (mgr == 1) {
            int const sfc = gi->scalefac_compress | (gi->scalefac_compress_hi << 8), sh = gi->block_type == LG_SHORT;
            int slen[5] = { 0, 0, 0, 0, 0 };
            if (gi->preflag) { slen[0] = (sfc - 500) / 3; slen[1] = (sfc - 500) % 3; }
            else { slen[0] = (sfc >> 4) / 5; slen[1] = (sfc >> 4) % 5; slen[2] = (sfc >> 2) & 3; slen[3] = sfc & 3; }
            for (int r = 0; r < 2; r++) {
                int const sfb = lane + 32 * r;
                int n = 0, v = 0;
                if (sfb < 39) { n = slen[lg_lsf_partition(sh, gi->preflag, sfb)]; v = gi->scalefac[sfb] > 0 ? gi->scalefac[sfb] : 0; }
                int tot;
                int const off = lg_warp_excl_scan(n, lane, &tot);
                lg_put(sm->img, pos + off, (unsigned) v, n);
                pos += tot;
            }
        }
        else {
            int const slen1 = LG_SLEN1_TAB[gi->scalefac_compress], slen2 = LG_SLEN2_TAB[gi->scalefac_compress];
            for (int r = 0; r < 2; r++) {
                int const sfb = lane + 32 * r;
                int n = 0, v = 0;
                if (sfb < gi->sfbmax) { v = gi->scalefac[sfb]; n = (v == -1) ? 0 : (sfb < gi->sfbdivide ? slen1 : slen2); }
                int tot;
                int const off = lg_warp_excl_scan(n, lane, &tot);
                lg_put(sm->img, pos + off, (unsigned) v, n);
                pos += tot;
            }
        }
        /* ---- big values: three regions (Huffmancode), a lane owns nine consecutive pairs */
        int const bigv = gi->big_values;
        int r1s, r2s;
        if (gi->block_type == LG_SHORT) {
            r1s = 3 * c->sfb_s[3];
            if (r1s > bigv) r1s = bigv;
            r2s = bigv;
        }
        else {
            int i = gi->region0_count + 1;
            r1s = c->sfb_l[i];
            i += gi->region1_count + 1;
            r2s = c->sfb_l[i];
            if (r1s > bigv) r1s = bigv;
            if (r2s > bigv) r2s = bigv;
        }
        int const t0 = gi->table_select[0], t1 = gi->table_select[1], t2 = gi->table_select[2];
        const unsigned *ixp = reinterpret_cast<const unsigned *>(gi->ix);
        int mylen = 0;
        for (int k = 0; k < 9; k++) {
            int const p = 9 * lane + k, i = 2 * p;
            if (i < bigv) {
                int const t = i < r1s ? t0 : (i < r2s ? t1 : t2);
                if (t) {
                    unsigned const u = __ldg(&ixp[p]);
                    unsigned code, ext; int cb, xb;
                    mylen += lg_pair_code(c, t, (int) (short) (u & 0xffffu), (int) (short) (u >> 16), &code, &cb, &ext, &xb);
                }
            }
        }
        int tot;
        int o = pos + lg_warp_excl_scan(mylen, lane, &tot);
        for (int k = 0; k < 9; k++) {
            int const p = 9 * lane + k, i = 2 * p;
            if (i < bigv) {
                int const t = i < r1s ? t0 : (i < r2s ? t1 : t2);
                if (t) {
                    unsigned const u = __ldg(&ixp[p]);
                    unsigned code, ext; int cb, xb;
                    (void) lg_pair_code(c, t, (int) (short) (u & 0xffffu), (int) (short) (u >> 16), &code, &cb, &ext, &xb);
                    lg_put(sm->img, o, code, cb);
                    lg_put(sm->img, o + cb, ext, xb);
                    o += cb + xb;
                }
            }
        }
        pos += tot;
        /* ---- count1 quadruples (huffman_coder_count1 bitstream.c:482) */
        {
            int const tq = gi->count1table_select + 32;
            const uint8_t *hlen = c->huff_len + c->huff_off[tq];
            const uint16_t *hcode = c->huff_code + c->huff_off[tq];
            int const nq = (gi->count1 - bigv) / 4;
            int const per = (nq + 31) >> 5;
            int const q0 = lane * per, q1 = (q0 + per < nq) ? q0 + per : nq;
            int len = 0;
            for (int q = q0; q < q1; q++) {
                const int16_t *ix = &gi->ix[bigv + 4 * q];
                int const p = (ix[0] ? 8 : 0) + (ix[1] ? 4 : 0) + (ix[2] ? 2 : 0) + (ix[3] ? 1 : 0);
                len += __ldg(&hlen[p]);
            }
            int o2 = pos + lg_warp_excl_scan(len, lane, &tot);
            for (int q = q0; q < q1; q++) {
                const int16_t *ix = &gi->ix[bigv + 4 * q];
                int huffbits = 0, p = 0;
                if (ix[0]) { p += 8; if (ix[0] < 0) huffbits++; }
                if (ix[1]) { p += 4; huffbits *= 2; if (ix[1] < 0) huffbits++; }
                if (ix[2]) { p += 2; huffbits *= 2; if (ix[2] < 0) huffbits++; }
                if (ix[3]) { p++; huffbits *= 2; if (ix[3] < 0) huffbits++; }
                int const n = __ldg(&hlen[p]);
                lg_put(sm->img, o2, (unsigned) (huffbits + __ldg(&hcode[p])), n);
                o2 += n;
            }
            pos += tot;
        }
        /* the packer must land exactly where kernel D's bit count said it would (bitstream.c:728 assert) */
        if (lane == 0 && pos != sm->gstart[warp] + gi->part2_3_length + gi->part2_length) {
#ifdef LG_EMULATE
            abort();
#else
            __trap();
#endif
        }
        /* ---- this granule.channel's 59 bits of side info (encodeSideInfo2 bitstream.c:409-460) */
        if (lane == 0 && mgr == 1) {
            /* bitstream.c:415-466: MPEG-2/2.5 side info of one gr.ch = 63 bits behind the 8-bit main_data_begin and the private bits */
            int so = 32 + crc + 8 + nch + 63 * ch;
            int a0 = t0, a1 = t1, a2 = t2;
            if (a0 == 14) a0 = 16;
            if (a1 == 14) a1 = 16;
            if (a2 == 14) a2 = 16;
            lg_put(sm->hdr, so, (unsigned) (gi->part2_3_length + gi->part2_length), 12); so += 12;
            lg_put(sm->hdr, so, (unsigned) (bigv / 2), 9); so += 9;
            lg_put(sm->hdr, so, gi->global_gain, 8); so += 8;
            lg_put(sm->hdr, so, (unsigned) (gi->scalefac_compress | (gi->scalefac_compress_hi << 8)), 9); so += 9;
            if (gi->block_type != LG_NORM) {
                lg_put(sm->hdr, so, 1u, 1); so += 1;
                lg_put(sm->hdr, so, gi->block_type, 2); so += 2;
                lg_put(sm->hdr, so, gi->mixed_block_flag, 1); so += 1;
                lg_put(sm->hdr, so, (unsigned) a0, 5); so += 5;
                lg_put(sm->hdr, so, (unsigned) a1, 5); so += 5;
                lg_put(sm->hdr, so, gi->subblock_gain[0], 3); so += 3;
                lg_put(sm->hdr, so, gi->subblock_gain[1], 3); so += 3;
                lg_put(sm->hdr, so, gi->subblock_gain[2], 3); so += 3;
            }
            else {
                so += 1;
                lg_put(sm->hdr, so, (unsigned) a0, 5); so += 5;
                lg_put(sm->hdr, so, (unsigned) a1, 5); so += 5;
                lg_put(sm->hdr, so, (unsigned) a2, 5); so += 5;
                lg_put(sm->hdr, so, gi->region0_count, 4); so += 4;
                lg_put(sm->hdr, so, gi->region1_count, 3); so += 3;
            }
            lg_put(sm->hdr, so, gi->scalefac_scale, 1); so += 1;
            lg_put(sm->hdr, so, gi->count1table_select, 1);
        }
        else if (lane == 0) {
            int so = 32 + crc + 9 + (nch == 2 ? 3 : 5) + 4 * nch + 59 * (gr * nch + ch);
            int a0 = t0, a1 = t1, a2 = t2;
            if (a0 == 14) a0 = 16;
            if (a1 == 14) a1 = 16;
            if (a2 == 14) a2 = 16;
            lg_put(sm->hdr, so, (unsigned) (gi->part2_3_length + gi->part2_length), 12); so += 12;
            lg_put(sm->hdr, so, (unsigned) (bigv / 2), 9); so += 9;
            lg_put(sm->hdr, so, gi->global_gain, 8); so += 8;
            lg_put(sm->hdr, so, gi->scalefac_compress, 4); so += 4;
            if (gi->block_type != LG_NORM) {
                lg_put(sm->hdr, so, 1u, 1); so += 1;
                lg_put(sm->hdr, so, gi->block_type, 2); so += 2;
                lg_put(sm->hdr, so, gi->mixed_block_flag, 1); so += 1;
                lg_put(sm->hdr, so, (unsigned) a0, 5); so += 5;
                lg_put(sm->hdr, so, (unsigned) a1, 5); so += 5;
                lg_put(sm->hdr, so, gi->subblock_gain[0], 3); so += 3;
                lg_put(sm->hdr, so, gi->subblock_gain[1], 3); so += 3;
                lg_put(sm->hdr, so, gi->subblock_gain[2], 3); so += 3;
            }
            else {
                so += 1;
                lg_put(sm->hdr, so, (unsigned) a0, 5); so += 5;
                lg_put(sm->hdr, so, (unsigned) a1, 5); so += 5;
                lg_put(sm->hdr, so, (unsigned) a2, 5); so += 5;
                lg_put(sm->hdr, so, gi->region0_count, 4); so += 4;
                lg_put(sm->hdr, so, gi->region1_count, 3); so += 3;
            }
            lg_put(sm->hdr, so, gi->preflag, 1); so += 1;
            lg_put(sm->hdr, so, gi->scalefac_scale, 1); so += 1;
            lg_put(sm->hdr, so, gi->count1table_select, 1);
        }
    }
    /* bytes 40..43 of the record, behind the longest side info: the four block types (4 = mixed, 0xff = no such channel) for
     * the host's statistics (encoder.c:156 updateStats) */
    if (lane == 0) lg_put(sm->hdr, 320 + 8 * warp, (ch < nch && gr < mgr) ? (g4[warp].mixed_block_flag ? 4u : (unsigned) g4[warp].block_type) : 0xffu, 8);
    /* ---- ancillary drains and the frame header (warps 2 and 3 are the lighter ones in joint stereo) */
    if (warp == 3) lg_put_drain(sm->img, 0, fo->drain_pre, fo->anc_pre, !c->disable_reservoir, lane);
    if (warp == 2) {
        int flag = fo->anc_pre;
        if (!c->disable_reservoir) {
            int rem = fo->drain_pre;
            for (int i = 0; i < 4; i++) if (rem >= 8) rem -= 8;
            if (rem >= 32) for (int i = 0; i < 6 && rem >= 8; ++i) rem -= 8;
            flag ^= rem & 1;
        }
        lg_put_drain(sm->img, 8 * pay_bytes - fo->drain_post, fo->drain_post, flag, !c->disable_reservoir, lane);
    }
    if (threadIdx.x == 33) {
        /* bitstream.c:330-372: header, main_data_begin, private bits, scfsi */
        int so = 0;
        lg_put(sm->hdr, so, c->samplerate < 16000 ? 0xffeu : 0xfffu, 12); so += 12;     /* MPEG-2.5 clears the last sync bit */
        lg_put(sm->hdr, so, (unsigned) c->version, 1); so += 1;
        lg_put(sm->hdr, so, 4 - 3, 2); so += 2;
        lg_put(sm->hdr, so, !c->error_protection, 1); so += 1;
        lg_put(sm->hdr, so, (unsigned) fo->bitrate_index, 4); so += 4;
        lg_put(sm->hdr, so, (unsigned) c->samplerate_index, 2); so += 2;
        lg_put(sm->hdr, so, (unsigned) fo->padding, 1); so += 1;
        lg_put(sm->hdr, so, (unsigned) c->extension, 1); so += 1;
        lg_put(sm->hdr, so, (unsigned) c->mode, 2); so += 2;
        lg_put(sm->hdr, so, (unsigned) fo->mode_ext, 2); so += 2;
        lg_put(sm->hdr, so, (unsigned) c->copyright, 1); so += 1;
        lg_put(sm->hdr, so, (unsigned) c->original, 1); so += 1;
        lg_put(sm->hdr, so, (unsigned) c->emphasis, 2); so += 2;
        so += crc;
        if (mgr == 1) lg_put(sm->hdr, so, (unsigned) fo->main_data_begin, 8);          /* then nch private bits (0), no scfsi */
        else {
            lg_put(sm->hdr, so, (unsigned) fo->main_data_begin, 9); so += 9;
            so += (nch == 2 ? 3 : 5);
            for (int k = 0; k < nch; k++)
                for (int band = 0; band < 4; band++) { lg_put(sm->hdr, so, fo->scfsi[k][band], 1); so += 1; }
        }
    }
    __syncthreads();
    /* ---- image -> HBM (big-endian bit order = byte-swapped words) */
    {
        unsigned char *dst = pay + (size_t) stream * pay_stride + fo->pay_off;
        for (int i = threadIdx.x; i < pay_bytes; i += 128) dst[i] = (unsigned char) (sm->img[i >> 2] >> (24 - 8 * (i & 3)));
        unsigned char *hd = hdr_out + ((size_t) stream * nframes + frame) * LG_HDR_STRIDE;
        if (threadIdx.x < LG_HDR_STRIDE) hd[threadIdx.x] = (unsigned char) (sm->hdr[threadIdx.x >> 2] >> (24 - 8 * (threadIdx.x & 3)));
    }
}
