// lg_bitstream.h - host bit packer state (one per stream).  Mirrors what the reference keeps in
// Bit_stream_struc + the header ring of EncStateVar_t (util.h:134, :272-282).
#pragma once
#include <stdint.h>
#include <vector>
#include "lg_types.h"

#define LG_MAX_HEADER_BUF 256

struct LgBitWriter {
    std::vector<unsigned char> buf;   /* bytes produced and not yet handed to the caller */
    int  bit_idx;                     /* free bits in the last byte of buf (0 = start a new byte) */
    long totbit;
    struct { long write_timing; int ptr; unsigned char buf[40]; } header[LG_MAX_HEADER_BUF];
    int  h_ptr, w_ptr, ancillary_flag;
    void reset();
};

void lg_header_crc(unsigned char *header, int sideinfo_len);
void lg_merge_frame(LgBitWriter *bw, const LgDevCfg *cfg, const LgFrameOut *fo, const unsigned char *hdr, const unsigned char *pay);
int  lg_pack_flush(LgBitWriter *bw, const LgDevCfg *cfg, int last_bitrate_index, int last_padding);
