// lg_api_stubs.cpp - libmp3lame exports that are OUTSIDE the accelerated path (the ID3 tag writer
// and the mpglib decoder), present so that programs written against include/lame.h - the reference's own frontend among them - link
// against liblamegpu.so unchanged (include/libmp3lame.sym).  They do not pretend: the tag setters say once that no tag will be
// written, the decoder entry points report failure the way a libmp3lame built without mpglib cannot even be asked.
#include <stdio.h>
#include <stddef.h>
#include <stdarg.h>
#include "../../include/lamegpu.h"

typedef struct hip_global_struct *hip_t;                      /* lame.h:1027 */
typedef struct { int header_parsed, stereo, samplerate, bitrate, mode, mode_ext, framesize; unsigned long nsamp; int totalframes, framenum; } mp3data_struct;   /* lame.h:998-1013 */

static void id3_notice(void)
{
    static int said = 0;
    if (!said) { said = 1; fprintf(stderr, "lamegpu: ID3 tags are not written by this library (the tag fields are ignored)\n"); }
}

extern "C" {
/* id3tag.h / lame.h:1176-1275 */
void id3tag_genre_list(void (*)(int, const char *, void *), void *) { }
void id3tag_init(lame_t) { }
void id3tag_add_v2(lame_t) { }
void id3tag_v1_only(lame_t) { }
void id3tag_v2_only(lame_t) { }
void id3tag_space_v1(lame_t) { }
void id3tag_pad_v2(lame_t) { }
void id3tag_set_pad(lame_t, size_t) { }
void id3tag_set_title(lame_t, const char *) { id3_notice(); }
void id3tag_set_artist(lame_t, const char *) { id3_notice(); }
void id3tag_set_album(lame_t, const char *) { id3_notice(); }
void id3tag_set_year(lame_t, const char *) { id3_notice(); }
void id3tag_set_comment(lame_t, const char *) { id3_notice(); }
int  id3tag_set_track(lame_t, const char *) { id3_notice(); return 0; }
int  id3tag_set_genre(lame_t, const char *) { id3_notice(); return 0; }
int  id3tag_set_fieldvalue(lame_t, const char *) { id3_notice(); return 0; }
int  id3tag_set_albumart(lame_t, const char *, size_t) { id3_notice(); return 0; }
int  id3tag_set_textinfo_latin1(lame_t, char const *, char const *) { id3_notice(); return 0; }
int  id3tag_set_comment_latin1(lame_t, char const *, char const *, char const *) { id3_notice(); return 0; }
int  id3tag_set_textinfo_ucs2(lame_t, char const *, unsigned short const *) { id3_notice(); return 0; }
int  id3tag_set_comment_ucs2(lame_t, char const *, unsigned short const *, unsigned short const *) { id3_notice(); return 0; }
int  id3tag_set_fieldvalue_ucs2(lame_t, const unsigned short *) { id3_notice(); return 0; }
int  id3tag_set_fieldvalue_utf16(lame_t, const unsigned short *) { id3_notice(); return 0; }
int  id3tag_set_textinfo_utf16(lame_t, char const *, unsigned short const *) { id3_notice(); return 0; }
int  id3tag_set_comment_utf16(lame_t, char const *, unsigned short const *, unsigned short const *) { id3_notice(); return 0; }
size_t lame_get_id3v1_tag(lame_t, unsigned char *, size_t) { return 0; }          /* lame.h:1257: 0 = no tag */
size_t lame_get_id3v2_tag(lame_t, unsigned char *, size_t) { return 0; }          /* lame.h:1245 */

/* mpglib_interface.c (decoder): lame.h:1032-1102 and the obsolete lame_decode_* forms :1105-1165 */
hip_t hip_decode_init(void) { fprintf(stderr, "lamegpu: no MP3 decoder in this library\n"); return NULL; }
int  hip_decode_exit(hip_t) { return 0; }
void hip_set_errorf(hip_t, lame_report_function) { }
void hip_set_debugf(hip_t, lame_report_function) { }
void hip_set_msgf(hip_t, lame_report_function) { }
int  hip_decode(hip_t, unsigned char *, size_t, short[], short[]) { return -1; }
int  hip_decode_headers(hip_t, unsigned char *, size_t, short[], short[], mp3data_struct *) { return -1; }
int  hip_decode1(hip_t, unsigned char *, size_t, short[], short[]) { return -1; }
int  hip_decode1_headers(hip_t, unsigned char *, size_t, short[], short[], mp3data_struct *) { return -1; }
int  hip_decode1_headersB(hip_t, unsigned char *, size_t, short[], short[], mp3data_struct *, int *, int *) { return -1; }
int  lame_decode_init(void) { return -1; }
int  lame_decode_exit(void) { return 0; }
int  lame_decode(unsigned char *, int, short[], short[]) { return -1; }
int  lame_decode_headers(unsigned char *, int, short[], short[], mp3data_struct *) { return -1; }
int  lame_decode1(unsigned char *, int, short[], short[]) { return -1; }
int  lame_decode1_headers(unsigned char *, int, short[], short[], mp3data_struct *) { return -1; }
int  lame_decode1_headersB(unsigned char *, int, short[], short[], mp3data_struct *, int *, int *) { return -1; }
}
