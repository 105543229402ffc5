/* hb_harness.h -- see hb_harness.c.  Plain-C interface for tests and bench. */
#ifndef HBCU_HB_HARNESS_H
#define HBCU_HB_HARNESS_H
#include "handbrake/handbrake.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_harness_io_s
{
    /* input */
    int             pix_fmt, width, height;
    int             n_in;
    const uint8_t  *in;          /* n_in packed planar frames */
    const uint16_t *in_flags;    /* per-frame s.flags, NULL = progressive */
    const uint8_t  *in_combed;   /* per-frame s.combed, NULL = HB_COMB_NONE */
    /* output */
    uint8_t        *out;         /* out_capacity packed planar frames (may be NULL) */
    int             out_capacity;
    uint8_t        *out_combed;
    uint16_t       *out_flags;
    int64_t        *out_start;
    int64_t        *out_stop;
    double         *out_duration;
    int             n_out;
    int             n_dropped;   /* outputs beyond out_capacity */
    int             saw_eof;
    int             init_failed; /* bit k set: filter k's init() returned non-zero */
    int             vrate_num_out, vrate_den_out;
} hb_harness_io_t;

size_t       hb_harness_frame_bytes(int pix_fmt, int w, int h);
hb_buffer_t *hb_harness_frame_from_packed(int pix_fmt, int w, int h, const uint8_t *src);
void         hb_harness_frame_to_packed(const hb_buffer_t *b, uint8_t *dst);
int hb_harness_run(hb_filter_object_t *proto, const char *settings, hb_harness_io_t *io);
int hb_harness_run_chain(int n_filters, hb_filter_object_t *const *protos,
                         const char *const *settings, hb_harness_io_t *io);

#ifdef __cplusplus
}
#endif
#endif
