/* hb_bench.h -- see hb_bench.c */
#ifndef HBCU_HB_BENCH_H
#define HBCU_HB_BENCH_H
#include "handbrake/handbrake.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct hb_bench_s hb_bench_t;
typedef struct
{
    double   seconds;      /* wall clock, first work() call to last output consumed */
    int64_t  frames_out;
    int64_t  bytes_in;     /* host bytes handed to work() */
    int64_t  bytes_out;    /* host bytes received from work() */
    uint64_t checksum;
    int64_t  ring_misses;  /* input frames that found every ring payload still inside the chain (built the slow way) */
} hb_bench_stats_t;
/* one filter / a chain of filters in libhb's order; frame_flags = s.flags of every input frame */
hb_bench_t *hb_bench_open(hb_filter_object_t *proto, const char *settings, int pix_fmt, int w, int h);
hb_bench_t *hb_bench_open_chain(int n_filters, hb_filter_object_t *const *protos, const char *const *settings,
                                int pix_fmt, int w, int h, int frame_flags);
/* feed n_frames more frames of a running stream (inputs: decoder-style headers over a ring of `ring` pre-filled
 * payloads, 0 = default), consume the outputs; may be called repeatedly -- timestamps continue */
int hb_bench_stream(hb_bench_t *b, const uint8_t *src, int n_unique, int n_frames, int ring, hb_bench_stats_t *st);
/* fill the ring of input payloads now (before the first hb_bench_stream), e.g. while a different allocator mode is in force */
int hb_bench_prefill(hb_bench_t *b, const uint8_t *src, int n_unique, int ring);
/* EOF, flush, close and free */
int hb_bench_finish(hb_bench_t *b, hb_bench_stats_t *st);
/* close and free without the EOF flush (buffered frames are dropped) */
int hb_bench_abort(hb_bench_t *b);
/* stream + finish */
int hb_bench_run(hb_bench_t *b, const uint8_t *src, int n_unique, int n_frames, hb_bench_stats_t *st);
/* a chain of filters in libhb's order (init, stream, EOF, close); frame_flags = s.flags of every input frame */
int hb_bench_run_chain(int n_filters, hb_filter_object_t *const *protos, const char *const *settings,
                       int pix_fmt, int w, int h, int frame_flags,
                       const uint8_t *src, int n_unique, int n_frames, hb_bench_stats_t *st);
#ifdef __cplusplus
}
#endif
#endif
