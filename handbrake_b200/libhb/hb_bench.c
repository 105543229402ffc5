/* hb_bench.c -- streaming benchmark driver over the libhb filter interface.
 *
 * Measures a filter the way libhb runs it (filter_loop, work.c:2527-2600): a
 * stream of host hb_buffer_t frames goes through work(), outputs are consumed
 * in order, EOF flushes.  The same driver times the CUDA objects and -- linked
 * into oracle/_ref/libhbref.so -- the reference's own CPU objects, so both arms
 * of bench.py run the identical host protocol.
 *
 * Input buffers are created (and filled) before the clock starts: in libhb the
 * decoder writes straight into the hb_buffer_t, so that copy is not part of
 * the filter.  Everything after -- host->device, kernels, device->host into a
 * fresh output hb_buffer_t, buffer release -- is inside the timed region.
 */
#define _GNU_SOURCE
#include "handbrake/handbrake.h"
#include "hb_harness.h"
#include "hb_bench.h"

#include <time.h>

struct hb_bench_s
{
    hb_filter_object_t *f;
    int pix_fmt, w, h;
    int volatile done;
};

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

hb_bench_t *hb_bench_open(hb_filter_object_t *proto, const char *settings, int pix_fmt, int w, int h)
{
    hb_bench_t *b = calloc(1, sizeof(*b));
    b->f = malloc(sizeof(*b->f));
    memcpy(b->f, proto, sizeof(*b->f));
    b->f->settings = settings ? hb_parse_filter_settings(settings) : NULL;
    b->f->done = &b->done;
    b->pix_fmt = pix_fmt;
    b->w = w;
    b->h = h;
    hb_filter_init_t init;
    memset(&init, 0, sizeof(init));
    init.pix_fmt = pix_fmt;
    init.geometry.width = w;
    init.geometry.height = h;
    init.geometry.par.num = init.geometry.par.den = 1;
    init.vrate.num = 30000;
    init.vrate.den = 1001;
    init.time_base.num = 1;
    init.time_base.den = 90000;
    if (b->f->init(b->f, &init) != 0)
    {
        if (b->f->settings) hb_dict_free(&b->f->settings);
        free(b->f);
        free(b);
        return NULL;
    }
    return b;
}

static void consume(hb_buffer_t *list, hb_bench_stats_t *st)
{
    while (list != NULL)
    {
        hb_buffer_t *b = list;
        list = b->next;
        b->next = NULL;
        if (!(b->s.flags & HB_BUF_FLAG_EOF) && b->storage_type != HBCU_DEVICE)
        {
            /* read the result on the host: one sample per plane row start */
            for (int p = 0; p <= b->f.max_plane; p++)
                for (int y = 0; y < b->plane[p].height; y += 16)
                    st->checksum += b->plane[p].data[(size_t)y * b->plane[p].stride + (y & 31)];
            st->bytes_out += hb_harness_frame_bytes(b->f.fmt, b->f.width, b->f.height);
            st->frames_out++;
        }
        hb_buffer_close(&b);
    }
}

/* feeds n_frames (cycling over n_unique packed source frames) and an EOF; closes the filter */
int hb_bench_run(hb_bench_t *b, const uint8_t *src, int n_unique, int n_frames, hb_bench_stats_t *st)
{
    memset(st, 0, sizeof(*st));
    const size_t fb = hb_harness_frame_bytes(b->pix_fmt, b->w, b->h);
    hb_buffer_t **in = calloc(n_frames, sizeof(*in));
    for (int i = 0; i < n_frames; i++)
    {
        in[i] = hb_harness_frame_from_packed(b->pix_fmt, b->w, b->h, src + (size_t)(i % n_unique) * fb);
        if (in[i] == NULL)
        {
            for (int j = 0; j < i; j++) hb_buffer_close(&in[j]);
            free(in);
            return -1;
        }
        in[i]->s.start = (int64_t)i * 3003;
        in[i]->s.stop = in[i]->s.start + 3003;
        in[i]->s.duration = 3003;
        in[i]->s.flags = PIC_FLAG_PROGRESSIVE_FRAME;
    }
    hb_buffer_t *eof = hb_buffer_eof_init();
    int rc = 0;

    const double t0 = now_s();
    for (int i = 0; i <= n_frames && rc == 0; i++)
    {
        hb_buffer_t *buf = i < n_frames ? in[i] : eof;
        hb_buffer_t *out = NULL;
        if (i < n_frames)
        {
            in[i] = NULL;
            st->bytes_in += fb;
        }
        else
        {
            eof = NULL;
        }
        int status = b->f->work(b->f, &buf, &out);
        if (buf != NULL) hb_buffer_close(&buf);
        if (status == HB_FILTER_FAILED) rc = -1;
        consume(out, st);
    }
    st->seconds = now_s() - t0;

    for (int i = 0; i < n_frames; i++) if (in[i]) hb_buffer_close(&in[i]);
    if (eof) hb_buffer_close(&eof);
    free(in);
    b->f->close(b->f);
    if (b->f->settings) hb_dict_free(&b->f->settings);
    free(b->f);
    free(b);
    return rc;
}

/* ------------------------------------------------------------------ */
/* chains: the filters in libhb's order, each fed by the previous one    */
/* ------------------------------------------------------------------ */
typedef struct
{
    int n;
    hb_filter_object_t **f;
    int *done;
    int failed;
    hb_bench_stats_t *st;
} bench_chain_t;

static void chain_feed(bench_chain_t *c, int k, hb_buffer_t *list)
{
    if (k >= c->n)
    {
        consume(list, c->st);
        return;
    }
    while (list != NULL)
    {
        hb_buffer_t *in = list;
        list = in->next;
        in->next = NULL;
        if (c->done[k])
        {
            hb_buffer_close(&in);
            continue;
        }
        hb_buffer_t *out = NULL;
        const int status = c->f[k]->work(c->f[k], &in, &out);
        if (in != NULL) hb_buffer_close(&in);
        if (status == HB_FILTER_FAILED) c->failed = 1;
        if (out != NULL) chain_feed(c, k + 1, out);
        if (status == HB_FILTER_DONE) c->done[k] = 1;
    }
}

int hb_bench_run_chain(int n_filters, hb_filter_object_t *const *protos, const char *const *settings,
                       int pix_fmt, int w, int h, int frame_flags,
                       const uint8_t *src, int n_unique, int n_frames, hb_bench_stats_t *st)
{
    memset(st, 0, sizeof(*st));
    bench_chain_t c;
    memset(&c, 0, sizeof(c));
    c.f = calloc(n_filters, sizeof(*c.f));
    c.done = calloc(n_filters, sizeof(int));
    c.st = st;
    int volatile done_flag = 0;
    hb_filter_init_t init;
    memset(&init, 0, sizeof(init));
    init.pix_fmt = pix_fmt;
    init.geometry.width = w;
    init.geometry.height = h;
    init.geometry.par.num = init.geometry.par.den = 1;
    init.vrate.num = 30000;
    init.vrate.den = 1001;
    for (int k = 0; k < n_filters; k++)
    {
        hb_filter_object_t *f = malloc(sizeof(*f));
        memcpy(f, protos[k], sizeof(*f));
        f->settings = settings && settings[k] ? hb_parse_filter_settings(settings[k]) : NULL;
        f->done = &done_flag;
        if (f->sub_filter != NULL)
        {
            hb_filter_object_t *sub = malloc(sizeof(*sub));
            memcpy(sub, f->sub_filter, sizeof(*sub));
            sub->settings = settings && settings[k] ? hb_parse_filter_settings(settings[k]) : NULL;
            f->sub_filter = sub;
        }
        if (f->init(f, &init) != 0)
        {
            free(f);
            free(c.f);
            free(c.done);
            return -2;
        }
        c.f[c.n++] = f;
    }
    const size_t fb = hb_harness_frame_bytes(pix_fmt, w, h);
    hb_buffer_t **in = calloc(n_frames, sizeof(*in));
    for (int i = 0; i < n_frames; i++)
    {
        in[i] = hb_harness_frame_from_packed(pix_fmt, w, h, src + (size_t)(i % n_unique) * fb);
        in[i]->s.start = (int64_t)i * 3003;
        in[i]->s.stop = in[i]->s.start + 3003;
        in[i]->s.duration = 3003;
        in[i]->s.flags = (uint16_t)frame_flags;
    }
    hb_buffer_t *eof = hb_buffer_eof_init();
    const double t0 = now_s();
    for (int i = 0; i < n_frames && !c.failed; i++)
    {
        hb_buffer_t *b = in[i];
        in[i] = NULL;
        st->bytes_in += fb;
        chain_feed(&c, 0, b);
    }
    if (!c.failed)
    {
        chain_feed(&c, 0, eof);
        eof = NULL;
    }
    st->seconds = now_s() - t0;
    for (int i = 0; i < n_frames; i++) if (in[i]) hb_buffer_close(&in[i]);
    if (eof) hb_buffer_close(&eof);
    free(in);
    for (int k = 0; k < c.n; k++)
    {
        c.f[k]->close(c.f[k]);
        if (c.f[k]->settings) hb_dict_free(&c.f[k]->settings);
        if (c.f[k]->sub_filter) { if (c.f[k]->sub_filter->settings) hb_dict_free(&c.f[k]->sub_filter->settings); free(c.f[k]->sub_filter); }
        free(c.f[k]);
    }
    free(c.f);
    free(c.done);
    return c.failed ? -1 : 0;
}
