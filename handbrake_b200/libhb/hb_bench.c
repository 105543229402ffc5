/* hb_bench.c -- streaming benchmark driver over the libhb filter interface.
 *
 * Measures a filter, or a chain of filters in libhb's order, the way libhb runs
 * them (filter_loop, work.c:2527-2600): a stream of host hb_buffer_t frames goes
 * through work(), outputs are consumed in order, EOF flushes.  The same driver
 * times the CUDA objects and -- linked into oracle/_ref/libhbref.so -- the
 * reference's own CPU objects, so both arms of bench.py run the identical host
 * protocol.
 *
 * Input frames are decoder-style buffers: a bounded ring of frame payloads is
 * filled once before the clock starts (in libhb the decoder writes straight
 * into the hb_buffer_t, so that copy is not part of any filter) and every
 * timed frame is a fresh hb_buffer_t header over the next free payload;
 * hb_buffer_close() -- wherever in the chain it happens -- hands the payload
 * back to the ring (hb_shim_buffer_set_release, the way libhb returns wrapped
 * AVFrame memory to the decoder).  A stream can therefore run for seconds
 * without one pinned buffer per frame.  Everything after the hand-over --
 * host->device, kernels, device->host into a fresh output hb_buffer_t, buffer
 * release -- is inside the timed region.
 */
#define _GNU_SOURCE
#include "handbrake/handbrake.h"
#include "hb_harness.h"
#include "hb_bench.h"

#include <pthread.h>
#include <time.h>

#define BENCH_RING_MAX 256

typedef struct
{
    hb_buffer_t     *master[BENCH_RING_MAX];   /* owns the payload (header never enters a filter) */
    void            *base[BENCH_RING_MAX];
    hb_shim_free_fn  orig_free[BENCH_RING_MAX];
    int              free_idx[BENCH_RING_MAX]; /* stack of free payloads */
    int              n, n_free;
    pthread_mutex_t  lock;
} bench_ring_t;

struct hb_bench_s
{
    int n;
    hb_filter_object_t **f;
    int *done;
    int failed;
    int pix_fmt, w, h, frame_flags;
    int volatile done_flag;
    int64_t next_index;                        /* frames fed so far (timestamps continue across hb_bench_stream calls) */
    bench_ring_t ring;
    int64_t ring_misses;
    hb_bench_stats_t *st;
};

/* hb_buffer_close() passes only the allocation's base pointer: one bench stream at a time per process owns the ring
 * (the benchmark processes run one stream at a time; chains share the ring of their stream) */
static bench_ring_t *g_ring = NULL;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void ring_release(void *base)
{
    bench_ring_t *r = g_ring;
    if (r == NULL) return;
    pthread_mutex_lock(&r->lock);
    for (int i = 0; i < r->n; i++)
        if (r->base[i] == base)
        {
            r->free_idx[r->n_free++] = i;
            break;
        }
    pthread_mutex_unlock(&r->lock);
}

static int ring_fill(hb_bench_t *b, const uint8_t *src, int n_unique, int want)
{
    bench_ring_t *r = &b->ring;
    const size_t fb = hb_harness_frame_bytes(b->pix_fmt, b->w, b->h);
    /* a whole number of passes over the source frames, so that payload i always holds source frame i % n_unique */
    if (want > BENCH_RING_MAX) want = BENCH_RING_MAX;
    want = want / n_unique * n_unique;
    if (want < n_unique) want = n_unique;
    while (r->n < want)
    {
        const int i = r->n;
        hb_buffer_t *m = hb_harness_frame_from_packed(b->pix_fmt, b->w, b->h, src + (size_t)(i % n_unique) * fb);
        if (m == NULL) return -1;
        r->master[i] = m;
        r->base[i] = hb_shim_buffer_set_release(m, ring_release, &r->orig_free[i]);
        r->free_idx[r->n_free++] = i;
        r->n++;
    }
    return 0;
}

static void ring_destroy(hb_bench_t *b)
{
    bench_ring_t *r = &b->ring;
    if (g_ring == r) g_ring = NULL;
    for (int i = 0; i < r->n; i++)
    {
        hb_shim_buffer_set_release(r->master[i], r->orig_free[i], NULL);
        hb_buffer_close(&r->master[i]);
    }
    r->n = r->n_free = 0;
}

/* next input frame: header over a free payload holding source frame (index % n_unique); when every payload is still
 * inside the chain (ring smaller than the chain's appetite) a fresh frame is built the slow way and counted */
static hb_buffer_t *next_input(hb_bench_t *b, const uint8_t *src, int n_unique)
{
    bench_ring_t *r = &b->ring;
    const int want = (int)(b->next_index % n_unique);
    hb_buffer_t *buf = NULL;
    pthread_mutex_lock(&r->lock);
    for (int k = r->n_free - 1; k >= 0; k--)
        if (r->free_idx[k] % n_unique == want)
        {
            const int i = r->free_idx[k];
            r->free_idx[k] = r->free_idx[--r->n_free];
            buf = hb_shim_frame_header_dup(r->master[i]);
            break;
        }
    pthread_mutex_unlock(&r->lock);
    if (buf == NULL)
    {
        const size_t fb = hb_harness_frame_bytes(b->pix_fmt, b->w, b->h);
        buf = hb_harness_frame_from_packed(b->pix_fmt, b->w, b->h, src + (size_t)want * fb);
        b->ring_misses++;
    }
    if (buf == NULL) return NULL;
    memset(&buf->s, 0, sizeof(buf->s));
    buf->s.type     = FRAME_BUF;
    buf->s.start    = b->next_index * 3003;
    buf->s.stop     = buf->s.start + 3003;
    buf->s.duration = 3003;
    buf->s.flags    = (uint16_t)b->frame_flags;
    b->next_index++;
    return buf;
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

static void chain_feed(hb_bench_t *c, int k, hb_buffer_t *list)
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

static void free_filter(hb_filter_object_t *f)
{
    if (f->settings) hb_dict_free(&f->settings);
    if (f->sub_filter)
    {
        if (f->sub_filter->settings) hb_dict_free(&f->sub_filter->settings);
        free(f->sub_filter);
    }
    free(f);
}

hb_bench_t *hb_bench_open_chain(int n_filters, hb_filter_object_t *const *protos, const char *const *settings,
                                int pix_fmt, int w, int h, int frame_flags)
{
    hb_bench_t *c = calloc(1, sizeof(*c));
    c->f = calloc(n_filters, sizeof(*c->f));
    c->done = calloc(n_filters, sizeof(int));
    c->pix_fmt = pix_fmt;
    c->w = w;
    c->h = h;
    c->frame_flags = frame_flags;
    pthread_mutex_init(&c->ring.lock, NULL);
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
    for (int k = 0; k < n_filters; k++)
    {
        hb_filter_object_t *f = malloc(sizeof(*f));
        memcpy(f, protos[k], sizeof(*f));
        f->settings = settings && settings[k] ? hb_parse_filter_settings(settings[k]) : NULL;
        f->done = &c->done_flag;
        if (f->sub_filter != NULL)
        {
            hb_filter_object_t *sub = malloc(sizeof(*sub));
            memcpy(sub, f->sub_filter, sizeof(*sub));
            sub->settings = settings && settings[k] ? hb_parse_filter_settings(settings[k]) : NULL;
            f->sub_filter = sub;
        }
        if (f->init(f, &init) != 0)
        {
            free_filter(f);
            for (int j = 0; j < c->n; j++)
            {
                c->f[j]->close(c->f[j]);
                free_filter(c->f[j]);
            }
            free(c->f);
            free(c->done);
            free(c);
            return NULL;
        }
        c->f[c->n++] = f;
    }
    return c;
}

hb_bench_t *hb_bench_open(hb_filter_object_t *proto, const char *settings, int pix_fmt, int w, int h)
{
    hb_filter_object_t *protos[1] = { proto };
    const char *sets[1] = { settings };
    return hb_bench_open_chain(1, protos, sets, pix_fmt, w, h, PIC_FLAG_PROGRESSIVE_FRAME);
}

/* fills the input ring without feeding anything (so that the caller can choose the ring's memory type around the call) */
int hb_bench_prefill(hb_bench_t *b, const uint8_t *src, int n_unique, int ring)
{
    if (b == NULL || b->failed) return -1;
    return ring_fill(b, src, n_unique, ring > 0 ? ring : 48);
}

/* feeds n_frames more frames (cycling over n_unique packed source frames; the stream's timestamps continue) and
 * consumes whatever comes out; `ring` payloads back the inputs (0 = default) */
int hb_bench_stream(hb_bench_t *b, const uint8_t *src, int n_unique, int n_frames, int ring, hb_bench_stats_t *st)
{
    memset(st, 0, sizeof(*st));
    if (b == NULL || b->failed) return -1;
    if (ring_fill(b, src, n_unique, ring > 0 ? ring : 48) != 0) return -1;
    g_ring = &b->ring;
    b->st = st;
    const size_t fb = hb_harness_frame_bytes(b->pix_fmt, b->w, b->h);
    const int64_t miss0 = b->ring_misses;
    const double t0 = now_s();
    for (int i = 0; i < n_frames && !b->failed; i++)
    {
        hb_buffer_t *buf = next_input(b, src, n_unique);
        if (buf == NULL) { b->failed = 1; break; }
        st->bytes_in += fb;
        chain_feed(b, 0, buf);
    }
    st->seconds = now_s() - t0;
    st->ring_misses = b->ring_misses - miss0;
    b->st = NULL;
    return b->failed ? -1 : 0;
}

/* EOF through the chain, outputs consumed, filters closed, handle freed */
int hb_bench_finish(hb_bench_t *b, hb_bench_stats_t *st)
{
    hb_bench_stats_t local;
    if (st == NULL) st = &local;
    memset(st, 0, sizeof(*st));
    if (b == NULL) return -1;
    b->st = st;
    g_ring = &b->ring;
    const double t0 = now_s();
    if (!b->failed) chain_feed(b, 0, hb_buffer_eof_init());
    st->seconds = now_s() - t0;
    const int rc = b->failed ? -1 : 0;
    for (int k = 0; k < b->n; k++)
    {
        b->f[k]->close(b->f[k]);
        free_filter(b->f[k]);
    }
    ring_destroy(b);
    pthread_mutex_destroy(&b->ring.lock);
    free(b->f);
    free(b->done);
    free(b);
    return rc;
}

/* ends the stream WITHOUT the EOF flush: filters are closed with whatever they still buffer (what a cancelled libhb job does).
 * For throughput samples of the CPU reference, whose flush of a partly filled taskset cycle runs serially and can take
 * minutes at 4K / 8K without being part of any measurement. */
int hb_bench_abort(hb_bench_t *b)
{
    if (b == NULL) return -1;
    g_ring = &b->ring;
    for (int k = 0; k < b->n; k++)
    {
        b->f[k]->close(b->f[k]);
        free_filter(b->f[k]);
    }
    ring_destroy(b);
    pthread_mutex_destroy(&b->ring.lock);
    free(b->f);
    free(b->done);
    free(b);
    return 0;
}

static void add_stats(hb_bench_stats_t *a, const hb_bench_stats_t *b)
{
    a->seconds += b->seconds;
    a->frames_out += b->frames_out;
    a->bytes_in += b->bytes_in;
    a->bytes_out += b->bytes_out;
    a->checksum += b->checksum;
    a->ring_misses += b->ring_misses;
}

/* one whole stream: n_frames, EOF, close */
int hb_bench_run(hb_bench_t *b, const uint8_t *src, int n_unique, int n_frames, hb_bench_stats_t *st)
{
    hb_bench_stats_t fin;
    int rc = hb_bench_stream(b, src, n_unique, n_frames, 0, st);
    if (hb_bench_finish(b, &fin) != 0) rc = -1;
    add_stats(st, &fin);
    return rc;
}

int hb_bench_run_chain(int n_filters, hb_filter_object_t *const *protos, const char *const *settings,
                       int pix_fmt, int w, int h, int frame_flags,
                       const uint8_t *src, int n_unique, int n_frames, hb_bench_stats_t *st)
{
    memset(st, 0, sizeof(*st));
    hb_bench_t *b = hb_bench_open_chain(n_filters, protos, settings, pix_fmt, w, h, frame_flags);
    if (b == NULL) return -2;
    return hb_bench_run(b, src, n_unique, n_frames, st);
}
