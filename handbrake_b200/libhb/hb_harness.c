/* hb_harness.c -- drives hb_filter_object_t instances the way libhb does.
 *
 * Restates the filter part of do_job()/filter_loop() (libhb/work.c:1840-1870,
 * 2527-2600): init each filter with a running hb_filter_init_t, then for every
 * input buffer call work(), close *buf_in if the filter left it set, forward
 * the ->next-linked output list to the next filter, stop a filter after it
 * returned HB_FILTER_DONE, finally close().  Works for any object exposing the
 * libhb filter interface, so the same harness runs the reference objects
 * (oracle/_ref/libhbref.so) and the CUDA objects (libhbcu_filters.so).
 *
 * Frames cross this API as tightly packed planar arrays (plane after plane,
 * row pitch = width*bps) so that Python/numpy callers need no struct mirror.
 */
#include "handbrake/handbrake.h"
#include "hb_harness.h"

typedef struct
{
    int                  n;
    hb_filter_object_t **f;
    int                 *done;
    hb_harness_io_t     *io;
    int                  failed;
    size_t               frame_bytes_out;
    int                  out_pix_fmt, out_w, out_h;
} chain_t;

size_t hb_harness_frame_bytes(int pix_fmt, int w, int h)
{
    const AVPixFmtDescriptor *d = av_pix_fmt_desc_get(pix_fmt);
    if (d == NULL) return 0;
    size_t total = 0;
    int nplanes = d->nb_components;
    for (int p = 0; p < nplanes; p++)
        total += (size_t)av_image_get_linesize(pix_fmt, w, p) * hb_image_height(pix_fmt, h, p);
    return total;
}

hb_buffer_t *hb_harness_frame_from_packed(int pix_fmt, int w, int h, const uint8_t *src)
{
    hb_buffer_t *b = hb_frame_buffer_init(pix_fmt, w, h);
    if (b == NULL) return NULL;
    for (int p = 0; p <= b->f.max_plane; p++)
    {
        const int line = av_image_get_linesize(pix_fmt, w, p);
        for (int y = 0; y < b->plane[p].height; y++)
        {
            memcpy(b->plane[p].data + (size_t)y * b->plane[p].stride, src, line);
            src += line;
        }
    }
    return b;
}

void hb_harness_frame_to_packed(const hb_buffer_t *b, uint8_t *dst)
{
    for (int p = 0; p <= b->f.max_plane; p++)
    {
        const int line = av_image_get_linesize(b->f.fmt, b->f.width, p);
        for (int y = 0; y < b->plane[p].height; y++)
        {
            memcpy(dst, b->plane[p].data + (size_t)y * b->plane[p].stride, line);
            dst += line;
        }
    }
}

static void sink(chain_t *c, hb_buffer_t *list)
{
    hb_harness_io_t *io = c->io;
    while (list != NULL)
    {
        hb_buffer_t *b = list;
        list = b->next;
        b->next = NULL;
        if (b->s.flags & HB_BUF_FLAG_EOF)
        {
            io->saw_eof = 1;
        }
        else if (b->storage_type == HBCU_DEVICE)
        {
            /* a device frame reached the end of the chain: the chain lacks its download adapter */
            hb_error("harness: HBCU_DEVICE buffer at the sink (no hb_filter_hbcu_download at the end of the chain)");
            c->failed = 1;
            io->n_dropped++;
        }
        else if (io->n_out < io->out_capacity)
        {
            const int i = io->n_out;
            if (io->out != NULL)
                hb_harness_frame_to_packed(b, io->out + (size_t)i * c->frame_bytes_out);
            if (io->out_combed) io->out_combed[i] = b->s.combed;
            if (io->out_flags)  io->out_flags[i]  = b->s.flags;
            if (io->out_start)  io->out_start[i]  = b->s.start;
            if (io->out_stop)   io->out_stop[i]   = b->s.stop;
            if (io->out_duration) io->out_duration[i] = b->s.duration;
            io->n_out++;
        }
        else
        {
            io->n_dropped++;
        }
        hb_buffer_close(&b);
    }
}

/* feed one buffer (or list) into filter k; mirrors one filter_loop iteration per buffer */
static void feed(chain_t *c, int k, hb_buffer_t *list)
{
    if (k >= c->n)
    {
        sink(c, list);
        return;
    }
    while (list != NULL)
    {
        hb_buffer_t *in = list;
        list = in->next;
        in->next = NULL;

        if (c->done[k])
        {
            hb_buffer_close(&in);   /* loop has exited; nothing consumes further input */
            continue;
        }
        hb_buffer_t *out = NULL;
        int status = c->f[k]->work(c->f[k], &in, &out);
        c->f[k]->status = status;
        if (in != NULL)
            hb_buffer_close(&in);                       /* work.c:2566 */
        if (status == HB_FILTER_FAILED)
            c->failed = 1;
        if (out != NULL)
            feed(c, k + 1, out);                        /* work.c:2574-2585 */
        if (status == HB_FILTER_DONE)
            c->done[k] = 1;                             /* work.c:2532 */
    }
}

int hb_harness_run_chain(int n_filters, hb_filter_object_t *const *protos,
                         const char *const *settings, hb_harness_io_t *io)
{
    chain_t c;
    memset(&c, 0, sizeof(c));
    c.io = io;
    c.f = calloc(n_filters, sizeof(*c.f));
    c.done = calloc(n_filters, sizeof(int));
    io->n_out = 0;
    io->n_dropped = 0;
    io->saw_eof = 0;

    hb_filter_init_t init;
    memset(&init, 0, sizeof(init));
    init.pix_fmt         = io->pix_fmt;
    init.geometry.width  = io->width;
    init.geometry.height = io->height;
    init.geometry.par.num = 1;
    init.geometry.par.den = 1;
    init.vrate.num = 30000;
    init.vrate.den = 1001;
    init.time_base.num = 1;
    init.time_base.den = 90000;
    init.color_prim = init.color_transfer = init.color_matrix = 1;
    init.color_range = 1;
    init.chroma_location = 1;

    int volatile done_flag = 0;
    int rc = 0;
    /* work.c:1857-1870: a filter whose init fails is dropped, the job goes on */
    for (int k = 0; k < n_filters; k++)
    {
        hb_filter_object_t *f = malloc(sizeof(*f));
        memcpy(f, protos[k], sizeof(*f));
        f->settings = settings && settings[k] ? hb_parse_filter_settings(settings[k]) : NULL;
        f->done = &done_flag;
        if (f->sub_filter != NULL)
        {
            /* wrapper filters (mt_frame): the wrapped filter gets its own copy of the settings (hb.c:1697-1700) */
            hb_filter_object_t *sub = malloc(sizeof(*sub));
            memcpy(sub, f->sub_filter, sizeof(*sub));
            sub->settings = settings && settings[k] ? hb_parse_filter_settings(settings[k]) : NULL;
            f->sub_filter = sub;
        }
        if (f->init(f, &init) != 0)
        {
            io->init_failed |= 1 << k;
            if (f->settings) hb_dict_free(&f->settings);
            if (f->sub_filter) { if (f->sub_filter->settings) hb_dict_free(&f->sub_filter->settings); free(f->sub_filter); }
            free(f);
            continue;
        }
        c.f[c.n++] = f;
    }
    io->vrate_num_out = init.vrate.num;
    io->vrate_den_out = init.vrate.den;
    c.out_pix_fmt = init.pix_fmt;
    c.out_w = init.geometry.width;
    c.out_h = init.geometry.height;
    c.frame_bytes_out = hb_harness_frame_bytes(c.out_pix_fmt, c.out_w, c.out_h);

    const size_t frame_bytes_in = hb_harness_frame_bytes(io->pix_fmt, io->width, io->height);
    for (int i = 0; i < io->n_in && !c.failed; i++)
    {
        hb_buffer_t *b = hb_harness_frame_from_packed(io->pix_fmt, io->width, io->height,
                                                      io->in + (size_t)i * frame_bytes_in);
        b->s.start    = (int64_t)i * 3003;
        b->s.stop     = b->s.start + 3003;
        b->s.duration = 3003;
        b->s.flags    = io->in_flags  ? io->in_flags[i]  : PIC_FLAG_PROGRESSIVE_FRAME;
        b->s.combed   = io->in_combed ? io->in_combed[i] : HB_COMB_NONE;
        b->s.new_chap = i;   /* lets tests check that props travel with the right frame */
        b->f.color_prim = init.color_prim;
        feed(&c, 0, b);
    }
    if (!c.failed)
        feed(&c, 0, hb_buffer_eof_init());
    else
        rc = -1;

    for (int k = 0; k < c.n; k++)
    {
        c.f[k]->close(c.f[k]);
        if (c.f[k]->settings) hb_dict_free(&c.f[k]->settings);
        if (c.f[k]->sub_filter) { if (c.f[k]->sub_filter->settings) hb_dict_free(&c.f[k]->sub_filter->settings); free(c.f[k]->sub_filter); }
        free(c.f[k]);
    }
    free(c.f);
    free(c.done);
    return rc;
}

int hb_harness_run(hb_filter_object_t *proto, const char *settings, hb_harness_io_t *io)
{
    hb_filter_object_t *protos[1] = { proto };
    const char *sets[1] = { settings };
    return hb_harness_run_chain(1, protos, sets, io);
}
