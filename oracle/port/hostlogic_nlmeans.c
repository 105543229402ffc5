/* hostlogic_nlmeans.c -- TEST INFRASTRUCTURE (see oracle_port.h).
 *
 * The hbcu_nlmeans_* calls the product's host filter (handbrake_b200/libhb/nlmeans_cuda.c) makes, implemented on the CPU
 * with the plain-C restatement of NLMeans (nlmeans_port.c).  oracle/Makefile links this with the UNTOUCHED host filter
 * (its hbcu_* references renamed to oracle_hbcu_*) into _ref/libhostlogic.so, so that everything the host side owns --
 * settings cascade and sanitising, strength scaling and the weight tables, look-ahead buffering, the shrinking window at
 * EOF, output order and properties -- is pinned against the compiled reference filter on a
 * machine without a GPU.  The weight table used here is the one the host filter hands over, not a recomputed one.
 * Never linked into the product.
 */
#include "../../include/hbcu.h"
#include "oracle_port.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct hbcu_nlmeans_s
{
    hbcu_nlmeans_config_t cfg;
    int bps, w[3], h[3];
    size_t off[3], frame_bytes;
    int ring;
    uint8_t *frames;            /* ring of tightly packed frames */
    int64_t *index;             /* which frame a ring slot holds */
    int mid_stream;             /* hbcu_nlmeans_set_stream_slice: index 0 is not the stream's first frame */
};

void oracle_hostlogic_set_error(const char *fmt, ...);

int oracle_hbcu_nlmeans_create(hbcu_nlmeans_t **out, const hbcu_nlmeans_config_t *cfg)
{
    struct hbcu_nlmeans_s *h = calloc(1, sizeof(*h));
    h->cfg = *cfg;
    h->bps = cfg->depth > 8 ? 2 : 1;
    for (int c = 0; c < 3; c++)
    {
        h->w[c] = c ? -((-cfg->width) >> cfg->chroma_shift_w) : cfg->width;
        h->h[c] = c ? -((-cfg->height) >> cfg->chroma_shift_h) : cfg->height;
        h->off[c] = h->frame_bytes;
        h->frame_bytes += (size_t)h->w[c] * h->h[c] * h->bps;
    }
    h->ring = cfg->ring_frames > 0 ? cfg->ring_frames : 8;
    h->frames = calloc(h->ring, h->frame_bytes);
    h->index = malloc(sizeof(int64_t) * h->ring);
    for (int i = 0; i < h->ring; i++) h->index[i] = -1;
    *out = h;
    return 0;
}

void oracle_hbcu_nlmeans_destroy(hbcu_nlmeans_t *h)
{
    if (h == NULL) return;
    free(h->frames); free(h->index); free(h);
}

int oracle_hbcu_nlmeans_upload(hbcu_nlmeans_t *h, int64_t index, const void *const planes[3], const int strides[3])
{
    const int slot = (int)(index % h->ring);
    uint8_t *dst = h->frames + (size_t)slot * h->frame_bytes;
    for (int c = 0; c < 3; c++)
        for (int y = 0; y < h->h[c]; y++)
            memcpy(dst + h->off[c] + (size_t)y * h->w[c] * h->bps, (const uint8_t *)planes[c] + (size_t)y * strides[c], (size_t)h->w[c] * h->bps);
    h->index[slot] = index;
    return 0;
}

/* multi-device dealing: the halo frame is taken from the peer handle's ring (the peer copy of the product) */
int oracle_hbcu_nlmeans_upload_peer(hbcu_nlmeans_t *dst, int64_t dst_index, hbcu_nlmeans_t *src, int64_t src_index)
{
    const int sslot = (int)(src_index % src->ring), dslot = (int)(dst_index % dst->ring);
    if (dst == src || dst->frame_bytes != src->frame_bytes || src->index[sslot] != src_index)
    {
        oracle_hostlogic_set_error("upload_peer: frame %lld is not in the source ring", (long long)src_index);
        return -1;
    }
    memcpy(dst->frames + (size_t)dslot * dst->frame_bytes, src->frames + (size_t)sslot * src->frame_bytes, src->frame_bytes);
    dst->index[dslot] = dst_index;
    return 0;
}

int oracle_hbcu_nlmeans_set_stream_slice(hbcu_nlmeans_t *h, int mid_stream) { h->mid_stream = mid_stream != 0; return 0; }
int oracle_hbcu_nlmeans_sync(hbcu_nlmeans_t *h) { (void)h; return 0; }

int oracle_hbcu_nlmeans_wait_upload(hbcu_nlmeans_t *h, int64_t index) { (void)h; (void)index; return 0; }

int oracle_hbcu_nlmeans_filter(hbcu_nlmeans_t *h, int64_t index, int navail, void *const planes[3], const int strides[3])
{
    for (int c = 0; c < 3; c++)
    {
        const hbcu_nlmeans_plane_t *pp = &h->cfg.plane[c];
        const int nf = pp->nframes < navail ? pp->nframes : navail;
        const void *frames[32];
        for (int f = 0; f < nf; f++)
        {
            const int slot = (int)((index + f) % h->ring);
            if (h->index[slot] != index + f)
            {   /* the host filter let a frame it still needs be overwritten: a ring-sizing bug */
                oracle_hostlogic_set_error("frame %lld is no longer in the ring", (long long)(index + f));
                return -1;
            }
            frames[f] = h->frames + (size_t)slot * h->frame_bytes + h->off[c];
        }
        uint8_t *tight = malloc((size_t)h->w[c] * h->h[c] * h->bps);
        /* the reference's stale source-patch pointer (see nlmeans_port.c): frame 0 of the stream, or no temporal window */
        oracle_nlmeans_plane_with_table(frames, nf, h->w[c], h->h[c], h->cfg.depth, pp->patch_size, pp->range, pp->origin_tune,
                                        pp->bypass, pp->prefilter, pp->weight_fact, pp->diff_max, pp->exptable,
                                        (index == 0 && !h->mid_stream) || pp->nframes < 2, tight);
        for (int y = 0; y < h->h[c]; y++)
            memcpy((uint8_t *)planes[c] + (size_t)y * strides[c], tight + (size_t)y * h->w[c] * h->bps, (size_t)h->w[c] * h->bps);
        free(tight);
    }
    return 0;
}

int oracle_hbcu_nlmeans_wait(hbcu_nlmeans_t *h, int64_t index) { (void)h; (void)index; return 0; }
int oracle_hbcu_nlmeans_poll(hbcu_nlmeans_t *h, int64_t index) { (void)h; (void)index; return 1; }

/* device frames: see hostlogic_frames.c */
const void *const *oracle_hostlogic_frame_planes(const hbcu_frame_t *f);
const int *oracle_hostlogic_frame_strides(const hbcu_frame_t *f);

int oracle_hbcu_nlmeans_upload_frame(hbcu_nlmeans_t *h, int64_t index, hbcu_frame_t *in)
{
    return oracle_hbcu_nlmeans_upload(h, index, oracle_hostlogic_frame_planes(in), oracle_hostlogic_frame_strides(in));
}
int oracle_hbcu_nlmeans_filter_frame(hbcu_nlmeans_t *h, int64_t index, int navail, hbcu_frame_t *out)
{
    return oracle_hbcu_nlmeans_filter(h, index, navail, (void *const *)oracle_hostlogic_frame_planes(out), oracle_hostlogic_frame_strides(out));
}
