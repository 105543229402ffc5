/* detelecine_port.c -- TEST INFRASTRUCTURE (see oracle_port.h).
 *
 * Plain-C restatement of the data-parallel half of pullup (libhb/detelecine.c): the three block metrics
 * (pullup_diff_y / pullup_licomb_y / pullup_var_y, :159-207, walked by pullup_compute_metric, :230-265), the
 * max-reductions inside pullup_compute_breaks (:369-374) and pullup_compute_affinity (:405-418) and pullup_copy_field
 * (:298-317) -- behind the SAME call interface as the device implementation (include/hbcu.h, hbcu_detelecine_*), with
 * the prefix oracle_hbcu_detelecine_.
 *
 * Purpose: oracle/Makefile builds _ref/libhostlogic.so from (among others) the product's host-side state machine
 * (handbrake_b200/libhb/detelecine_cuda.c, compiled with the hbcu_detelecine_* names redirected here) so that the state
 * machine + this restatement can be pinned, on a machine without a GPU, against the compiled reference filter.  The
 * product library never links this file; on the GPU the same host code drives the CUDA implementation and the GPU tests
 * compare that against the compiled reference directly.
 */
#include "../../include/hbcu.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct hbcu_detelecine_s
{
    hbcu_detelecine_config_t cfg;
    int bps, w[3], h[3], pitch[3];          /* pitch in bytes */
    size_t off[3], picture_bytes;
    int mw, mh, mlen;
    size_t moff;
    uint8_t *pictures;
    int *metrics, *results;
};

void oracle_hostlogic_set_error(const char *fmt, ...);

static inline int iabs(int a) { return a < 0 ? -a : a; }
static inline int sample(const struct hbcu_detelecine_s *h, const uint8_t *p, ptrdiff_t i)
{
    return h->bps == 2 ? ((const uint16_t *)p)[i] : p[i];
}
static uint8_t *plane_of(const struct hbcu_detelecine_s *h, int picture, int p)
{
    return h->pictures + (size_t)picture * h->picture_bytes + h->off[p];
}
static int *metric_of(const struct hbcu_detelecine_s *h, int field, int which)
{
    return h->metrics + ((size_t)field * 3 + which) * h->mlen;
}

int oracle_hbcu_detelecine_create(hbcu_detelecine_t **out, const hbcu_detelecine_config_t *cfg)
{
    struct hbcu_detelecine_s *h = calloc(1, sizeof(*h));
    h->cfg = *cfg;
    h->bps = cfg->depth > 8 ? 2 : 1;
    for (int p = 0; p < 3; p++)
    {
        h->w[p] = p ? -((-cfg->width) >> cfg->chroma_shift_w) : cfg->width;
        h->h[p] = p ? -((-cfg->height) >> cfg->chroma_shift_h) : cfg->height;
        h->pitch[p] = (h->w[p] * h->bps + 63) / 64 * 64;
        h->off[p] = h->picture_bytes;
        h->picture_bytes += (size_t)h->pitch[p] * h->h[p];
    }
    const int mp = cfg->metric_plane;
    h->mw = (h->w[mp] - ((cfg->junk_left + cfg->junk_right) << 3)) >> 3;
    h->mh = (h->h[mp] - ((cfg->junk_top + cfg->junk_bottom) << 1)) >> 3;
    if (h->mw < 1 || h->mh < 1)
    {
        oracle_hostlogic_set_error("no metric blocks");
        free(h);
        return -1;
    }
    h->mlen = h->mw * h->mh;
    h->moff = (size_t)cfg->junk_left * 8 * h->bps + (size_t)(cfg->junk_top << 1) * h->pitch[mp];
    h->pictures = calloc(cfg->pictures, h->picture_bytes);
    h->metrics = calloc((size_t)cfg->fields * 3 * h->mlen, sizeof(int));
    h->results = calloc((size_t)cfg->results * 2, sizeof(int));
    *out = h;
    return 0;
}

void oracle_hbcu_detelecine_destroy(hbcu_detelecine_t *h)
{
    if (h == NULL) return;
    free(h->pictures); free(h->metrics); free(h->results); free(h);
}

int oracle_hbcu_detelecine_upload(hbcu_detelecine_t *h, int picture, const void *const planes[3], const int strides[3])
{
    for (int p = 0; p < 3; p++)
    {
        const int row = strides[p] < h->pitch[p] ? strides[p] : h->pitch[p];
        for (int y = 0; y < h->h[p]; y++)
            memcpy(plane_of(h, picture, p) + (size_t)y * h->pitch[p], (const uint8_t *)planes[p] + (size_t)y * strides[p], row);
    }
    return 0;
}

/* a, b: first sample of the block in the respective field; fs = field line stride in samples */
static int block_diff(const struct hbcu_detelecine_s *h, const uint8_t *a, const uint8_t *b, ptrdiff_t fs)
{
    int d = 0;
    for (int line = 0; line < 4; line++)
        for (int j = 0; j < 8; j++) d += iabs(sample(h, a, line * fs + j) - sample(h, b, line * fs + j));
    return d;
}
static int block_comb(const struct hbcu_detelecine_s *h, const uint8_t *a, const uint8_t *b, ptrdiff_t fs)
{
    int c = 0;
    for (int line = 0; line < 4; line++)
        for (int j = 0; j < 8; j++)
        {
            const ptrdiff_t o = line * fs + j;
            c += iabs(2 * sample(h, a, o) - sample(h, b, o - fs) - sample(h, b, o))
               + iabs(2 * sample(h, b, o) - sample(h, a, o) - sample(h, a, o + fs));
        }
    return c;
}
static int block_var(const struct hbcu_detelecine_s *h, const uint8_t *a, ptrdiff_t fs)
{
    int v = 0;
    for (int line = 0; line < 3; line++)
        for (int j = 0; j < 8; j++) v += iabs(sample(h, a, line * fs + j) - sample(h, a, (line + 1) * fs + j));
    return 4 * v;
}

int oracle_hbcu_detelecine_metrics(hbcu_detelecine_t *h, int field, int picture, int parity, int diff_picture, int comb_top_picture, int comb_bottom_picture)
{
    const int mp = h->cfg.metric_plane, pitch = h->pitch[mp];
    const ptrdiff_t fs = 2 * (pitch / h->bps);
    int *diffs = metric_of(h, field, 0), *comb = metric_of(h, field, 1), *var = metric_of(h, field, 2);
    for (int by = 0; by < h->mh; by++)
        for (int bx = 0; bx < h->mw; bx++)
        {
            const size_t blk = h->moff + (size_t)by * 8 * pitch + (size_t)bx * 8 * h->bps;
            const int o = by * h->mw + bx;
            const uint8_t *cur = plane_of(h, picture, mp) + (size_t)parity * pitch + blk;
            var[o] = block_var(h, cur, fs);
            if (diff_picture == picture) diffs[o] = 0;
            else if (diff_picture >= 0) diffs[o] = block_diff(h, cur, plane_of(h, diff_picture, mp) + (size_t)parity * pitch + blk, fs);
            if (comb_top_picture >= 0)
                comb[o] = block_comb(h, plane_of(h, comb_top_picture, mp) + blk, plane_of(h, comb_bottom_picture, mp) + pitch + blk, fs);
        }
    return 0;
}

int oracle_hbcu_detelecine_breaks(hbcu_detelecine_t *h, int field2, int field3, int slot)
{
    const int *d2 = metric_of(h, field2, 0), *d3 = metric_of(h, field3, 0);
    int max_l = 0, max_r = 0;
    for (int i = 0; i < h->mlen; i++)
    {
        const int l = d2[i] - d3[i];
        if (l > max_l) max_l = l;
        if (-l > max_r) max_r = -l;
    }
    h->results[2 * slot] = max_l;
    h->results[2 * slot + 1] = max_r;
    return 0;
}

int oracle_hbcu_detelecine_affinity(hbcu_detelecine_t *h, int field_prev, int field, int field_next, int slot)
{
    const int *vp = metric_of(h, field_prev, 2), *vc = metric_of(h, field, 2), *vn = metric_of(h, field_next, 2);
    const int *cc = metric_of(h, field, 1), *cn = metric_of(h, field_next, 1);
    int max_l = 0, max_r = 0;
    for (int i = 0; i < h->mlen; i++)
    {
        int lc = cc[i] - (vc[i] + vp[i]) + iabs(vc[i] - vp[i]);
        int rc = cn[i] - (vc[i] + vn[i]) + iabs(vc[i] - vn[i]);
        if (lc < 0) lc = 0;
        if (rc < 0) rc = 0;
        const int l = lc - rc;
        if (l > max_l) max_l = l;
        if (-l > max_r) max_r = -l;
    }
    h->results[2 * slot] = max_l;
    h->results[2 * slot + 1] = max_r;
    return 0;
}

int oracle_hbcu_detelecine_fetch(hbcu_detelecine_t *h, int *dst, int nslots)
{
    if (nslots > 0) memcpy(dst, h->results, sizeof(int) * 2 * nslots);
    return 0;
}

int oracle_hbcu_detelecine_copy_field(hbcu_detelecine_t *h, int dst_picture, int src_picture, int parity)
{
    if (dst_picture == src_picture) return 0;
    for (int p = 0; p < 3; p++)
        for (int line = 0; line < h->h[p] >> 1; line++)
        {
            const size_t o = (size_t)(2 * line + parity) * h->pitch[p];
            memcpy(plane_of(h, dst_picture, p) + o, plane_of(h, src_picture, p) + o, h->pitch[p]);
        }
    return 0;
}

int oracle_hbcu_detelecine_download(hbcu_detelecine_t *h, int picture, void *const planes[3], const int strides[3])
{
    for (int p = 0; p < 3; p++)
    {
        const int row = strides[p] < h->pitch[p] ? strides[p] : h->pitch[p];
        for (int y = 0; y < h->h[p]; y++)
            memcpy((uint8_t *)planes[p] + (size_t)y * strides[p], plane_of(h, picture, p) + (size_t)y * h->pitch[p], row);
    }
    return 0;
}

/* the deferred copy: _begin parks the picture (as the product's staging buffer does), _end delivers it -- the latest the
 * product's asynchronous copy can land, so a caller that touched the host planes early would be caught */
static uint8_t *g_staging;
static size_t g_staging_bytes;
static struct { hbcu_detelecine_t *h; void *planes[3]; int strides[3]; int armed; } g_pending_copy;
int oracle_hbcu_detelecine_download_begin(hbcu_detelecine_t *h, int picture, void *const planes[3], const int strides[3])
{
    size_t total = 0, off[3];
    for (int p = 0; p < 3; p++) { off[p] = total; total += (size_t)h->pitch[p] * h->h[p]; }
    if (g_staging_bytes < total) { free(g_staging); g_staging = malloc(total); g_staging_bytes = total; }
    for (int p = 0; p < 3; p++) memcpy(g_staging + off[p], plane_of(h, picture, p), (size_t)h->pitch[p] * h->h[p]);
    g_pending_copy.h = h;
    for (int p = 0; p < 3; p++) { g_pending_copy.planes[p] = planes[p]; g_pending_copy.strides[p] = strides[p]; }
    g_pending_copy.armed = 1;
    return 0;
}
int oracle_hbcu_detelecine_download_end(hbcu_detelecine_t *h)
{
    if (!g_pending_copy.armed || g_pending_copy.h != h) return 0;
    size_t off = 0;
    for (int p = 0; p < 3; p++)
    {
        const int row = g_pending_copy.strides[p] < h->pitch[p] ? g_pending_copy.strides[p] : h->pitch[p];
        for (int y = 0; y < h->h[p]; y++)
            memcpy((uint8_t *)g_pending_copy.planes[p] + (size_t)y * g_pending_copy.strides[p], g_staging + off + (size_t)y * h->pitch[p], row);
        off += (size_t)h->pitch[p] * h->h[p];
    }
    g_pending_copy.armed = 0;
    return 0;
}

/* device frames: see hostlogic_frames.c */
const void *const *oracle_hostlogic_frame_planes(const hbcu_frame_t *f);
const int *oracle_hostlogic_frame_strides(const hbcu_frame_t *f);
int oracle_hbcu_detelecine_upload_frame(hbcu_detelecine_t *h, int picture, hbcu_frame_t *in)
{
    return oracle_hbcu_detelecine_upload(h, picture, oracle_hostlogic_frame_planes(in), oracle_hostlogic_frame_strides(in));
}
int oracle_hbcu_detelecine_download_frame(hbcu_detelecine_t *h, int picture, hbcu_frame_t *out)
{
    return oracle_hbcu_detelecine_download(h, picture, (void *const *)oracle_hostlogic_frame_planes(out), oracle_hostlogic_frame_strides(out));
}

int oracle_hbcu_detelecine_mark(hbcu_detelecine_t *h, int which) { (void)h; (void)which; return 0; }
int oracle_hbcu_detelecine_elapsed_ms(hbcu_detelecine_t *h, float *ms) { (void)h; *ms = 0.f; return 0; }
