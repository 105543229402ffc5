/* hostlogic_filters.c -- TEST INFRASTRUCTURE (see hostlogic_nlmeans.c for the idea).
 *
 * CPU stand-ins for the hbcu_comb_detect_*, hbcu_decomb_*, hbcu_lapsharp_*, hbcu_unsharp_* and hbcu_hqdn3d_* calls of the product's host filters
 * (handbrake_b200/libhb/comb_detect_cuda.c, decomb_cuda.c, lapsharp_cuda.c), built on the plain-C restatement.  What the
 * host side owns and what is therefore pinned against the compiled reference through these: comb-detect's three-frame
 * window (first frame duplicated, HB_FILTER_DELAY, exhaustive check at both ends, the same hb_buffer_t passed through with
 * its verdict); decomb's per-frame mode selection from the combed tag, parity / field-order logic, bob timestamps and frame
 * rate, the EEDI2 call per field; lapsharp's settings cascade and sanitising.  Never linked into the product.
 */
#include "../../include/hbcu.h"
#include "oracle_port.h"

#include <stdlib.h>
#include <string.h>

void oracle_hostlogic_set_error(const char *fmt, ...);
const void *const *oracle_hostlogic_frame_planes(const hbcu_frame_t *f);      /* hostlogic_frames.c */
const int *oracle_hostlogic_frame_strides(const hbcu_frame_t *f);

/* ------------------------------------------------------------------------------------------ a ring of packed frames */
typedef struct
{
    int slots, bps, w[3], h[3], planes;
    size_t off[3], bytes;
    uint8_t *mem;
    int64_t *index;
} ring_t;

static void ring_init(ring_t *r, int slots, int planes, int width, int height, int depth, int sw, int sh)
{
    memset(r, 0, sizeof(*r));
    r->slots = slots > 0 ? slots : 8;
    r->planes = planes;
    r->bps = depth > 8 ? 2 : 1;
    for (int c = 0; c < planes; c++)
    {
        r->w[c] = c ? -((-width) >> sw) : width;
        r->h[c] = c ? -((-height) >> sh) : height;
        r->off[c] = r->bytes;
        r->bytes += (size_t)r->w[c] * r->h[c] * r->bps;
    }
    r->mem = calloc(r->slots, r->bytes);
    r->index = malloc(sizeof(int64_t) * r->slots);
    for (int i = 0; i < r->slots; i++) r->index[i] = -1;
}
static void ring_free(ring_t *r) { free(r->mem); free(r->index); }
static void ring_put(ring_t *r, int64_t index, const void *const planes[3], const int strides[3])
{
    const int slot = (int)(index % r->slots);
    for (int c = 0; c < r->planes; c++)
        for (int y = 0; y < r->h[c]; y++)
            memcpy(r->mem + (size_t)slot * r->bytes + r->off[c] + (size_t)y * r->w[c] * r->bps,
                   (const uint8_t *)planes[c] + (size_t)y * strides[c], (size_t)r->w[c] * r->bps);
    r->index[slot] = index;
}
static const uint8_t *ring_get(const ring_t *r, int64_t index)
{
    const int slot = (int)(index % r->slots);
    if (index < 0 || r->index[slot] != index)
    {
        oracle_hostlogic_set_error("frame %lld is no longer in the ring", (long long)index);
        return NULL;
    }
    return r->mem + (size_t)slot * r->bytes;
}
static void unpack(const ring_t *r, const uint8_t *packed, void *const planes[3], const int strides[3])
{
    for (int c = 0; c < r->planes; c++)
        for (int y = 0; y < r->h[c]; y++)
            memcpy((uint8_t *)planes[c] + (size_t)y * strides[c], packed + r->off[c] + (size_t)y * r->w[c] * r->bps, (size_t)r->w[c] * r->bps);
}

/* ------------------------------------------------------------------------------------------ comb detect */
struct hbcu_comb_detect_s
{
    hbcu_comb_detect_config_t cfg;
    oracle_comb_params_t params;
    ring_t ring;
    int64_t verdict_of[64];     /* several runs may be queued before their results are collected */
    int verdict[64];
};

int oracle_hbcu_comb_detect_create(hbcu_comb_detect_t **out, const hbcu_comb_detect_config_t *cfg)
{
    struct hbcu_comb_detect_s *h = calloc(1, sizeof(*h));
    h->cfg = *cfg;
    const int shift = cfg->depth - 8;
    /* the restatement takes the thresholds as the settings give them and derives the rest itself; the host's own derived
     * values must agree with that derivation */
    h->params = (oracle_comb_params_t){ cfg->mode, cfg->spatial_metric, cfg->motion_threshold >> shift, cfg->spatial_threshold >> shift,
                                        cfg->filter_mode, cfg->block_threshold, cfg->block_width, cfg->block_height };
    const int n = 1 << cfg->depth;
    float *lut = malloc(sizeof(float) * n);
    oracle_comb_gamma_lut(cfg->depth, lut);
    const int lut_ok = cfg->gamma_lut != NULL && memcmp(lut, cfg->gamma_lut, sizeof(float) * n) == 0;
    free(lut);
    if (!lut_ok || (h->params.motion_threshold << shift) != cfg->motion_threshold || (h->params.spatial_threshold << shift) != cfg->spatial_threshold)
    {
        oracle_hostlogic_set_error("comb_detect: the host's gamma table / shifted thresholds differ from comb_detect.c:1074-1081, 1152-1153");
        free(h);
        return -1;
    }
    ring_init(&h->ring, cfg->slots, 1, cfg->width, cfg->height, cfg->depth, 0, 0);
    for (int i = 0; i < 64; i++) h->verdict_of[i] = -1;
    *out = h;
    return 0;
}
void oracle_hbcu_comb_detect_destroy(hbcu_comb_detect_t *h) { if (h) { ring_free(&h->ring); free(h); } }
int oracle_hbcu_comb_detect_upload(hbcu_comb_detect_t *h, int64_t index, const void *luma, int stride)
{
    const void *const planes[3] = { luma, NULL, NULL };
    const int strides[3] = { stride, 0, 0 };
    ring_put(&h->ring, index, planes, strides);
    return 0;
}
int oracle_hbcu_comb_detect_upload_frame(hbcu_comb_detect_t *h, int64_t index, hbcu_frame_t *in)
{
    return oracle_hbcu_comb_detect_upload(h, index, oracle_hostlogic_frame_planes(in)[0], oracle_hostlogic_frame_strides(in)[0]);
}
int oracle_hbcu_comb_detect_run(hbcu_comb_detect_t *h, int64_t prev, int64_t cur, int64_t next, int force_exhaustive)
{
    const uint8_t *p = ring_get(&h->ring, prev), *c = ring_get(&h->ring, cur), *n = ring_get(&h->ring, next);
    if (p == NULL || c == NULL || n == NULL) return -1;
    h->verdict[cur & 63] = oracle_comb_detect(p, c, n, h->cfg.width, h->cfg.height, h->cfg.depth, &h->params, force_exhaustive, NULL, NULL);
    h->verdict_of[cur & 63] = cur;
    return 0;
}
int oracle_hbcu_comb_detect_result(hbcu_comb_detect_t *h, int64_t cur, int *combed)
{
    if (h->verdict_of[cur & 63] != cur) { oracle_hostlogic_set_error("comb_detect: no verdict for frame %lld", (long long)cur); return -1; }
    *combed = h->verdict[cur & 63];
    return 0;
}

/* ------------------------------------------------------------------------------------------ decomb */
struct hbcu_decomb_s
{
    hbcu_decomb_config_t cfg;
    ring_t ring;
    void *eedi;
    uint8_t *eedi_frame, *out;
};

int oracle_hbcu_decomb_create(hbcu_decomb_t **out, const hbcu_decomb_config_t *cfg)
{
    struct hbcu_decomb_s *h = calloc(1, sizeof(*h));
    h->cfg = *cfg;
    ring_init(&h->ring, cfg->slots, 3, cfg->width, cfg->height, cfg->depth, cfg->chroma_shift_w, cfg->chroma_shift_h);
    if (cfg->mode & HBCU_DECOMB_EEDI2)
    {
        h->eedi = oracle_eedi2_create(cfg->width, cfg->height, cfg->depth, cfg->magnitude_threshold, cfg->variance_threshold,
                                      cfg->laplacian_threshold, cfg->dilation_threshold, cfg->erosion_threshold, cfg->noise_threshold,
                                      cfg->maximum_search_distance, cfg->post_processing);
        h->eedi_frame = malloc(h->ring.bytes);
    }
    h->out = malloc(h->ring.bytes);
    *out = h;
    return 0;
}
void oracle_hbcu_decomb_destroy(hbcu_decomb_t *h)
{
    if (h == NULL) return;
    if (h->eedi) oracle_eedi2_destroy(h->eedi);
    free(h->eedi_frame); free(h->out); ring_free(&h->ring); free(h);
}
int oracle_hbcu_decomb_upload(hbcu_decomb_t *h, int64_t index, const void *const planes[3], const int strides[3]) { ring_put(&h->ring, index, planes, strides); return 0; }
int oracle_hbcu_decomb_upload_frame(hbcu_decomb_t *h, int64_t index, hbcu_frame_t *in)
{
    return oracle_hbcu_decomb_upload(h, index, oracle_hostlogic_frame_planes(in), oracle_hostlogic_frame_strides(in));
}
int oracle_hbcu_decomb_wait_upload(hbcu_decomb_t *h, int64_t index) { (void)h; (void)index; return 0; }
int oracle_hbcu_decomb_filter(hbcu_decomb_t *h, int64_t ticket, int64_t prev, int64_t cur, int64_t next,
                              int frame_mode, int parity, int tff, void *const planes[3], const int strides[3])
{
    (void)ticket;
    const uint8_t *p = ring_get(&h->ring, prev), *c = ring_get(&h->ring, cur), *n = ring_get(&h->ring, next);
    if (p == NULL || c == NULL || n == NULL) return -1;
    memset(h->out, 0, h->ring.bytes);                           /* libhb-shim output buffers start zeroed */
    if (frame_mode == 0) memcpy(h->out, c, h->ring.bytes);      /* "just passing through": hb_buffer_copy (decomb template :893-896) */
    else if (frame_mode & HBCU_DECOMB_EEDI2)
    {
        if (h->eedi == NULL) { oracle_hostlogic_set_error("decomb: EEDI2 asked of a handle created without it"); return -1; }
        oracle_eedi2_field(h->eedi, c, !parity, h->eedi_frame);  /* pv->tff = !parity (decomb.c:539-542) */
        oracle_decomb_field_eedi2(p, c, n, h->eedi_frame, h->out, h->cfg.width, h->cfg.height, h->cfg.depth, frame_mode, parity, tff);
    }
    else oracle_decomb_field_eedi2(p, c, n, NULL, h->out, h->cfg.width, h->cfg.height, h->cfg.depth, frame_mode, parity, tff);
    unpack(&h->ring, h->out, planes, strides);
    return 0;
}
int oracle_hbcu_decomb_filter_frame(hbcu_decomb_t *h, int64_t ticket, int64_t prev, int64_t cur, int64_t next, int frame_mode, int parity, int tff, hbcu_frame_t *out)
{
    return oracle_hbcu_decomb_filter(h, ticket, prev, cur, next, frame_mode, parity, tff, (void *const *)oracle_hostlogic_frame_planes(out),
                                     oracle_hostlogic_frame_strides(out));
}
int oracle_hbcu_decomb_wait(hbcu_decomb_t *h, int64_t ticket) { (void)h; (void)ticket; return 0; }
int oracle_hbcu_decomb_poll(hbcu_decomb_t *h, int64_t ticket) { (void)h; (void)ticket; return 1; }

/* ------------------------------------------------------------------------------------------ lapsharp */
struct hbcu_lapsharp_s { hbcu_lapsharp_config_t cfg; };

int oracle_hbcu_lapsharp_create(hbcu_lapsharp_t **out, const hbcu_lapsharp_config_t *cfg)
{
    struct hbcu_lapsharp_s *h = calloc(1, sizeof(*h));
    h->cfg = *cfg;
    *out = h;
    return 0;
}
void oracle_hbcu_lapsharp_destroy(hbcu_lapsharp_t *h) { free(h); }
int oracle_hbcu_lapsharp_filter_frames(hbcu_lapsharp_t *h, int64_t ticket, hbcu_frame_t *in_frame, const void *const in_planes[3], const int in_strides[3],
                                       hbcu_frame_t *out_frame, void *const out_planes[3], const int out_strides[3])
{
    (void)ticket;
    if (in_frame != NULL)
    {
        /* the host mirrors the stride region of a host buffer before the call (lapsharp.c:333); for a device frame the
         * device side does it: hb_frame_buffer_mirror_stride's word arithmetic (fifo.c:906-959), restated */
        in_planes = oracle_hostlogic_frame_planes(in_frame); in_strides = oracle_hostlogic_frame_strides(in_frame);
        for (int c = 0; c < 3; c++)
        {
            const int w = c ? -((-h->cfg.width) >> h->cfg.chroma_shift_w) : h->cfg.width;
            const int ht = c ? -((-h->cfg.height) >> h->cfg.chroma_shift_h) : h->cfg.height;
            uint16_t *d = (uint16_t *)in_planes[c];
            const int stride = in_strides[c] / 2, margin = stride - w, front = margin / 2, back = margin - front;
            for (int y = 0; y < ht; y++)
            {
                int pos = y * stride + w;
                for (int i = 0; i < back; i++) d[pos + i] = d[pos - i - 1];
                pos = (y + 1) * stride - 1;
                for (int i = 0; i < front; i++) d[pos - i] = d[pos + i + 1];
            }
        }
    }
    if (out_frame != NULL) { out_planes = (void *const *)oracle_hostlogic_frame_planes(out_frame); out_strides = oracle_hostlogic_frame_strides(out_frame); }
    for (int c = 0; c < 3; c++)
    {
        const int w = c ? -((-h->cfg.width) >> h->cfg.chroma_shift_w) : h->cfg.width;
        const int ht = c ? -((-h->cfg.height) >> h->cfg.chroma_shift_h) : h->cfg.height;
        oracle_lapsharp_plane(in_planes[c], out_planes[c], w, ht, in_strides[c], out_strides[c], h->cfg.depth, h->cfg.kernel[c], h->cfg.strength[c]);
    }
    return 0;
}
int oracle_hbcu_lapsharp_wait(hbcu_lapsharp_t *h, int64_t ticket) { (void)h; (void)ticket; return 0; }
int oracle_hbcu_lapsharp_poll(hbcu_lapsharp_t *h, int64_t ticket) { (void)h; (void)ticket; return 1; }

/* ------------------------------------------------------------------------------------------ unsharp / chroma smooth, hqdn3d */
static void plane_dims(int width, int height, int sw, int sh, int c, int *w, int *h)
{
    *w = c ? -((-width) >> sw) : width;
    *h = c ? -((-height) >> sh) : height;
}
static uint8_t *tight_copy(const void *plane, int stride, int w, int h, int bps)
{
    uint8_t *t = malloc((size_t)w * h * bps);
    for (int y = 0; y < h; y++) memcpy(t + (size_t)y * w * bps, (const uint8_t *)plane + (size_t)y * stride, (size_t)w * bps);
    return t;
}
static void strided_copy(void *plane, int stride, const uint8_t *t, int w, int h, int bps)
{
    for (int y = 0; y < h; y++) memcpy((uint8_t *)plane + (size_t)y * stride, t + (size_t)y * w * bps, (size_t)w * bps);
}

struct hbcu_unsharp_s { hbcu_unsharp_config_t cfg; };

int oracle_hbcu_unsharp_create(hbcu_unsharp_t **out, const hbcu_unsharp_config_t *cfg)
{
    struct hbcu_unsharp_s *h = calloc(1, sizeof(*h));
    h->cfg = *cfg;
    *out = h;
    return 0;
}
void oracle_hbcu_unsharp_destroy(hbcu_unsharp_t *h) { free(h); }
int oracle_hbcu_unsharp_filter_frames(hbcu_unsharp_t *h, int64_t ticket, hbcu_frame_t *in_frame, const void *const in_planes[3], const int in_strides[3],
                                      hbcu_frame_t *out_frame, void *const out_planes[3], const int out_strides[3])
{
    (void)ticket;
    if (in_frame != NULL) { in_planes = oracle_hostlogic_frame_planes(in_frame); in_strides = oracle_hostlogic_frame_strides(in_frame); }
    if (out_frame != NULL) { out_planes = (void *const *)oracle_hostlogic_frame_planes(out_frame); out_strides = oracle_hostlogic_frame_strides(out_frame); }
    const int bps = h->cfg.depth > 8 ? 2 : 1;
    for (int c = 0; c < 3; c++)
    {
        int w, ht;
        plane_dims(h->cfg.width, h->cfg.height, h->cfg.chroma_shift_w, h->cfg.chroma_shift_h, c, &w, &ht);
        uint8_t *src = tight_copy(in_planes[c], in_strides[c], w, ht, bps), *dst = malloc((size_t)w * ht * bps);
        /* the restatement takes (strength, size); amount / 65536 and 2 steps + 1 map back onto the host's (amount, steps)
         * exactly, and its own sanitising is idempotent on sanitised values */
        oracle_unsharp_plane(src, dst, w, ht, h->cfg.depth, h->cfg.amount[c] / 65536.0, 2 * h->cfg.steps[c] + 1, h->cfg.smooth, 1);
        strided_copy(out_planes[c], out_strides[c], dst, w, ht, bps);
        free(src); free(dst);
    }
    return 0;
}
int oracle_hbcu_unsharp_wait(hbcu_unsharp_t *h, int64_t ticket) { (void)h; (void)ticket; return 0; }
int oracle_hbcu_unsharp_poll(hbcu_unsharp_t *h, int64_t ticket) { (void)h; (void)ticket; return 1; }

struct hbcu_hqdn3d_s
{
    hbcu_hqdn3d_config_t cfg;
    int16_t *coef[6];
    uint16_t *ant[3];
    int ant_valid[3];
};

int oracle_hbcu_hqdn3d_create(hbcu_hqdn3d_t **out, const hbcu_hqdn3d_config_t *cfg)
{
    struct hbcu_hqdn3d_s *h = calloc(1, sizeof(*h));
    h->cfg = *cfg;
    const size_t entries = (size_t)512 << (cfg->depth == 16 ? 8 : 4);
    for (int i = 0; i < 6; i++)
    {
        h->coef[i] = malloc(entries * sizeof(int16_t));            /* the host frees its tables after create() */
        memcpy(h->coef[i], cfg->coef[i], entries * sizeof(int16_t));
    }
    for (int c = 0; c < 3; c++)
    {
        int w, ht;
        plane_dims(cfg->width, cfg->height, cfg->chroma_shift_w, cfg->chroma_shift_h, c, &w, &ht);
        h->ant[c] = calloc((size_t)w * ht, sizeof(uint16_t));
    }
    *out = h;
    return 0;
}
void oracle_hbcu_hqdn3d_destroy(hbcu_hqdn3d_t *h)
{
    if (h == NULL) return;
    for (int i = 0; i < 6; i++) free(h->coef[i]);
    for (int c = 0; c < 3; c++) free(h->ant[c]);
    free(h);
}
int oracle_hbcu_hqdn3d_filter_frames(hbcu_hqdn3d_t *h, int64_t ticket, hbcu_frame_t *in_frame, const void *const in_planes[3], const int in_strides[3],
                                     hbcu_frame_t *out_frame, void *const out_planes[3], const int out_strides[3])
{
    (void)ticket;
    if (in_frame != NULL) { in_planes = oracle_hostlogic_frame_planes(in_frame); in_strides = oracle_hostlogic_frame_strides(in_frame); }
    if (out_frame != NULL) { out_planes = (void *const *)oracle_hostlogic_frame_planes(out_frame); out_strides = oracle_hostlogic_frame_strides(out_frame); }
    const int bps = h->cfg.depth > 8 ? 2 : 1;
    for (int c = 0; c < 3; c++)
    {
        int w, ht;
        plane_dims(h->cfg.width, h->cfg.height, h->cfg.chroma_shift_w, h->cfg.chroma_shift_h, c, &w, &ht);
        uint8_t *src = tight_copy(in_planes[c], in_strides[c], w, ht, bps), *dst = malloc((size_t)w * ht * bps);
        oracle_hqdn3d_plane(src, dst, h->ant[c], &h->ant_valid[c], w, ht, h->cfg.depth, h->coef[2 * c], h->coef[2 * c + 1]);
        strided_copy(out_planes[c], out_strides[c], dst, w, ht, bps);
        free(src); free(dst);
    }
    return 0;
}
int oracle_hbcu_hqdn3d_wait(hbcu_hqdn3d_t *h, int64_t ticket) { (void)h; (void)ticket; return 0; }
int oracle_hbcu_hqdn3d_poll(hbcu_hqdn3d_t *h, int64_t ticket) { (void)h; (void)ticket; return 1; }
