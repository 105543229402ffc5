/* decomb_port.c -- TEST INFRASTRUCTURE (see oracle_port.h).
 *
 * Restates HandBrake's decomb line filters and frame logic without EEDI2:
 *   cubic_interpolate_pixel/line   templates/decomb_template.c:43-107
 *   blend_filter_pixel/line        templates/decomb_template.c:279-361
 *   yadif_filter_line, YADIF_CHECK templates/decomb_template.c:482-710
 *   yadif_decomb_filter_work       templates/decomb_template.c:714-808
 *   filter_{8,16}                  templates/decomb_template.c:810-898
 *   process_frame, hb_decomb_work  decomb.c:500-612
 * Formulated per output pixel (the reference walks row segments per CPU; rows are independent).
 */
#include "oracle_port.h"

#include <stdlib.h>
#include <string.h>

#define ABSI(a) ((a) > 0 ? (a) : -(a))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))
#define MAX2(a, b) ((a) > (b) ? (a) : (b))
#define MIN3(a, b, c) MIN2(MIN2(a, b), c)
#define MAX3(a, b, c) MAX2(MAX2(a, b), c)

typedef struct
{
    const uint8_t *base;
    int w, h, depth;
} plane_t;

static inline int px(const plane_t *p, int x, int y)
{
    const size_t i = (size_t)y * p->w + x;
    return p->depth > 8 ? ((const uint16_t *)p->base)[i] : p->base[i];
}

/* crop_table[v + 1024] (template :23-41): clamp to [0, max] for v in [-1024, max + 1023] */
static inline int crop(int v, int maxv)
{
    return v < 0 ? 0 : v > maxv ? maxv : v;
}

/* template :43-49; C division truncates toward zero */
static inline int cubic_px(int maxv, int y0, int y1, int y2, int y3)
{
    const int r = (y0 * -3) + (y1 * 23) + (y2 * 23) + (y3 * -3);
    return crop(r / 40, maxv);
}

static int cubic_line_px(const plane_t *c, int x, int y, int maxv)
{
    int a = 0, b = 0, cc = 0, d = 0;
    const int h = c->h;
    if (y >= 3)                 { a = px(c, x, y - 3); b = px(c, x, y - 1); }
    else if (y == 2 || y == 1)  { a = px(c, x, y - 1); b = a; }
    else if (y == 0)            { a = px(c, x, y + 1); b = a; }
    if (y <= h - 4)                       { cc = px(c, x, y + 1); d = px(c, x, y + 3); }
    else if (y == h - 3 || y == h - 2)    { cc = px(c, x, y + 1); d = cc; }
    else if (y == h - 1)                  { cc = px(c, x, y - 1); d = cc; }
    return cubic_px(maxv, a, b, cc, d);
}

static int blend_line_px(const plane_t *c, int x, int y, int maxv)
{
    const int h = c->h;
    int u1, u2, d1, d2;
    if (y > 1 && y < h - 2)  { u1 = -1; u2 = -2; d1 = 1; d2 = 2; }
    else if (y == 0)         { u1 = u2 = 0; d1 = 1; d2 = 2; }
    else if (y == 1)         { u1 = u2 = -1; d1 = 1; d2 = 2; }
    else if (y == h - 2)     { u1 = -1; u2 = -2; d1 = d2 = 1; }
    else                     { u1 = -1; u2 = -2; d1 = d2 = 0; }
    int r = -px(c, x, y + u2) + 2 * px(c, x, y + u1) + 6 * px(c, x, y) + 2 * px(c, x, y + d1) - px(c, x, y + d2);
    r >>= 3;
    return crop(r, maxv);
}

static int yadif_px(const plane_t *prev, const plane_t *cur, const plane_t *next, const plane_t *eedi,
                    int x, int y, int par /* parity ^ tff */, int mode, int maxv)
{
    const int w = cur->w, h = cur->h;
    const plane_t *prev2 = par ? prev : cur, *next2 = par ? cur : next;
    const int yp = y ? y - 1 : y + 1;                 /* mirrored at the first / last line (:606-611) */
    const int yn = y + 1 < h ? y + 1 : y - 1;
    const int vertical_edge = (y < 3) || (y > h - 4);
    const int cubic = (mode & ORACLE_DECOMB_CUBIC) != 0;
    const int margin = cubic ? 3 : 2;

    const int c = px(cur, x, yp);
    const int d = (px(prev2, x, y) + px(next2, x, y)) >> 1;
    const int e = px(cur, x, yn);
    const int td0 = ABSI(px(prev2, x, y) - px(next2, x, y));
    const int td1 = (ABSI(px(prev, x, yp) - c) + ABSI(px(prev, x, yn) - e)) >> 1;
    const int td2 = (ABSI(px(next, x, yp) - c) + ABSI(px(next, x, yn) - e)) >> 1;
    int diff = MAX3(td0 >> 1, td1, td2);
    int spatial_pred;

    if (eedi != NULL)
    {
        spatial_pred = px(eedi, x, y);
    }
    else
    {
        if (cubic && !vertical_edge)
            spatial_pred = cubic_px(maxv, px(cur, x, y - 3), px(cur, x, y - 1), px(cur, x, y + 1), px(cur, x, y + 3));
        else
            spatial_pred = (c + e) >> 1;

        if (x > margin && x < w - (margin + 1))
        {
            int score = ABSI(px(cur, x - 1, yp) - px(cur, x - 1, yn)) + ABSI(c - e) +
                        ABSI(px(cur, x + 1, yp) - px(cur, x + 1, yn)) - 1;
            for (int dir = -1; dir <= 1; dir += 2)       /* YADIF_CHECK(-1){(-2)} then YADIF_CHECK(1){(2)} */
            {
                for (int step = 1; step <= 2; step++)
                {
                    const int j = dir * step;
                    const int s = ABSI(px(cur, x - 1 + j, yp) - px(cur, x - 1 - j, yn)) +
                                  ABSI(px(cur, x + j, yp) - px(cur, x - j, yn)) +
                                  ABSI(px(cur, x + 1 + j, yp) - px(cur, x + 1 - j, yn));
                    if (!(s < score)) break;             /* the +-2 check is nested inside a successful +-1 check */
                    score = s;
                    if (cubic && !vertical_edge)
                    {
                        if (step == 1)
                            spatial_pred = cubic_px(maxv, px(cur, x + 3 * j, y - 3), px(cur, x + j, y - 1),
                                                    px(cur, x - j, y + 1), px(cur, x - 3 * j, y + 3));
                        else   /* j = +-2: outer taps average two rows at x +- 4 */
                            spatial_pred = cubic_px(maxv,
                                                    (px(cur, x + 2 * j, y - 3) + px(cur, x + 2 * j, y - 1)) / 2,
                                                    px(cur, x + j, y - 1), px(cur, x - j, y + 1),
                                                    (px(cur, x - 2 * j, y + 3) + px(cur, x - 2 * j, y + 1)) / 2);
                    }
                    else
                    {
                        spatial_pred = (px(cur, x + j, yp) + px(cur, x - j, yn)) >> 1;
                    }
                }
            }
        }
    }

    if (!vertical_edge)
    {
        const int b = (px(prev2, x, y - 2) + px(next2, x, y - 2)) >> 1;
        const int f = (px(prev2, x, y + 2) + px(next2, x, y + 2)) >> 1;
        const int mx = MAX3(d - e, d - c, MIN2(b - c, f - e));
        const int mn = MIN3(d - e, d - c, MAX2(b - c, f - e));
        diff = MAX3(diff, mn, -mx);
    }
    if (spatial_pred > d + diff) spatial_pred = d + diff;
    else if (spatial_pred < d - diff) spatial_pred = d - diff;
    return spatial_pred;
}

static void put(uint8_t *base, int w, int depth, int x, int y, int v)
{
    const size_t i = (size_t)y * w + x;
    if (depth > 8) ((uint16_t *)base)[i] = (uint16_t)v;
    else base[i] = (uint8_t)v;
}

void *oracle_eedi2_create(int width, int height, int depth, int mthresh, int vthresh, int lthresh, int dstr, int estr,
                          int nt, int maxd, int pp);
void oracle_eedi2_destroy(void *e);
void oracle_eedi2_field(void *e, const uint8_t *cur, int tff, uint8_t *out);

/* `eedi`: packed EEDI2 interpolation of this field (or NULL).  With it, mode carries the EEDI2 bit:
 * yadif takes its spatial prediction from it (:648-653); without yadif the picture IS the interpolation (:855-875) */
static void decomb_field_ex(const uint8_t *prev, const uint8_t *cur, const uint8_t *next, const uint8_t *eedi, uint8_t *dst,
                            int width, int height, int depth, int mode, int parity, int tff)
{
    const int bps = depth > 8 ? 2 : 1, maxv = (1 << depth) - 1;
    const int cw = -((-width) >> 1), ch = -((-height) >> 1);
    const int pw[3] = { width, cw, cw }, ph[3] = { height, ch, ch };
    size_t off = 0;
    for (int pp = 0; pp < 3; pp++)
    {
        plane_t P = { prev + off, pw[pp], ph[pp], depth }, C = { cur + off, pw[pp], ph[pp], depth },
                N = { next + off, pw[pp], ph[pp], depth }, E = { eedi ? eedi + off : NULL, pw[pp], ph[pp], depth };
        uint8_t *D = dst + off;
        if (eedi != NULL && !(mode & ORACLE_DECOMB_YADIF))
        {
            memcpy(D, E.base, (size_t)pw[pp] * ph[pp] * bps);
            off += (size_t)pw[pp] * ph[pp] * bps;
            continue;
        }
        for (int y = 0; y < ph[pp]; y++)
        {
            /* parity 1 filters the even rows, parity 0 the odd rows (template :744, :796) */
            const int filtered = parity ? !(y & 1) : (y & 1);
            for (int x = 0; x < pw[pp]; x++)
            {
                if (!filtered)
                    put(D, pw[pp], depth, x, y, px(&C, x, y));
                else if (mode == ORACLE_DECOMB_BLEND)
                    put(D, pw[pp], depth, x, y, blend_line_px(&C, x, y, maxv));
                else if (mode == ORACLE_DECOMB_CUBIC)
                    put(D, pw[pp], depth, x, y, cubic_line_px(&C, x, y, maxv));
                else if (mode & ORACLE_DECOMB_YADIF)
                    put(D, pw[pp], depth, x, y, yadif_px(&P, &C, &N, eedi ? &E : NULL, x, y, parity ^ tff, mode, maxv));
                /* else: no line filter runs; the reference leaves the row unwritten */
            }
        }
        off += (size_t)pw[pp] * ph[pp] * bps;
    }
}

void oracle_decomb_field(const uint8_t *prev, const uint8_t *cur, const uint8_t *next, uint8_t *dst,
                         int width, int height, int depth, int filter_mode_unused, int mode, int parity, int tff)
{
    (void)filter_mode_unused;
    decomb_field_ex(prev, cur, next, NULL, dst, width, height, depth, mode, parity, tff);
}

/* one output picture with the EEDI2 picture of the same field given (mode may carry the EEDI2 bit) */
void oracle_decomb_field_eedi2(const uint8_t *prev, const uint8_t *cur, const uint8_t *next, const uint8_t *eedi, uint8_t *dst,
                               int width, int height, int depth, int mode, int parity, int tff)
{
    decomb_field_ex(prev, cur, next, eedi, dst, width, height, depth, mode, parity, tff);
}

int oracle_decomb_clip(const uint8_t *in, int n_in, const uint16_t *flags, const uint8_t *combed,
                       int width, int height, int depth, int mode, int parity_setting,
                       uint8_t *out, int *out_src)
{
    return oracle_decomb_clip_pp(in, n_in, flags, combed, width, height, depth, mode, parity_setting, 1, out, out_src);
}

int oracle_decomb_clip_pp(const uint8_t *in, int n_in, const uint16_t *flags, const uint8_t *combed,
                          int width, int height, int depth, int mode, int parity_setting, int postproc,
                          uint8_t *out, int *out_src)
{
    const int bps = depth > 8 ? 2 : 1;
    const int cw = -((-width) >> 1), ch = -((-height) >> 1);
    const size_t fb = ((size_t)width * height + 2 * (size_t)cw * ch) * bps;
    int n_out = 0;
    /* EEDI2 with the filter's default thresholds (decomb.c:234-243); its edge-mask state lives for the whole clip */
    void *eedi_state = (mode & ORACLE_DECOMB_EEDI2) ? oracle_eedi2_create(width, height, depth, 10, 20, 20, 4, 2, 50, 24, postproc) : NULL;
    uint8_t *eedi_frame = eedi_state ? malloc(fb) : NULL;
    for (int t = 0; t < n_in; t++)
    {
        const uint8_t *cur = in + (size_t)t * fb;
        const uint8_t *prev = t > 0 ? cur - fb : cur;          /* first frame is its own predecessor */
        const uint8_t *next = t + 1 < n_in ? cur + fb : cur;   /* last frame its own successor */
        const int is_combed_tag = combed ? combed[t] : 0;
        if ((mode & ORACLE_DECOMB_SELECTIVE) && is_combed_tag == 0)
        {
            memcpy(out + (size_t)n_out * fb, cur, fb);          /* decomb.c:502-510 */
            if (out_src) out_src[n_out] = t;
            n_out++;
            continue;
        }
        int tff;
        if (parity_setting < 0)
        {
            const int fl = flags ? flags[t] : 0x10;
            tff = ((fl & 0x10) == 0) ? !!(fl & 0x08) : 1;       /* decomb.c:515-519 */
        }
        else
            tff = (parity_setting & 1) ^ 1;
        const int nfields = (mode & ORACLE_DECOMB_BOB) ? 2 : 1;
        for (int frame = 0; frame < nfields; frame++)
        {
            const int parity = frame ^ tff ^ 1;
            int is_combed = 2, fmode = 0;
            if (mode & ORACLE_DECOMB_SELECTIVE) is_combed = is_combed_tag;
            if ((mode & ORACLE_DECOMB_BLEND) && is_combed == 1) fmode = ORACLE_DECOMB_BLEND;
            else if (is_combed != 0) fmode = mode & ~ORACLE_DECOMB_SELECTIVE;
            uint8_t *dst = out + (size_t)n_out * fb;
            memset(dst, 0, fb);                                  /* shim buffers start zeroed */
            if (fmode == 0) memcpy(dst, cur, fb);
            else if (fmode & ORACLE_DECOMB_EEDI2)
            {
                oracle_eedi2_field(eedi_state, cur, !parity, eedi_frame);        /* pv->tff = !parity (decomb.c:539-542) */
                decomb_field_ex(prev, cur, next, eedi_frame, dst, width, height, depth, fmode, parity, tff);
            }
            else decomb_field_ex(prev, cur, next, NULL, dst, width, height, depth, fmode, parity, tff);
            if (out_src) out_src[n_out] = t;
            n_out++;
        }
    }
    if (eedi_state) oracle_eedi2_destroy(eedi_state);
    free(eedi_frame);
    return n_out;
}
