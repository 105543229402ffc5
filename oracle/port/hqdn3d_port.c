/* hqdn3d_port.c -- TEST INFRASTRUCTURE (see oracle_port.h).
 *
 * Restates libhb/denoise.c (hqdn3d, "high quality 3-D denoise") as three separate passes over a plane instead of the
 * reference's single interleaved sweep (denoise.c:126-165):
 *   H  every row on its own:      h(0) = LOAD(0)  [first row: lowpass(LOAD(0), LOAD(0))],  h(x) = lowpass(h(x-1), LOAD(x))
 *   V  every column on its own:   v(x,0) = h(x,0),  v(x,y) = lowpass(v(x,y-1), h(x,y))
 *   T  every sample on its own:   ant = lowpass(ant, v)  -> output ant >> (16 - depth); `ant` persists from frame to frame and
 *                                 starts as LOAD of the first frame (denoise.c:175-189)
 * lowpass(prev, cur) = cur + coef[(prev - cur) >> (8 - LUT_BITS)] in 16-bit fixed point (denoise.c:96-100), the table
 * from hqdn3d_precalc_coef (:78-94).  With no spatial strength (table[0] == 0) only T runs (:192-200).
 */
#include "oracle_port.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LUT_BITS_OF(depth) ((depth) == 16 ? 8 : 4)

void oracle_hqdn3d_coef(int16_t *ct, int depth, double dist25)
{
    const int lb = LUT_BITS_OF(depth);
    const double gamma = log(0.25) / log(1.0 - (dist25 > 252.0 ? 252.0 : dist25) / 255.0 - 0.00001);
    for (int i = -(256 << lb); i < 256 << lb; i++)
    {
        const double f = (i * (1 << (9 - lb)) + (1 << (8 - lb)) - 1) / 512.0;
        double simil = 1.0 - fabs(f) / 255.0;
        if (simil < 0) simil = 0;
        const double C = pow(simil, gamma) * 256.0 * f;
        ct[(256 << lb) + i] = lrint(C);
    }
    ct[0] = !!dist25;
}

static inline unsigned lowpass(int prev, int cur, const int16_t *coef, int lb)
{
    const int d = (prev - cur) >> (8 - lb);
    return cur + coef[d];
}

/* one plane, tightly packed samples in and out; `ant` (w*h uint16) is the filter's state, *ant_valid says whether it holds
 * a previous frame */
void oracle_hqdn3d_plane(const void *src_, void *dst_, uint16_t *ant, int *ant_valid, int w, int h, int depth,
                         const int16_t *spatial_tab, const int16_t *temporal_tab)
{
    const int lb = LUT_BITS_OF(depth), sh = 16 - depth, bias = ((1 << sh) - 1) >> 1;
    const uint8_t *s8 = src_;
    const uint16_t *s16 = src_;
    uint8_t *d8 = dst_;
    uint16_t *d16 = dst_;
    const int16_t *spatial = spatial_tab + (256 << lb), *temporal = temporal_tab + (256 << lb);
#define LOADP(i) ((int)(((depth == 8 ? s8[i] : s16[i]) << sh) + bias))
    if (!*ant_valid)
    {
        for (int i = 0; i < w * h; i++) ant[i] = (uint16_t)LOADP(i);
        *ant_valid = 1;
    }
    uint16_t *v = malloc((size_t)w * h * sizeof(uint16_t));
    if (spatial_tab[0])
    {
        uint16_t *hh = malloc((size_t)w * h * sizeof(uint16_t));
        for (int y = 0; y < h; y++)
        {
            unsigned p = LOADP(y * w);
            if (y == 0) p = lowpass(p, LOADP(0), spatial, lb);
            hh[y * w] = (uint16_t)p;            /* line_ant / tmp are uint16_t / uint32_t in the reference: values stay below 2^16 */
            for (int x = 1; x < w; x++)
            {
                p = lowpass(p, LOADP(y * w + x), spatial, lb);
                hh[y * w + x] = (uint16_t)p;
            }
        }
        for (int x = 0; x < w; x++)
        {
            unsigned p = hh[x];
            v[x] = (uint16_t)p;
            for (int y = 1; y < h; y++)
            {
                p = lowpass((uint16_t)p, hh[y * w + x], spatial, lb);
                v[y * w + x] = (uint16_t)p;
            }
        }
        free(hh);
    }
    else
        for (int i = 0; i < w * h; i++) v[i] = (uint16_t)LOADP(i);
    for (int i = 0; i < w * h; i++)
    {
        const unsigned t = lowpass(ant[i], spatial_tab[0] ? v[i] : LOADP(i), temporal, lb);
        ant[i] = (uint16_t)t;
        if (depth == 8) d8[i] = (uint8_t)(t >> sh);
        else            d16[i] = (uint16_t)(t >> sh);
    }
#undef LOADP
    free(v);
}

/* n packed yuv420p frames; strengths[6] = y-spatial, y-temporal, cb-spatial, cb-temporal, cr-spatial, cr-temporal after the
 * defaults of denoise.c:237-266 */
void oracle_hqdn3d_clip(const uint8_t *in, int n, int width, int height, int depth, const double strengths[6], uint8_t *out)
{
    const int bps = depth > 8 ? 2 : 1, lb = LUT_BITS_OF(depth);
    const int cw = -((-width) >> 1), chh = -((-height) >> 1);
    const int pw[3] = { width, cw, cw }, ph[3] = { height, chh, chh };
    size_t off[3], fb = 0;
    for (int c = 0; c < 3; c++) { off[c] = fb; fb += (size_t)pw[c] * ph[c] * bps; }
    int16_t *tab[6];
    for (int i = 0; i < 6; i++)
    {
        tab[i] = malloc((512 << lb) * sizeof(int16_t));
        oracle_hqdn3d_coef(tab[i], depth, strengths[i]);
    }
    uint16_t *ant[3];
    int valid[3] = { 0, 0, 0 };
    for (int c = 0; c < 3; c++) ant[c] = malloc((size_t)pw[c] * ph[c] * sizeof(uint16_t));
    for (int t = 0; t < n; t++)
        for (int c = 0; c < 3; c++)
            oracle_hqdn3d_plane(in + (size_t)t * fb + off[c], out + (size_t)t * fb + off[c], ant[c], &valid[c], pw[c], ph[c], depth,
                                tab[2 * c], tab[2 * c + 1]);
    for (int c = 0; c < 3; c++) free(ant[c]);
    for (int i = 0; i < 6; i++) free(tab[i]);
}
