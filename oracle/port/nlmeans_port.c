/* nlmeans_port.c -- TEST INFRASTRUCTURE (see oracle_port.h).
 *
 * Restates HandBrake's NLMeans (libhb/nlmeans.c + templates/nlmeans_template.c)
 * as a direct per-pixel computation:
 *   - the mirror border of nlmeans_border (template :20-43) becomes an index map;
 *   - the integral image of build_integral_scalar (template :545-591) plus the
 *     four-corner lookup (template :682) is the exact n x n sum of squared
 *     differences, computed here by summing the n*n terms directly;
 *   - weights / accumulation / output follow template :644-713 literally,
 *     because their floating-point order IS the contract.
 */
#include "oracle_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EXPSIZE 128

void oracle_nlmeans_table(double strength, int patch_size, int depth,
                          float *weight_fact, int *diff_max, float exptable[128])
{
    /* nlmeans.c:343 */
    strength *= depth > 8 ? (depth - 8) * (depth - 8) : 1;
    /* nlmeans.c:349-358; the float/double mix below is deliberate and identical */
    const float weight_factor       = 1.0 / patch_size / patch_size / (strength * strength);
    const float min_weight_in_table = 0.0005;
    const float stretch             = EXPSIZE / (-log(min_weight_in_table));
    *weight_fact = weight_factor * stretch;
    *diff_max    = EXPSIZE / *weight_fact;
    for (int i = 0; i < EXPSIZE; i++)
        exptable[i] = exp(-i / stretch);
    exptable[EXPSIZE - 1] = 0;
}

/* template :20-43: img[-1-x] = img[x]; img[w+x] = img[w-1-x] (rows likewise) */
static inline int mirror(int v, int n)
{
    if (v < 0)  return -1 - v;
    if (v >= n) return 2 * n - 1 - v;
    return v;
}

#define DEFINE_PLANE(NAME, PIXEL)                                                                   \
static void NAME(const void *const *frames, int nframes, int w, int h, int depth,                  \
                 const oracle_nlmeans_plane_params_t *pp, void *dst_v)                             \
{                                                                                                   \
    PIXEL *dst = dst_v;                                                                             \
    const PIXEL *src = frames[0];                                                                   \
    if (pp->strength == 0)                                                                          \
    {   /* nlmeans.c:493-499 */                                                                     \
        memcpy(dst, src, (size_t)w * h * sizeof(PIXEL));                                            \
        return;                                                                                     \
    }                                                                                               \
    float wfact, exptable[EXPSIZE];                                                                 \
    int diff_max;                                                                                   \
    oracle_nlmeans_table(pp->strength, pp->patch_size, depth, &wfact, &diff_max, exptable);         \
    const int nh = (pp->patch_size - 1) / 2, rh = (pp->range - 1) / 2;                              \
    const double origin_tune = pp->origin_tune;                                                     \
    for (int y = 0; y < h; y++)                                                                     \
    {                                                                                               \
        for (int x = 0; x < w; x++)                                                                 \
        {                                                                                           \
            float weight_sum = 0, pixel_sum = 0;                                                    \
            for (int f = 0; f < nframes; f++)                                                       \
            {                                                                                       \
                const PIXEL *cmp = frames[f];                                                       \
                for (int dy = -rh; dy <= rh; dy++)                                                  \
                {                                                                                   \
                    for (int dx = -rh; dx <= rh; dx++)                                              \
                    {                                                                               \
                        if (dx == 0 && dy == 0 && f == 0)                                           \
                        {   /* template :644-655 */                                                 \
                            weight_sum += origin_tune;                                              \
                            pixel_sum  += origin_tune * src[y * w + x];                             \
                            continue;                                                               \
                        }                                                                           \
                        uint32_t ssd = 0;                                                           \
                        for (int j = -nh; j <= nh; j++)                                             \
                        {                                                                           \
                            /* the compare patch is read from the BORDERED image, i.e. the     */   \
                            /* mirror applies to the displaced coordinate as a whole           */   \
                            const int ya = mirror(y + j, h), yb = mirror(y + j + dy, h);            \
                            for (int k = -nh; k <= nh; k++)                                         \
                            {                                                                       \
                                const int xa = mirror(x + k, w), xb = mirror(x + k + dx, w);        \
                                const int d = (int)src[ya * w + xa] - (int)cmp[yb * w + xb];        \
                                ssd += (uint32_t)(d * d);                                           \
                            }                                                                       \
                        }                                                                           \
                        const int diff = (int)ssd;                                                  \
                        if (diff < diff_max)                                                        \
                        {   /* template :685-694 */                                                 \
                            const int diffidx = diff * wfact;                                       \
                            const float weight = exptable[diffidx];                                 \
                            weight_sum += weight;                                                   \
                            pixel_sum  += weight * cmp[mirror(y + dy, h) * w + mirror(x + dx, w)];  \
                        }                                                                           \
                    }                                                                               \
                }                                                                                   \
            }                                                                                       \
            /* template :706-713 */                                                                 \
            const PIXEL result = (PIXEL)(pixel_sum / weight_sum);                                   \
            dst[y * w + x] = result ? result : src[y * w + x];                                      \
        }                                                                                           \
    }                                                                                               \
}

DEFINE_PLANE(plane_u8, uint8_t)
DEFINE_PLANE(plane_u16, uint16_t)

void oracle_nlmeans_plane(const void *const *frames, int nframes, int w, int h, int depth,
                          const oracle_nlmeans_plane_params_t *pp, void *dst)
{
    if (depth > 8) plane_u16(frames, nframes, w, h, depth, pp, dst);
    else           plane_u8(frames, nframes, w, h, depth, pp, dst);
}

int oracle_nlmeans_clip(const uint8_t *in, int n_in, int width, int height, int depth,
                        const oracle_nlmeans_plane_params_t pp[3], uint8_t *out)
{
    const int bps = depth > 8 ? 2 : 1;
    const int cw = -((-width) >> 1), ch = -((-height) >> 1);
    const int pw[3] = { width, cw, cw }, ph[3] = { height, ch, ch };
    size_t off[3], frame_bytes = 0;
    for (int c = 0; c < 3; c++)
    {
        off[c] = frame_bytes;
        frame_bytes += (size_t)pw[c] * ph[c] * bps;
    }
    for (int t = 0; t < n_in; t++)
    {
        for (int c = 0; c < 3; c++)
        {
            /* output t reads frames t .. t+nframes-1; at EOF the window shrinks (nlmeans.c:636-640) */
            int nf = pp[c].nframes;
            if (nf > n_in - t) nf = n_in - t;
            const void *frames[32];
            for (int f = 0; f < nf; f++)
                frames[f] = in + (size_t)(t + f) * frame_bytes + off[c];
            oracle_nlmeans_plane(frames, nf, pw[c], ph[c], depth, &pp[c], out + (size_t)t * frame_bytes + off[c]);
        }
    }
    return 0;
}
