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

typedef struct { float weight_fact; int diff_max; const float *exptable; } nlm_table_t;

/* template :20-43: img[-1-x] = img[x]; img[w+x] = img[w-1-x] (rows likewise) */
static inline int mirror(int v, int n)
{
    if (v < 0)  return -1 - v;
    if (v >= n) return 2 * n - 1 - v;
    return v;
}

#define DEFINE_PLANE(NAME, PIXEL)                                                                   \
static void NAME(const void *const *frames, const void *const *pres, int nframes, int w, int h, int depth, \
                 const oracle_nlmeans_plane_params_t *pp, void *dst_v, const void *src_pre_v,       \
                 const nlm_table_t *given)                                                          \
{                                                                                                   \
    PIXEL *dst = dst_v;                                                                             \
    const PIXEL *src = frames[0];                                                                   \
    const PIXEL *src_pre = src_pre_v;                                                               \
    if (pp->strength == 0)                                                                          \
    {   /* nlmeans.c:493-499 */                                                                     \
        memcpy(dst, src, (size_t)w * h * sizeof(PIXEL));                                            \
        return;                                                                                     \
    }                                                                                               \
    float wfact, exptable[EXPSIZE];                                                                 \
    int diff_max;                                                                                   \
    if (given != NULL)                                                                              \
    {   /* the caller's table (the C-ABI hands the device exactly this, hbcu_nlmeans_plane_t) */    \
        wfact = given->weight_fact; diff_max = given->diff_max;                                     \
        memcpy(exptable, given->exptable, sizeof(exptable));                                        \
    }                                                                                               \
    else oracle_nlmeans_table(pp->strength, pp->patch_size, depth, &wfact, &diff_max, exptable);    \
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
                const PIXEL *cmp_pre = pres[f];   /* distances on image_pre, pixels from image */     \
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
                                const int d = (int)src_pre[ya * w + xa] - (int)cmp_pre[yb * w + xb]; \
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

/* ---------------- prefilters (templates/nlmeans_template.c:103-543) ---------------- */
#define DEFINE_PREFILTER(NAME, PIXEL, PIXEL2)                                                       \
static int NAME(const PIXEL *src, int w, int h, int filter_type, PIXEL *pre)                        \
{                                                                                                   \
    if (!(filter_type & (1 | 2 | 4 | 8 | 16 | 32))) return 0;                                       \
    /* priority csm5 > csm3 > median5 > median3 > mean5 > mean3 (:479-508) */                       \
    const int kind = (filter_type & (16 | 32)) ? 2 : (filter_type & (4 | 8)) ? 1 : 0;              \
    const int size = kind == 2 ? ((filter_type & 32) ? 5 : 3) : kind == 1 ? ((filter_type & 8) ? 5 : 3) \
                                                                           : ((filter_type & 2) ? 5 : 3); \
    const int lo = -((size - 1) / 2), hi = (size + 1) / 2;                                          \
    for (int y = 0; y < h; y++)                                                                     \
        for (int x = 0; x < w; x++)                                                                 \
        {                                                                                           \
            PIXEL v[25];                                                                            \
            int n = 0;                                                                              \
            for (int k = lo; k < hi; k++)                                                           \
                for (int j = lo; j < hi; j++)                                                       \
                    v[n++] = src[mirror(y + j, h) * w + mirror(x + k, w)];                          \
            const PIXEL c = src[y * w + x];                                                         \
            PIXEL out = c;                                                                          \
            if (kind == 0)                                                                          \
            {   /* mean (:115-129): pixel_2 sum times a double weight, truncated */                 \
                PIXEL2 sum = 0;                                                                     \
                for (int i = 0; i < n; i++) sum = sum + v[i];                                       \
                const double pixel_weight = 1.0 / (size * size);                                    \
                out = (PIXEL)(sum * pixel_weight);                                                  \
            }                                                                                       \
            else if (kind == 1)                                                                     \
            {   /* median (:135-198): the sorting networks return the true median */                \
                for (int i = 1; i < n; i++)                                                         \
                {                                                                                   \
                    const PIXEL t = v[i];                                                           \
                    int q = i;                                                                      \
                    while (q > 0 && v[q - 1] > t) { v[q] = v[q - 1]; q--; }                        \
                    v[q] = t;                                                                       \
                }                                                                                   \
                out = v[n / 2];                                                                     \
            }                                                                                       \
            else                                                                                    \
            {   /* csm (:232-323): clamp towards the range of the neighbours (origin excluded) */   \
                /* the reference leaves the inner (row) loop with `goto end` both after initialising   */ \
                /* min/max from the first neighbour and at the origin (:253-266): column k = offset_min */ \
                /* contributes only its first sample, column k = 0 only the samples above the origin   */ \
                int mn = v[0], mx = v[0];                                                           \
                for (int k = 1; k < size; k++)                                                      \
                    for (int j = 0; j < size; j++)                                                  \
                    {                                                                               \
                        if (k == size / 2 && j == size / 2) break;                                  \
                        const int pv = v[k * size + j];                                             \
                        if (pv < mn) mn = pv;                                                       \
                        if (pv > mx) mx = pv;                                                       \
                    }                                                                               \
                const PIXEL min = (PIXEL)mn, max = (PIXEL)mx;                                       \
                const PIXEL median = (min + max) / 2;                                               \
                const PIXEL min2 = (min + median) / 2, max2 = (max + median) / 2;                   \
                const PIXEL min3 = (min2 + median) / 2, max3 = (max2 + median) / 2;                 \
                if (c < min) out = min; else if (c > max) out = max;                                \
                else if (c < min2) out = min2; else if (c > max2) out = max2;                       \
                else if (c < min3) out = min3; else if (c > max3) out = max3;                       \
            }                                                                                       \
            pre[y * w + x] = out;                                                                   \
        }                                                                                           \
    if (filter_type & 1024)                                                                         \
    {   /* edgeboost (:325-426), raster order: a cleared mask sample changes the counts after it */ \
        static const int kernel[3][3] = { { -31, 0, 31 }, { -44, 0, 44 }, { -31, 0, 31 } };        \
        const double kernel_coef = 1.0 / 126.42;                                                    \
        const int bw = w + 2, bh = h + 2;                 /* the mask is zero outside the picture */ \
        PIXEL *mask_mem = calloc((size_t)bw * bh, sizeof(PIXEL));                                   \
        PIXEL *mask = mask_mem + bw + 1;                                                            \
        for (int y = 0; y < h; y++)                                                                 \
            for (int x = 0; x < w; x++)                                                             \
            {                                                                                       \
                PIXEL2 p1 = 0, p2 = 0;                                                              \
                for (int k = -1; k < 2; k++)                                                        \
                    for (int j = -1; j < 2; j++)                                                    \
                    {                                                                               \
                        const PIXEL sv = src[mirror(y + j, h) * w + mirror(x + k, w)];              \
                        p1 += kernel[j + 1][k + 1] * sv;                                            \
                        p2 += kernel[k + 1][j + 1] * sv;                                            \
                    }                                                                               \
                p1 = p1 > 0 ? p1 : -p1;                                                             \
                p2 = p2 > 0 ? p2 : -p2;                                                             \
                p1 = (PIXEL2)(((double)p1 * kernel_coef) + 128);                                    \
                p2 = (PIXEL2)(((double)p2 * kernel_coef) + 128);                                    \
                PIXEL m = (PIXEL)(p1 + p2);                                                         \
                m = m > 160 ? 235 : m > 16 ? 128 : 16;                                              \
                mask[y * bw + x] = m;                                                               \
            }                                                                                       \
        for (int y = 0; y < h; y++)                                                                 \
            for (int x = 0; x < w; x++)                                                             \
            {                                                                                       \
                if (mask[y * bw + x] <= 16) continue;                                               \
                int pixels = 0;                                                                     \
                for (int k = -1; k < 2; k++)                                                        \
                    for (int j = -1; j < 2; j++)                                                    \
                        if (mask[(y + j) * bw + (x + k)] > 16) pixels++;                            \
                if (pixels < 3) mask[y * bw + x] = 16;                                              \
                if (mask[y * bw + x] > 16)                                                          \
                {                                                                                   \
                    if (mask[y * bw + x] == 235) pre[y * w + x] = (3 * src[y * w + x] + 1 * pre[y * w + x]) / 4; \
                    else                         pre[y * w + x] = (2 * src[y * w + x] + 3 * pre[y * w + x]) / 5; \
                }                                                                                   \
            }                                                                                       \
        free(mask_mem);                                                                             \
    }                                                                                               \
    int wet = 1, dry = 0;                                  /* reduce (:510-533) */                  \
    if ((filter_type & 512) && (filter_type & 256)) { wet = 1; dry = 3; }                           \
    else if (filter_type & 512) { wet = 1; dry = 1; }                                               \
    else if (filter_type & 256) { wet = 3; dry = 1; }                                               \
    if (dry > 0)                                                                                    \
        for (int i = 0; i < w * h; i++) pre[i] = (PIXEL)((wet * pre[i] + dry * src[i]) / (wet + dry)); \
    return 1;                                                                                       \
}

DEFINE_PREFILTER(prefilter_u8, uint8_t, uint16_t)
DEFINE_PREFILTER(prefilter_u16, uint16_t, uint32_t)

int oracle_nlmeans_prefilter(const void *src, int w, int h, int depth, int filter_type, void *pre)
{
    return depth > 8 ? prefilter_u16(src, w, h, filter_type, pre) : prefilter_u8(src, w, h, filter_type, pre);
}

/* stale_src: nlmeans_plane reads frame[0].image_pre BEFORE it prefilters frame 0 (template :612 vs :628), so a frame
 * that was never a compare frame of an earlier output -- the first frame of the stream, every frame when the temporal
 * window is 1 -- contributes its UNFILTERED image as the source patch while the compare patches are prefiltered.
 * (With more than one worker thread the reference races on this; the contract restated here is threads=1.) */
static void nlmeans_plane_ex(const void *const *frames, int nframes, int w, int h, int depth,
                             const oracle_nlmeans_plane_params_t *pp, void *dst, int stale_src, const nlm_table_t *given)
{
    const int bps = depth > 8 ? 2 : 1;
    const size_t bytes = (size_t)w * h * bps;
    void *pre_mem[32];
    const void *pres[32];
    for (int f = 0; f < nframes; f++)
    {
        pre_mem[f] = NULL;
        pres[f] = frames[f];
        if (pp->prefilter)
        {
            pre_mem[f] = malloc(bytes);
            if (oracle_nlmeans_prefilter(frames[f], w, h, depth, pp->prefilter, pre_mem[f])) pres[f] = pre_mem[f];
        }
    }
    if (pp->prefilter & 2048)
    {   /* passthru (nlmeans.c:485-491): the prefiltered plane is the output, NLMeans itself does not run */
        memcpy(dst, pres[0], bytes);
    }
    else
    {
        /* the source patch pointer may be stale; the compare patch of frame 0 is always the prefiltered image */
        if (depth > 8) plane_u16(frames, pres, nframes, w, h, depth, pp, dst, stale_src ? frames[0] : pres[0], given);
        else           plane_u8(frames, pres, nframes, w, h, depth, pp, dst, stale_src ? frames[0] : pres[0], given);
    }
    for (int f = 0; f < nframes; f++) free(pre_mem[f]);
}

void oracle_nlmeans_plane(const void *const *frames, int nframes, int w, int h, int depth,
                          const oracle_nlmeans_plane_params_t *pp, void *dst)
{
    nlmeans_plane_ex(frames, nframes, w, h, depth, pp, dst, 0, NULL);
}

/* the same with the weight table GIVEN (what the C-ABI's hbcu_nlmeans_plane_t carries) instead of derived from a strength:
 * lets the host filter's own table arithmetic be part of what is checked */
void oracle_nlmeans_plane_with_table(const void *const *frames, int nframes, int w, int h, int depth,
                                     int patch_size, int range, double origin_tune, int bypass, int prefilter,
                                     float weight_fact, int diff_max, const float *exptable, int stale_src, void *dst)
{
    const oracle_nlmeans_plane_params_t pp = { bypass ? 0.0 : 1.0, origin_tune, patch_size, range, nframes, prefilter };
    const nlm_table_t tab = { weight_fact, diff_max, exptable };
    nlmeans_plane_ex(frames, nframes, w, h, depth, &pp, dst, stale_src, &tab);
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
            nlmeans_plane_ex(frames, nf, pw[c], ph[c], depth, &pp[c], out + (size_t)t * frame_bytes + off[c],
                             t == 0 || pp[c].nframes < 2, NULL);
        }
    }
    return 0;
}
