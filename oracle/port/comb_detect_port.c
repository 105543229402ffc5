/* comb_detect_port.c -- TEST INFRASTRUCTURE (see oracle_port.h).
 *
 * Restates HandBrake's comb detection for one (prev, cur, next) luma triple:
 *   raw mask        templates/comb_detect_template.c:288-402 (gamma) and :789-933 (integer)
 *   mask filtering  comb_detect.c:901-966 (filter), :726-792 (erode), :556-622 (dilate), order :1059-1068
 *   block scoring   comb_detect.c:221-276 (filtered) / :384-454 (unfiltered), verdict :1029-1049
 * Whole-image formulation: the reference's cpu_count segments tile the image
 * exactly (segment heights are multiples of the block height), so the verdict
 * is "HEAVY if any block > threshold, else LIGHT if any block >= threshold/2".
 */
#include "oracle_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct
{
    int w, h;
    const oracle_comb_params_t *p;
    int depth, max_value;
    int mthresh, athresh;                 /* depth-scaled */
    float g_mthresh, g_athresh, g_athresh6;
    float *gamma_lut;
} ctx_t;

#define PIXEL_AT(plane, depth, idx) ((depth) > 8 ? (int)((const uint16_t *)(plane))[idx] : (int)((const uint8_t *)(plane))[idx])

static void raw_mask(const ctx_t *c, const void *prev, const void *cur, const void *next,
                     int force, uint8_t *mask)
{
    const int w = c->w, h = c->h, depth = c->depth;
    const oracle_comb_params_t *p = c->p;
    memset(mask, 0, (size_t)w * h);
    const int gamma = p->mode & 1;
    const int athresh_sq = c->athresh * c->athresh, athresh6 = 6 * c->athresh;
    const int c32min = 10 << (depth - 8), c32max = 15 << (depth - 8);
    for (int y = 2; y < h - 2; y++)
    {
        for (int x = 0; x < w; x++)
        {
            const int i = y * w + x;
            const int pc = PIXEL_AT(cur, depth, i), pu1 = PIXEL_AT(cur, depth, i - w), pd1 = PIXEL_AT(cur, depth, i + w);
            const int pu2 = PIXEL_AT(cur, depth, i - 2 * w), pd2 = PIXEL_AT(cur, depth, i + 2 * w);
            const int qc = PIXEL_AT(prev, depth, i), qu1 = PIXEL_AT(prev, depth, i - w), qd1 = PIXEL_AT(prev, depth, i + w);
            const int nc = PIXEL_AT(next, depth, i), nu1 = PIXEL_AT(next, depth, i - w), nd1 = PIXEL_AT(next, depth, i + w);
            if (gamma)
            {
                const float *g = c->gamma_lut;
                const float up_diff = g[pc] - g[pu1], down_diff = g[pc] - g[pd1];
                if (!((up_diff > c->g_athresh && down_diff > c->g_athresh) ||
                      (up_diff < -c->g_athresh && down_diff < -c->g_athresh)))
                    continue;
                int motion = 0;
                if (c->g_mthresh > 0)
                {
                    if (fabs(g[qc] - g[pc]) > c->g_mthresh && fabs(g[pu1] - g[nu1]) > c->g_mthresh &&
                        fabs(g[pd1] - g[nd1]) > c->g_mthresh)
                        motion++;
                    if (fabs(g[nc] - g[pc]) > c->g_mthresh && fabs(g[qu1] - g[pu1]) > c->g_mthresh &&
                        fabs(g[qd1] - g[pd1]) > c->g_mthresh)
                        motion++;
                }
                else
                    motion = 1;
                if (motion || force)
                {
                    float combing = fabs(g[pu2] + (4 * g[pc]) + g[pd2] - (3 * (g[pu1] + g[pd1])));
                    if (combing > c->g_athresh6) mask[i] = 1;
                }
            }
            else
            {
                const int up_diff = pc - pu1, down_diff = pc - pd1;
                if (!((up_diff > c->athresh && down_diff > c->athresh) ||
                      (up_diff < -c->athresh && down_diff < -c->athresh)))
                    continue;
                int motion = 0;
                if (c->mthresh > 0)
                {
                    if (abs(qc - pc) > c->mthresh && abs(pu1 - nu1) > c->mthresh && abs(pd1 - nd1) > c->mthresh) motion++;
                    if (abs(nc - pc) > c->mthresh && abs(qu1 - pu1) > c->mthresh && abs(qd1 - pd1) > c->mthresh) motion++;
                }
                else
                    motion = 1;
                if (motion || force)
                {
                    if (p->spatial_metric == 0)
                    {
                        if (abs(pc - pd2) < c32min && abs(pc - pd1) > c32max) mask[i] = 1;
                    }
                    else if (p->spatial_metric == 1)
                    {
                        if ((pu1 - pc) * (pd1 - pc) > athresh_sq) mask[i] = 1;
                    }
                    else if (p->spatial_metric == 2)
                    {
                        if (abs(pu2 + 4 * pc + pd2 - 3 * (pu1 + pd1)) > athresh6) mask[i] = 1;
                    }
                }
            }
        }
    }
}

/* dst interior <- f(src 3x3); the one-pixel frame of every intermediate mask stays 0 (init memset :1193-1195) */
enum { OP_FILTER_CLASSIC, OP_FILTER_HV, OP_ERODE, OP_DILATE };
static void stencil(const uint8_t *src, uint8_t *dst, int w, int h, int op)
{
    memset(dst, 0, (size_t)w * h);
    for (int y = 1; y < h - 1; y++)
        for (int x = 1; x < w - 1; x++)
        {
            const uint8_t *s = src + y * w + x;
            int v;
            if (op == OP_FILTER_CLASSIC || op == OP_FILTER_HV)
            {
                const int hc = s[-1] & s[0] & s[1], vc = s[-w] & s[0] & s[w];
                v = op == OP_FILTER_CLASSIC ? hc : (hc & vc);
            }
            else
            {
                const int count = s[-w - 1] + s[-w] + s[-w + 1] + s[-1] + s[1] + s[w - 1] + s[w] + s[w + 1];
                if (op == OP_ERODE) v = s[0] ? count >= 2 : 0;
                else                v = s[0] ? 1 : count >= 4;
            }
            dst[y * w + x] = (uint8_t)v;
        }
}

/* comb_detect.c:1074-1081 */
void oracle_comb_gamma_lut(int depth, float *out)
{
    const int max = (1 << depth) - 1;
    for (int i = 0; i < max + 1; i++)
        out[i] = pow(((float)i / (float)max), 2.2f);
}

int oracle_comb_detect(const void *prev, const void *cur, const void *next, int w, int h, int depth,
                       const oracle_comb_params_t *p, int force, uint8_t *mask_out, uint8_t *filtered_out)
{
    ctx_t c;
    c.w = w; c.h = h; c.p = p; c.depth = depth; c.max_value = (1 << depth) - 1;
    c.mthresh = p->motion_threshold << (depth - 8);
    c.athresh = p->spatial_threshold << (depth - 8);
    c.g_mthresh  = (float)c.mthresh / (float)c.max_value;
    c.g_athresh  = (float)c.athresh / (float)c.max_value;
    c.g_athresh6 = 6 * c.g_athresh;
    c.gamma_lut = malloc(sizeof(float) * (c.max_value + 1));
    oracle_comb_gamma_lut(depth, c.gamma_lut);

    uint8_t *mask = malloc((size_t)w * h), *t1 = malloc((size_t)w * h), *t2 = malloc((size_t)w * h);
    raw_mask(&c, prev, cur, next, force, mask);
    const uint8_t *scored = mask;
    const int filtered = (p->mode & 2) != 0;
    if (filtered)
    {
        if (p->filter_mode == 1)
        {
            stencil(mask, t1, w, h, OP_FILTER_CLASSIC);
            scored = t1;
        }
        else
        {
            stencil(mask, t2, w, h, OP_FILTER_HV);
            if (p->filter_mode == 2)
            {
                stencil(t2, t1, w, h, OP_ERODE);
                stencil(t1, t2, w, h, OP_DILATE);
                stencil(t2, t1, w, h, OP_ERODE);
                scored = t1;
            }
            else
            {
                /* any other filter-mode: mask_filter writes mask_temp, nothing fills mask_filtered (stays 0) */
                memset(t1, 0, (size_t)w * h);
                scored = t1;
            }
        }
    }
    int bw = p->block_width > w ? w : p->block_width, bh = p->block_height > h ? h : p->block_height;
    int light = 0, heavy = 0;
    for (int y = 0; y + bh <= h; y += bh)
        for (int x = 0; x < w - bw; x += bw)
        {
            int score = 0;
            for (int by = 0; by < bh; by++)
                for (int bx = 0; bx < bw; bx++)
                {
                    const int xx = x + bx;
                    const uint8_t *m = scored + (y + by) * w + xx;
                    if (filtered) score += m[0];
                    else if (xx == 0) score += m[0] & m[1];
                    else if (xx == w - 1) score += m[-1] & m[0];
                    else score += m[-1] & m[0] & m[1];
                }
            if (score >= p->block_threshold / 2) light = 1;
            if (score > p->block_threshold) heavy = 1;
        }
    if (mask_out) memcpy(mask_out, mask, (size_t)w * h);
    if (filtered_out) memcpy(filtered_out, scored, (size_t)w * h);
    free(mask); free(t1); free(t2); free(c.gamma_lut);
    return heavy ? 2 : light ? 1 : 0;
}

int oracle_comb_detect_clip(const uint8_t *in, int n_in, int width, int height, int depth,
                            const oracle_comb_params_t *p, uint8_t *verdicts)
{
    const int bps = depth > 8 ? 2 : 1;
    const int cw = -((-width) >> 1), ch = -((-height) >> 1);
    const size_t frame_bytes = ((size_t)width * height + 2 * (size_t)cw * ch) * bps;
    for (int t = 0; t < n_in; t++)
    {
        /* first frame: prev is a duplicate of itself; last frame: next is a duplicate of itself;
         * both of those passes run with force_exaustive_check (comb_detect.c:1111,1534,1552) */
        const uint8_t *cur = in + (size_t)t * frame_bytes;
        const uint8_t *prev = t > 0 ? cur - frame_bytes : cur;
        const uint8_t *next = t + 1 < n_in ? cur + frame_bytes : cur;
        const int force = (t == 0) || (t + 1 == n_in);
        verdicts[t] = (uint8_t)oracle_comb_detect(prev, cur, next, width, height, depth, p, force, NULL, NULL);
    }
    return 0;
}
