/* unsharp_port.c -- TEST INFRASTRUCTURE (see oracle_port.h).
 *
 * Restates libhb/unsharp.c:88-168 and libhb/chroma_smooth.c:86-168 without the running-sum cascade: the cascade of
 * `steps` pairs of [1 1] accumulators per axis (SR[] along x, SC[][] along y) is the convolution with the binomial
 * row of order 2*steps in each direction, i.e. the sum over a size x size window of C(2s,i) C(2s,j) src[..] with edge
 * replication (the reference clamps x to [0,w-1] and keeps the last row for y >= h), total weight 2^(4*steps) =
 * 1 << scalebits.  All sums are uint32 and wrap (255 * 2^28 overflows for size 15), exactly like the cascade.
 *   unsharp       res = src + (((src - blur) * amount) >> 16), clamped to [0, (int16_t)max]
 *   chroma_smooth res = src - (((src - blur) * amount) >> 16), clamped to [(int16_t)max/16, (int16_t)(max - max/16)],
 *                 luma copied (amount 0)
 * The clamp bounds are declared int16_t in the reference (unsharp.c:108, chroma_smooth.c:112-113): at depth 16 they
 * wrap negative; restated as such.
 */
#include "oracle_port.h"
#include <string.h>

static uint32_t binom(int n, int k)
{
    uint64_t r = 1;
    for (int i = 1; i <= k; i++) r = r * (uint64_t)(n - k + i) / (uint64_t)i;
    return (uint32_t)r;
}

void oracle_unsharp_plane(const void *src_, void *dst_, int w, int h, int depth, double strength, int size, int smooth, int is_chroma)
{
    const int bps = depth > 8 ? 2 : 1;
    /* parameter sanitising: unsharp.c:262-272, chroma_smooth.c:243-268 */
    if (strength < 0) strength = 0;
    if (strength > (smooth ? 3.0 : 1.5)) strength = smooth ? 3.0 : 1.5;
    if (size % 2 == 0) size--;
    if (size < 3) size = 3;
    if (size > 15) size = 15;
    int amount = (int)(strength * 65536.0);
    if (smooth && !is_chroma) amount = 0;
    if (!amount)
    {
        memcpy(dst_, src_, (size_t)w * h * bps);
        return;
    }
    const int steps = size / 2, scalebits = steps * 4;
    const uint32_t halfscale = 1u << (scalebits - 1);
    const int maxi = 1 << depth;
    const int16_t max_value = (int16_t)(smooth ? maxi - maxi / 16 : maxi - 1);
    const int16_t min_value = (int16_t)(smooth ? maxi / 16 : 0);
    uint32_t B[16];
    for (int k = 0; k <= 2 * steps; k++) B[k] = binom(2 * steps, k);
    const uint8_t *s8 = src_;
    const uint16_t *s16 = src_;
    uint8_t *d8 = dst_;
    uint16_t *d16 = dst_;
    for (int y = 0; y < h; y++)
    {
        for (int x = 0; x < w; x++)
        {
            uint32_t sum = 0;
            for (int j = -steps; j <= steps; j++)
            {
                int yy = y + j;
                if (yy < 0) yy = 0;
                if (yy > h - 1) yy = h - 1;
                uint32_t row = 0;
                for (int i = -steps; i <= steps; i++)
                {
                    int xx = x + i;
                    if (xx < 0) xx = 0;
                    if (xx > w - 1) xx = w - 1;
                    const uint32_t v = bps == 1 ? s8[(size_t)yy * w + xx] : s16[(size_t)yy * w + xx];
                    row += B[i + steps] * v;
                }
                sum += B[j + steps] * row;
            }
            const int32_t v = bps == 1 ? s8[(size_t)y * w + x] : s16[(size_t)y * w + x];
            const int32_t blur = (int32_t)((sum + halfscale) >> scalebits);
            const int32_t delta = (int32_t)((uint32_t)(v - blur) * (uint32_t)amount) >> 16;
            const int32_t res = smooth ? v - delta : v + delta;
            const int32_t o = res > max_value ? max_value : res < min_value ? min_value : res;
            if (bps == 1) d8[(size_t)y * w + x] = (uint8_t)o;
            else          d16[(size_t)y * w + x] = (uint16_t)o;
        }
    }
}

/* a whole yuv420p clip, packed frames in and out; per-plane parameters after the cascade of unsharp.c:232-260 /
 * chroma_smooth.c:215-241 has been applied by the caller */
void oracle_unsharp_clip(const uint8_t *in, int n, int width, int height, int depth, const double strength[3], const int size[3],
                         int smooth, uint8_t *out)
{
    const int bps = depth > 8 ? 2 : 1;
    const int cw = (width + 1) >> 1, chh = (height + 1) >> 1;
    const size_t ybytes = (size_t)width * height * bps, cbytes = (size_t)cw * chh * bps, fb = ybytes + 2 * cbytes;
    for (int t = 0; t < n; t++)
    {
        const uint8_t *f = in + (size_t)t * fb;
        uint8_t *o = out + (size_t)t * fb;
        oracle_unsharp_plane(f, o, width, height, depth, strength[0], size[0], smooth, 0);
        oracle_unsharp_plane(f + ybytes, o + ybytes, cw, chh, depth, strength[1], size[1], smooth, 1);
        oracle_unsharp_plane(f + ybytes + cbytes, o + ybytes + cbytes, cw, chh, depth, strength[2], size[2], smooth, 1);
    }
}
