/* lapsharp_port.c -- TEST INFRASTRUCTURE (see oracle_port.h).
 * Restates HandBrake's lapsharp (libhb/lapsharp.c:36-101 kernels, :125-182 DEF_LAPSHARP_FUNC) for
 * one plane given WITH its stride: the filter's border rule depends on (stride - width) / 2 and its
 * last columns read the stride region, so the port takes the strided plane as libhb lays it out
 * (hb_frame_buffer_mirror_stride already applied by the caller, lapsharp.c:333).
 */
#include "oracle_port.h"

static const int k_lap[]    = { 0, -1, 0, -1, 5, -1, 0, -1, 0 };
static const int k_isolap[] = { -1, -4, -1, -4, 25, -4, -1, -4, -1 };
static const int k_log[]    = { 0, 0, -1, 0, 0, 0, -1, -2, -1, 0, -1, -2, 21, -2, -1, 0, -1, -2, -1, 0, 0, 0, -1, 0, 0 };
static const int k_isolog[] = { 0, -1, -1, -1, 0, -1, -3, -4, -3, -1, -1, -4, 55, -4, -1, -1, -3, -4, -3, -1, 0, -1, -1, -1, 0 };
static const struct { const int *mem; int size; double coef; } kernels[4] = {
    { k_lap, 3, 1.0 }, { k_isolap, 3, 1.0 / 5 }, { k_log, 5, 1.0 / 5 }, { k_isolog, 5, 1.0 / 15 } };

void oracle_lapsharp_plane(const void *src_v, void *dst_v, int width, int height, int stride_src, int stride_dst,
                           int depth, int kernel_id, double strength)
{
    const int bps = depth > 8 ? 2 : 1, max_value = (1 << depth) - 1;
    const int size = kernels[kernel_id].size;
    const double coef = kernels[kernel_id].coef;
    const int ss = stride_src / bps, ds = stride_dst / bps;
    const int offset_min = -((size - 1) / 2), offset_max = (size + 1) / 2;
    const int stride_border = (ss - width) / 2;
    for (int y = 0; y < height; y++)
    {
        for (int x = 0; x < width; x++)
        {
#define SRC(yy, xx) (bps == 2 ? (int)((const uint16_t *)src_v)[(size_t)(yy) * ss + (xx)] : (int)((const uint8_t *)src_v)[(size_t)(yy) * ss + (xx)])
            const int s0 = SRC(y, x);
            int out;
            if (y < offset_max || y > height - offset_max || x < stride_border + offset_max || x > width + stride_border - offset_max)
            {
                out = s0;
            }
            else
            {
                int acc = 0;
                for (int k = offset_min; k < offset_max; k++)
                    for (int j = offset_min; j < offset_max; j++)
                        acc += kernels[kernel_id].mem[(j - offset_min) * size + k - offset_min] * SRC(y + j, x + k);
                if (bps == 1)
                {
                    int16_t pixel = (int16_t)acc;
                    pixel = (int16_t)(((pixel * coef) - s0) * strength) + s0;
                    out = pixel;
                }
                else
                {
                    int32_t pixel = acc;
                    pixel = (int32_t)(((pixel * coef) - s0) * strength) + s0;
                    out = pixel;
                }
                out = out < 0 ? 0 : out;
                out = out > max_value ? max_value : out;
            }
#undef SRC
            if (bps == 2) ((uint16_t *)dst_v)[(size_t)y * ds + x] = (uint16_t)out;
            else          ((uint8_t *)dst_v)[(size_t)y * ds + x] = (uint8_t)out;
        }
    }
}
