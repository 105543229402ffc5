/* hostlogic_frames.c -- TEST INFRASTRUCTURE (see hostlogic_nlmeans.c for the idea).
 *
 * Stand-ins for the device-frame half of the C-ABI (hbcu_frame_*, hbcu_xfer_*): a "device frame" is a reference-counted
 * block of host memory with the planes back to back at the given strides.  With these the host side of a device-resident
 * chain -- HBCU_DEVICE hb_buffer_t backing, shallow dups and closes, the upload / download adapter filters, hw_pix_fmt
 * propagation through init(), filters reading and emitting device buffers (handbrake_b200/libhb/hbcu_device_frames.c and
 * the *_cuda.c filters) -- runs on a machine without a GPU and is compared with the reference's plain host chain.
 * Never linked into the product.
 */
#include "../../include/hbcu.h"

#include <stdlib.h>
#include <string.h>

void oracle_hostlogic_set_error(const char *fmt, ...);

struct hbcu_frame_s
{
    int refs, device;
    int row_bytes[3], rows[3], strides[3];
    uint8_t *base;
    void *planes[3];
    hbcu_frame_release_fn ext_release;      /* wrapped frames: somebody else's memory */
    void *ext_opaque;
    int external;
};

static long frames_alive = 0;

int oracle_hbcu_frame_alloc(hbcu_frame_t **out, int device, const int row_bytes[3], const int rows[3], const int strides[3])
{
    size_t bytes = 0;
    for (int p = 0; p < 3; p++)
    {
        if (row_bytes[p] <= 0 || rows[p] <= 0 || strides[p] < row_bytes[p] || strides[p] % 16)
        {
            oracle_hostlogic_set_error("frame_alloc: bad geometry of plane %d", p);
            return -1;
        }
        bytes += (size_t)strides[p] * rows[p];
    }
    struct hbcu_frame_s *f = calloc(1, sizeof(*f));
    f->refs = 1;
    f->device = device;
    f->base = calloc(1, bytes + 64);                    /* device frames start zeroed, like the real pool's */
    size_t off = 0;
    for (int p = 0; p < 3; p++)
    {
        f->row_bytes[p] = row_bytes[p]; f->rows[p] = rows[p]; f->strides[p] = strides[p];
        f->planes[p] = f->base + off;
        off += (size_t)strides[p] * rows[p];
    }
    frames_alive++;
    *out = f;
    return 0;
}
void oracle_hbcu_frame_retain(hbcu_frame_t *f) { if (f) f->refs++; }
void oracle_hbcu_frame_release(hbcu_frame_t *f)
{
    if (f == NULL || --f->refs > 0) return;
    frames_alive--;
    if (f->external)
    {
        if (f->ext_release) f->ext_release(f->ext_opaque);
    }
    else free(f->base);
    free(f);
}

int oracle_hbcu_frame_wrap(hbcu_frame_t **out, int device, void *const dplanes[3], const int row_bytes[3], const int rows[3],
                           const int strides[3], size_t readable_tail_bytes, void *producer_stream,
                           hbcu_frame_release_fn release, void *opaque)
{
    (void)producer_stream;
    if (readable_tail_bytes < 256)
    {
        oracle_hostlogic_set_error("frame_wrap: no readable tail");
        return -1;
    }
    struct hbcu_frame_s *f = calloc(1, sizeof(*f));
    f->refs = 1;
    f->device = device;
    f->external = 1;
    f->ext_release = release;
    f->ext_opaque = opaque;
    for (int p = 0; p < 3; p++)
    {
        f->row_bytes[p] = row_bytes[p]; f->rows[p] = rows[p]; f->strides[p] = strides[p];
        f->planes[p] = dplanes[p];
    }
    frames_alive++;
    *out = f;
    return 0;
}
int oracle_hbcu_frame_acquire(hbcu_frame_t *f, void *s) { (void)f; (void)s; return 0; }
int oracle_hbcu_frame_done(hbcu_frame_t *f, void *s) { (void)f; (void)s; return 0; }
void *oracle_hbcu_frame_plane(const hbcu_frame_t *f, int plane) { return f->planes[plane]; }
int   oracle_hbcu_frame_stride(const hbcu_frame_t *f, int plane) { return f->strides[plane]; }
int   oracle_hbcu_frame_device(const hbcu_frame_t *f) { return f->device; }
long  oracle_hbcu_frames_alive(void) { return frames_alive; }
void  oracle_hbcu_frame_trim(void) {}

const void *const *oracle_hostlogic_frame_planes(const hbcu_frame_t *f) { return (const void *const *)f->planes; }
const int *oracle_hostlogic_frame_strides(const hbcu_frame_t *f) { return f->strides; }

struct hbcu_xfer_s { int device; };

int oracle_hbcu_xfer_create(hbcu_xfer_t **x, int device, int depth)
{
    (void)depth;
    *x = calloc(1, sizeof(**x));
    (*x)->device = device;
    return 0;
}
void oracle_hbcu_xfer_destroy(hbcu_xfer_t *x) { free(x); }
static void copy_planes(const hbcu_frame_t *f, void *const dst[3], const int dstride[3], const void *const src[3], const int sstride[3])
{
    for (int p = 0; p < 3; p++)
        for (int y = 0; y < f->rows[p]; y++)
            memcpy((uint8_t *)dst[p] + (size_t)y * dstride[p], (const uint8_t *)src[p] + (size_t)y * sstride[p], (size_t)f->row_bytes[p]);
}
int oracle_hbcu_xfer_upload(hbcu_xfer_t *x, int64_t ticket, hbcu_frame_t *f, const void *const planes[3], const int strides[3])
{
    (void)x; (void)ticket;
    copy_planes(f, f->planes, f->strides, planes, strides);
    return 0;
}
int oracle_hbcu_xfer_download(hbcu_xfer_t *x, int64_t ticket, hbcu_frame_t *f, void *const planes[3], const int strides[3])
{
    (void)x; (void)ticket;
    copy_planes(f, planes, strides, (const void *const *)f->planes, f->strides);
    return 0;
}
int oracle_hbcu_xfer_wait(hbcu_xfer_t *x, int64_t ticket) { (void)x; (void)ticket; return 0; }
int oracle_hbcu_xfer_poll(hbcu_xfer_t *x, int64_t ticket) { (void)x; (void)ticket; return 1; }
