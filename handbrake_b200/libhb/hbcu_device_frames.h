/* hbcu_device_frames.h -- HBCU_DEVICE backing of hb_buffer_t and the two adapter filters of a device-resident
 * filter chain.  See hbcu_device_frames.c. */
#ifndef HBCU_DEVICE_FRAMES_H
#define HBCU_DEVICE_FRAMES_H

#include "handbrake/handbrake.h"
#include "hbcu.h"

/* adapter filter ids (outside libhb's own range, like the VT pre/post adapters are outside the user-visible ids) */
#define HB_FILTER_HBCU_UPLOAD   (HB_FILTER_MT_FRAME + 1)
#define HB_FILTER_HBCU_DOWNLOAD (HB_FILTER_MT_FRAME + 2)

extern hb_filter_object_t hb_filter_hbcu_upload;     /* host hb_buffer_t  -> HBCU_DEVICE hb_buffer_t */
extern hb_filter_object_t hb_filter_hbcu_download;   /* HBCU_DEVICE hb_buffer_t -> pinned host hb_buffer_t */

/* a frame buffer whose planes live in HBM on `device` (hb_frame_buffer_init's device twin, fifo.c:839-881):
 * same f.fmt/width/height, plane[].width/height/stride as the host buffer, plane[].data = DEVICE pointers,
 * data = NULL, storage_type = HBCU_DEVICE, storage = the hbcu_frame_t */
hb_buffer_t  *hbcu_device_frame_buffer_init(int pix_fmt, int width, int height, int device);
/* the device frame behind a buffer, NULL for host buffers */
hbcu_frame_t *hbcu_buffer_frame(const hb_buffer_t *b);
/* an AV_PIX_FMT_CUDA AVFrame's planes (data / linesize / device / stream) as an HBCU_DEVICE buffer, no copy */
hb_buffer_t  *hbcu_wrap_cuda_frame(int pix_fmt, int width, int height, int device, void *const data[3], const int linesize[3],
                                   size_t readable_tail_bytes, void *cuda_stream, void (*release)(void *), void *opaque);
/* does this filter instance hand its output on in HBM?  (init->hw_pix_fmt == AV_PIX_FMT_CUDA, the way libhb marks
 * a hardware-frame pipeline: nvenc_common.c:329-336, hwaccel.c:15-60) */
int           hbcu_init_wants_device_output(const hb_filter_init_t *init);
int           hbcu_env_device(void);
/* the devices a frame-parallel filter deals its stream to: setting `devices=0,1,..` (not part of the reference's templates:
 * a front end that validates settings appends the key), else HBCU_DEVICES, else the one device of
 * HBCU_DEVICE.  Returns the count (>= 1), or -1 for a malformed list.  An ordinal may repeat (two handles on one GPU). */
#define HBCU_MAX_DEVICES 16
int           hbcu_settings_devices(const hb_dict_t *settings, int devices[HBCU_MAX_DEVICES]);

#endif
