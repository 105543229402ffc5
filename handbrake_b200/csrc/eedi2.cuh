// eedi2.cuh -- interface between decomb.cu and eedi2.cu (EEDI2 edge-directed interpolation).
#pragma once
#include <cuda_runtime.h>

namespace hbcu {

struct Eedi2Config
{
    int depth;
    int w[3], h[3], pitch[3];        // full-height plane geometry (pitch in elements = the reference's stride / bps)
    int half_frame_height;           // height of the half-height frame buffers (decomb.c:291-296: geometry.height / 2)
    int chroma_shift_h;
    int mthresh, vthresh, lthresh, dstr, estr, nt, maxd, pp;
};

struct Eedi2;

Eedi2 *eedi2_create(const Eedi2Config &cfg);
void   eedi2_destroy(Eedi2 *e);
// runs the whole stage chain (decomb template :366-441) for the three planes of `cur`;
// `tff` is pv->tff at that point (= !parity); results stay in the DST2PF planes
int    eedi2_run(Eedi2 *e, const void *const planes[3], int tff, cudaStream_t st);
const void *eedi2_output(const Eedi2 *e, int plane);
int    eedi2_debug_read(const Eedi2 *e, int which, void *host, size_t host_bytes);

}  // namespace hbcu
