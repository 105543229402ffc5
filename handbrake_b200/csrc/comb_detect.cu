// comb_detect.cu -- comb detection for sm_100a behind the C-ABI of include/hbcu.h.
//
// Replaces (reference /root/reference/libhb):
//   detect_gamma_combed_segment / detect_combed_segment   templates/comb_detect_template.c:288-402, 789-933
//   mask_filter_work / mask_erode_work / mask_dilate_work comb_detect.c:901-966, 726-792, 556-622
//   check_filtered_combing_mask / check_combing_mask      comb_detect.c:221-276, 384-454
//   check_combing_results                                 comb_detect.c:1029-1049
// The reference runs five fork/join tasksets over row segments; here one frame is
// three kernels on one stream:
//   comb_mask_kernel    raw mask from the prev/cur/next luma planes (gamma values come from the
//                       host-built table staged in shared memory, so the float results are the
//                       reference's bit for bit; no FMA contraction: -fmad=false + explicit rn ops)
//   comb_filter_kernel  filter -> erode -> dilate -> erode fused in shared memory (4-pixel halo),
//                       intermediate masks never touch HBM
//   comb_score_kernel   block sums -> LIGHT/HEAVY flags (atomicOr), verdict read back through
//                       pinned memory
// Integer work is bit-exact by construction; the segment decomposition of the reference does
// not influence its result (block grid is globally aligned).
#include "hbcu_common.h"
#include "hbcu_frames.h"
#include "../../include/hbcu.h"

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace {

using hbcu::set_error;

struct CombParams
{
    int w, h, pitch;              // luma plane, pitch in elements
    int mode, spatial_metric, filter_mode;
    int mthresh, athresh, athresh_sq, athresh6;
    int c32min, c32max;
    float g_mthresh, g_athresh, g_athresh6;
    int force;
    int lut_size;
    const float *gamma_lut;
    int block_threshold, block_width, block_height;
};

// ---------------------------------------------------------------------------
// raw mask: one thread per pixel, rows [2, h-2); everything else is 0
// ---------------------------------------------------------------------------
template <typename PIX, bool GAMMA>
__global__ void __launch_bounds__(256) comb_mask_kernel(const PIX *__restrict__ prev, const PIX *__restrict__ cur,
                                                       const PIX *__restrict__ next, uint8_t *__restrict__ mask,
                                                       int mpitch, CombParams p)
{
    extern __shared__ float s_lut[];
    if (GAMMA)
    {
        for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < p.lut_size; i += blockDim.x * blockDim.y)
            s_lut[i] = p.gamma_lut[i];
        __syncthreads();
    }
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= p.w || y >= p.h) return;
    uint8_t m = 0;
    if (y >= 2 && y < p.h - 2)
    {
        const size_t i = (size_t)y * p.pitch + x;
        const int pc = cur[i], pu1 = cur[i - p.pitch], pd1 = cur[i + p.pitch];
        if (GAMMA)
        {
            const float gc = s_lut[pc], gu1 = s_lut[pu1], gd1 = s_lut[pd1];
            const float up = __fsub_rn(gc, gu1), down = __fsub_rn(gc, gd1);
            if ((up > p.g_athresh && down > p.g_athresh) || (up < -p.g_athresh && down < -p.g_athresh))
            {
                int motion = 0;
                if (p.g_mthresh > 0)
                {
                    const float qc = s_lut[prev[i]], nc = s_lut[next[i]];
                    const float qu1 = s_lut[prev[i - p.pitch]], qd1 = s_lut[prev[i + p.pitch]];
                    const float nu1 = s_lut[next[i - p.pitch]], nd1 = s_lut[next[i + p.pitch]];
                    if (fabsf(__fsub_rn(qc, gc)) > p.g_mthresh && fabsf(__fsub_rn(gu1, nu1)) > p.g_mthresh &&
                        fabsf(__fsub_rn(gd1, nd1)) > p.g_mthresh)
                        motion++;
                    if (fabsf(__fsub_rn(nc, gc)) > p.g_mthresh && fabsf(__fsub_rn(qu1, gu1)) > p.g_mthresh &&
                        fabsf(__fsub_rn(qd1, gd1)) > p.g_mthresh)
                        motion++;
                }
                else
                    motion = 1;
                if (motion || p.force)
                {
                    const float gu2 = s_lut[cur[i - 2 * p.pitch]], gd2 = s_lut[cur[i + 2 * p.pitch]];
                    // fabs(up2 + 4*c + down2 - 3*(up1 + down1)), left to right, no contraction
                    const float lhs = __fadd_rn(__fadd_rn(gu2, __fmul_rn(4.0f, gc)), gd2);
                    const float rhs = __fmul_rn(3.0f, __fadd_rn(gu1, gd1));
                    if (fabsf(__fsub_rn(lhs, rhs)) > p.g_athresh6) m = 1;
                }
            }
        }
        else
        {
            const int up = pc - pu1, down = pc - pd1;
            if ((up > p.athresh && down > p.athresh) || (up < -p.athresh && down < -p.athresh))
            {
                int motion = 0;
                if (p.mthresh > 0)
                {
                    const int qc = prev[i], nc = next[i];
                    const int qu1 = prev[i - p.pitch], qd1 = prev[i + p.pitch];
                    const int nu1 = next[i - p.pitch], nd1 = next[i + p.pitch];
                    if (abs(qc - pc) > p.mthresh && abs(pu1 - nu1) > p.mthresh && abs(pd1 - nd1) > p.mthresh) motion++;
                    if (abs(nc - pc) > p.mthresh && abs(qu1 - pu1) > p.mthresh && abs(qd1 - pd1) > p.mthresh) motion++;
                }
                else
                    motion = 1;
                if (motion || p.force)
                {
                    const int pu2 = cur[i - 2 * p.pitch], pd2 = cur[i + 2 * p.pitch];
                    if (p.spatial_metric == 0)
                        m = (abs(pc - pd2) < p.c32min) && (abs(pc - pd1) > p.c32max);
                    else if (p.spatial_metric == 1)
                        m = (pu1 - pc) * (pd1 - pc) > p.athresh_sq;
                    else if (p.spatial_metric == 2)
                        m = abs(pu2 + 4 * pc + pd2 - 3 * (pu1 + pd1)) > p.athresh6;
                }
            }
        }
    }
    mask[(size_t)y * mpitch + x] = m;
}

// ---------------------------------------------------------------------------
// mask filter chain fused in shared memory.  Output tile FT_W x FT_H, halo 4.
// Every intermediate mask is 0 outside [1,w-2] x [1,h-2] (the reference never writes there).
// ---------------------------------------------------------------------------
constexpr int FT_W = 64, FT_H = 32, FT_HALO = 4;
constexpr int FS_W = FT_W + 2 * FT_HALO, FS_H = FT_H + 2 * FT_HALO;

enum { OP_CLASSIC = 0, OP_HV = 1, OP_ERODE = 2, OP_DILATE = 3 };

__device__ __forceinline__ void filter_stage(const uint8_t *src, uint8_t *dst, int op, int shrink,
                                             int gx0, int gy0, int w, int h)
{
    // computes dst on the tile region shrunk by `shrink` pixels per side from src (valid one pixel wider)
    const int rw = FS_W - 2 * shrink, rh = FS_H - 2 * shrink;
    for (int i = threadIdx.x; i < rw * rh; i += blockDim.x)
    {
        const int lx = shrink + i % rw, ly = shrink + i / rw;
        const int gx = gx0 + lx, gy = gy0 + ly;
        uint8_t v = 0;
        if (gx >= 1 && gx <= w - 2 && gy >= 1 && gy <= h - 2)
        {
            const uint8_t *s = src + ly * FS_W + lx;
            if (op == OP_CLASSIC || op == OP_HV)
            {
                const int hc = s[-1] & s[0] & s[1];
                const int vc = s[-FS_W] & s[0] & s[FS_W];
                v = op == OP_CLASSIC ? hc : (hc & vc);
            }
            else
            {
                const int count = s[-FS_W - 1] + s[-FS_W] + s[-FS_W + 1] + s[-1] + s[1] + s[FS_W - 1] + s[FS_W] + s[FS_W + 1];
                if (op == OP_ERODE) v = s[0] ? (count >= 2) : 0;
                else                v = s[0] ? 1 : (count >= 4);
            }
        }
        dst[ly * FS_W + lx] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) comb_filter_kernel(const uint8_t *__restrict__ mask, uint8_t *__restrict__ out,
                                                         int mpitch, int w, int h, int filter_mode)
{
    __shared__ uint8_t a[FS_W * FS_H], b[FS_W * FS_H];
    const int gx0 = blockIdx.x * FT_W - FT_HALO, gy0 = blockIdx.y * FT_H - FT_HALO;
    for (int i = threadIdx.x; i < FS_W * FS_H; i += blockDim.x)
    {
        const int gx = gx0 + i % FS_W, gy = gy0 + i / FS_W;
        a[i] = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? mask[(size_t)gy * mpitch + gx] : 0;
        b[i] = 0;
    }
    __syncthreads();
    const uint8_t *res;
    if (filter_mode == 1)
    {
        filter_stage(a, b, OP_CLASSIC, 1, gx0, gy0, w, h);
        res = b;
    }
    else
    {
        filter_stage(a, b, OP_HV, 1, gx0, gy0, w, h);          // mask -> temp
        if (filter_mode == 2)
        {
            filter_stage(b, a, OP_ERODE, 2, gx0, gy0, w, h);   // temp -> filtered
            filter_stage(a, b, OP_DILATE, 3, gx0, gy0, w, h);  // filtered -> temp
            filter_stage(b, a, OP_ERODE, 4, gx0, gy0, w, h);   // temp -> filtered
            res = a;
        }
        else
        {
            res = nullptr;                                      // nothing ever writes mask_filtered: stays 0
        }
    }
    for (int i = threadIdx.x; i < FT_W * FT_H; i += blockDim.x)
    {
        const int lx = FT_HALO + i % FT_W, ly = FT_HALO + i / FT_W;
        const int gx = gx0 + lx, gy = gy0 + ly;
        if (gx < w && gy < h) out[(size_t)gy * mpitch + gx] = res ? res[ly * FS_W + lx] : 0;
    }
}

// ---------------------------------------------------------------------------
// block scores -> flags: bit0 some block >= threshold/2, bit1 some block > threshold
// one warp per block
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) comb_score_kernel(const uint8_t *__restrict__ m, int mpitch, int w, int h,
                                                        int bw, int bh, int nbx, int nby, int threshold, int filtered,
                                                        int *__restrict__ flags)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nbx * nby) return;
    const int x0 = (warp % nbx) * bw, y0 = (warp / nbx) * bh;
    int score = 0;
    for (int i = lane; i < bw * bh; i += 32)
    {
        const int x = x0 + i % bw, y = y0 + i / bw;
        const uint8_t *p = m + (size_t)y * mpitch + x;
        if (filtered) score += p[0];
        else if (x == 0) score += p[0] & p[1];
        else if (x == w - 1) score += p[-1] & p[0];
        else score += p[-1] & p[0] & p[1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) score += __shfl_xor_sync(0xffffffffu, score, o);
    if (lane == 0)
    {
        int f = 0;
        if (score >= threshold / 2) f |= 1;
        if (score > threshold) f |= 2;
        if (f) atomicOr(flags, f);
    }
}

}  // namespace

struct hbcu_comb_detect_s
{
    hbcu_comb_detect_config_t cfg;
    int bps, pitch, mpitch;
    size_t plane_bytes;
    int slots;
    std::vector<uint8_t *> luma;
    std::vector<int64_t> index;
    std::vector<cudaEvent_t> ev_upload, ev_readers;
    uint8_t *d_mask, *d_scored;
    float *d_lut;
    int nres;
    int *d_flags;            // nres ints
    int *h_flags;            // pinned
    std::vector<int64_t> res_index;
    std::vector<cudaEvent_t> ev_result;
    int next_res;
    cudaStream_t s_h2d, s_compute;
    cudaEvent_t ev_mark[2];
};

extern "C" {

int hbcu_comb_detect_create(hbcu_comb_detect_t **out, const hbcu_comb_detect_config_t *cfg)
{
    if (out == nullptr || cfg == nullptr || cfg->gamma_lut == nullptr)
    {
        set_error("comb_detect_create: null argument");
        return -1;
    }
    *out = nullptr;
    if (cfg->width < 4 || cfg->height < 5 || cfg->depth < 8 || cfg->depth > 16 || cfg->block_width < 1 || cfg->block_height < 1)
    {
        set_error("comb_detect_create: unsupported geometry %dx%d depth %d", cfg->width, cfg->height, cfg->depth);
        return -1;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev)
    {
        cudaGetLastError();
        set_error("comb_detect_create: CUDA device %d not available (%d devices); there is no CPU fallback", cfg->device, ndev);
        return -1;
    }
    HBCU_CHECK(cudaSetDevice(cfg->device));
    hbcu_comb_detect_s *h = new (std::nothrow) hbcu_comb_detect_s();
    if (h == nullptr) { set_error("comb_detect_create: out of memory"); return -1; }
    h->cfg = *cfg;
    h->bps = cfg->depth > 8 ? 2 : 1;
    h->pitch = (cfg->width + 127) / 128 * 128;
    h->mpitch = h->pitch;
    h->plane_bytes = (size_t)h->pitch * cfg->height * h->bps;
    h->slots = cfg->slots >= 4 ? cfg->slots : 4;
    h->nres = 16;
    h->next_res = 0;
    h->d_mask = h->d_scored = nullptr;
    h->d_lut = nullptr;
    h->d_flags = nullptr;
    h->h_flags = nullptr;
    h->s_h2d = h->s_compute = nullptr;
    h->ev_mark[0] = h->ev_mark[1] = nullptr;
#define CK(expr)                                                                  \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) {                                                  \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));            \
            hbcu_comb_detect_destroy(h);                                          \
            return -1;                                                            \
        }                                                                         \
    } while (0)
    CK(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking));
    h->luma.assign(h->slots, nullptr);
    h->index.assign(h->slots, -1);
    h->ev_upload.assign(h->slots, nullptr);
    h->ev_readers.assign(h->slots, nullptr);
    for (int s = 0; s < h->slots; s++)
    {
        CK(cudaMalloc(&h->luma[s], h->plane_bytes));
        CK(cudaEventCreateWithFlags(&h->ev_upload[s], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_readers[s], cudaEventDisableTiming));
    }
    const size_t mbytes = (size_t)h->mpitch * cfg->height;
    CK(cudaMalloc(&h->d_mask, mbytes));
    CK(cudaMalloc(&h->d_scored, mbytes));
    CK(cudaMemset(h->d_mask, 0, mbytes));
    CK(cudaMemset(h->d_scored, 0, mbytes));
    const int lut_size = 1 << cfg->depth;
    CK(cudaMalloc(&h->d_lut, lut_size * sizeof(float)));
    CK(cudaMemcpy(h->d_lut, cfg->gamma_lut, lut_size * sizeof(float), cudaMemcpyHostToDevice));
    h->cfg.gamma_lut = nullptr;   // the caller's table is not kept
    CK(cudaMalloc(&h->d_flags, h->nres * sizeof(int)));
    CK(cudaHostAlloc(&h->h_flags, h->nres * sizeof(int), cudaHostAllocPortable));
    h->res_index.assign(h->nres, -1);
    h->ev_result.assign(h->nres, nullptr);
    for (int r = 0; r < h->nres; r++) CK(cudaEventCreateWithFlags(&h->ev_result[r], cudaEventDisableTiming));
    CK(cudaEventCreate(&h->ev_mark[0]));
    CK(cudaEventCreate(&h->ev_mark[1]));
#undef CK
    *out = h;
    return 0;
}

void hbcu_comb_detect_destroy(hbcu_comb_detect_t *h)
{
    if (h == nullptr) return;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (auto p : h->luma) if (p) cudaFree(p);
    for (auto e : h->ev_upload) if (e) cudaEventDestroy(e);
    for (auto e : h->ev_readers) if (e) cudaEventDestroy(e);
    for (auto e : h->ev_result) if (e) cudaEventDestroy(e);
    if (h->d_mask) cudaFree(h->d_mask);
    if (h->d_scored) cudaFree(h->d_scored);
    if (h->d_lut) cudaFree(h->d_lut);
    if (h->d_flags) cudaFree(h->d_flags);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    if (h->ev_mark[0]) cudaEventDestroy(h->ev_mark[0]);
    if (h->ev_mark[1]) cudaEventDestroy(h->ev_mark[1]);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_compute) cudaStreamDestroy(h->s_compute);
    delete h;
}

static int comb_upload(hbcu_comb_detect_t *h, int64_t index, const void *luma, int stride, cudaMemcpyKind kind)
{
    if (h == nullptr || luma == nullptr || index < 0) { set_error("comb_detect_upload: bad argument"); return -1; }
    HBCU_CHECK(cudaSetDevice(h->cfg.device));
    const int slot = (int)(index % h->slots);
    HBCU_CHECK(cudaStreamWaitEvent(h->s_h2d, h->ev_readers[slot], 0));
    HBCU_CHECK(cudaMemcpy2DAsync(h->luma[slot], (size_t)h->pitch * h->bps, luma, (size_t)stride,
                                 (size_t)h->cfg.width * h->bps, (size_t)h->cfg.height, kind, h->s_h2d));
    HBCU_CHECK(cudaEventRecord(h->ev_upload[slot], h->s_h2d));
    h->index[slot] = index;
    return 0;
}

int hbcu_comb_detect_upload(hbcu_comb_detect_t *h, int64_t index, const void *luma, int stride)
{
    return comb_upload(h, index, luma, stride, cudaMemcpyHostToDevice);
}

int hbcu_comb_detect_upload_device(hbcu_comb_detect_t *h, int64_t index, const void *dluma, int stride)
{
    return comb_upload(h, index, dluma, stride, cudaMemcpyDeviceToDevice);
}

int hbcu_comb_detect_upload_frame(hbcu_comb_detect_t *h, int64_t index, hbcu_frame_t *in)
{
    if (h == nullptr || in == nullptr || in->device != h->cfg.device || in->row_bytes[0] != h->cfg.width * h->bps || in->rows[0] != h->cfg.height)
    {
        set_error("comb_detect_upload_frame: bad argument or frame geometry");
        return -1;
    }
    HBCU_CHECK(cudaSetDevice(h->cfg.device));
    if (hbcu::frame_begin_read(in, h->s_h2d) != 0) return -1;
    if (comb_upload(h, index, in->plane[0], in->stride[0], cudaMemcpyDeviceToDevice) != 0) return -1;
    return hbcu::frame_end_read(in, h->s_h2d);
}

int hbcu_comb_detect_run(hbcu_comb_detect_t *h, int64_t prev, int64_t cur, int64_t next, int force)
{
    if (h == nullptr || prev < 0 || cur < 0 || next < 0) { set_error("comb_detect_run: bad argument"); return -1; }
    HBCU_CHECK(cudaSetDevice(h->cfg.device));
    const int64_t idx[3] = { prev, cur, next };
    const uint8_t *pl[3];
    for (int k = 0; k < 3; k++)
    {
        const int slot = (int)(idx[k] % h->slots);
        if (h->index[slot] != idx[k])
        {
            set_error("comb_detect_run: frame %lld is not resident", (long long)idx[k]);
            return -1;
        }
        HBCU_CHECK(cudaStreamWaitEvent(h->s_compute, h->ev_upload[slot], 0));
        pl[k] = h->luma[slot];
    }
    const int r = h->next_res;
    h->next_res = (h->next_res + 1) % h->nres;
    // the result slot is reused only after its previous verdict has been copied out (stream order)
    const hbcu_comb_detect_config_t &c = h->cfg;
    CombParams p;
    p.w = c.width; p.h = c.height; p.pitch = h->pitch;
    p.mode = c.mode; p.spatial_metric = c.spatial_metric; p.filter_mode = c.filter_mode;
    p.mthresh = c.motion_threshold; p.athresh = c.spatial_threshold;
    p.athresh_sq = c.spatial_threshold * c.spatial_threshold; p.athresh6 = 6 * c.spatial_threshold;
    p.c32min = c.comb32detect_min; p.c32max = c.comb32detect_max;
    p.g_mthresh = c.gamma_motion_threshold; p.g_athresh = c.gamma_spatial_threshold; p.g_athresh6 = c.gamma_spatial_threshold6;
    p.force = force;
    p.lut_size = 1 << c.depth;
    p.gamma_lut = h->d_lut;
    p.block_threshold = c.block_threshold; p.block_width = c.block_width; p.block_height = c.block_height;

    HBCU_CHECK(cudaMemsetAsync(h->d_flags + r, 0, sizeof(int), h->s_compute));
    dim3 blk(64, 4), grid((c.width + 63) / 64, (c.height + 3) / 4);
    const bool gamma = (c.mode & 1) != 0;
    const size_t lut_bytes = gamma ? (size_t)p.lut_size * sizeof(float) : 0;
    if (lut_bytes > 48 * 1024)
    {
        set_error("comb_detect: gamma table for depth %d does not fit shared memory", c.depth);
        return -1;
    }
    if (h->bps == 1)
    {
        if (gamma) comb_mask_kernel<uint8_t, true><<<grid, blk, lut_bytes, h->s_compute>>>(pl[0], pl[1], pl[2], h->d_mask, h->mpitch, p);
        else       comb_mask_kernel<uint8_t, false><<<grid, blk, 0, h->s_compute>>>(pl[0], pl[1], pl[2], h->d_mask, h->mpitch, p);
    }
    else
    {
        const uint16_t *a = (const uint16_t *)pl[0], *b = (const uint16_t *)pl[1], *d = (const uint16_t *)pl[2];
        if (gamma) comb_mask_kernel<uint16_t, true><<<grid, blk, lut_bytes, h->s_compute>>>(a, b, d, h->d_mask, h->mpitch, p);
        else       comb_mask_kernel<uint16_t, false><<<grid, blk, 0, h->s_compute>>>(a, b, d, h->d_mask, h->mpitch, p);
    }
    hbcu::count_launch();
    HBCU_CHECK(cudaGetLastError());
    const bool filtered = (c.mode & 2) != 0;
    const uint8_t *scored = h->d_mask;
    if (filtered)
    {
        dim3 fgrid((c.width + FT_W - 1) / FT_W, (c.height + FT_H - 1) / FT_H);
        comb_filter_kernel<<<fgrid, 256, 0, h->s_compute>>>(h->d_mask, h->d_scored, h->mpitch, c.width, c.height, c.filter_mode);
        hbcu::count_launch();
        HBCU_CHECK(cudaGetLastError());
        scored = h->d_scored;
    }
    // block grid: y = k*bh while y + bh <= height; x = j*bw while x < width - bw   (comb_detect.c:238-240)
    const int bw = c.block_width < c.width ? c.block_width : c.width;
    const int bh = c.block_height < c.height ? c.block_height : c.height;
    const int nby = c.height / bh;
    const int nbx = (c.width - bw + bw - 1) / bw;       // number of j with j*bw < width - bw
    if (nbx > 0 && nby > 0)
    {
        const int warps = nbx * nby;
        comb_score_kernel<<<(warps * 32 + 127) / 128, 128, 0, h->s_compute>>>(scored, h->mpitch, c.width, c.height, bw, bh, nbx, nby,
                                                                               c.block_threshold, filtered ? 1 : 0, h->d_flags + r);
        hbcu::count_launch();
        HBCU_CHECK(cudaGetLastError());
    }
    HBCU_CHECK(cudaMemcpyAsync(h->h_flags + r, h->d_flags + r, sizeof(int), cudaMemcpyDeviceToHost, h->s_compute));
    HBCU_CHECK(cudaEventRecord(h->ev_result[r], h->s_compute));
    h->res_index[r] = cur;
    // prev is never needed again once this run is done (the window moves forward)
    HBCU_CHECK(cudaEventRecord(h->ev_readers[(int)(prev % h->slots)], h->s_compute));
    return 0;
}

int hbcu_comb_detect_result(hbcu_comb_detect_t *h, int64_t cur, int *combed)
{
    if (h == nullptr || combed == nullptr) { set_error("comb_detect_result: bad argument"); return -1; }
    for (int r = 0; r < h->nres; r++)
    {
        if (h->res_index[r] == cur)
        {
            HBCU_CHECK(cudaEventSynchronize(h->ev_result[r]));
            const int f = h->h_flags[r];
            *combed = (f & 2) ? 2 : (f & 1) ? 1 : 0;     // check_combing_results, comb_detect.c:1029-1049
            return 0;
        }
    }
    set_error("comb_detect_result: no run pending for frame %lld", (long long)cur);
    return -1;
}

int hbcu_comb_detect_masks(hbcu_comb_detect_t *h, uint8_t *raw, uint8_t *scored)
{
    if (h == nullptr) { set_error("comb_detect_masks: null handle"); return -1; }
    HBCU_CHECK(cudaSetDevice(h->cfg.device));
    HBCU_CHECK(cudaStreamSynchronize(h->s_compute));
    const size_t w = h->cfg.width, hh = h->cfg.height;
    if (raw) HBCU_CHECK(cudaMemcpy2D(raw, w, h->d_mask, h->mpitch, w, hh, cudaMemcpyDeviceToHost));
    if (scored)
        HBCU_CHECK(cudaMemcpy2D(scored, w, (h->cfg.mode & 2) ? h->d_scored : h->d_mask, h->mpitch, w, hh, cudaMemcpyDeviceToHost));
    return 0;
}

int hbcu_comb_detect_sync(hbcu_comb_detect_t *h)
{
    if (h == nullptr) { set_error("comb_detect_sync: null handle"); return -1; }
    HBCU_CHECK(cudaSetDevice(h->cfg.device));
    HBCU_CHECK(cudaStreamSynchronize(h->s_h2d));
    HBCU_CHECK(cudaStreamSynchronize(h->s_compute));
    return 0;
}

int hbcu_comb_detect_mark(hbcu_comb_detect_t *h, int which)
{
    if (h == nullptr || which < 0 || which > 1) { set_error("comb_detect_mark: bad argument"); return -1; }
    HBCU_CHECK(cudaSetDevice(h->cfg.device));
    HBCU_CHECK(cudaEventRecord(h->ev_mark[which], h->s_compute));
    return 0;
}

int hbcu_comb_detect_elapsed_ms(hbcu_comb_detect_t *h, float *ms)
{
    if (h == nullptr || ms == nullptr) { set_error("comb_detect_elapsed_ms: bad argument"); return -1; }
    HBCU_CHECK(cudaEventSynchronize(h->ev_mark[1]));
    HBCU_CHECK(cudaEventElapsedTime(ms, h->ev_mark[0], h->ev_mark[1]));
    return 0;
}

}  // extern "C"
