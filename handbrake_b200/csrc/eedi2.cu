// eedi2.cu -- EEDI2 edge-directed interpolation of one field to a full frame, for sm_100a.
//
// Replaces the stage chain eedi2_interpolate_plane (reference libhb/templates/decomb_template.c:366-441)
// and the stages it calls (libhb/templates/eedi2_template.c, HandBrake's fork of tritical's EEDI2):
//   fill_half_height_buffer_plane :79-92     upscale_by_2 :101-111      build_edge_mask :122-195
//   dilate :207-247   erode :259-293   remove_small_gaps :308-342       calc_directions :358-525
//   filter_map :538-635   filter_dir_map :649-709   expand_dir_map :722-773
//   mark_directions_2x :787-858   filter_dir_map_2x :872-939   expand_dir_map_2x :953-1011
//   fill_gaps_2x :1025-1132   interpolate_lattice :1148-1335   post_process :1349-1378
//   gaussian_blur1 :1402-1527   gaussian_blur_sqrt2 :1540-1745   calc_derivatives :1756-1845   post_process_corner :1864-1900
// The reference runs the three planes on three CPU threads; here every stage is one kernel per
// plane on the decomb stream (planes are independent, stages are ordered by the stream).
//
// Bit-exactness notes (all reproduced on purpose):
//   * buffers keep the reference's LINEAR layout: one allocation per "frame buffer", planes
//     back to back, row pitch = hb_image_stride; several stages index x-1-u / x+1+u without
//     clamping, i.e. they read the neighbouring rows (or the neighbouring plane) through the
//     linear address, and interpolate_lattice reads dmskp[-1] / dmskp[width];
//   * the edge mask buffer is cleared only in its top half on each call (:132), the bottom half
//     carries the previous field's final mask: the buffers are persistent and fields are
//     processed in stream order;
//   * constants declared `pixel` wrap at 8 bits (nt13 = 138, nt19 = 182, nt4 = 200, nt7 = 94,
//     nt8 = 144) while `min == 7*nt` compares against the unwrapped 350;
//   * build_edge_mask is called with (magnitude, variance, laplacian) but declares
//     (mthresh, lthresh, vthresh): the two are swapped inside, as in the reference;
//   * mark_directions_2x compares dmskp[x+1] with dmskpn[x-1] (:835);
//   * interpolate_lattice rewrites the direction map in place and pixel x reads the rewritten
//     dmskp[x-1]: a parallel pass computes both possible outcomes per pixel, a per-row
//     sequential pass resolves the chain;
//   * float/double expressions ((float)sum/(float)count, j*step+0.5) are evaluated with explicit
//     round-to-nearest operations, no FMA contraction.
// Memory outside the planes (before the first plane, after the last) reads as zero, which is
// what the zero-initialised shim buffers of the oracle give (libhb proper leaves it undefined).
#include "hbcu_common.h"
#include "eedi2.cuh"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace hbcu {

namespace {

constexpr int kLeadSlack = 1024;          // zero bytes in front of every frame buffer
constexpr int kTailRows  = 4;             // zero rows (of the luma pitch) behind every frame buffer

enum { SRCPF = 0, MSKPF = 1, TMPPF = 2, DSTPF = 3 };                       // decomb.c:64-68
enum { DST2PF = 0, TMP2PF2 = 1, MSK2PF = 2, TMP2PF = 3, DST2MPF = 4 };     // decomb.c:70-74

struct Lim { int v[33]; };                // eedi_limlut, pixel-typed values widened to int

__device__ __forceinline__ int iabs(int a) { return a < 0 ? -a : a; }

// eedi2_sort_metrics (eedi2.c:66-80)
__device__ __forceinline__ void sort_metrics(int *order, int length)
{
    for (int i = 1; i < length; ++i)
    {
        int j = i;
        const int temp = order[j];
        while (j > 0 && order[j - 1] > temp)
        {
            order[j] = order[j - 1];
            --j;
        }
        order[j] = temp;
    }
}

__device__ __forceinline__ int median_of(const int *order, int n)
{
    return (n & 1) ? order[n >> 1] : (order[(n - 1) >> 1] + order[n >> 1] + 1) >> 1;
}

// Branch-free variants for the per-sample filters (the insertion sort above walks a dynamically indexed array --
// local memory -- with data-dependent loops that diverge per lane).  Candidates sit in FIXED slots; a missing candidate is
// kNone, which sorts behind every real value, so after the network the n real values are o[0 .. n-1] in ascending order,
// exactly what eedi2_sort_metrics leaves in order[0 .. n-1] (equal values are indistinguishable).
constexpr int kNone = 0x3fffffff;      // above every sample and direction value; differences with it cannot overflow

template <int M>
__device__ __forceinline__ void sort_net(int (&o)[M])
{
#pragma unroll
    for (int i = 0; i < M - 1; i++)
#pragma unroll
        for (int j = 0; j + 1 < M - i; j++)
        {
            const int lo = min(o[j], o[j + 1]), hi = max(o[j], o[j + 1]);
            o[j] = lo;
            o[j + 1] = hi;
        }
}

template <int M>
__device__ __forceinline__ int pick_net(const int (&o)[M], int idx)
{
    int r = o[0];
#pragma unroll
    for (int i = 1; i < M; i++) r = idx == i ? o[i] : r;
    return r;
}

template <int M>
__device__ __forceinline__ int median_net(const int (&o)[M], int n)
{
    const int a = pick_net(o, (n - 1) >> 1), b = pick_net(o, n >> 1);
    return (n & 1) ? b : (a + b + 1) >> 1;
}

// count and sum of the first n (sorted) values within `l` of `mid`
template <int M>
__device__ __forceinline__ void near_net(const int (&o)[M], int n, int mid, int l, int &count, int &sum)
{
    count = 0;
    sum = 0;
#pragma unroll
    for (int i = 0; i < M; i++)
    {
        const bool in = i < n && iabs(o[i] - mid) <= l;
        count += in ? 1 : 0;
        sum += in ? o[i] : 0;
    }
}

// (int)(((float)(sum + mid) / (float)(count + 1)) + 0.5f)
__device__ __forceinline__ int avg_round(int sum, int mid, int count)
{
    return (int)__fadd_rn(__fdiv_rn((float)(sum + mid), (float)(count + 1)), 0.5f);
}

template <typename PIX>
struct K            // per-depth constants
{
    int depth, shift, shift2, peak, neutral;
    __host__ __device__ explicit K(int d) : depth(d), shift(d - 8), shift2(2 + d - 8), peak((1 << d) - 1), neutral(1 << (d - 1)) {}
};

// ---------------------------------------------------------------------------------------------
// 16 bytes (8 or 16 samples) per thread.  Most stencil stages only do work where the edge mask is set, and the mask is
// sparse (edges): a thread first looks at the 16-byte vectors of the gating planes and, when no sample of its group
// can be active, finishes with one vector copy (or nothing); only groups that contain an active sample run the
// per-sample code -- which is the reference's arithmetic, untouched.  Rows are 16-byte aligned (64-byte strides).
// ---------------------------------------------------------------------------------------------
template <typename PIX> struct Vec { static constexpr int N = 16 / (int)sizeof(PIX); };
__device__ __forceinline__ uint4 ld16(const void *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void st16(void *p, const uint4 &v) { *reinterpret_cast<uint4 *>(p) = v; }
template <typename PIX> __device__ __forceinline__ uint32_t splat(int v)
{
    return sizeof(PIX) == 1 ? (uint32_t)(v & 0xff) * 0x01010101u : (uint32_t)(v & 0xffff) * 0x00010001u;
}
template <typename PIX> __device__ __forceinline__ uint32_t cmpeq(uint32_t a, uint32_t b) { return sizeof(PIX) == 1 ? __vcmpeq4(a, b) : __vcmpeq2(a, b); }
template <typename PIX> __device__ __forceinline__ bool any_eq(const uint4 &v, int val)
{
    const uint32_t p = splat<PIX>(val);
    return (cmpeq<PIX>(v.x, p) | cmpeq<PIX>(v.y, p) | cmpeq<PIX>(v.z, p) | cmpeq<PIX>(v.w, p)) != 0u;
}
template <typename PIX> __device__ __forceinline__ bool any_ne(const uint4 &v, int val)
{
    const uint32_t p = splat<PIX>(val);
    return (cmpeq<PIX>(v.x, p) & cmpeq<PIX>(v.y, p) & cmpeq<PIX>(v.z, p) & cmpeq<PIX>(v.w, p)) != 0xffffffffu;
}

// bit i set <=> sample i of the 16-byte vector equals val
template <typename PIX> __device__ __forceinline__ uint32_t eq_bits(const uint4 &v, int val)
{
    const uint32_t p = splat<PIX>(val);
    const uint32_t m[4] = { cmpeq<PIX>(v.x, p), cmpeq<PIX>(v.y, p), cmpeq<PIX>(v.z, p), cmpeq<PIX>(v.w, p) };
    uint32_t bits = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        if (sizeof(PIX) == 2)
            bits |= ((m[k] & 1u) | ((m[k] >> 15) & 2u)) << (2 * k);
        else
        {
            uint32_t t = m[k] & 0x01010101u;
            t = (t | (t >> 7) | (t >> 14) | (t >> 21)) & 0xfu;
            bits |= t << (4 * k);
        }
    }
    return bits;
}
// bits of the samples x0 .. x0+N-1 that lie in [lo, hi]
template <int N> __device__ __forceinline__ uint32_t range_bits(int x0, int lo, int hi)
{
    const int a = max(lo - x0, 0), b = min(hi - x0, N - 1);
    if (a > b) return 0u;
    return (0xffffffffu >> (31 - b)) & (0xffffffffu << a);
}

// The active samples of a stage cluster along edges: a vertical edge puts them into a few lanes of a warp, a horizontal
// one into every lane of the few warps that own that row.  Handing every thread the samples of its own group serialises
// the first case (measured: calc_directions 3.5x slower), pooling per warp serialises the second (a warp owning a fully
// active 256-sample row segment needs 8 rounds while the rest of the GPU idles: 218 us vs 66 us, profiles/r01m).  So
// the whole CTA (256 threads = 4 rows x 512 samples at 16 bit) pools its active samples in shared memory and deals them
// out one per thread.  Every thread of the CTA must call this (no early returns before it).
constexpr int kStageThreads = 256;             // stage kernels run blocks of 64 x 4 threads
template <int N, typename F>
__device__ __forceinline__ void block_deal(uint32_t bits, uint16_t *list, int *warp_sums, F fn)
{
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int cnt = __popc(bits);
    int pre = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const int t = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += t;
    }
    if (lane == 31) warp_sums[warp] = pre;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kStageThreads / 32; w++)
    {
        const int v = warp_sums[w];
        if (w < warp) base += v;
        total += v;
    }
    if (total == 0) return;                    // uniform over the CTA
    pre += base - cnt;
    while (bits)
    {
        const int j = __ffs(bits) - 1;
        bits &= bits - 1;
        list[pre++] = (uint16_t)(tid * N + j);
    }
    __syncthreads();
    for (int k = tid; k < total; k += kStageThreads)
    {
        const int e = list[k], t = e / N, j = e - t * N;
        // owner thread t = (tx, ty) of the 64 x 4 block -> its group's first sample and its row index
        fn((int)((blockIdx.x * blockDim.x + (t % blockDim.x)) * N + j), (int)(blockIdx.y * blockDim.y + t / blockDim.x));
    }
}
#define STAGE_LIST(PIX) __shared__ uint16_t list[kStageThreads * Vec<PIX>::N]; __shared__ int warp_sums[kStageThreads / 32]

// ---------------------------------------------------------------------------------------------
// copies
// ---------------------------------------------------------------------------------------------
// The copy-like stages move 16 bytes per thread.  Every row starts 16-byte aligned: plane bases are 64-byte
// aligned (own allocations / 64-byte plane offsets) and the pitch is the reference's 64-byte-rounded stride.
template <typename PIX>
__global__ void __launch_bounds__(256) k_fill_half(const PIX *__restrict__ src, PIX *__restrict__ dst, int pitch, int rows)
{
    // row r of the field buffer = row 2r of src (src already points at the field's first line); whole pitch
    constexpr int EPC = 16 / (int)sizeof(PIX);
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * EPC, r = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < pitch && r < rows)
        *reinterpret_cast<uint4 *>(dst + (size_t)r * pitch + x) = *reinterpret_cast<const uint4 *>(src + (size_t)2 * r * pitch + x);
}

template <typename PIX>
__global__ void __launch_bounds__(256) k_upscale2(const PIX *__restrict__ src, PIX *__restrict__ dst, int pitch, int rows)
{
    constexpr int EPC = 16 / (int)sizeof(PIX);
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * EPC, r = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < pitch && r < rows)
    {
        const uint4 v = *reinterpret_cast<const uint4 *>(src + (size_t)r * pitch + x);
        *reinterpret_cast<uint4 *>(dst + (size_t)(2 * r) * pitch + x) = v;
        *reinterpret_cast<uint4 *>(dst + (size_t)(2 * r + 1) * pitch + x) = v;
    }
}

// bit_blit over `width` samples per row (not the stride): whole 16-byte chunks as vectors, the ragged tail by element
template <typename PIX>
__global__ void __launch_bounds__(256) k_blit(const PIX *__restrict__ src, PIX *__restrict__ dst, int pitch, int width, int rows)
{
    constexpr int EPC = 16 / (int)sizeof(PIX);
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * EPC, r = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || r >= rows) return;
    const PIX *s = src + (size_t)r * pitch + x;
    PIX *d = dst + (size_t)r * pitch + x;
    if (x + EPC <= width)
        *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(s);
    else
        for (int i = 0; x + i < width; i++) d[i] = s[i];
}

// n is a multiple of the pitch (whole rows), so of 16 bytes
template <typename PIX>
__global__ void __launch_bounds__(256) k_fill(PIX *__restrict__ dst, size_t n, int value)
{
    constexpr int EPC = 16 / (int)sizeof(PIX);
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * EPC;
    if (i >= n) return;
    const uint32_t w = sizeof(PIX) == 1 ? (uint32_t)(value & 0xff) * 0x01010101u : (uint32_t)(value & 0xffff) * 0x00010001u;
    *reinterpret_cast<uint4 *>(dst + i) = make_uint4(w, w, w, w);
}

// ---------------------------------------------------------------------------------------------
// edge mask (:122-195).  NB parameter names follow the callee: lthresh receives the variance
// setting and vthresh the laplacian setting.
// ---------------------------------------------------------------------------------------------
template <typename PIX>
__device__ __forceinline__ void edge_mask_eval(PIX *__restrict__ dstp, size_t o, const K<PIX> &k, int ten, int mthresh10, int lthresh, int vthresh81,
                                               int pm, int pc, int pp, int cm, int cc, int cp, int nm, int nc, int np)
{
    if ((iabs(pc - cc) < ten && iabs(cc - nc) < ten && iabs(pc - nc) < ten) ||
        (iabs(pm - cm) < ten && iabs(cm - nm) < ten && iabs(pm - nm) < ten &&
         iabs(pp - cp) < ten && iabs(cp - np) < ten && iabs(pp - np) < ten))
        return;
    const int s = k.shift;
    const int sum = (pm + pc + pp + cm + cc + cp + nm + nc + np) >> s;
    const int sumsq = (pm >> s) * (pm >> s) + (pc >> s) * (pc >> s) + (pp >> s) * (pp >> s) +
                      (cm >> s) * (cm >> s) + (cc >> s) * (cc >> s) + (cp >> s) * (cp >> s) +
                      (nm >> s) * (nm >> s) + (nc >> s) * (nc >> s) + (np >> s) * (np >> s);
    if (9 * sumsq - sum * sum < vthresh81) return;
    const int Ix = (cp - cm) >> s;
    const int Iy = max(max(iabs(pc - nc), iabs(pc - cc)), iabs(cc - nc)) >> s;
    if (Ix * Ix + Iy * Iy >= mthresh10)
    {
        dstp[o] = (PIX)k.peak;
        return;
    }
    const int Ixx = (cm - 2 * cc + cp) >> s;
    const int Iyy = (pc - 2 * cc + nc) >> s;
    if (iabs(Ixx) + iabs(Iyy) >= lthresh) dstp[o] = (PIX)k.peak;
}

// one thread = N samples of one row; the three source rows are read once into a register window of N + 2 samples
template <typename PIX>
__global__ void __launch_bounds__(256) k_edge_mask(PIX *__restrict__ dstp, const PIX *__restrict__ srcp, int pitch, int width, int height,
                                                   int mthresh10, int lthresh, int vthresh81, int depth)
{
    constexpr int N = Vec<PIX>::N;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * N, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= width || y < 1 || y >= height - 1) return;
    const K<PIX> k(depth);
    const int ten = (int)(PIX)(10 << k.shift);
    const PIX *c = srcp + (size_t)y * pitch, *p = c - pitch, *n = c + pitch;
    if (x0 + N <= width)
    {
        int w[3][N + 2];
        const PIX *rows[3] = { p, c, n };
#pragma unroll
        for (int r = 0; r < 3; r++)
        {
            const uint4 v = ld16(rows[r] + x0);
            const uint32_t u[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int i = 0; i < N; i++)
                w[r][i + 1] = sizeof(PIX) == 1 ? (int)((u[i >> 2] >> (8 * (i & 3))) & 0xffu) : (int)((u[i >> 1] >> (16 * (i & 1))) & 0xffffu);
            w[r][0] = x0 >= 1 ? (int)rows[r][x0 - 1] : 0;
            w[r][N + 1] = x0 + N < width ? (int)rows[r][x0 + N] : 0;
        }
#pragma unroll
        for (int i = 0; i < N; i++)
        {
            const int x = x0 + i;
            if (x < 1 || x >= width - 1) continue;
            edge_mask_eval<PIX>(dstp, (size_t)y * pitch + x, k, ten, mthresh10, lthresh, vthresh81,
                                w[0][i], w[0][i + 1], w[0][i + 2], w[1][i], w[1][i + 1], w[1][i + 2], w[2][i], w[2][i + 1], w[2][i + 2]);
        }
        return;
    }
    for (int i = 0; i < N && x0 + i < width; i++)
    {
        const int x = x0 + i;
        if (x < 1 || x >= width - 1) continue;
        edge_mask_eval<PIX>(dstp, (size_t)y * pitch + x, k, ten, mthresh10, lthresh, vthresh81,
                            p[x - 1], p[x], p[x + 1], c[x - 1], c[x], c[x + 1], n[x - 1], n[x], n[x + 1]);
    }
}

// erode (:259-293) / dilate (:207-247): dst = copy of src over `width`, interior rule applied
template <typename PIX, bool DILATE>
__device__ __forceinline__ int morph_px(const PIX *__restrict__ mskp, int pitch, int width, int height, int str, int peak, int x, int y)
{
    const PIX *c = mskp + (size_t)y * pitch;
    int v = c[x];
    if (x >= 1 && x < width - 1 && y >= 1 && y < height - 1 && (DILATE ? v == 0 : v == peak))
    {
        const PIX *p = c - pitch, *n = c + pitch;
        int count = 0;
        count += p[x - 1] == peak; count += p[x] == peak; count += p[x + 1] == peak;
        count += c[x - 1] == peak; count += c[x + 1] == peak;
        count += n[x - 1] == peak; count += n[x] == peak; count += n[x + 1] == peak;
        if (DILATE) { if (count >= str) v = peak; }
        else        { if (count < str) v = 0; }
    }
    return v;
}

template <typename PIX, bool DILATE>
__global__ void __launch_bounds__(256) k_morph(const PIX *__restrict__ mskp, PIX *__restrict__ dstp, int pitch, int width, int height, int str, int depth)
{
    constexpr int N = Vec<PIX>::N;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * N, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= width || y >= height) return;
    const int peak = (1 << depth) - 1;
    PIX *out = dstp + (size_t)y * pitch;
    if (x0 + N <= width)
    {
        const PIX *c = mskp + (size_t)y * pitch + x0;
        const uint4 cv = ld16(c);
        bool any = false;
        if (y >= 1 && y < height - 1)
        {
            if (DILATE)
            {
                // a clear sample is set only if set samples surround it: nothing set in the 3 x (N+2) neighbourhood -> copy
                any = any_eq<PIX>(cv, peak) || any_eq<PIX>(ld16(c - pitch), peak) || any_eq<PIX>(ld16(c + pitch), peak);
                if (x0 >= 1)        any = any || c[-1] == peak || c[-1 - pitch] == peak || c[-1 + pitch] == peak;
                if (x0 + N < width) any = any || c[N] == peak || c[N - pitch] == peak || c[N + pitch] == peak;
            }
            else
                any = any_eq<PIX>(cv, peak);             // only set samples can be eroded
        }
        if (!any)
        {
            st16(out + x0, cv);
            return;
        }
    }
    for (int i = 0; i < N && x0 + i < width; i++)
        out[x0 + i] = (PIX)morph_px<PIX, DILATE>(mskp, pitch, width, height, str, peak, x0 + i, y);
}

// remove_small_gaps (:308-342)
template <typename PIX>
__device__ __forceinline__ int gaps_px(const PIX *__restrict__ mskp, int pitch, int width, int height, int peak, int x, int y)
{
    const PIX *m = mskp + (size_t)y * pitch;
    int v = m[x];
    if (y >= 1 && y < height - 1 && x >= 3 && x < width - 3)
    {
        if (m[x])
        {
            if (!(m[x - 3] || m[x - 2] || m[x - 1] || m[x + 1] || m[x + 2] || m[x + 3])) v = 0;
        }
        else
        {
            if ((m[x + 1] && (m[x - 1] || m[x - 2] || m[x - 3])) || (m[x + 2] && (m[x - 1] || m[x - 2])) || (m[x + 3] && m[x - 1]))
                v = peak;
        }
    }
    return v;
}

template <typename PIX>
__global__ void __launch_bounds__(256) k_gaps(const PIX *__restrict__ mskp, PIX *__restrict__ dstp, int pitch, int width, int height, int depth)
{
    constexpr int N = Vec<PIX>::N;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * N, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= width || y >= height) return;
    const int peak = (1 << depth) - 1;
    PIX *out = dstp + (size_t)y * pitch;
    if (x0 + N <= width)
    {
        const PIX *m = mskp + (size_t)y * pitch + x0;
        const uint4 mv = ld16(m);
        // nothing set in the group: samples change only if set samples lie within 3 to both sides -- one non-zero
        // neighbour within reach is enough to look closer
        bool any = any_ne<PIX>(mv, 0);
        if (!any && y >= 1 && y < height - 1)
        {
            for (int j = 1; j <= 3; j++)
            {
                if (x0 - j >= 0)        any = any || m[-j] != 0;
                if (x0 + N - 1 + j < width) any = any || m[N - 1 + j] != 0;
            }
        }
        if (!any)
        {
            st16(out + x0, mv);
            return;
        }
    }
    for (int i = 0; i < N && x0 + i < width; i++)
        out[x0 + i] = (PIX)gaps_px<PIX>(mskp, pitch, width, height, peak, x0 + i, y);
}

// ---------------------------------------------------------------------------------------------
// calc_directions (:358-525): dst pre-filled with peak (whole pitch); one thread per pixel
// ---------------------------------------------------------------------------------------------
template <typename PIX>
__device__ __forceinline__ void calc_directions_px(int plane, const PIX *__restrict__ mskp, const PIX *__restrict__ srcp, PIX *__restrict__ dstp,
                                                   int pitch, int width, int height, int maxd, int nt, int depth, const K<PIX> &k, const Lim &lim,
                                                   int x, int y)
{
    if (x < 1 || x >= width - 1) return;
    const PIX *mc = mskp + (size_t)y * pitch, *mp = mc - pitch, *mn = mc + pitch;
    if (mc[x] != k.peak || (mc[x - 1] != k.peak && mc[x + 1] != k.peak)) return;
    const PIX *sc = srcp + (size_t)y * pitch, *sp = sc - pitch, *sn = sc + pitch, *s2p = sc - 2 * pitch, *s2n = sc + 2 * pitch;
    const int nt13 = (int)(PIX)((nt << (depth - 8)) * 13);
    const int nt19 = (int)(PIX)((nt << (depth - 8)) * 19);
    const int maxdt = plane == 0 ? maxd : (maxd >> 1);
    const int startu = max(-x + 1, -maxdt), stopu = min(width - 2 - x, maxdt);
    const int base = iabs((int)sc[x] - (int)sn[x]) + iabs((int)sc[x] - (int)sp[x]);
    int minb = min(nt13, base * 6), mina = min(nt19, base * 9);
    int minc = mina, mind = minb, mine = minb;
    int dira = -5000, dirb = -5000, dirc = -5000, dird = -5000, dire = -5000;
    const int c0 = sc[x - 1], c1 = sc[x], c2 = sc[x + 1];
    const int p0 = sp[x - 1], p1 = sp[x], p2 = sp[x + 1];
    const int n0 = sn[x - 1], n1 = sn[x], n2 = sn[x + 1];
    for (int u = startu; u <= stopu; ++u)
    {
        if (!(y == 1 || mp[x - 1 + u] == k.peak || mp[x + u] == k.peak || mp[x + 1 + u] == k.peak)) continue;
        if (!(y == height - 2 || mn[x - 1 - u] == k.peak || mn[x - u] == k.peak || mn[x + 1 - u] == k.peak)) continue;
        const int diffsn = iabs(c0 - (int)sn[x - 1 - u]) + iabs(c1 - (int)sn[x - u]) + iabs(c2 - (int)sn[x + 1 - u]);
        const int diffsp = iabs(c0 - (int)sp[x - 1 + u]) + iabs(c1 - (int)sp[x + u]) + iabs(c2 - (int)sp[x + 1 + u]);
        const int diffps = iabs(p0 - (int)sc[x - 1 - u]) + iabs(p1 - (int)sc[x - u]) + iabs(p2 - (int)sc[x + 1 - u]);
        const int diffns = iabs(n0 - (int)sc[x - 1 + u]) + iabs(n1 - (int)sc[x + u]) + iabs(n2 - (int)sc[x + 1 + u]);
        const int diff = diffsn + diffsp + diffps + diffns;
        int diffd = diffsp + diffns, diffe = diffsn + diffps;
        if (diff < minb) { dirb = u; minb = diff; }
        if (y > 1)
        {
            const int diff2pp = iabs((int)s2p[x - 1] - (int)sp[x - 1 - u]) + iabs((int)s2p[x] - (int)sp[x - u]) + iabs((int)s2p[x + 1] - (int)sp[x + 1 - u]);
            const int diffp2p = iabs(p0 - (int)s2p[x - 1 + u]) + iabs(p1 - (int)s2p[x + u]) + iabs(p2 - (int)s2p[x + 1 + u]);
            const int diffa = diff + diff2pp + diffp2p;
            diffd += diffp2p;
            diffe += diff2pp;
            if (diffa < mina) { dira = u; mina = diffa; }
        }
        if (y < height - 2)
        {
            const int diff2nn = iabs((int)s2n[x - 1] - (int)sn[x - 1 + u]) + iabs((int)s2n[x] - (int)sn[x + u]) + iabs((int)s2n[x + 1] - (int)sn[x + 1 + u]);
            const int diffn2n = iabs(n0 - (int)s2n[x - 1 - u]) + iabs(n1 - (int)s2n[x - u]) + iabs(n2 - (int)s2n[x + 1 - u]);
            const int diffc = diff + diff2nn + diffn2n;
            diffd += diff2nn;
            diffe += diffn2n;
            if (diffc < minc) { dirc = u; minc = diffc; }
        }
        if (diffd < mind) { dird = u; mind = diffd; }
        if (diffe < mine) { dire = u; mine = diffe; }
    }
    int order[5] = { dira != -5000 ? dira : kNone, dirb != -5000 ? dirb : kNone, dirc != -5000 ? dirc : kNone,
                     dird != -5000 ? dird : kNone, dire != -5000 ? dire : kNone };
    const int n = (dira != -5000) + (dirb != -5000) + (dirc != -5000) + (dird != -5000) + (dire != -5000);
    int out = k.neutral;
    if (n > 1)
    {
        sort_net(order);
        const int mid = median_net(order, n);
        const int tlim = max(lim.v[iabs(mid)] >> 2, 2);
        int sum, count;
        near_net(order, n, mid, tlim, count, sum);
        if (count > 1) out = k.neutral + ((int)__fdiv_rn((float)sum, (float)count)) * (1 << k.shift2);
    }
    dstp[(size_t)y * pitch + x] = (PIX)out;
}

template <typename PIX>
__global__ void __launch_bounds__(256) k_calc_directions(int plane, const PIX *__restrict__ mskp, const PIX *__restrict__ srcp, PIX *__restrict__ dstp,
                                                         int pitch, int width, int height, int maxd, int nt, int depth, Lim lim)
{
    constexpr int N = Vec<PIX>::N;
    STAGE_LIST(PIX);
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * N, y = blockIdx.y * blockDim.y + threadIdx.y;
    const K<PIX> k(depth);
    uint32_t bits = 0;
    if (y < 1 || y >= height - 1) bits = 0;
    else if (x0 + N <= width)  bits = eq_bits<PIX>(ld16(mskp + (size_t)y * pitch + x0), k.peak) & range_bits<N>(x0, 1, width - 2);
    else if (x0 < width)  bits = range_bits<N>(x0, 1, width - 2);
    // (splitting one sample's 49-step search over 8 lanes was measured slower: 163 vs 119 us -- the densely active CTAs
    //  that set the kernel's duration need 8x the rounds and 30 shuffles per round)
    block_deal<N>(bits, list, warp_sums, [&](int x, int yy) {
        calc_directions_px<PIX>(plane, mskp, srcp, dstp, pitch, width, height, maxd, nt, depth, k, lim, x, yy);   // dst keeps its peak fill elsewhere
    });
}

// ---------------------------------------------------------------------------------------------
// filter_dir_map (:649-709) and expand_dir_map (:722-773); also their 2x variants (:872-1011)
//   TWOX = false: rows 1..h-2, neighbours at +-pitch
//   TWOX = true : rows y = 2-field, +2 ...; neighbours at +-2*pitch, gated by y>1 / y<h-2; the
//                 mask test uses rows y-1 and y+1 of the (line-doubled) edge mask
// ---------------------------------------------------------------------------------------------
template <typename PIX, bool EXPAND, bool TWOX>
__device__ __forceinline__ int dir_map_px(const PIX *__restrict__ mskp, const PIX *__restrict__ dmskp,
                                          int pitch, int width, int height, int field, const K<PIX> &k, const Lim &lim, int x, int y)
{
    const PIX *dc = dmskp + (size_t)y * pitch;
    int v = dc[x];                                   // bit_blit: dst starts as a copy of dmsk (over width)
    const bool row_ok = TWOX ? (y >= 2 - field && y < height - 1 && ((y - (2 - field)) & 1) == 0) : (y >= 1 && y < height - 1);
    if (row_ok && x >= 1 && x < width - 1)
    {
        bool active;
        if (TWOX)
        {
            const PIX *m0 = mskp + (size_t)(y - 1) * pitch, *m1 = mskp + (size_t)(y + 1) * pitch;
            active = !(m0[x] != k.peak && m1[x] != k.peak);
        }
        else
        {
            active = mskp[(size_t)y * pitch + x] == k.peak;
        }
        if (EXPAND) active = active && dc[x] == k.peak;
        if (active)
        {
            const int step = TWOX ? 2 * pitch : pitch;
            const PIX *dp = dc - step, *dn = dc + step;
            const bool up = !TWOX || y > 1, down = !TWOX || y < height - 2;
            // nine fixed candidate slots (three rows x three columns); kNone where the reference skips the sample
            // (a row that does not exist is never read: `take` guards the load)
            auto cand = [&](bool take, const PIX *row, int i) { const int value = take ? (int)row[i] : k.peak; return value != k.peak ? value : kNone; };
            int order[9] = { cand(up, dp, x - 1), cand(up, dp, x), cand(up, dp, x + 1),
                             cand(true, dc, x - 1), cand(!EXPAND, dc, x), cand(true, dc, x + 1),
                             cand(down, dn, x - 1), cand(down, dn, x), cand(down, dn, x + 1) };
            int u = 0;
#pragma unroll
            for (int i = 0; i < 9; i++) u += order[i] != kNone;
            if (EXPAND)
            {
                if (u >= 5)
                {
                    sort_net(order);
                    const int mid = median_net(order, u);
                    const int l = lim.v[iabs(mid - k.neutral) >> k.shift2];
                    int sum, count;
                    near_net(order, u, mid, l, count, sum);
                    if (count >= 5) v = (int)(PIX)avg_round(sum, mid, count);
                }
            }
            else
            {
                if (u < 4)
                    v = k.peak;
                else
                {
                    sort_net(order);
                    const int mid = median_net(order, u);
                    const int l = lim.v[iabs(mid - k.neutral) >> k.shift2];
                    int sum, count;
                    near_net(order, u, mid, l, count, sum);
                    if (count < 4 || (count < 5 && dc[x] == k.peak)) v = k.peak;
                    else v = (int)(PIX)avg_round(sum, mid, count);
                }
            }
        }
    }
    return v;
}

template <typename PIX, bool EXPAND, bool TWOX>
__global__ void __launch_bounds__(256) k_dir_map(const PIX *__restrict__ mskp, const PIX *__restrict__ dmskp, PIX *__restrict__ dstp,
                                                 int pitch, int width, int height, int field, int depth, Lim lim)
{
    constexpr int N = Vec<PIX>::N;
    STAGE_LIST(PIX);
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * N, y = blockIdx.y * blockDim.y + threadIdx.y;
    const K<PIX> k(depth);
    PIX *out = dstp + (size_t)y * pitch;
    uint32_t bits = 0;
    if (x0 < width && y < height)
    {
        const bool row_ok = TWOX ? (y >= 2 - field && y < height - 1 && ((y - (2 - field)) & 1) == 0) : (y >= 1 && y < height - 1);
        if (x0 + N <= width)
        {
            const uint4 dv = ld16(dmskp + (size_t)y * pitch + x0);
            st16(out + x0, dv);                               // bit_blit: dst = dmsk; active samples are overwritten below
            if (row_ok)
            {
                if (TWOX) bits = eq_bits<PIX>(ld16(mskp + (size_t)(y - 1) * pitch + x0), k.peak) | eq_bits<PIX>(ld16(mskp + (size_t)(y + 1) * pitch + x0), k.peak);
                else      bits = eq_bits<PIX>(ld16(mskp + (size_t)y * pitch + x0), k.peak);
                if (EXPAND) bits &= eq_bits<PIX>(dv, k.peak);
                bits &= range_bits<N>(x0, 1, width - 2);
            }
        }
        else
            bits = range_bits<N>(x0, 0, width - 1);           // ragged row end: every sample goes through the per-sample code
    }
    block_deal<N>(bits, list, warp_sums, [&](int x, int yy) {
        dstp[(size_t)yy * pitch + x] = (PIX)dir_map_px<PIX, EXPAND, TWOX>(mskp, dmskp, pitch, width, height, field, k, lim, x, yy);
    });
}

// filter_map (:538-635)
template <typename PIX>
__device__ __forceinline__ int filter_map_px(const PIX *__restrict__ mskp, const PIX *__restrict__ dmskp,
                                             int pitch, int width, int height, const K<PIX> &k, int x, int y)
{
    const PIX *dc = dmskp + (size_t)y * pitch;
    int v = dc[x];
    if (y >= 1 && y < height - 1 && x >= 1 && x < width - 1 && !(dc[x] == k.peak || mskp[(size_t)y * pitch + x] != k.peak))
    {
        const PIX *dp = dc - pitch, *dn = dc + pitch;
        const int shift = k.shift2;                       // `shift` in the callee is 2 + (depth - 8)
        const int twelve = 12 << shift;
        const int cur = dc[x];
        int dir = (cur - k.neutral) >> 2;
        const int lm = max(iabs(dir) * 2, twelve);
        dir >>= shift;
        int ict = 0, icb = 0;
#define EEDI_BAD(row, j) ((iabs((int)(row)[x + (j)] - cur) > lm && (row)[x + (j)] != k.peak))
        if (dir < 0)
        {
            const int dirt = max(-x, dir);
            for (int j = dirt; j <= 0; ++j)
                if (EEDI_BAD(dp, j) || (dc[x + j] == k.peak && dp[x + j] == k.peak) || EEDI_BAD(dc, j)) { ict = 1; break; }
        }
        else
        {
            const int dirt = min(width - x - 1, dir);
            for (int j = 0; j <= dirt; ++j)
                if (EEDI_BAD(dp, j) || (dc[x + j] == k.peak && dp[x + j] == k.peak) || EEDI_BAD(dc, j)) { ict = 1; break; }
        }
        if (ict)
        {
            if (dir < 0)
            {
                const int dirt = min(width - x - 1, iabs(dir));
                for (int j = 0; j <= dirt; ++j)
                    if (EEDI_BAD(dn, j) || (dn[x + j] == k.peak && dc[x + j] == k.peak) || EEDI_BAD(dc, j)) { icb = 1; break; }
            }
            else
            {
                const int dirt = max(-x, -dir);
                for (int j = dirt; j <= 0; ++j)
                    if (EEDI_BAD(dn, j) || (dn[x + j] == k.peak && dc[x + j] == k.peak) || EEDI_BAD(dc, j)) { icb = 1; break; }
            }
            if (icb) v = k.peak;
        }
#undef EEDI_BAD
    }
    return v;
}

template <typename PIX>
__global__ void __launch_bounds__(256) k_filter_map(const PIX *__restrict__ mskp, const PIX *__restrict__ dmskp, PIX *__restrict__ dstp,
                                                    int pitch, int width, int height, int depth)
{
    constexpr int N = Vec<PIX>::N;
    STAGE_LIST(PIX);
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * N, y = blockIdx.y * blockDim.y + threadIdx.y;
    const K<PIX> k(depth);
    PIX *out = dstp + (size_t)y * pitch;
    uint32_t bits = 0;
    if (x0 < width && y < height)
    {
        if (x0 + N <= width)
        {
            const uint4 dv = ld16(dmskp + (size_t)y * pitch + x0);
            st16(out + x0, dv);
            // a sample is looked at only where the edge mask is set and a direction exists
            if (y >= 1 && y < height - 1)
                bits = eq_bits<PIX>(ld16(mskp + (size_t)y * pitch + x0), k.peak) & ~eq_bits<PIX>(dv, k.peak) & range_bits<N>(x0, 1, width - 2);
        }
        else
            bits = range_bits<N>(x0, 0, width - 1);
    }
    block_deal<N>(bits, list, warp_sums, [&](int x, int yy) {
        dstp[(size_t)yy * pitch + x] = (PIX)filter_map_px<PIX>(mskp, dmskp, pitch, width, height, k, x, yy);
    });
}

// mark_directions_2x (:787-858): dst pre-filled with peak (whole pitch)
template <typename PIX>
__device__ __forceinline__ void mark_directions_2x_px(const PIX *__restrict__ mskp, const PIX *__restrict__ dmskp, PIX *__restrict__ dstp,
                                                      int pitch, int width, int height, const K<PIX> &k, const Lim &lim, int x, int y)
{
    if (x < 1 || x >= width - 1) return;
    const PIX *m0 = mskp + (size_t)(y - 1) * pitch, *m1 = m0 + 2 * pitch;
    if (m0[x] != k.peak && m1[x] != k.peak) return;
    const PIX *d0 = dmskp + (size_t)(y - 1) * pitch, *d1 = d0 + 2 * pitch;
    int order[6] = { d0[x - 1] != k.peak ? (int)d0[x - 1] : kNone, d0[x] != k.peak ? (int)d0[x] : kNone, d0[x + 1] != k.peak ? (int)d0[x + 1] : kNone,
                     d1[x - 1] != k.peak ? (int)d1[x - 1] : kNone, d1[x] != k.peak ? (int)d1[x] : kNone, d1[x + 1] != k.peak ? (int)d1[x + 1] : kNone };
    int v = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) v += order[i] != kNone;
    if (v < 3) return;
    sort_net(order);
    const int mid = median_net(order, v);
    const int l = lim.v[iabs(mid - k.neutral) >> k.shift2];
    int u = 0;
    if (iabs((int)d0[x - 1] - (int)d1[x - 1]) <= l || d0[x - 1] == k.peak || d1[x - 1] == k.peak) ++u;
    if (iabs((int)d0[x] - (int)d1[x]) <= l || d0[x] == k.peak || d1[x] == k.peak) ++u;
    if (iabs((int)d0[x + 1] - (int)d1[x - 1]) <= l || d0[x + 1] == k.peak || d1[x + 1] == k.peak) ++u;    // sic (:835)
    if (u < 2) return;
    int count, sum;
    near_net(order, v, mid, l, count, sum);
    if (count < v - 2 || count < 2) return;
    dstp[(size_t)y * pitch + x] = (PIX)avg_round(sum, mid, count);
}

template <typename PIX>
__global__ void __launch_bounds__(256) k_mark_directions_2x(const PIX *__restrict__ mskp, const PIX *__restrict__ dmskp, PIX *__restrict__ dstp,
                                                            int pitch, int width, int height, int tff, int depth, Lim lim)
{
    constexpr int N = Vec<PIX>::N;
    STAGE_LIST(PIX);
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * N;
    const int y = 2 - tff + 2 * (blockIdx.y * blockDim.y + threadIdx.y);
    const K<PIX> k(depth);
    uint32_t bits = 0;
    if (y >= height - 1) bits = 0;
    else if (x0 + N <= width)
        bits = (eq_bits<PIX>(ld16(mskp + (size_t)(y - 1) * pitch + x0), k.peak) | eq_bits<PIX>(ld16(mskp + (size_t)(y + 1) * pitch + x0), k.peak)) &
               range_bits<N>(x0, 1, width - 2);
    else if (x0 < width)
        bits = range_bits<N>(x0, 1, width - 2);
    block_deal<N>(bits, list, warp_sums, [&](int x, int row) {
        mark_directions_2x_px<PIX>(mskp, dmskp, dstp, pitch, width, height, k, lim, x, 2 - tff + 2 * row);   // dst keeps its peak fill elsewhere
    });
}

// fill_gaps_2x (:1025-1132): the copy of dmsk into dst is part of the kernel; every gap sample writes its own position
template <typename PIX>
__device__ __forceinline__ void fill_gaps_2x_px(const PIX *__restrict__ mskp, const PIX *__restrict__ dmskp, PIX *__restrict__ dstp,
                                                int pitch, int width, int height, const K<PIX> &k, int x, int y)
{
    if (x < 1 || x >= width - 1) return;
    const int eight = 8 << k.shift, twenty = 20 << k.shift, fiveHundred = 500 << k.shift;
    const PIX *dc = dmskp + (size_t)y * pitch, *dp = dc - 2 * pitch, *dn = dc + 2 * pitch;
    const PIX *mc = mskp + (size_t)(y - 1) * pitch, *mpp = mc - 2 * pitch, *mn = mc + 2 * pitch, *mnn = mn + 2 * pitch;
    if (dc[x] != k.peak || (mc[x] != k.peak && mn[x] != k.peak)) return;
    int u = x - 1, back = fiveHundred, forward = -fiveHundred;
    while (u)
    {
        if (dc[u] != k.peak) { back = dc[u]; break; }
        if (mc[u] != k.peak && mn[u] != k.peak) break;
        --u;
    }
    int v = x + 1;
    while (v < width)
    {
        if (dc[v] != k.peak) { forward = dc[v]; break; }
        if (mc[v] != k.peak && mn[v] != k.peak) break;
        ++v;
    }
    int tc = 1, bc = 1;
    int mint = fiveHundred, maxt = -twenty, minb = fiveHundred, maxb = -twenty;
    for (int j = u; j <= v; ++j)
    {
        if (tc)
        {
            if (y <= 2 || dp[j] == k.peak || (mpp[j] != k.peak && mc[j] != k.peak)) { tc = 0; mint = maxt = twenty; }
            else { if ((int)dp[j] < mint) mint = dp[j]; if ((int)dp[j] > maxt) maxt = dp[j]; }
        }
        if (bc)
        {
            if (y >= height - 3 || dn[j] == k.peak || (mn[j] != k.peak && mnn[j] != k.peak)) { bc = 0; minb = maxb = twenty; }
            else { if ((int)dn[j] < minb) minb = dn[j]; if ((int)dn[j] > maxb) maxb = dn[j]; }
        }
    }
    if (maxt == -twenty) maxt = mint = twenty;
    if (maxb == -twenty) maxb = minb = twenty;
    const int fb = max(iabs(forward - k.neutral), iabs(back - k.neutral));
    const int thresh = max(max(fb >> 2, eight), max(iabs(mint - maxt), iabs(minb - maxb)));
    const int flim = min(fb >> k.shift2, 6);
    if (iabs(forward - back) <= thresh && (v - u - 1 <= flim || tc || bc))
    {
        // The reference fills the whole gap u+1 .. v-1 from its first sample; every sample of the gap is itself a start
        // sample that finds the same (u, v, back, forward) and takes the same decision, so each writes just its own
        // position -- a gather instead of v-u-1 identical scatters per gap.
        const double step = __ddiv_rn((double)(forward - back), (double)(v - u));
        dstp[(size_t)y * pitch + x] = (PIX)(back + (int)__dadd_rn(__dmul_rn((double)(x - u - 1), step), 0.5));
    }
}

template <typename PIX>
__global__ void __launch_bounds__(256) k_fill_gaps_2x(const PIX *__restrict__ mskp, const PIX *__restrict__ dmskp, PIX *__restrict__ dstp,
                                                      int pitch, int width, int height, int field, int depth)
{
    constexpr int N = Vec<PIX>::N;
    STAGE_LIST(PIX);
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * N;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;            // index of the processed row
    const int y = 2 - field + 2 * r;
    const K<PIX> k(depth);
    // eedi2_bit_blit (:1036) folded in: dst = dmsk over `width`.  A thread copies its vector on its processed row and the
    // row above it; the first processed row also takes the rows above that pair, the last one the rows below it.  The
    // barrier inside block_deal orders these stores before the gap samples' own stores.
    if (x0 < width && y < height - 1)
    {
        const int lo = r == 0 ? 0 : y - 1;
        const int hi = y + 2 >= height - 1 ? height - 1 : y;
        for (int yy = lo; yy <= hi; yy++)
        {
            const PIX *sp = dmskp + (size_t)yy * pitch + x0;
            PIX *dp = dstp + (size_t)yy * pitch + x0;
            if (x0 + N <= width) st16(dp, ld16(sp));
            else
                for (int i = 0; x0 + i < width; i++) dp[i] = sp[i];
        }
    }
    // a gap starts at a sample without direction (peak) that lies on the (line-doubled) edge mask
    uint32_t bits = 0;
    if (y >= height - 1) bits = 0;
    else if (x0 + N <= width)
        bits = eq_bits<PIX>(ld16(dmskp + (size_t)y * pitch + x0), k.peak) &
               (eq_bits<PIX>(ld16(mskp + (size_t)(y - 1) * pitch + x0), k.peak) | eq_bits<PIX>(ld16(mskp + (size_t)(y + 1) * pitch + x0), k.peak)) &
               range_bits<N>(x0, 1, width - 2);
    else if (x0 < width)
        bits = range_bits<N>(x0, 1, width - 2);
    block_deal<N>(bits, list, warp_sums, [&](int x, int row) {
        fill_gaps_2x_px<PIX>(mskp, dmskp, dstp, pitch, width, height, k, x, 2 - field + 2 * row);
    });
}

// ---------------------------------------------------------------------------------------------
// interpolate_lattice (:1148-1335)
//   pass A (parallel): per pixel, everything that does not depend on the rewritten dmskp[x-1]:
//     codeA  = what happens when the first test fires (always: vertical average; mask -> neutral unless dir==peak)
//     resB   = (pixel value, new mask) of the remaining path
//     gate   = dir==peak (fires regardless) / cond on dmskp[x+1] / lim
//   pass B (one thread per row): walks x = 0..w-1 carrying the rewritten dmskp[x-1]
// ---------------------------------------------------------------------------------------------
struct LatticeTmp            // per pixel, 8 bytes
{
    uint16_t valB;           // pixel value of the non-early path
    uint16_t mskB;           // rewritten mask of the non-early path
    uint16_t lim;            // lim for the first test
    uint8_t  nextfar;        // |dmskp[x] - dmskp[x+1]| > lim (uses the ORIGINAL x+1)
    uint8_t  pad;
};

template <typename PIX>
__device__ __forceinline__ void lattice_a_px(int plane, const PIX *__restrict__ dmskp, const PIX *__restrict__ dstp_base, const PIX *__restrict__ omsk_base,
                                             LatticeTmp *__restrict__ tmp, int pitch, int width, int height, int nt, int depth,
                                             const K<PIX> &k, const Lim &lim, int x, int y, int row)
{
    const int three = (int)(PIX)(3 << k.shift), nine = (int)(PIX)(9 << k.shift);
    const int nt4 = (int)(PIX)((nt << (depth - 8)) * 4);
    const int nt7 = (int)(PIX)((nt << (depth - 8)) * 7);
    const int nt8 = (int)(PIX)((nt << (depth - 8)) * 8);
    const PIX *dm = dmskp + (size_t)y * pitch;
    const PIX *dstp = dstp_base + (size_t)(y - 1) * pitch, *dstpnn = dstp + 2 * pitch;
    const PIX *omskp = omsk_base + (size_t)(y - 1) * pitch, *omskn = omskp + 2 * pitch;
    LatticeTmp t;
    int dir = dm[x];
    const int cur = dir;
    const int l = lim.v[iabs(dir - k.neutral) >> k.shift2];
    t.lim = (uint16_t)l;
    t.nextfar = iabs(cur - (int)dm[x + 1]) > l;
    t.pad = 0;
    const int avg = ((int)dstp[x] + (int)dstpnn[x] + 1) >> 1;
    int val = avg, msk = 0;
    bool done = false;
    if (dir == k.peak)
    {
        done = true;            // early path regardless; values unused
        msk = k.peak;
    }
    if (!done && l < nine)
    {
        const int s = k.shift;
        const int a0 = dstp[x - 1], a1 = dstp[x], a2 = dstp[x + 1], b0 = dstpnn[x - 1], b1 = dstpnn[x], b2 = dstpnn[x + 1];
        const int sum = (a0 + a1 + a2 + b0 + b1 + b2) >> s;
        const int sumsq = (a0 >> s) * (a0 >> s) + (a1 >> s) * (a1 >> s) + (a2 >> s) * (a2 >> s) +
                          (b0 >> s) * (b0 >> s) + (b1 >> s) * (b1 >> s) + (b2 >> s) * (b2 >> s);
        if (6 * sumsq - sum * sum < 576) { val = avg; msk = k.peak; done = true; }
    }
    if (!done && x > 1 && x < width - 2)
    {
        const int a = dstp[x], b = dstpnn[x];
        if ((a < max((int)dstp[x - 2], (int)dstp[x - 1]) - three && a < max((int)dstp[x + 2], (int)dstp[x + 1]) - three &&
             b < max((int)dstpnn[x - 2], (int)dstpnn[x - 1]) - three && b < max((int)dstpnn[x + 2], (int)dstpnn[x + 1]) - three) ||
            (a > min((int)dstp[x - 2], (int)dstp[x - 1]) + three && a > min((int)dstp[x + 2], (int)dstp[x + 1]) + three &&
             b > min((int)dstpnn[x - 2], (int)dstpnn[x - 1]) + three && b > min((int)dstpnn[x + 2], (int)dstpnn[x + 1]) + three))
        {
            val = avg; msk = k.neutral; done = true;
        }
    }
    if (!done)
    {
        dir = (dir - k.neutral + (1 << (k.shift2 - 1))) >> k.shift2;
        const int startu = (dir - 2 < 0) ? max(-x + 1, max(dir - 2, -width + 2 + x)) : min(x - 1, min(dir - 2, width - 2 - x));
        const int stopu  = (dir + 2 < 0) ? max(-x + 1, max(dir + 2, -width + 2 + x)) : min(x - 1, min(dir + 2, width - 2 - x));
        int mn = nt8;
        for (int u = startu; u <= stopu; ++u)
        {
            const int diff = iabs((int)dstp[x - 1] - (int)dstpnn[x - u - 1]) + iabs((int)dstp[x] - (int)dstpnn[x - u]) +
                             iabs((int)dstp[x + 1] - (int)dstpnn[x - u + 1]) + iabs((int)dstpnn[x - 1] - (int)dstp[x + u - 1]) +
                             iabs((int)dstpnn[x] - (int)dstp[x + u]) + iabs((int)dstpnn[x + 1] - (int)dstp[x + u + 1]);
#define NEAR(arr, idx) ((arr)[(idx)] != k.peak && iabs((int)(arr)[(idx)] - cur) <= l)
            if (diff < mn &&
                (NEAR(omskp, x - 1 + u) || NEAR(omskp, x + u) || NEAR(omskp, x + 1 + u)) &&
                (NEAR(omskn, x - 1 - u) || NEAR(omskn, x - u) || NEAR(omskn, x + 1 - u)))
            {
                const int h0 = u >> 1, h1 = (u + 1) >> 1;
                const int diff2 = iabs((int)dstp[x + h0 - 1] - (int)dstpnn[x - h0 - 1]) + iabs((int)dstp[x + h0] - (int)dstpnn[x - h0]) +
                                  iabs((int)dstp[x + h0 + 1] - (int)dstpnn[x - h0 + 1]);
                if (diff2 < nt4 &&
                    (((iabs((int)omskp[x + h0] - (int)omskn[x - h0]) <= l || iabs((int)omskp[x + h0] - (int)omskn[x - h1]) <= l) && omskp[x + h0] != k.peak) ||
                     ((iabs((int)omskp[x + h1] - (int)omskn[x - h0]) <= l || iabs((int)omskp[x + h1] - (int)omskn[x - h1]) <= l) && omskp[x + h1] != k.peak)))
                {
                    if ((iabs(cur - (int)omskp[x + h0]) <= l || iabs(cur - (int)omskp[x + h1]) <= l) &&
                        (iabs(cur - (int)omskn[x - h0]) <= l || iabs(cur - (int)omskn[x - h1]) <= l))
                    {
                        val = ((int)dstp[x + h0] + (int)dstp[x + h1] + (int)dstpnn[x - h0] + (int)dstpnn[x - h1] + 2) >> 2;
                        mn = diff;
                        dir = u;
                    }
                }
            }
#undef NEAR
        }
        if (mn != nt8)
        {
            msk = k.neutral + dir * (1 << k.shift2);
        }
        else
        {
            const int minm = min((int)dstp[x], (int)dstpnn[x]), maxm = max((int)dstp[x], (int)dstpnn[x]);
            const int d = plane == 0 ? 4 : 2;
            const int su = max(-x + 1, -d), eu = min(width - 2 - x, d);
            mn = nt7;
            for (int u = su; u <= eu; ++u)
            {
                const int h0 = u >> 1, h1 = (u + 1) >> 1;
                const int p1 = (int)dstp[x + h0] + (int)dstp[x + h1];
                const int p2 = (int)dstpnn[x - h0] + (int)dstpnn[x - h1];
                const int diff = iabs((int)dstp[x - 1] - (int)dstpnn[x - u - 1]) + iabs((int)dstp[x] - (int)dstpnn[x - u]) +
                                 iabs((int)dstp[x + 1] - (int)dstpnn[x - u + 1]) + iabs((int)dstpnn[x - 1] - (int)dstp[x + u - 1]) +
                                 iabs((int)dstpnn[x] - (int)dstp[x + u]) + iabs((int)dstpnn[x + 1] - (int)dstp[x + u + 1]) + iabs(p1 - p2);
                if (diff < mn)
                {
                    const int valt = (p1 + p2 + 2) >> 2;
                    if (valt >= minm && valt <= maxm) { val = valt; mn = diff; dir = u; }
                }
            }
            if (mn == 7 * nt) msk = k.neutral;                  // unwrapped constant, as in the reference (:1324)
            else msk = k.neutral + dir * (1 << k.shift2);
        }
    }
    t.valB = (uint16_t)(PIX)val;
    t.mskB = (uint16_t)(PIX)msk;
    tmp[(size_t)row * width + x] = t;
}

// Samples without a direction (dmsk == peak) take the vertical average in pass B whatever their LatticeTmp holds
// (lat_pixel() and the final select both test cur == peak first), so such samples write nothing.
template <typename PIX>
__global__ void __launch_bounds__(256) k_lattice_a(int plane, const PIX *__restrict__ dmskp, const PIX *__restrict__ dstp_base, const PIX *__restrict__ omsk_base,
                                                   LatticeTmp *__restrict__ tmp, int pitch, int width, int height, int field, int nt, int depth, Lim lim)
{
    constexpr int N = Vec<PIX>::N;
    STAGE_LIST(PIX);
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * N;
    const int row = blockIdx.y * blockDim.y + threadIdx.y;
    const int y = 2 - field + 2 * row;
    const K<PIX> k(depth);
    uint32_t bits = 0;
    if (y >= height - 1) bits = 0;
    else if (x0 + N <= width)  bits = ~eq_bits<PIX>(ld16(dmskp + (size_t)y * pitch + x0), k.peak) & range_bits<N>(x0, 0, width - 1);
    else if (x0 < width)  bits = range_bits<N>(x0, 0, width - 1);
    block_deal<N>(bits, list, warp_sums, [&](int x, int r) {
        lattice_a_px<PIX>(plane, dmskp, dstp_base, omsk_base, tmp, pitch, width, height, nt, depth, k, lim, x, 2 - field + 2 * r, r);
    });
}

// Pass B as a parallel scan.  Pixel x maps the rewritten mask of its left neighbour m to its own
// rewritten mask:  f_x(m) = peak                         if cur == peak
//                          B_x                          if |cur - dm[x+1]| <= lim          (first test cannot fire)
//                          (|cur - m| > lim) ? neutral : B_x   otherwise
// i.e. every f_x, and every composition of them, has the form  F(m) = (lo <= m <= hi) ? vin : vout.
// That family is closed under composition, so the chain m_x = f_x(m_{x-1}) is an associative scan:
// each thread composes a contiguous chunk, a block-wide Hillis-Steele scan composes the chunks, and a
// second walk applies the now-known incoming value.  One CTA per row; the row is staged in shared memory.
struct LatFn { int lo, hi, vin, vout; };
__device__ __forceinline__ int lat_apply(const LatFn &f, int m) { return (m >= f.lo && m <= f.hi) ? f.vin : f.vout; }
__device__ __forceinline__ LatFn lat_compose(const LatFn &later, const LatFn &earlier)
{
    return LatFn{ earlier.lo, earlier.hi, lat_apply(later, earlier.vin), lat_apply(later, earlier.vout) };
}
__device__ __forceinline__ LatFn lat_pixel(int cur, const LatticeTmp &e, int peak, int neutral)
{
    if (cur == peak) return LatFn{ 1, 0, peak, peak };                       // empty interval: constant
    if (!e.nextfar) return LatFn{ 1, 0, (int)e.mskB, (int)e.mskB };
    return LatFn{ cur - (int)e.lim, cur + (int)e.lim, (int)e.mskB, neutral };
}

constexpr int kLatThreads = 256;

template <typename PIX>
__global__ void __launch_bounds__(kLatThreads) k_lattice_b(PIX *__restrict__ dmskp, PIX *__restrict__ dst_base, const LatticeTmp *__restrict__ tmp,
                                                           int pitch, int width, int height, int field, int depth)
{
    extern __shared__ __align__(16) unsigned char lat_smem[];
    LatticeTmp *s_t = reinterpret_cast<LatticeTmp *>(lat_smem);                       // width entries
    int *s_cur = reinterpret_cast<int *>(lat_smem + (size_t)width * sizeof(LatticeTmp));      // width
    int *s_prevm = s_cur + width;                                                    // width: rewritten mask of x-1
    __shared__ LatFn s_fn[kLatThreads];

    const int row = blockIdx.x;
    const int y = 2 - field + 2 * row;
    if (y >= height - 1) return;
    const int peak = (1 << depth) - 1, neutral = 1 << (depth - 1);
    PIX *dm = dmskp + (size_t)y * pitch;
    PIX *dn = dst_base + (size_t)y * pitch;
    const PIX *up = dn - pitch, *down = dn + pitch;
    const LatticeTmp *t = tmp + (size_t)row * width;
    const int tid = threadIdx.x;
    for (int x = tid; x < width; x += kLatThreads) s_cur[x] = dm[x];
    __syncthreads();
    const int chunk = (width + kLatThreads - 1) / kLatThreads;
    const int x0 = tid * chunk, x1 = min(x0 + chunk, width);

    // Fast path.  A sample without direction (cur == peak) maps every incoming mask to peak, so the chain only runs
    // inside runs of samples that have a direction, and every run starts from a known value.  When every thread's
    // chunk contains such a reset sample no run is longer than two chunks: the thread that owns a run's first sample
    // walks it serially.  (Textured rows with long runs take the scan below.)
    {
        bool full = x0 < x1;
        for (int x = x0; x < x1 && full; ++x) full = s_cur[x] != peak;
        if (!__syncthreads_or(full ? 1 : 0))
        {
            for (int x = tid; x < width; x += kLatThreads)
                if (s_cur[x] == peak) dn[x] = (PIX)(((int)up[x] + (int)down[x] + 1) >> 1);       // mask stays peak
            for (int x = x0; x < x1; ++x)
            {
                if (s_cur[x] == peak || (x > 0 && s_cur[x - 1] != peak)) continue;              // not the head of a run
                int prev = x > 0 ? peak : (int)dm[-1];
                for (int xx = x; xx < width && s_cur[xx] != peak; ++xx)
                {
                    const LatticeTmp e = t[xx];
                    const int cur = s_cur[xx];
                    int newm, val;
                    if (abs(cur - prev) > (int)e.lim && e.nextfar)
                    {
                        val = ((int)up[xx] + (int)down[xx] + 1) >> 1;
                        newm = neutral;
                    }
                    else
                    {
                        val = e.valB;
                        newm = e.mskB;
                    }
                    dn[xx] = (PIX)val;
                    dm[xx] = (PIX)newm;
                    prev = newm;
                }
            }
            return;
        }
    }
    for (int x = tid; x < width; x += kLatThreads) s_t[x] = t[x];
    __syncthreads();
    LatFn f{ 1, 0, 0, 0 };
    bool have = false;
    for (int x = x0; x < x1; ++x)
    {
        const LatFn g = lat_pixel(s_cur[x], s_t[x], peak, neutral);
        f = have ? lat_compose(g, f) : g;
        have = true;
    }
    // identity for empty chunks: F(m) = m cannot be written as an interval test, so mark them and skip
    s_fn[tid] = f;
    __shared__ unsigned char s_has[kLatThreads];
    s_has[tid] = have;
    __syncthreads();
    // inclusive scan of compositions (later o earlier) over the non-empty chunks
    for (int off = 1; off < kLatThreads; off <<= 1)
    {
        LatFn mine = s_fn[tid];
        bool mh = s_has[tid];
        LatFn other{ 1, 0, 0, 0 };
        bool oh = false;
        if (tid >= off) { other = s_fn[tid - off]; oh = s_has[tid - off]; }
        __syncthreads();
        if (oh)
        {
            s_fn[tid] = mh ? lat_compose(mine, other) : other;
            s_has[tid] = true;
        }
        __syncthreads();
    }
    const int m_init = dm[-1];                                  // linear predecessor of the row (not rewritten by this pass)
    int prev = m_init;
    if (tid > 0 && s_has[tid - 1]) prev = lat_apply(s_fn[tid - 1], m_init);
    for (int x = x0; x < x1; ++x)
    {
        s_prevm[x] = prev;
        prev = lat_apply(lat_pixel(s_cur[x], s_t[x], peak, neutral), prev);
    }
    __syncthreads();
    for (int x = tid; x < width; x += kLatThreads)
    {
        const int cur = s_cur[x];
        const LatticeTmp e = s_t[x];
        int newm, val;
        if (cur == peak || (abs(cur - s_prevm[x]) > (int)e.lim && e.nextfar))
        {
            val = ((int)up[x] + (int)down[x] + 1) >> 1;
            newm = cur != peak ? neutral : cur;
        }
        else
        {
            val = e.valB;
            newm = e.mskB;
        }
        dn[x] = (PIX)val;
        dm[x] = (PIX)newm;
    }
}

// post_process (:1349-1378)
template <typename PIX>
__global__ void __launch_bounds__(256) k_post_process(const PIX *__restrict__ nmskp, const PIX *__restrict__ omskp, PIX *__restrict__ dstp,
                                                      int pitch, int width, int height, int field, int depth, Lim lim)
{
    constexpr int N = Vec<PIX>::N;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * N;
    const int y = 2 - field + 2 * (blockIdx.y * blockDim.y + threadIdx.y);
    if (x0 >= width || y >= height - 1) return;
    const K<PIX> k(depth);
    // only samples whose OLD direction exists (not peak) can change
    if (x0 + N <= width && !any_ne<PIX>(ld16(omskp + (size_t)y * pitch + x0), k.peak)) return;
    for (int i = 0; i < N && x0 + i < width; i++)
    {
        const int x = x0 + i;
        const size_t o = (size_t)y * pitch + x;
        const int nm = nmskp[o], om = omskp[o];
        const int l = lim.v[iabs(nm - k.neutral) >> k.shift2];
        if (iabs(nm - om) > l && om != k.peak && om != k.neutral)
            dstp[o] = (PIX)(((int)dstp[o - pitch] + (int)dstp[o + pitch] + 1) >> 1);
    }
}

// ---------------------------------------------------------------------------------------------
// postproc 2/3: junctions and corners (:1391-1904, called from decomb template :431-440)
//
// The reference writes its two blurs out as one expression per edge case.  Read together: a symmetric kernel whose tap
// at distance k that would fall outside [0,n) is replaced by its point reflection through the centre sample
// (x-k <-> x+k, hence the doubled weights) -- with one exception, reproduced: the horizontal pass of
// gaussian_blur_sqrt2 at x = width-2 reads srcp[x+3] for both distance-3 taps (:1625), one element past the row
// (stride padding, or the next row's second element when pitch == width).
// The reference shares ONE set of scratch arrays between its three concurrently running plane threads (decomb.c:396-403:
// a data race) and that exception reads elements nothing ever wrote.  Implemented is the race-free reading: every plane
// owns its four scratch arrays, zero-filled at create.
struct BlurTaps { int v[5]; int radius; };

template <typename Load>
__device__ __forceinline__ int blur_sample(Load ld, int i, int n, const BlurTaps &t, int typo_at)
{
    int acc = ld(0) * t.v[0] + 32768;
#pragma unroll 4
    for (int k = 1; k <= t.radius; k++)
    {
        int a = i - k >= 0 ? -k : k, b = i + k <= n - 1 ? k : -k;
        if (k == 3 && i == typo_at) a = b = k;
        acc += (ld(a) + ld(b)) * t.v[k];
    }
    return acc;
}

// T = pixel type (gaussian_blur1, :1402-1527) or int (gaussian_blur_sqrt2, :1540-1745); VERT selects the pass
template <typename T, bool VERT>
__global__ void __launch_bounds__(256) k_blur(const T *__restrict__ src, T *__restrict__ dst, int pitch, int width, int height,
                                              BlurTaps taps, int typo_at, int shift)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    const T *p = src + (size_t)y * pitch + x;
    const int acc = VERT ? blur_sample([&](int d) { return (int)p[(ptrdiff_t)d * pitch]; }, y, height, taps, -1)
                         : blur_sample([&](int d) { return (int)p[d]; }, x, width, taps, typo_at);
    dst[(size_t)y * pitch + x] = (T)(acc >> shift);
}

// calc_derivatives (:1756-1845): central differences, one-sided at the plane's edges, scaled back to 8 bits
template <typename PIX>
__global__ void __launch_bounds__(256) k_derivatives(const PIX *__restrict__ src, int *__restrict__ x2, int *__restrict__ y2, int *__restrict__ xy,
                                                     int pitch, int width, int height, int shift)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    const PIX *row = src + (size_t)y * pitch;
    const int Ix = ((int)row[min(x + 1, width - 1)] - (int)row[max(x - 1, 0)]) >> shift;
    const int Iy = ((int)src[(size_t)max(y - 1, 0) * pitch + x] - (int)src[(size_t)min(y + 1, height - 1) * pitch + x]) >> shift;
    const size_t o = (size_t)y * pitch + x;
    x2[o] = (Ix * Ix) >> 1;
    y2[o] = (Iy * Iy) >> 1;
    xy[o] = (Ix * Iy) >> 1;
}

__device__ __forceinline__ int corner_response(int a, int b, int c)      // :1882-1885, double arithmetic without contraction
{
    const double s = (double)(a + b);
    return __double2int_rz(__dsub_rn((double)(a * b - c * c), __dmul_rn(__dmul_rn(0.09, s), s)));
}

// post_process_corner (:1864-1900): picture rows y = 8-field, +2, ... < height-7 against derivative rows 3, 4, ...
template <typename PIX>
__global__ void __launch_bounds__(256) k_corner(const int *__restrict__ x2, const int *__restrict__ y2, const int *__restrict__ xy,
                                                const PIX *__restrict__ mskp, PIX *__restrict__ dstp, int pitch, int width, int height,
                                                int field, int depth)
{
    const int x = 4 + blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
    const int y = 8 - field + 2 * r;
    if (x >= width - 4 || y >= height - 7) return;
    const size_t o = (size_t)y * pitch + x;
    const int m = mskp[o];
    if (m == (1 << depth) - 1 || m == 1 << (depth - 1)) return;
    const size_t d = (size_t)(r + 3) * pitch + x;
    if (corner_response(x2[d], y2[d], xy[d]) > 775 || corner_response(x2[d + pitch], y2[d + pitch], xy[d + pitch]) > 775)
        dstp[o] = (PIX)(((int)dstp[o - pitch] + (int)dstp[o + pitch] + 1) >> 1);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Eedi2
{
    Eedi2Config cfg;
    int bps;
    int half_h[3];                 // field-buffer plane heights
    size_t half_off[3], full_off[3], half_bytes, full_bytes;
    uint8_t *half_mem[4], *full_mem[5];      // allocation starts (incl. lead slack)
    LatticeTmp *lattice_tmp, *lattice_tmp_pl[3];   // one allocation, a private region per plane (the planes run side by side)
    int *deriv_mem, *deriv[3][4];                   // postproc 2/3: x2, y2, xy, tmp per plane (pitch x (field rows + 1), zeroed)
    cudaStream_t s_aux[2];                          // chroma planes' branches of the captured graph
    cudaEvent_t ev_fork, ev_join[2];
    Lim lim;
    int stop_after;                // debug: number of stage launches to run per plane (0 = all); HBCU_EEDI2_STOP
    // HBCU_EEDI2_TIMING=1: CUDA-event pair around every stage launch, per-stage totals printed at destroy (warm,
    // in-pipeline times; an ncu launch list gives cold, serialised ones)
    int timing;
    cudaEvent_t tev[2];
    double stage_ms[3][40];
    const char *stage_name[40];
    long stage_calls;
    // The ~90 stage launches of a field are captured once per (source frame slot, field order) into a CUDA graph and
    // replayed: the chroma stages are a few microseconds each, launch-bound when issued one by one.
    struct GraphEntry { const void *planes[3]; int tff; cudaGraphExec_t exec; int launches; };
    std::vector<GraphEntry> graphs;
    int use_graphs;                // HBCU_EEDI2_GRAPHS=0 turns the replay off (A/B)
};

namespace {

template <typename PIX>
int run_plane(Eedi2 *e, int pl, const PIX *cur_plane, int tff, cudaStream_t st)
{
    const Eedi2Config &c = e->cfg;
    const int pitch = c.pitch[pl], width = c.w[pl], height = c.h[pl], hh = e->half_h[pl], depth = c.depth;
    auto H = [&](int k) { return reinterpret_cast<PIX *>(e->half_mem[k] + kLeadSlack + e->half_off[pl]); };
    auto F = [&](int k) { return reinterpret_cast<PIX *>(e->full_mem[k] + kLeadSlack + e->full_off[pl]); };
    PIX *srcp = H(SRCPF), *mskp = H(MSKPF), *tmpp = H(TMPPF), *dstp = H(DSTPF);
    PIX *dst2p = F(DST2PF), *tmp2p2 = F(TMP2PF2), *msk2p = F(MSK2PF), *tmp2p = F(TMP2PF), *dst2mp = F(DST2MPF);
    const dim3 blk(64, 4);
    auto grid2 = [&](int w, int h) { return dim3((w + 63) / 64, (h + 3) / 4); };
    const int vn = 16 / (int)sizeof(PIX);                         // samples per thread of the vector-gated stages
    auto gridv = [&](int w, int h) { return dim3(((w + vn - 1) / vn + 63) / 64, (h + 3) / 4); };
    const dim3 rowblk(64, 4);                                    // 16-byte chunks x rows
    const int epc = 16 / (int)sizeof(PIX);
    auto gridrows = [&](int w, int rows) { return dim3(((w + epc - 1) / epc + 63) / 64, (rows + 3) / 4); };
    int launches = 0;
#define LAUNCH(...) do { if (e->stop_after == 0 || launches < e->stop_after) {                                        \
        if (e->timing && launches < 40) { cudaEventRecord(e->tev[0], st); __VA_ARGS__; cudaEventRecord(e->tev[1], st);    \
            cudaEventSynchronize(e->tev[1]); float ms_ = 0.f; cudaEventElapsedTime(&ms_, e->tev[0], e->tev[1]);          \
            e->stage_ms[pl][launches] += ms_; e->stage_name[launches] = #__VA_ARGS__; }                                  \
        else { __VA_ARGS__; } } ++launches; } while (0)

    if ((uintptr_t)cur_plane % 16)
    {
        set_error("eedi2: source plane %d is not 16-byte aligned", pl);
        return -1;
    }
    // eedi2_planer: field start_line = !tff of the current frame, whole strides (:455-466, :79-92)
    const int field_rows = (height + 1) / 2;
    LAUNCH((k_fill_half<PIX><<<gridrows(pitch, field_rows), rowblk, 0, st>>>(cur_plane + (size_t)pitch * (!tff), srcp, pitch, field_rows)));

    // edge mask: top half cleared, bottom half keeps the previous field's mask (:132)
    if (e->stop_after == 0 || launches < e->stop_after) cudaMemsetAsync(mskp, 0, (size_t)(hh / 2) * pitch * sizeof(PIX), st);
    LAUNCH((k_edge_mask<PIX><<<gridv(width, hh), blk, 0, st>>>(mskp, srcp, pitch, width, hh, c.mthresh * 10, c.vthresh /* lthresh <- variance */,
                                                              c.lthresh * 81 /* vthresh <- laplacian */, depth)));
    LAUNCH((k_morph<PIX, false><<<gridv(width, hh), blk, 0, st>>>(mskp, tmpp, pitch, width, hh, c.estr, depth)));
    LAUNCH((k_morph<PIX, true><<<gridv(width, hh), blk, 0, st>>>(tmpp, mskp, pitch, width, hh, c.dstr, depth)));
    LAUNCH((k_morph<PIX, false><<<gridv(width, hh), blk, 0, st>>>(mskp, tmpp, pitch, width, hh, c.estr, depth)));
    LAUNCH((k_gaps<PIX><<<gridv(width, hh), blk, 0, st>>>(tmpp, mskp, pitch, width, hh, depth)));

    // direction mask
    const int peak = (1 << depth) - 1;
    {
        const size_t n = (size_t)pitch * hh;
        LAUNCH((k_fill<PIX><<<(unsigned)((n / epc + 255) / 256), 256, 0, st>>>(tmpp, n, peak)));
    }
    LAUNCH((k_calc_directions<PIX><<<gridv(width, hh), blk, 0, st>>>(pl, mskp, srcp, tmpp, pitch, width, hh, c.maxd, c.nt, depth, e->lim)));
    LAUNCH((k_dir_map<PIX, false, false><<<gridv(width, hh), blk, 0, st>>>(mskp, tmpp, dstp, pitch, width, hh, 0, depth, e->lim)));
    LAUNCH((k_dir_map<PIX, true, false><<<gridv(width, hh), blk, 0, st>>>(mskp, dstp, tmpp, pitch, width, hh, 0, depth, e->lim)));
    LAUNCH((k_filter_map<PIX><<<gridv(width, hh), blk, 0, st>>>(mskp, tmpp, dstp, pitch, width, hh, depth)));

    // upscale 2x vertically (whole strides)
    LAUNCH((k_upscale2<PIX><<<gridrows(pitch, hh), rowblk, 0, st>>>(srcp, dst2p, pitch, hh)));
    LAUNCH((k_upscale2<PIX><<<gridrows(pitch, hh), rowblk, 0, st>>>(dstp, tmp2p2, pitch, hh)));
    LAUNCH((k_upscale2<PIX><<<gridrows(pitch, hh), rowblk, 0, st>>>(mskp, msk2p, pitch, hh)));

    // direction mask at frame height
    const int rows2 = (height - 1 - (2 - tff) + 1) / 2 > 0 ? (height - 1 - (2 - tff) + 1) / 2 : 0;   // y = 2-tff, +2, ... < height-1
    {
        const size_t n = (size_t)pitch * height;
        LAUNCH((k_fill<PIX><<<(unsigned)((n / epc + 255) / 256), 256, 0, st>>>(tmp2p, n, peak)));
    }
    LAUNCH((k_mark_directions_2x<PIX><<<gridv(width, rows2), blk, 0, st>>>(msk2p, tmp2p2, tmp2p, pitch, width, height, tff, depth, e->lim)));
    LAUNCH((k_dir_map<PIX, false, true><<<gridv(width, height), blk, 0, st>>>(msk2p, tmp2p, dst2mp, pitch, width, height, tff, depth, e->lim)));
    LAUNCH((k_dir_map<PIX, true, true><<<gridv(width, height), blk, 0, st>>>(msk2p, dst2mp, tmp2p, pitch, width, height, tff, depth, e->lim)));
    // fill_gaps_2x twice (:391-392); its eedi2_bit_blit is fused into the kernel (a plane too small to have a processed
    // row still gets the plain copy)
    if (rows2 > 0) LAUNCH((k_fill_gaps_2x<PIX><<<gridv(width, rows2), blk, 0, st>>>(msk2p, tmp2p, dst2mp, pitch, width, height, tff, depth)));
    else           LAUNCH((k_blit<PIX><<<gridrows(width, height), rowblk, 0, st>>>(tmp2p, dst2mp, pitch, width, height)));
    if (rows2 > 0) LAUNCH((k_fill_gaps_2x<PIX><<<gridv(width, rows2), blk, 0, st>>>(msk2p, dst2mp, tmp2p, pitch, width, height, tff, depth)));
    else           LAUNCH((k_blit<PIX><<<gridrows(width, height), rowblk, 0, st>>>(dst2mp, tmp2p, pitch, width, height)));

    // interpolate the missing lines (:1148-1335): first copy one border row, then the two passes
    if (tff == 1) LAUNCH((k_blit<PIX><<<gridrows(width, 1), rowblk, 0, st>>>(dst2p + (size_t)(height - 2) * pitch, dst2p + (size_t)(height - 1) * pitch, pitch, width, 1)));
    else          LAUNCH((k_blit<PIX><<<gridrows(width, 1), rowblk, 0, st>>>(dst2p + pitch, dst2p, pitch, width, 1)));
    LAUNCH((k_lattice_a<PIX><<<gridv(width, rows2), blk, 0, st>>>(pl, tmp2p, dst2p, tmp2p2, e->lattice_tmp_pl[pl], pitch, width, height, tff, c.nt, depth, e->lim)));
    LAUNCH((k_lattice_b<PIX><<<rows2, kLatThreads, (size_t)width * (sizeof(LatticeTmp) + 2 * sizeof(int)), st>>>(tmp2p, dst2p, e->lattice_tmp_pl[pl], pitch, width, height, tff, depth)));

    if (c.pp == 1 || c.pp == 3)
    {
        LAUNCH((k_blit<PIX><<<gridrows(width, height), rowblk, 0, st>>>(tmp2p, tmp2p2, pitch, width, height)));
        LAUNCH((k_dir_map<PIX, false, true><<<gridv(width, height), blk, 0, st>>>(msk2p, tmp2p, dst2mp, pitch, width, height, tff, depth, e->lim)));
        LAUNCH((k_dir_map<PIX, true, true><<<gridv(width, height), blk, 0, st>>>(msk2p, dst2mp, tmp2p, pitch, width, height, tff, depth, e->lim)));
        LAUNCH((k_post_process<PIX><<<gridv(width, rows2), blk, 0, st>>>(tmp2p, tmp2p2, dst2p, pitch, width, height, tff, depth, e->lim)));
    }
    if (c.pp == 2 || c.pp == 3)
    {
        // filter junctions and corners (decomb template :431-440): blur the field in place, structure tensor of the
        // blurred field, blur its three components, then replace flagged interpolated samples by the vertical average
        const BlurTaps t1 = { { 26152, 15862, 3539, 291, 0 }, 3 }, t2 = { { 18508, 14415, 6809, 1951, 339 }, 4 };
        int *cx2 = e->deriv[pl][0], *cy2 = e->deriv[pl][1], *cxy = e->deriv[pl][2], *tmpc = e->deriv[pl][3];
        LAUNCH((k_blur<PIX, false><<<grid2(width, hh), blk, 0, st>>>(srcp, tmpp, pitch, width, hh, t1, -1, 16)));
        LAUNCH((k_blur<PIX, true><<<grid2(width, hh), blk, 0, st>>>(tmpp, srcp, pitch, width, hh, t1, -1, 16)));
        LAUNCH((k_derivatives<PIX><<<grid2(width, hh), blk, 0, st>>>(srcp, cx2, cy2, cxy, pitch, width, hh, depth - 8)));
        int *const comp[3] = { cx2, cy2, cxy };
        for (int k = 0; k < 3; k++)
        {
            LAUNCH((k_blur<int, false><<<grid2(width, hh), blk, 0, st>>>(comp[k], tmpc, pitch, width, hh, t2, width - 2, 16)));
            LAUNCH((k_blur<int, true><<<grid2(width, hh), blk, 0, st>>>(tmpc, comp[k], pitch, width, hh, t2, -1, 18)));
        }
        const int crows = (height - 7 - (8 - tff) + 1) / 2;
        if (crows > 0 && width > 8)
            LAUNCH((k_corner<PIX><<<grid2(width - 8, crows), blk, 0, st>>>(cx2, cy2, cxy, tmp2p2, dst2p, pitch, width, height, tff, depth)));
    }
#undef LAUNCH
    hbcu::count_launch(launches);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
    {
        set_error("eedi2: kernel launch failed: %s", cudaGetErrorString(err));
        return -1;
    }
    return 0;
}

}  // namespace

Eedi2 *eedi2_create(const Eedi2Config &cfg)
{
    for (int pl = 0; pl < 3; pl++)
    {
        if (cfg.h[pl] & 1)
        {
            // odd plane heights make the reference copy one line beyond the plane into the field buffer (:79-92)
            set_error("eedi2: plane %d has odd height %d; EEDI2 needs a frame height that keeps every plane even", pl, cfg.h[pl]);
            return nullptr;
        }
    }
    if (cfg.pp < 0 || cfg.pp > 3)
    {
        set_error("eedi2: postproc %d is out of range", cfg.pp);
        return nullptr;
    }
    for (int pl = 0; pl < 3 && cfg.pp > 1; pl++)
    {
        if (cfg.w[pl] < 16 || cfg.h[pl] < 32)
        {
            // the reference's blurs spell out 4 edge columns / rows on either side; smaller planes run them into each other
            set_error("eedi2: postproc %d needs planes of at least 16x32, plane %d is %dx%d", cfg.pp, pl, cfg.w[pl], cfg.h[pl]);
            return nullptr;
        }
    }
    for (int pl = 0; pl < 3; pl++)
    {
        if (((size_t)cfg.pitch[pl] * (cfg.depth > 8 ? 2 : 1)) % 64)
        {
            set_error("eedi2: plane %d pitch %d is not the reference's 64-byte rounded stride", pl, cfg.pitch[pl]);
            return nullptr;
        }
    }
    Eedi2 *e = new (std::nothrow) Eedi2();
    if (e == nullptr) { set_error("eedi2: out of memory"); return nullptr; }
    e->cfg = cfg;
    e->bps = cfg.depth > 8 ? 2 : 1;
    for (int k = 0; k < 4; k++) e->half_mem[k] = nullptr;
    for (int k = 0; k < 5; k++) e->full_mem[k] = nullptr;
    e->lattice_tmp = nullptr;
    e->deriv_mem = nullptr;
    e->s_aux[0] = e->s_aux[1] = nullptr;
    e->ev_fork = e->ev_join[0] = e->ev_join[1] = nullptr;
    e->stop_after = 0;
    if (const char *sa = getenv("HBCU_EEDI2_STOP")) e->stop_after = atoi(sa);   // test hook: stage-by-stage parity
    e->timing = getenv("HBCU_EEDI2_TIMING") != nullptr;
    e->use_graphs = 1;
    if (const char *g = getenv("HBCU_EEDI2_GRAPHS")) e->use_graphs = atoi(g) != 0;
    e->tev[0] = e->tev[1] = nullptr;
    memset(e->stage_ms, 0, sizeof(e->stage_ms));
    memset(e->stage_name, 0, sizeof(e->stage_name));
    e->stage_calls = 0;
    if (e->timing) { cudaEventCreate(&e->tev[0]); cudaEventCreate(&e->tev[1]); }
    // field buffers are frames of height frame_height/2 (decomb.c:291-296): chroma rounds up from that
    e->half_h[0] = cfg.half_frame_height;
    e->half_h[1] = e->half_h[2] = -((-cfg.half_frame_height) >> cfg.chroma_shift_h);
    size_t ho = 0, fo = 0;
    for (int pl = 0; pl < 3; pl++)
    {
        if (e->half_h[pl] * 2 != cfg.h[pl])
        {
            set_error("eedi2: plane %d: field buffer height %d is not half of %d", pl, e->half_h[pl], cfg.h[pl]);
            delete e;
            return nullptr;
        }
        e->half_off[pl] = ho;
        e->full_off[pl] = fo;
        ho += (size_t)cfg.pitch[pl] * e->half_h[pl] * e->bps;
        fo += (size_t)cfg.pitch[pl] * cfg.h[pl] * e->bps;
    }
    e->half_bytes = ho;
    e->full_bytes = fo;
    const size_t tail = (size_t)kTailRows * cfg.pitch[0] * e->bps;
    bool ok = true;
    for (int k = 0; k < 4 && ok; k++)
    {
        ok = cudaMalloc(&e->half_mem[k], kLeadSlack + ho + tail) == cudaSuccess &&
             cudaMemset(e->half_mem[k], 0, kLeadSlack + ho + tail) == cudaSuccess;
    }
    for (int k = 0; k < 5 && ok; k++)
    {
        ok = cudaMalloc(&e->full_mem[k], kLeadSlack + fo + tail) == cudaSuccess &&
             cudaMemset(e->full_mem[k], 0, kLeadSlack + fo + tail) == cudaSuccess;
    }
    {
        const int need = cfg.w[0] * (int)(sizeof(LatticeTmp) + 2 * sizeof(int));
        if (need > 48 * 1024)
        {
            cudaFuncSetAttribute(k_lattice_b<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
            cudaFuncSetAttribute(k_lattice_b<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
        }
    }
    {
        size_t n[3], total = 0;
        for (int pl = 0; pl < 3; pl++) { n[pl] = (size_t)cfg.w[pl] * (cfg.h[pl] / 2 + 1); total += n[pl]; }
        if (ok) ok = cudaMalloc(&e->lattice_tmp, sizeof(LatticeTmp) * total) == cudaSuccess;
        // entries of direction-less samples are never written (and never used): keep them defined
        if (ok) ok = cudaMemset(e->lattice_tmp, 0, sizeof(LatticeTmp) * total) == cudaSuccess;
        e->lattice_tmp_pl[0] = e->lattice_tmp;
        e->lattice_tmp_pl[1] = e->lattice_tmp_pl[0] + n[0];
        e->lattice_tmp_pl[2] = e->lattice_tmp_pl[1] + n[1];
    }
    if (cfg.pp > 1)
    {
        size_t n[3], total = 0;
        for (int pl = 0; pl < 3; pl++) { n[pl] = ((size_t)cfg.pitch[pl] * (e->half_h[pl] + 1) + 31) / 32 * 32; total += 4 * n[pl]; }
        if (ok) ok = cudaMalloc(&e->deriv_mem, sizeof(int) * total) == cudaSuccess &&
                     cudaMemset(e->deriv_mem, 0, sizeof(int) * total) == cudaSuccess;
        int *p = e->deriv_mem;
        for (int pl = 0; pl < 3; pl++)
            for (int k = 0; k < 4; k++) { e->deriv[pl][k] = p; p += n[pl]; }
    }
    if (ok) ok = cudaStreamCreateWithFlags(&e->s_aux[0], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaStreamCreateWithFlags(&e->s_aux[1], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&e->ev_join[0], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&e->ev_join[1], cudaEventDisableTiming) == cudaSuccess;
    if (!ok)
    {
        set_error("eedi2: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        eedi2_destroy(e);
        return nullptr;
    }
    // eedi2_init_limlut (:23-33): ((pixel)limlut[i]) << shift, -1 entries wrap in the pixel type
    static const int base[33] = { 6, 6, 7, 7, 8, 8, 9, 9, 9, 10, 10, 11, 11, 12, 12, 12, 12, 12, 12, 12,
                                  12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, -1, -1 };
    const unsigned shift = cfg.depth - 8;
    for (int i = 0; i < 33; i++)
    {
        if (e->bps == 1) e->lim.v[i] = (int)(uint8_t)(((uint8_t)base[i]) << shift);
        else             e->lim.v[i] = (int)(uint16_t)(((uint16_t)base[i]) << shift);
    }
    return e;
}

void eedi2_destroy(Eedi2 *e)
{
    if (e == nullptr) return;
    if (e->timing && e->stage_calls > 0)
    {
        double tot = 0;
        for (int i = 0; i < 40; i++) tot += e->stage_ms[0][i] + e->stage_ms[1][i] + e->stage_ms[2][i];
        fprintf(stderr, "eedi2 stage times over %ld fields (us per field: luma / cb / cr), total %.1f us per field\n", e->stage_calls, 1e3 * tot / e->stage_calls);
        for (int i = 0; i < 40; i++)
        {
            if (e->stage_name[i] == nullptr) continue;
            char name[40];
            snprintf(name, sizeof(name), "%s", e->stage_name[i] + 1);
            for (char *c = name; *c; c++) if (*c == '<' && c[1] == '<') { *c = 0; break; }
            fprintf(stderr, "  %2d %-34s %8.1f %8.1f %8.1f\n", i, name, 1e3 * e->stage_ms[0][i] / e->stage_calls, 1e3 * e->stage_ms[1][i] / e->stage_calls,
                    1e3 * e->stage_ms[2][i] / e->stage_calls);
        }
    }
    if (e->tev[0]) cudaEventDestroy(e->tev[0]);
    if (e->tev[1]) cudaEventDestroy(e->tev[1]);
    for (auto &g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (int k = 0; k < 4; k++) if (e->half_mem[k]) cudaFree(e->half_mem[k]);
    for (int k = 0; k < 5; k++) if (e->full_mem[k]) cudaFree(e->full_mem[k]);
    if (e->lattice_tmp) cudaFree(e->lattice_tmp);
    if (e->deriv_mem) cudaFree(e->deriv_mem);
    for (int i = 0; i < 2; i++)
    {
        if (e->s_aux[i]) cudaStreamDestroy(e->s_aux[i]);
        if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
    }
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    delete e;
}

static int run_planes(Eedi2 *e, const void *const planes[3], int tff, cudaStream_t st)
{
    for (int pl = 0; pl < 3; pl++)
    {
        const int rc = e->bps == 1 ? run_plane<uint8_t>(e, pl, (const uint8_t *)planes[pl], tff, st)
                                   : run_plane<uint16_t>(e, pl, (const uint16_t *)planes[pl], tff, st);
        if (rc != 0) return rc;
    }
    return 0;
}

int eedi2_run(Eedi2 *e, const void *const planes[3], int tff, cudaStream_t st)
{
    e->stage_calls++;
    if (!e->use_graphs || e->timing || e->stop_after != 0) return run_planes(e, planes, tff, st);
    for (auto &g : e->graphs)
    {
        if (g.planes[0] == planes[0] && g.planes[1] == planes[1] && g.planes[2] == planes[2] && g.tff == tff)
        {
            if (cudaGraphLaunch(g.exec, st) != cudaSuccess)
            {
                set_error("eedi2: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return -1;
            }
            hbcu::count_launch(g.launches);
            return 0;
        }
    }
    // first time for this (source slot, field order): record the launches while they are issued
    if (e->graphs.size() >= 64) return run_planes(e, planes, tff, st);         // callers rotate a handful of slots; never hit
    const uint64_t before = hbcu::g_kernel_launches.load();
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
    {
        cudaGetLastError();
        return run_planes(e, planes, tff, st);
    }
    // The three planes never read outside their own plane of any work buffer (every stage clamps its horizontal reach to
    // the row and guards rows +-1 / +-2 / +-3), so they are independent: luma stays on `st`, the chroma planes fork into
    // two auxiliary streams and join again -- parallel branches of the captured graph.
    int rc = 0;
    bool forked = cudaEventRecord(e->ev_fork, st) == cudaSuccess;
    for (int pl = 1; pl < 3 && forked && rc == 0; pl++)
    {
        forked = cudaStreamWaitEvent(e->s_aux[pl - 1], e->ev_fork, 0) == cudaSuccess;
        if (!forked) break;
        rc = e->bps == 1 ? run_plane<uint8_t>(e, pl, (const uint8_t *)planes[pl], tff, e->s_aux[pl - 1])
                         : run_plane<uint16_t>(e, pl, (const uint16_t *)planes[pl], tff, e->s_aux[pl - 1]);
        forked = cudaEventRecord(e->ev_join[pl - 1], e->s_aux[pl - 1]) == cudaSuccess;
    }
    if (forked && rc == 0)
        rc = e->bps == 1 ? run_plane<uint8_t>(e, 0, (const uint8_t *)planes[0], tff, st) : run_plane<uint16_t>(e, 0, (const uint16_t *)planes[0], tff, st);
    for (int pl = 1; pl < 3 && forked; pl++) forked = cudaStreamWaitEvent(st, e->ev_join[pl - 1], 0) == cudaSuccess;
    if (!forked && rc == 0) rc = -2;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != 0 || ce != cudaSuccess || graph == nullptr)
    {
        if (graph) cudaGraphDestroy(graph);
        if (rc == 0) set_error("eedi2: stream capture failed: %s", cudaGetErrorString(ce));
        cudaGetLastError();
        return -1;
    }
    Eedi2::GraphEntry g;
    for (int pl = 0; pl < 3; pl++) g.planes[pl] = planes[pl];
    g.tff = tff;
    g.exec = nullptr;
    g.launches = (int)(hbcu::g_kernel_launches.load() - before);
    const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess)
    {
        set_error("eedi2: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
        cudaGetLastError();
        return -1;
    }
    e->graphs.push_back(g);
    if (cudaGraphLaunch(g.exec, st) != cudaSuccess)
    {
        set_error("eedi2: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    return 0;
}

// test hook: copies one EEDI2 work buffer (0-3 field buffers SRCPF..DSTPF, 4-8 frame buffers DST2PF..DST2MPF),
// all three planes with their strides, to the host
int eedi2_debug_read(const Eedi2 *e, int which, void *host, size_t host_bytes)
{
    const bool half = which < 4;
    const size_t n = half ? e->half_bytes : e->full_bytes;
    if (which < 0 || which > 8 || host_bytes < n)
    {
        set_error("eedi2_debug_read: bad buffer %d or size %zu < %zu", which, host_bytes, n);
        return -1;
    }
    const uint8_t *src = (half ? e->half_mem[which] : e->full_mem[which - 4]) + kLeadSlack;
    cudaError_t err = cudaMemcpy(host, src, n, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) { set_error("eedi2_debug_read: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

const void *eedi2_output(const Eedi2 *e, int plane)
{
    return e->full_mem[DST2PF] + kLeadSlack + e->full_off[plane];
}

}  // namespace hbcu
