// copy_bench.cu -- how fast do frame-sized host<->device copies go on this box, by pattern?
//   nvcc -O2 -o /tmp/copy_bench tools/copy_bench.cu && /tmp/copy_bench
// Patterns: one contiguous copy per 4K 8-bit frame (12.44 MB) vs three per-plane cudaMemcpy2DAsync (8.3 + 2.07 + 2.07 MB),
// one direction alone vs both directions at once (separate streams), 64 frames each.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main(int argc, char **argv)
{
    const int W = 3840, H = 2160;
    const int NBUF = argc > 1 ? atoi(argv[1]) : 8;      // distinct pinned frames (working set)
    const int N = NBUF > 64 ? NBUF : 64;
    const int NDEV = 8;
    const size_t Y = (size_t)W * H, Cb = Y / 4, F = Y + 2 * Cb;
    std::vector<unsigned char *> hin(NBUF), hout(NBUF), din(NDEV), dout(NDEV);
    for (int i = 0; i < NBUF; i++)
    {
        CK(cudaHostAlloc(&hin[i], F, cudaHostAllocDefault));
        CK(cudaHostAlloc(&hout[i], F, cudaHostAllocDefault));
        for (size_t k = 0; k < F; k += 4096) hin[i][k] = (unsigned char)k;
    }
    for (int i = 0; i < NDEV; i++)
    {
        CK(cudaMalloc(&din[i], F));
        CK(cudaMalloc(&dout[i], F));
    }
    printf("working set: %d + %d pinned frames of %.2f MB\n", NBUF, NBUF, F / 1e6);
    cudaStream_t s_in, s_out;
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, f0, f1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));

    auto copy_frame = [&](bool h2d, int i, int mode, cudaStream_t st) {
        unsigned char *h = h2d ? hin[i] : hout[i], *d = h2d ? din[i % NDEV] : dout[i % NDEV];
        const cudaMemcpyKind kind = h2d ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        void *dst = h2d ? (void *)d : (void *)h;
        const void *src = h2d ? (const void *)h : (const void *)d;
        if (mode == 0) { CK(cudaMemcpyAsync(dst, src, F, kind, st)); return; }
        const size_t off[3] = { 0, Y, Y + Cb };
        const int w[3] = { W, W / 2, W / 2 }, hh[3] = { H, H / 2, H / 2 };
        for (int p = 0; p < 3; p++)
        {
            if (mode == 1) CK(cudaMemcpy2DAsync((char *)dst + off[p], w[p], (const char *)src + off[p], w[p], w[p], hh[p], kind, st));
            else           CK(cudaMemcpyAsync((char *)dst + off[p], (const char *)src + off[p], (size_t)w[p] * hh[p], kind, st));
        }
    };
    const char *mname[3] = { "1 contiguous copy/frame", "3 x cudaMemcpy2DAsync/frame", "3 x cudaMemcpyAsync/frame" };
    for (int mode = 0; mode < 3; mode++)
    {
        for (int dir = 0; dir < 3; dir++)      // 0 h2d only, 1 d2h only, 2 both
        {
            for (int rep = 0; rep < 2; rep++)
            {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0, s_in)); CK(cudaEventRecord(f0, s_out));
                for (int i = 0; i < N; i++)
                {
                    if (dir != 1) copy_frame(true, i % NBUF, mode, s_in);
                    if (dir != 0) copy_frame(false, i % NBUF, mode, s_out);
                }
                CK(cudaEventRecord(e1, s_in)); CK(cudaEventRecord(f1, s_out));
                CK(cudaDeviceSynchronize());
                float a = 0, b = 0;
                CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, f0, f1));
                if (rep == 1)
                    printf("%-30s %-8s h2d %6.1f GB/s (%.3f ms/frame)  d2h %6.1f GB/s (%.3f ms/frame)\n", mname[mode],
                           dir == 0 ? "h2d" : dir == 1 ? "d2h" : "both",
                           dir != 1 ? N * F / a / 1e6 : 0.0, dir != 1 ? a / N : 0.0, dir != 0 ? N * F / b / 1e6 : 0.0, dir != 0 ? b / N : 0.0);
            }
        }
    }
    return 0;
}
