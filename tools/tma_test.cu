// tools/tma_test.cu -- minimal TMA 2-D tile load probe (debug aid, not product)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap map, uint8_t *out, int box_bytes, int x, int y)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = (uint64_t *)(smem + 200 * 1024);
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(smem)), "l"(&map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)) : "memory");
    for (int i = threadIdx.x; i < box_bytes; i += blockDim.x) out[i] = smem[i];
}

int main(int argc, char **argv)
{
    int bw = atoi(argv[1]), bh = atoi(argv[2]), esz = atoi(argv[3]);
    int W = 700, H = 400, pitch = 768 * esz;
    uint8_t *d, *o;
    cudaMalloc(&d, (size_t)pitch * H); cudaMalloc(&o, 256 * 1024);
    std::vector<uint8_t> hbuf((size_t)pitch * H);
    for (size_t i = 0; i < hbuf.size(); i++) hbuf[i] = (uint8_t)(i * 7 + i / pitch);
    cudaMemcpy(d, hbuf.data(), hbuf.size(), cudaMemcpyHostToDevice);
    void *fp; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fp;
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H}, gstr[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
    CUresult r = enc(&map, esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d esz %d encode=%d ", bw, bh, esz, (int)r);
    int bytes = bw * bh * esz;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    int cx = argc > 4 ? atoi(argv[4]) : 8, cy = argc > 5 ? atoi(argv[5]) : 8; k<<<1, 128, 201 * 1024>>>(map, o, bytes, cx, cy);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess)
    {
        std::vector<uint8_t> ho(bytes); cudaMemcpy(ho.data(), o, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < bh; yy++) for (int xx = 0; xx < bw * esz; xx++)
            if (ho[yy * bw * esz + xx] != hbuf[(size_t)(cy + yy) * pitch + cx * esz + xx]) bad++;
        printf("mismatch=%d", bad);
    }
    printf("\n");
    return 0;
}
