// hbcu_core.cu -- runtime part of the C-ABI (include/hbcu.h): errors, device
// discovery, pinned host memory, launch counter, TMA descriptor encoding.
#include "hbcu_common.h"
#include "../../include/hbcu.h"

#include <cstring>
#include <mutex>
#include <vector>

namespace hbcu {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_kernel_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tensor_map_2d(CUtensorMap *map, int elem_bytes, void *base,
                         uint64_t width_elems, uint64_t height, uint64_t pitch_bytes,
                         uint32_t box_w, uint32_t box_h)
{
    static encode_tiled_fn fn = nullptr;
    if (fn == nullptr)
    {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || p == nullptr || q != cudaDriverEntryPointSuccess)
        {
            set_error("cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
            return -1;
        }
        fn = (encode_tiled_fn)p;
    }
    CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
    cuuint64_t gdim[2]    = { width_elems, height };
    cuuint64_t gstride[1] = { pitch_bytes };          // stride of dim 1 in bytes (dim 0 is dense)
    cuuint32_t box[2]     = { box_w, box_h };
    cuuint32_t estr[2]    = { 1, 1 };
    CUresult r = fn(map, dt, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (w=%llu h=%llu pitch=%llu box=%ux%u)", (int)r,
                  (unsigned long long)width_elems, (unsigned long long)height,
                  (unsigned long long)pitch_bytes, box_w, box_h);
        return -1;
    }
    return 0;
}

}  // namespace hbcu

extern "C" {

int hbcu_abi_version(void) { return HBCU_ABI_VERSION; }

const char *hbcu_last_error(void) { return hbcu::t_error; }

int hbcu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// Pinned-memory pools: size classes 2^10 .. 2^31, one free list each, a 64-byte
// header in front of the payload remembers the class.
namespace {
struct PinHeader { uint32_t magic; uint32_t cls; PinHeader *next; char pad[64 - 16]; };
static_assert(sizeof(PinHeader) == 64, "header keeps the payload 64-byte aligned");
constexpr uint32_t kPinMagic = 0x48424355u;   // 'HBCU'
std::mutex g_pin_lock;
PinHeader *g_pin_free[2][32] = { { nullptr }, { nullptr } };      // [write-combined][size class]
constexpr uint32_t kPinWcBit = 0x100u;                            // header cls bit: the block is write-combined memory
int g_pin_wc = 0;                                                 // hbcu_host_set_write_combined
}

// Write-combined pinned memory for buffers the CPU only WRITES and the GPU only reads (a decoder's output frames): it is
// not snooped, so device reads over PCIe leave the host's caches and coherency traffic alone.  CPU reads from it are very
// slow, so it is a per-allocation choice made by the caller, never the default.
void hbcu_host_set_write_combined(int on)
{
    std::lock_guard<std::mutex> g(g_pin_lock);
    g_pin_wc = on != 0;
}

void *hbcu_host_alloc(size_t bytes)
{
    uint32_t cls = 10;
    while (((size_t)1 << cls) < bytes + sizeof(PinHeader) && cls < 31) cls++;
    if (((size_t)1 << cls) < bytes + sizeof(PinHeader))
    {
        hbcu::set_error("hbcu_host_alloc: %zu bytes is too large", bytes);
        return nullptr;
    }
    int wc;
    {
        std::lock_guard<std::mutex> g(g_pin_lock);
        wc = g_pin_wc;
        if (g_pin_free[wc][cls] != nullptr)
        {
            PinHeader *h = g_pin_free[wc][cls];
            g_pin_free[wc][cls] = h->next;
            return (void *)(h + 1);
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)1 << cls, cudaHostAllocPortable | (wc ? cudaHostAllocWriteCombined : 0)) != cudaSuccess)
    {
        hbcu::set_error("cudaHostAlloc(%zu) failed: %s", (size_t)1 << cls, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    PinHeader *h = (PinHeader *)p;
    h->magic = kPinMagic;
    h->cls = cls | (wc ? kPinWcBit : 0u);
    h->next = nullptr;
    return (void *)(h + 1);
}

void hbcu_host_free(void *p)
{
    if (p == nullptr) return;
    PinHeader *h = (PinHeader *)p - 1;
    if (h->magic != kPinMagic) return;   // not ours: refuse to touch it
    std::lock_guard<std::mutex> g(g_pin_lock);
    const int wc = (h->cls & kPinWcBit) != 0;
    h->next = g_pin_free[wc][h->cls & 31u];
    g_pin_free[wc][h->cls & 31u] = h;
}

int hbcu_host_reserve(size_t bytes, int count)
{
    std::vector<void *> blocks;
    for (int i = 0; i < count; i++)
    {
        // bypass the free list so that `count` NEW blocks are created
        uint32_t cls = 10;
        while (((size_t)1 << cls) < bytes + sizeof(PinHeader) && cls < 31) cls++;
        void *p = nullptr;
        if (cudaHostAlloc(&p, (size_t)1 << cls, cudaHostAllocPortable) != cudaSuccess)
        {
            cudaGetLastError();
            break;
        }
        PinHeader *h = (PinHeader *)p;
        h->magic = kPinMagic;
        h->cls = cls;
        h->next = nullptr;
        blocks.push_back((void *)(h + 1));
    }
    for (void *b : blocks) hbcu_host_free(b);
    return (int)blocks.size();
}

void hbcu_host_trim(void)
{
    std::lock_guard<std::mutex> g(g_pin_lock);
    for (int wc = 0; wc < 2; wc++)
        for (int c = 0; c < 32; c++)
        {
            while (g_pin_free[wc][c] != nullptr)
            {
                PinHeader *h = g_pin_free[wc][c];
                g_pin_free[wc][c] = h->next;
                cudaFreeHost(h);
            }
        }
}

uint64_t hbcu_kernel_launches(void) { return hbcu::g_kernel_launches.load(); }

}  // extern "C"
