// hbcu_common.h -- shared plumbing of the CUDA side (error reporting, launch
// counting, TMA descriptor helper).  Internal; the public surface is include/hbcu.h.
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <atomic>

namespace hbcu {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_kernel_launches;

inline void count_launch(int n = 1) { g_kernel_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency)
int encode_tensor_map_2d(CUtensorMap *map, int elem_bytes, void *base,
                         uint64_t width_elems, uint64_t height, uint64_t pitch_bytes,
                         uint32_t box_w, uint32_t box_h);

#define HBCU_CHECK(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            hbcu::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                            __FILE__, __LINE__);                                          \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

}  // namespace hbcu
