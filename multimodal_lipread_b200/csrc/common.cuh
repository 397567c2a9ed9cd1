// Shared helpers for the lipread_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/lipread_b200.h"

namespace lr {

// Thread-local message for the last non-zero status (lr_last_error()).
char* last_error_buf();
int fail(int code, const char* fmt, ...);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define LR_CHECK_ARG(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) return ::lr::fail(LR_EINVAL, __VA_ARGS__);   \
    } while (0)

#define LR_CHECK_ALIGN(ptr)                                                              \
    do {                                                                                 \
        if (!::lr::aligned16(ptr)) return ::lr::fail(LR_EALIGN, "%s is not 16-byte aligned", #ptr); \
    } while (0)

// Launch-error check that never synchronises (safe under CUDA-graph capture).
#define LR_CHECK_LAUNCH(name)                                                            \
    do {                                                                                 \
        cudaError_t e__ = cudaPeekAtLastError();                                         \
        if (e__ != cudaSuccess) {                                                        \
            (void)cudaGetLastError();                                                    \
            return ::lr::fail(LR_ECUDA, "%s: %s", name, cudaGetErrorString(e__));        \
        }                                                                                \
    } while (0)

void count_launch(int n = 1);
// Raise a kernel's opt-in dynamic shared-memory limit to `bytes` once per (device, kernel); thread-safe, and a later
// call with a smaller size never lowers a limit that captured graph nodes rely on.
cudaError_t ensure_max_dynamic_smem(const void* func, int bytes);
template <typename F> inline cudaError_t ensure_max_dynamic_smem(F* func, int bytes) {
    return ensure_max_dynamic_smem(reinterpret_cast<const void*>(func), bytes);
}
int sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device

// uint8 pixel -> float.  The reference divides (`lip_regions.astype(np.float32) / 255.0`,
// video/data_utils/dataset_loader.py:90): x / 255 and x * (1/255) differ in the last bit for some x, so the
// canonical scale 1/255 is applied as a true division; any other scale is a multiplication.
__device__ __forceinline__ float u8_scaled(unsigned char v, float scale) {
    return scale == (1.f / 255.f) ? __fdiv_rn((float)v, 255.f) : (float)v * scale;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace lr
