// Device helpers shared by the trunk / head kernels: activations, vector access, and the
// "fixed channel group per thread" mapping used by every [rows, C] channels-last kernel.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace nn {

// LR_ACT_* in include/lipread_b200.h
__device__ __forceinline__ float act_fwd(float u, int act) {
    switch (act) {
        case LR_ACT_RELU: return fmaxf(u, 0.f);
        case LR_ACT_HSWISH: return u * fminf(fmaxf(u + 3.f, 0.f), 6.f) * (1.f / 6.f);
        case LR_ACT_HSIGMOID: return fminf(fmaxf(u + 3.f, 0.f), 6.f) * (1.f / 6.f);
        case LR_ACT_RELU6: return fminf(fmaxf(u, 0.f), 6.f);
        default: return u;
    }
}
// d act(u) / du evaluated at the pre-activation u (torch: hardswish_backward, hardsigmoid_backward,
// threshold_backward conventions at the kinks).
__device__ __forceinline__ float act_grad(float u, int act) {
    switch (act) {
        case LR_ACT_RELU: return u > 0.f ? 1.f : 0.f;
        case LR_ACT_HSWISH: return u < -3.f ? 0.f : (u <= 3.f ? u * (1.f / 3.f) + 0.5f : 1.f);
        case LR_ACT_HSIGMOID: return (u > -3.f && u < 3.f) ? (1.f / 6.f) : 0.f;
        case LR_ACT_RELU6: return (u > 0.f && u < 6.f) ? 1.f : 0.f;
        default: return 1.f;
    }
}
// derivative expressed through the OUTPUT y = act(u) (only for monotone acts where that is possible)
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
    switch (act) {
        case LR_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case LR_ACT_HSIGMOID: return (y > 0.f && y < 1.f) ? (1.f / 6.f) : 0.f;
        case LR_ACT_RELU6: return (y > 0.f && y < 6.f) ? 1.f : 0.f;
        default: return 1.f;
    }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---- bf16 storage (precision "bf16": activations and their gradients live in HBM as bfloat16, all arithmetic stays
// fp32 in registers).  The kernels are templates over the storage type T in {float, bf16}; these overloads are the
// only place the type shows: 4 consecutive channels are one 8-byte access instead of one 16-byte access.
typedef __nv_bfloat16 bf16;
__device__ __forceinline__ float4 unpack4(uint2 u) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 pack4(float4 v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const unsigned*>(&a); u.y = *reinterpret_cast<const unsigned*>(&b);
    return u;
}
__device__ __forceinline__ float4 ld4(const bf16* p) { return unpack4(*reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ void st4(bf16* p, float4 v) { *reinterpret_cast<uint2*>(p) = pack4(v); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// L2 residency hints for two-pass kernels: the first pass loads with evict_last so that the second pass (the next
// kernel) finds the tensor in the 126 MB L2 instead of HBM; the second pass loads with evict_first.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld4_hint(const float* p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld4_hint(const bf16* p, unsigned long long pol) {
    uint2 u;
    asm volatile("ld.global.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(u.x), "=r"(u.y) : "l"(p), "l"(pol));
    return unpack4(u);
}

// Raw 4-channel vectors: what a streaming kernel keeps IN FLIGHT (float4 = 4 registers, bf16 = 2 registers), converted
// to float4 only at the point of use so that doubling the rows in flight for bf16 costs no extra registers.
template <typename T> struct Raw4 { typedef float4 type; };
template <> struct Raw4<bf16> { typedef uint2 type; };
__device__ __forceinline__ float4 ldraw(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ uint2 ldraw(const bf16* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ float4 ldraw_hint(const float* p, unsigned long long pol) { return ld4_hint(p, pol); }
__device__ __forceinline__ uint2 ldraw_hint(const bf16* p, unsigned long long pol) {
    uint2 u;
    asm volatile("ld.global.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(u.x), "=r"(u.y) : "l"(p), "l"(pol));
    return u;
}
__device__ __forceinline__ float4 cvt4(float4 v) { return v; }
__device__ __forceinline__ float4 cvt4(uint2 u) { return unpack4(u); }

// Channel-group mapping for a row-major [rows, C] tensor with C % 4 == 0: a thread owns ONE group of
// 4 consecutive channels (so per-channel parameters live in registers) and strides over rows.
// Consecutive threads touch consecutive float4s, so every warp access is fully coalesced.
struct CgMap {
    int ncg;        // channel groups per row = C / 4
    int cg;         // this thread's group (valid if active)
    int rlane;      // this thread's row lane
    int rpp;        // rows per pass of the block
    bool active;
    // maxw: channel groups per block column (0: as many as the block has threads).  A narrow column (32 groups = 128
    // channels) keeps all threads busy for wide layers (row lanes = threads / width) and shrinks the per-block
    // statistic atomics to the column's channels.
    __device__ __forceinline__ CgMap(int C, int cg0 /*first group handled by this block column*/, int maxw = 0) {
        ncg = C >> 2;
        const int w = min(ncg - cg0, maxw > 0 ? maxw : (int)blockDim.x);     // groups handled by this block column
        rpp = blockDim.x / w;
        cg = cg0 + threadIdx.x % w;
        rlane = threadIdx.x / w;
        active = rlane < rpp;
    }
};
// number of block columns needed so that every channel group is owned by some thread
inline int cg_block_cols(int C, int threads) { return ((C >> 2) + threads - 1) / threads; }
// column width (in channel groups) of the streaming [rows, C] kernels: the whole row up to 32 groups, else 32
inline int cg_col_width(int C) { const int ncg = C >> 2; return ncg <= 32 ? ncg : 32; }

__device__ __forceinline__ void atomic_add_double(double* p, double v) { atomicAdd(p, v); }

}  // namespace nn
