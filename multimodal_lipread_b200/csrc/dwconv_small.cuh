// Depthwise convolutions on SMALL images (H = W <= 6), the last seven of MobileNetV3-small's eleven depthwise layers
// at 88-pixel lip frames (6x6 and 3x3 maps under 5x5 windows).  There the window covers most of the image: the tiled
// kernels of dwconv.cu stage halos that are mostly padding and reach 0.6-1.2 TB/s.  Here one thread owns one
// (frame, channel): it loads the whole H x H map of its channel into registers (every load coalesced over 64
// consecutive channels), and the compile-time geometry unrolls into straight-line FMAs over the VALID taps only.
#pragma once
#include "nn_common.cuh"

namespace dws {

constexpr int TH = 256, CW = 64, FL = TH / CW;     // 64 channels x 4 frame lanes per block

template <int H, int K, int S>
struct G {
    static constexpr int P = K / 2, Ho = (H + 2 * P - K) / S + 1;
};

template <typename T, int H, int K, int S>
__global__ void __launch_bounds__(TH)
fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y, double* __restrict__ stats,
           int F, int C) {
    constexpr int P = G<H, K, S>::P, Ho = G<H, K, S>::Ho;
    __shared__ float ssum[FL][CW], ssq[FL][CW];
    const int cl = threadIdx.x % CW, fl = threadIdx.x / CW;
    const int c = blockIdx.y * CW + cl;
    const bool ok = c < C;
    float wr[K * K];
#pragma unroll
    for (int t = 0; t < K * K; ++t) wr[t] = ok ? w[c * K * K + t] : 0.f;
    float ls = 0.f, lq = 0.f;
    for (int f = blockIdx.x * FL + fl; f < F && ok; f += gridDim.x * FL) {
        const T* xf = x + (long long)f * H * H * C + c;
        float xr[H * H];
#pragma unroll
        for (int i = 0; i < H * H; ++i) xr[i] = nn::ld1(xf + (long long)i * C);
        T* yf = y + (long long)f * Ho * Ho * C + c;
#pragma unroll
        for (int ho = 0; ho < Ho; ++ho)
#pragma unroll
            for (int wo = 0; wo < Ho; ++wo) {
                float acc = 0.f;
#pragma unroll
                for (int kh = 0; kh < K; ++kh) {
                    const int hi = ho * S - P + kh;
                    if (hi < 0 || hi >= H) continue;
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) {
                        const int wi = wo * S - P + kw;
                        if (wi < 0 || wi >= H) continue;
                        acc = fmaf(xr[hi * H + wi], wr[kh * K + kw], acc);
                    }
                }
                nn::st1(yf + (long long)(ho * Ho + wo) * C, acc);
                if (sizeof(T) == 2) acc = __bfloat162float(__float2bfloat16_rn(acc));   // statistics of the STORED value
                ls += acc; lq = fmaf(acc, acc, lq);
            }
    }
    if (stats) {
        ssum[fl][cl] = ls; ssq[fl][cl] = lq;
        __syncthreads();
        if (fl == 0 && ok) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int i = 0; i < FL; ++i) { a += ssum[i][cl]; b += ssq[i][cl]; }
            nn::atomic_add_double(stats + c, (double)a);
            nn::atomic_add_double(stats + C + c, (double)b);
        }
    }
}

template <typename T, int H, int K, int S>
__global__ void __launch_bounds__(TH)
dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx, int F, int C) {
    constexpr int P = G<H, K, S>::P, Ho = G<H, K, S>::Ho;
    const int cl = threadIdx.x % CW, fl = threadIdx.x / CW;
    const int c = blockIdx.y * CW + cl;
    if (c >= C) return;
    float wr[K * K];
#pragma unroll
    for (int t = 0; t < K * K; ++t) wr[t] = w[c * K * K + t];
    for (int f = blockIdx.x * FL + fl; f < F; f += gridDim.x * FL) {
        const T* gf = dy + (long long)f * Ho * Ho * C + c;
        float gr[Ho * Ho];
#pragma unroll
        for (int i = 0; i < Ho * Ho; ++i) gr[i] = nn::ld1(gf + (long long)i * C);
        T* xf = dx + (long long)f * H * H * C + c;
#pragma unroll
        for (int hi = 0; hi < H; ++hi)
#pragma unroll
            for (int wi = 0; wi < H; ++wi) {
                float acc = 0.f;
#pragma unroll
                for (int ho = 0; ho < Ho; ++ho) {
                    const int kh = hi - ho * S + P;
                    if (kh < 0 || kh >= K) continue;
#pragma unroll
                    for (int wo = 0; wo < Ho; ++wo) {
                        const int kw = wi - wo * S + P;
                        if (kw < 0 || kw >= K) continue;
                        acc = fmaf(gr[ho * Ho + wo], wr[kh * K + kw], acc);
                    }
                }
                nn::st1(xf + (long long)(hi * H + wi) * C, acc);
            }
    }
}

template <typename T, int H, int K, int S>
__global__ void __launch_bounds__(TH)
wgrad_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* __restrict__ dwt, int F, int C) {
    constexpr int P = G<H, K, S>::P, Ho = G<H, K, S>::Ho;
    __shared__ float red[FL][CW];
    const int cl = threadIdx.x % CW, fl = threadIdx.x / CW;
    const int c = blockIdx.y * CW + cl;
    const bool ok = c < C;
    float acc[K * K];
#pragma unroll
    for (int t = 0; t < K * K; ++t) acc[t] = 0.f;
    for (int f = blockIdx.x * FL + fl; f < F && ok; f += gridDim.x * FL) {
        const T* xf = x + (long long)f * H * H * C + c;
        const T* gf = dy + (long long)f * Ho * Ho * C + c;
        float xr[H * H], gr[Ho * Ho];
#pragma unroll
        for (int i = 0; i < H * H; ++i) xr[i] = nn::ld1(xf + (long long)i * C);
#pragma unroll
        for (int i = 0; i < Ho * Ho; ++i) gr[i] = nn::ld1(gf + (long long)i * C);
#pragma unroll
        for (int kh = 0; kh < K; ++kh)
#pragma unroll
            for (int kw = 0; kw < K; ++kw) {
                float a = acc[kh * K + kw];
#pragma unroll
                for (int ho = 0; ho < Ho; ++ho) {
                    const int hi = ho * S - P + kh;
                    if (hi < 0 || hi >= H) continue;
#pragma unroll
                    for (int wo = 0; wo < Ho; ++wo) {
                        const int wi = wo * S - P + kw;
                        if (wi < 0 || wi >= H) continue;
                        a = fmaf(gr[ho * Ho + wo], xr[hi * H + wi], a);
                    }
                }
                acc[kh * K + kw] = a;
            }
    }
    // reduce the FL frame lanes of the block, then one atomic per (channel, tap) and block
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
        red[fl][cl] = acc[t];
        __syncthreads();
        if (fl == 0 && ok) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < FL; ++i) a += red[i][cl];
            atomicAdd(&dwt[c * K * K + t], a);
        }
        __syncthreads();
    }
}

// blocks_per_sm: 4 for the streaming kernels; 1 for wgrad, whose every block ends in C/64 * K*K atomics
inline dim3 grid_for(int F, int C, int blocks_per_sm = 4) {
    const int chunks = (C + CW - 1) / CW;
    int gx = (lr::sm_count() * blocks_per_sm + chunks - 1) / chunks;
    const int maxgx = (F + FL - 1) / FL;
    if (gx > maxgx) gx = maxgx;
    if (gx < 1) gx = 1;
    return dim3((unsigned)gx, (unsigned)chunks);
}

// mode 0 fwd (a = x, b = w, out = y, stats), 1 dgrad (a = dy, b = w, out = dx), 2 wgrad (a = dy, b = x, out = dw)
// (the activation operands are T, the weights and the weight gradient are always fp32)
template <typename T, int H, int K, int S>
inline void launch(int mode, const void* a, const void* b, void* out, double* stats, int F, int C, cudaStream_t st) {
    const dim3 grid = grid_for(F, C, mode == 2 ? 1 : 4);
    if (mode == 0) fwd_kernel<T, H, K, S><<<grid, TH, 0, st>>>(static_cast<const T*>(a), static_cast<const float*>(b), static_cast<T*>(out), stats, F, C);
    else if (mode == 1) dgrad_kernel<T, H, K, S><<<grid, TH, 0, st>>>(static_cast<const T*>(a), static_cast<const float*>(b), static_cast<T*>(out), F, C);
    else wgrad_kernel<T, H, K, S><<<grid, TH, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(b), static_cast<float*>(out), F, C);
}

// true if a specialised kernel exists (and was launched) for this geometry
template <typename T>
inline bool dispatch(int mode, const void* a, const void* b, void* out, double* stats, int F, int H, int W, int C, int k,
                     int stride, cudaStream_t st) {
    if (H != W) return false;
#define DWS_CASE(H_, K_, S_) if (H == H_ && k == K_ && stride == S_) { launch<T, H_, K_, S_>(mode, a, b, out, stats, F, C, st); return true; }
    DWS_CASE(6, 5, 1) DWS_CASE(6, 5, 2) DWS_CASE(3, 5, 1) DWS_CASE(3, 5, 2) DWS_CASE(2, 5, 1) DWS_CASE(6, 3, 1)
#undef DWS_CASE
    return false;
}

}  // namespace dws
