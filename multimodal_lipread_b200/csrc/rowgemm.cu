// Linear layers on at most 32 rows (the heads: one row per clip of the batch -- classifier MLPs, the reverse LSTM
// direction's single step, fusion layers; reference call sites audio_video/models/middle_fusion_fast.py:20-25,38-39).
// As tile GEMMs these are one or a few CTAs walking K in a chain of dependent tiles: 17-26 us per launch for
// 3-20 MFLOP, all of it latency, all of it on the step's critical path between the LSTM and the loss.  Here the rows
// live in shared memory and the OUTPUT columns are spread over many CTAs:
//   forward : a warp owns an output column, its lanes split K (coalesced float4 reads of the weight row), 32 row
//             accumulators per lane, transposing butterfly (31 shuffles) -> lane m holds row m.
//   dgrad   : a CTA owns 32 input columns, a thread 4 of them for 16 rows, the reduction over the N outputs is split 16
//             ways inside the CTA and summed in a fixed order.
// fp32 FMA throughout, deterministic.
#include "nn_common.cuh"

namespace rg {

constexpr int TH = 256;
constexpr int RM = 32;               // rows held per CTA

__device__ __forceinline__ float reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const bool hi = lane & w;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = hi ? v[i] : v[i + w], keep = hi ? v[i + w] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    return v[0];
}

// y[m, n] = act(b[n] + sum_k x[m * ldx + k] * w[n * K + k]),  m < M <= 32.  grid: ceil(N / 8) CTAs of 8 warps.
__global__ void __launch_bounds__(TH)
linear_small_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                        const float* __restrict__ b, float* __restrict__ y, long long ldy, int M, int N, int K, int act) {
    extern __shared__ __align__(16) float X[];            // [32][K], rows >= M zero
    const int K4 = K >> 2;
#pragma unroll 8
    for (int e = threadIdx.x; e < RM * K4; e += TH) {
        const int m = e / K4, k = (e - m * K4) * 4;
        nn::st4(X + m * K + k, m < M ? __ldg(reinterpret_cast<const float4*>(x + m * ldx + k)) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * (TH / 32) + warp;
    // the first weight vector travels while the rows are staged
    float4 cur = make_float4(0.f, 0.f, 0.f, 0.f), nxt = cur;
    if (n < N && lane * 4 < K) cur = __ldg(reinterpret_cast<const float4*>(w + (size_t)n * K + lane * 4));
    __syncthreads();
    if (n >= N) return;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 128) {
        const int kn = k0 + 128 + lane * 4;
        if (kn < K) nxt = __ldg(reinterpret_cast<const float4*>(w + (size_t)n * K + kn));
        const int k = min(k0 + lane * 4, K - 4);             // lanes past the row end multiply zeros by valid data
        const bool live = k0 + lane * 4 < K;
        const float4 wv = live ? cur : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int m = 0; m < RM; ++m) {
            const float4 xv = nn::ld4(X + m * K + k);
            float a = v[m];
            a = fmaf(xv.x, wv.x, a); a = fmaf(xv.y, wv.y, a); a = fmaf(xv.z, wv.z, a); a = fmaf(xv.w, wv.w, a);
            v[m] = a;
        }
        cur = nxt;
    }
    const float d = reduce32(v, lane);
    if (lane < M) y[lane * ldy + n] = nn::act_fwd(d + (b ? b[n] : 0.f), act);
}

// dx[m, k] = sum_n dy[m * ldy + n] * w[n * K + k] (+ r[m * ldr + k]),  m < M <= 32.  grid: K / 32 CTAs.
// thread: k4 = tid & 7 (4 columns), mh = (tid >> 3) & 1 (rows 16 mh .. 16 mh + 15), part = tid >> 4 (slice of N).
__global__ void __launch_bounds__(TH)
linear_small_dgrad_kernel(const float* __restrict__ dy, long long ldy, const float* __restrict__ w,
                          float* __restrict__ dx, long long ldx, const float* __restrict__ r, long long ldr, int M,
                          int N, int K) {
    extern __shared__ __align__(16) float smd[];
    float* D = smd;                                        // [32][N], rows >= M zero
    float* red = D + RM * N;                               // [16][32][32]
    const int N4 = N >> 2;
#pragma unroll 8
    for (int e = threadIdx.x; e < RM * N4; e += TH) {
        const int m = e / N4, n = (e - m * N4) * 4;
        nn::st4(D + m * N + n, m < M ? __ldg(reinterpret_cast<const float4*>(dy + m * ldy + n)) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    __syncthreads();
    const int k4 = threadIdx.x & 7, mh = (threadIdx.x >> 3) & 1, part = threadIdx.x >> 4;
    const int kc = blockIdx.x * 32 + k4 * 4;
    const int nchunk = ((N4 + 15) / 16) * 4;
    const int n0 = part * nchunk, n1 = min(N, n0 + nchunk);
    float acc[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    for (int n = n0; n < n1; n += 4) {
        const float* wp = w + (size_t)n * K + kc;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + K));
        const float4 w2 = __ldg(reinterpret_cast<const float4*>(wp + 2 * (size_t)K));
        const float4 w3 = __ldg(reinterpret_cast<const float4*>(wp + 3 * (size_t)K));
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 d = nn::ld4(D + (mh * 16 + i) * N + n);
            acc[i][0] = fmaf(d.x, w0.x, acc[i][0]); acc[i][1] = fmaf(d.x, w0.y, acc[i][1]);
            acc[i][2] = fmaf(d.x, w0.z, acc[i][2]); acc[i][3] = fmaf(d.x, w0.w, acc[i][3]);
            acc[i][0] = fmaf(d.y, w1.x, acc[i][0]); acc[i][1] = fmaf(d.y, w1.y, acc[i][1]);
            acc[i][2] = fmaf(d.y, w1.z, acc[i][2]); acc[i][3] = fmaf(d.y, w1.w, acc[i][3]);
            acc[i][0] = fmaf(d.z, w2.x, acc[i][0]); acc[i][1] = fmaf(d.z, w2.y, acc[i][1]);
            acc[i][2] = fmaf(d.z, w2.z, acc[i][2]); acc[i][3] = fmaf(d.z, w2.w, acc[i][3]);
            acc[i][0] = fmaf(d.w, w3.x, acc[i][0]); acc[i][1] = fmaf(d.w, w3.y, acc[i][1]);
            acc[i][2] = fmaf(d.w, w3.z, acc[i][2]); acc[i][3] = fmaf(d.w, w3.w, acc[i][3]);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
        nn::st4(red + ((part * RM + mh * 16 + i) * 32 + k4 * 4), make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    __syncthreads();
    const int m = threadIdx.x >> 3;                        // 32 rows x 8 column groups
    if (m < M) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 t = nn::ld4(red + ((q * RM + m) * 32 + k4 * 4));
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        if (r) {
            const float4 t = nn::ld4(r + m * ldr + kc);
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        nn::st4(dx + m * ldx + kc, s);
    }
}

}  // namespace rg

extern "C" int lr_linear_small_fwd(const float* x, long long ldx, const float* w, const float* b, float* y, long long ldy,
                                   int M, int N, int K, int act, lr_stream_t stream) {
    LR_CHECK_ARG(M >= 0 && M <= rg::RM && N > 0 && K >= 4 && (K & 3) == 0 && (ldx & 3) == 0 && ldx >= K && ldy >= N,
                 "lr_linear_small_fwd: bad shape (M <= 32, K and ldx multiples of 4)");
    if (M == 0) return LR_OK;
    LR_CHECK_ARG(x && w && y, "lr_linear_small_fwd: null pointer");
    LR_CHECK_ALIGN(x); LR_CHECK_ALIGN(w);
    const size_t smem = (size_t)rg::RM * K * sizeof(float);
    LR_CHECK_ARG(smem <= 200 * 1024, "lr_linear_small_fwd: K too large for the shared-memory row block (K <= 1600)");
    cudaError_t e = lr::ensure_max_dynamic_smem(rg::linear_small_fwd_kernel, (int)smem);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "lr_linear_small_fwd: %s", cudaGetErrorString(e));
    const int grid = (N + rg::TH / 32 - 1) / (rg::TH / 32);
    rg::linear_small_fwd_kernel<<<grid, rg::TH, smem, stream>>>(x, ldx, w, b, y, ldy, M, N, K, act);
    lr::count_launch();
    LR_CHECK_LAUNCH("linear_small_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_linear_small_dgrad(const float* dy, long long ldy, const float* w, float* dx, long long ldx,
                                     const float* r, long long ldr, int M, int N, int K, lr_stream_t stream) {
    LR_CHECK_ARG(M >= 0 && M <= rg::RM && N >= 4 && (N & 3) == 0 && K >= 32 && (K & 31) == 0 && (ldy & 3) == 0 &&
                 (ldx & 3) == 0 && ldy >= N && ldx >= K && (!r || ((ldr & 3) == 0 && ldr >= K)),
                 "lr_linear_small_dgrad: bad shape (M <= 32, N multiple of 4, K multiple of 32, pitches multiples of 4)");
    if (M == 0) return LR_OK;
    LR_CHECK_ARG(dy && w && dx, "lr_linear_small_dgrad: null pointer");
    LR_CHECK_ALIGN(dy); LR_CHECK_ALIGN(w); LR_CHECK_ALIGN(dx); LR_CHECK_ALIGN(r);
    const size_t smem = ((size_t)rg::RM * N + 16 * rg::RM * 32) * sizeof(float);
    LR_CHECK_ARG(smem <= 200 * 1024, "lr_linear_small_dgrad: N too large for the shared-memory row block (N <= 1088)");
    cudaError_t e = lr::ensure_max_dynamic_smem(rg::linear_small_dgrad_kernel, (int)smem);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "lr_linear_small_dgrad: %s", cudaGetErrorString(e));
    rg::linear_small_dgrad_kernel<<<K / 32, rg::TH, smem, stream>>>(dy, ldy, w, dx, ldx, r, ldr, M, N, K);
    lr::count_launch();
    LR_CHECK_LAUNCH("linear_small_dgrad_kernel");
    return LR_OK;
}
