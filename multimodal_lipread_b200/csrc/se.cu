// Squeeze-Excitation gate on the per-frame pooled vectors, one launch per direction:
//   forward : s = act2(W2 . act1(W1 . p + b1) + b2)                       (torchvision SqueezeExcitation._scale after the
//             avgpool: fc1 -> ReLU -> fc2 -> Hardsigmoid; reference call sites middle_fusion_fast.py:15-17,34)
//   backward: dz2 = ds * act2'(s),  dz1 = (dz2 . W2) * act1'(h1),  dp = dz1 . W1
// The two 1x1 convolutions are [frames x C] x [C x C/4] products on 928 rows: as GEMM launches they were two
// (forward) and four (backward: two activation derivatives, two input-gradient GEMMs) latency-bound ~10 us kernels per
// SE block on the step's critical path.  Here a CTA owns R = 8 frames, keeps their vectors in shared memory and walks
// the weights (L2-resident, <= 332 KB) once; everything is fp32 FMA in a fixed order, so the forward is bit-reproducible.
// The weight gradients (dW = dz^T x, db = colsum dz) stay separate launches on the parallel branches of the step graph:
// the backward kernel leaves dz2 (in place of ds) and dz1 in global memory for them.
#include "nn_common.cuh"

namespace se {

constexpr int R = 8;          // frames per CTA
constexpr int TH = 512;
constexpr int RED_FLOATS = 16384;     // 64 KB of partial sums for the split-K stages

// Every lane holds 32 partial sums v[r * 4 + o] (frame r < 8, output column o < 4).  Each exchange step halves the
// number of values a lane is responsible for; after five steps lane l holds the complete sum number l
// (frame l >> 2, column l & 3).  31 shuffles instead of 160, fixed order.
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

// Y[r][n] = act(bias[n] + sum_k X[r][k] * W[n * K + k]), r < R.  X, Y in shared memory (row pitches K and N), N % 4 == 0.
// A warp owns four output columns at a time: its lanes split k in float4 steps (coalesced weight reads, every x vector
// read from shared memory feeds 16 FMAs), then reduce32.  The (column group, k pass) items of a warp form one sequence
// whose next weight vectors are already in flight while the current ones are used: the loop is bound by the L2
// latency of the weights otherwise.
__device__ __forceinline__ void stage_nt(const float* __restrict__ X, const float* __restrict__ W,
                                         const float* __restrict__ bias, int N, int K, int act, float* __restrict__ Y) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = TH / 32;
    const int npass = (K + 127) >> 7, ngroups = N >> 2;
    if (warp >= ngroups) return;
    auto load = [&](int g, int q, float4 (&w)[4]) {
        const int k = q * 128 + lane * 4;
#pragma unroll
        for (int o = 0; o < 4; ++o)
            w[o] = k < K ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(g * 4 + o) * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 cur[4], nxt[4];
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    int g = warp, q = 0;
    load(g, q, cur);
    while (g < ngroups) {
        int gn = g, qn = q + 1;
        if (qn == npass) { qn = 0; gn = g + NW; }
        if (gn < ngroups) load(gn, qn, nxt);
        const int k = min(q * 128 + lane * 4, K - 4);          // lanes past the row end multiply zeros by valid data
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 x = nn::ld4(X + r * K + k);
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                float a = v[r * 4 + o];
                a = fmaf(x.x, cur[o].x, a); a = fmaf(x.y, cur[o].y, a);
                a = fmaf(x.z, cur[o].z, a); a = fmaf(x.w, cur[o].w, a);
                v[r * 4 + o] = a;
            }
        }
        if (qn == 0) {                                          // last pass of this column group
            const float d = reduce32(v, lane);
            const int n = g * 4 + (lane & 3);
            Y[(lane >> 2) * N + n] = nn::act_fwd(d + (bias ? bias[n] : 0.f), act);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) cur[o] = nxt[o];
        g = gn; q = qn;
    }
}

// Partial sums of Y[r][n] = sum_k X[r][k] * W[k * N + n] (W is [K, N] row-major): a thread owns 4 consecutive n, the
// k range is split over `parts` thread groups; red[(part * R + r) * N + n] receives the partials, the caller adds them
// in part order.  Returns the number of parts.
__device__ __forceinline__ int stage_nn_partials(const float* __restrict__ X, const float* __restrict__ W, int N, int K,
                                                 float* __restrict__ red) {
    const int N4 = N >> 2, K4 = K >> 2;
    int parts = min(min(TH / N4, RED_FLOATS / (R * N)), K4);
    const int kchunk = ((K4 + parts - 1) / parts) * 4;
    parts = (K + kchunk - 1) / kchunk;
    const int part = threadIdx.x / N4, n = (threadIdx.x % N4) * 4;
    if (part < parts) {
        float acc[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
        const int k0 = part * kchunk, k1 = min(K, k0 + kchunk);
        // the next four weight rows are in flight while the current ones are used
        float4 w0, w1, w2, w3, n0, n1, n2, n3;
        {
            const float* w = W + (size_t)k0 * N + n;
            w0 = __ldg(reinterpret_cast<const float4*>(w));
            w1 = __ldg(reinterpret_cast<const float4*>(w + N));
            w2 = __ldg(reinterpret_cast<const float4*>(w + 2 * N));
            w3 = __ldg(reinterpret_cast<const float4*>(w + 3 * N));
        }
        for (int k = k0; k < k1; k += 4) {
            if (k + 4 < k1) {
                const float* w = W + (size_t)(k + 4) * N + n;
                n0 = __ldg(reinterpret_cast<const float4*>(w));
                n1 = __ldg(reinterpret_cast<const float4*>(w + N));
                n2 = __ldg(reinterpret_cast<const float4*>(w + 2 * N));
                n3 = __ldg(reinterpret_cast<const float4*>(w + 3 * N));
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 x = nn::ld4(X + r * K + k);
                acc[r][0] = fmaf(x.x, w0.x, acc[r][0]); acc[r][1] = fmaf(x.x, w0.y, acc[r][1]);
                acc[r][2] = fmaf(x.x, w0.z, acc[r][2]); acc[r][3] = fmaf(x.x, w0.w, acc[r][3]);
                acc[r][0] = fmaf(x.y, w1.x, acc[r][0]); acc[r][1] = fmaf(x.y, w1.y, acc[r][1]);
                acc[r][2] = fmaf(x.y, w1.z, acc[r][2]); acc[r][3] = fmaf(x.y, w1.w, acc[r][3]);
                acc[r][0] = fmaf(x.z, w2.x, acc[r][0]); acc[r][1] = fmaf(x.z, w2.y, acc[r][1]);
                acc[r][2] = fmaf(x.z, w2.z, acc[r][2]); acc[r][3] = fmaf(x.z, w2.w, acc[r][3]);
                acc[r][0] = fmaf(x.w, w3.x, acc[r][0]); acc[r][1] = fmaf(x.w, w3.y, acc[r][1]);
                acc[r][2] = fmaf(x.w, w3.z, acc[r][2]); acc[r][3] = fmaf(x.w, w3.w, acc[r][3]);
            }
            w0 = n0; w1 = n1; w2 = n2; w3 = n3;
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            nn::st4(red + (size_t)(part * R + r) * N + n, make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
    }
    return parts;
}

// shared memory: P [R][C] | H [R][Cs] | S [R][C]
__global__ void __launch_bounds__(TH)
se_fc_fwd_kernel(const float* __restrict__ p, const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ h1,
                 float* __restrict__ s, int F, int C, int Cs, int act1, int act2) {
    extern __shared__ __align__(16) float sm[];
    float* P = sm;
    float* H = P + R * C;
    float* S = H + R * Cs;
    const int row0 = blockIdx.x * R, nrows = min(R, F - row0);
    const int C4 = C >> 2;
    for (int e = threadIdx.x; e < R * C4; e += TH) {
        const int r = e / C4, c = (e - r * C4) * 4;
        nn::st4(P + r * C + c, r < nrows ? nn::ld4(p + (size_t)(row0 + r) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    __syncthreads();
    stage_nt(P, w1, b1, Cs, C, act1, H);
    __syncthreads();
    for (int e = threadIdx.x; e < nrows * Cs; e += TH) h1[(size_t)row0 * Cs + e] = H[e];
    stage_nt(H, w2, b2, C, Cs, act2, S);
    __syncthreads();
    for (int e = threadIdx.x; e < nrows * C4; e += TH) nn::st4(s + (size_t)row0 * C + e * 4, nn::ld4(S + e * 4));
}

// shared memory: A [R][C] (dz2) | B [R][Cs] (dz1) | red [RED_FLOATS]
__global__ void __launch_bounds__(TH)
se_fc_bwd_kernel(float* __restrict__ ds, const float* __restrict__ s, const float* __restrict__ h1,
                 const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ dz1,
                 float* __restrict__ dp, int F, int C, int Cs, int act1, int act2) {
    extern __shared__ __align__(16) float sm[];
    float* A = sm;
    float* B = A + R * C;
    float* red = B + R * Cs;
    const int row0 = blockIdx.x * R, nrows = min(R, F - row0);
    const int C4 = C >> 2;
    for (int e = threadIdx.x; e < R * C4; e += TH) {
        const int r = e / C4, c = (e - r * C4) * 4;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows) {
            const size_t o = (size_t)(row0 + r) * C + c;
            g = nn::ld4(ds + o);
            const float4 y = nn::ld4(s + o);
            g.x *= nn::act_grad_from_out(y.x, act2); g.y *= nn::act_grad_from_out(y.y, act2);
            g.z *= nn::act_grad_from_out(y.z, act2); g.w *= nn::act_grad_from_out(y.w, act2);
            nn::st4(ds + o, g);
        }
        nn::st4(A + r * C + c, g);
    }
    __syncthreads();
    // dz1 = (dz2 . W2) * act1'(h1): W2 is [C, Cs] = [K][N]
    int parts = stage_nn_partials(A, w2, Cs, C, red);
    __syncthreads();
    for (int e = threadIdx.x; e < R * Cs; e += TH) {
        const int r = e / Cs;
        float v = 0.f;
        for (int q = 0; q < parts; ++q) v += red[(size_t)q * R * Cs + e];
        if (r < nrows) {
            v *= nn::act_grad_from_out(h1[(size_t)row0 * Cs + e], act1);
            dz1[(size_t)row0 * Cs + e] = v;
        } else {
            v = 0.f;
        }
        B[e] = v;
    }
    __syncthreads();
    // dp = dz1 . W1: W1 is [Cs, C] = [K][N]
    parts = stage_nn_partials(B, w1, C, Cs, red);
    __syncthreads();
    for (int e = threadIdx.x; e < nrows * C; e += TH) {
        float v = 0.f;
        for (int q = 0; q < parts; ++q) v += red[(size_t)q * R * C + e];
        dp[(size_t)row0 * C + e] = v;
    }
}

static bool act_ok(int a) { return a == LR_ACT_NONE || a == LR_ACT_RELU || a == LR_ACT_HSIGMOID || a == LR_ACT_RELU6; }

}  // namespace se

extern "C" int lr_se_fc_fwd(const float* p, const float* w1, const float* b1, const float* w2, const float* b2,
                            float* h1, float* s, int F, int C, int Cs, int act1, int act2, lr_stream_t stream) {
    LR_CHECK_ARG(F >= 0 && C > 0 && Cs > 0 && (C & 3) == 0 && (Cs & 3) == 0, "lr_se_fc_fwd: bad shape (C, Cs multiples of 4)");
    LR_CHECK_ARG(se::act_ok(act1) && se::act_ok(act2), "lr_se_fc_fwd: activation not supported");
    if (F == 0) return LR_OK;
    LR_CHECK_ARG(p && w1 && w2 && h1 && s, "lr_se_fc_fwd: null pointer");
    LR_CHECK_ALIGN(p); LR_CHECK_ALIGN(w1); LR_CHECK_ALIGN(w2); LR_CHECK_ALIGN(h1); LR_CHECK_ALIGN(s);
    const size_t smem = (size_t)se::R * (2 * C + Cs) * sizeof(float);
    LR_CHECK_ARG(smem <= 200 * 1024, "lr_se_fc_fwd: C too large for the shared-memory row block");
    cudaError_t e = lr::ensure_max_dynamic_smem(se::se_fc_fwd_kernel, (int)smem);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "lr_se_fc_fwd: %s", cudaGetErrorString(e));
    se::se_fc_fwd_kernel<<<(F + se::R - 1) / se::R, se::TH, smem, stream>>>(p, w1, b1, w2, b2, h1, s, F, C, Cs, act1, act2);
    lr::count_launch();
    LR_CHECK_LAUNCH("se_fc_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_se_fc_bwd(float* ds, const float* s, const float* h1, const float* w1, const float* w2, float* dz1,
                            float* dp, int F, int C, int Cs, int act1, int act2, lr_stream_t stream) {
    LR_CHECK_ARG(F >= 0 && C > 0 && Cs > 0 && (C & 3) == 0 && (Cs & 3) == 0, "lr_se_fc_bwd: bad shape (C, Cs multiples of 4)");
    LR_CHECK_ARG(se::act_ok(act1) && se::act_ok(act2), "lr_se_fc_bwd: activation has no output-form derivative");
    LR_CHECK_ARG(C / 4 <= se::TH && Cs / 4 <= se::TH && se::R * C <= se::RED_FLOATS, "lr_se_fc_bwd: C too large");
    if (F == 0) return LR_OK;
    LR_CHECK_ARG(ds && s && h1 && w1 && w2 && dz1 && dp, "lr_se_fc_bwd: null pointer");
    LR_CHECK_ALIGN(ds); LR_CHECK_ALIGN(s); LR_CHECK_ALIGN(h1); LR_CHECK_ALIGN(w1); LR_CHECK_ALIGN(w2);
    LR_CHECK_ALIGN(dz1); LR_CHECK_ALIGN(dp);
    const size_t smem = ((size_t)se::R * (C + Cs) + se::RED_FLOATS) * sizeof(float);
    cudaError_t e = lr::ensure_max_dynamic_smem(se::se_fc_bwd_kernel, (int)smem);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "lr_se_fc_bwd: %s", cudaGetErrorString(e));
    se::se_fc_bwd_kernel<<<(F + se::R - 1) / se::R, se::TH, smem, stream>>>(ds, s, h1, w1, w2, dz1, dp, F, C, Cs, act1, act2);
    lr::count_launch();
    LR_CHECK_LAUNCH("se_fc_bwd_kernel");
    return LR_OK;
}
