// nn.LayerNorm over the last dimension with the residual add of a post-norm TransformerEncoderLayer fused in
// (x = norm(x + sublayer(x)), video/models/resnet_trans.py:96-103, audio/models/lstm_resnet_trans_model.py:52-58),
// and the broadcast add that builds the audio model's input sequence (repeat + positional encoding, :91-94).
// Rows are few (B*T <= a few thousand) and short (D <= 1024): one warp per row, HBM-bound and launch-bound.
#include "common.cuh"

namespace ln {

constexpr int TH = 256;
constexpr int WARPS = TH / 32;
constexpr int MAXD = 1024;
constexpr int PER = MAXD / 32;

// y = (s - mean) * rstd * gamma + beta with s = a + b (b may be NULL); s and (mean, rstd) are saved for the backward.
__global__ void __launch_bounds__(TH) layernorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float eps, float* __restrict__ y, float* __restrict__ s_out,
                                                           float* __restrict__ stats, int rows, int D) {
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += gridDim.x * WARPS) {
        const long long o = (long long)r * D;
        float v[PER];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = lane + 32 * i;
            v[i] = 0.f;
            if (c < D) { v[i] = a[o + c] + (b ? b[o + c] : 0.f); sum += v[i]; }
        }
        const float mean = lr::warp_sum(sum) / (float)D;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = lane + 32 * i;
            if (c < D) { const float d = v[i] - mean; var = fmaf(d, d, var); }
        }
        const float rstd = rsqrtf(lr::warp_sum(var) / (float)D + eps);
        if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = lane + 32 * i;
            if (c < D) {
                s_out[o + c] = v[i];
                y[o + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
            }
        }
    }
}

// ds = rstd * (g - mean(g) - xhat * mean(g * xhat)) with g = dy * gamma;  dgamma += sum_r dy * xhat;  dbeta += sum_r dy
__global__ void __launch_bounds__(TH) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ s,
                                                           const float* __restrict__ stats, const float* __restrict__ gamma,
                                                           float* __restrict__ ds, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int rows, int D) {
    __shared__ float red[2][WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float pg[PER], pb[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) pg[i] = pb[i] = 0.f;
    for (int r = blockIdx.x * WARPS + warp; r < rows; r += gridDim.x * WARPS) {
        const long long o = (long long)r * D;
        const float mean = stats[2 * r], rstd = stats[2 * r + 1];
        float xh[PER], g[PER];
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = lane + 32 * i;
            xh[i] = g[i] = 0.f;
            if (c < D) {
                const float d = dy[o + c];
                xh[i] = (s[o + c] - mean) * rstd;
                g[i] = d * gamma[c];
                m1 += g[i];
                m2 = fmaf(g[i], xh[i], m2);
                pg[i] = fmaf(d, xh[i], pg[i]);
                pb[i] += d;
            }
        }
        m1 = lr::warp_sum(m1) / (float)D;
        m2 = lr::warp_sum(m2) / (float)D;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = lane + 32 * i;
            if (c < D) ds[o + c] = rstd * (g[i] - m1 - xh[i] * m2);
        }
    }
    // column partials: across the warps of the block through shared memory, across blocks with one atomic per column
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        if (32 * i >= D) break;
        red[0][warp][lane] = pg[i];
        red[1][warp][lane] = pb[i];
        __syncthreads();
        if (warp == 0) {
            float tg = 0.f, tb = 0.f;
            for (int w = 0; w < WARPS; ++w) { tg += red[0][w][lane]; tb += red[1][w][lane]; }
            const int c = lane + 32 * i;
            if (c < D) { atomicAdd(dgamma + c, tg); atomicAdd(dbeta + c, tb); }
        }
        __syncthreads();
    }
}

// out[f, t, c] = x[f, c] + r[t, c]   (r may be NULL): x.unsqueeze(1).repeat(1, T, 1) + pe[:, :T]
__global__ void __launch_bounds__(TH) add_bcast_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                       float* __restrict__ out, int F, int T, int C) {
    const long long n = (long long)F * T * C;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < n; i += (long long)gridDim.x * TH) {
        const int c = (int)(i % C);
        const long long ft = i / C;
        const int t = (int)(ft % T);
        const long long f = ft / T;
        out[i] = x[f * C + c] + (r ? r[(long long)t * C + c] : 0.f);
    }
}

}  // namespace ln

extern "C" int lr_layernorm_fwd(const float* a, const float* b, const float* gamma, const float* beta, float eps, float* y,
                                float* s, float* stats, int rows, int D, lr_stream_t stream) {
    LR_CHECK_ARG(rows >= 0 && D >= 1 && D <= ln::MAXD, "lr_layernorm_fwd: need 1 <= D <= %d (D %d)", ln::MAXD, D);
    if (rows == 0) return LR_OK;
    LR_CHECK_ARG(a && gamma && beta && y && s && stats, "lr_layernorm_fwd: null pointer");
    int grid = (rows + ln::WARPS - 1) / ln::WARPS;
    const int cap = lr::sm_count() * 8;
    if (grid > cap) grid = cap;
    ln::layernorm_fwd_kernel<<<grid, ln::TH, 0, stream>>>(a, b, gamma, beta, eps, y, s, stats, rows, D);
    lr::count_launch();
    LR_CHECK_LAUNCH("layernorm_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_layernorm_bwd(const float* dy, const float* s, const float* stats, const float* gamma, float* ds,
                                float* dgamma, float* dbeta, int rows, int D, lr_stream_t stream) {
    LR_CHECK_ARG(rows >= 0 && D >= 1 && D <= ln::MAXD, "lr_layernorm_bwd: need 1 <= D <= %d (D %d)", ln::MAXD, D);
    if (rows == 0) return LR_OK;
    LR_CHECK_ARG(dy && s && stats && gamma && ds && dgamma && dbeta, "lr_layernorm_bwd: null pointer");
    int grid = (rows + ln::WARPS - 1) / ln::WARPS;
    const int cap = lr::sm_count();
    if (grid > cap) grid = cap;
    ln::layernorm_bwd_kernel<<<grid, ln::TH, 0, stream>>>(dy, s, stats, gamma, ds, dgamma, dbeta, rows, D);
    lr::count_launch();
    LR_CHECK_LAUNCH("layernorm_bwd_kernel");
    return LR_OK;
}

extern "C" int lr_add_bcast(const float* x, const float* r, float* out, int F, int T, int C, lr_stream_t stream) {
    LR_CHECK_ARG(F >= 0 && T >= 1 && C >= 1, "lr_add_bcast: bad shape");
    if (F == 0) return LR_OK;
    LR_CHECK_ARG(x && out, "lr_add_bcast: null pointer");
    long long g = ((long long)F * T * C + ln::TH - 1) / ln::TH;
    const long long cap = (long long)lr::sm_count() * 8;
    if (g > cap) g = cap;
    ln::add_bcast_kernel<<<(unsigned)g, ln::TH, 0, stream>>>(x, r, out, F, T, C);
    lr::count_launch();
    LR_CHECK_LAUNCH("add_bcast_kernel");
    return LR_OK;
}
