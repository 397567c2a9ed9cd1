// Attention fusion of the late triple-fusion model (audio_cues_video/models/late_fusion_mobile.py:6-19):
//   stacked = stack(feats, dim=1) [B,S,C]; scores = attn(stacked) [B,S]; w = softmax(scores, dim=1);
//   fused[b,:] = sum_s w[b,s] * stacked[b,s,:]
// (the attn MLP itself is two lr_gemm calls on the [B*S, C] rows).  One CTA per clip; S <= 8.
#include "nn_common.cuh"

namespace fu {

constexpr int TH = 128;
constexpr int SMAX = 8;

__global__ void __launch_bounds__(TH)
attn_fuse_fwd_kernel(const float* __restrict__ stacked, const float* __restrict__ scores, float* __restrict__ w,
                     float* __restrict__ fused, int S, int C) {
    const int b = blockIdx.x;
    float p[SMAX];
    float mx = -INFINITY;
    for (int s = 0; s < S; ++s) { p[s] = scores[b * S + s]; mx = fmaxf(mx, p[s]); }
    float den = 0.f;
    for (int s = 0; s < S; ++s) { p[s] = expf(p[s] - mx); den += p[s]; }
    const float inv = 1.f / den;
    for (int s = 0; s < S; ++s) p[s] *= inv;
    if (threadIdx.x < S) w[b * S + threadIdx.x] = p[threadIdx.x];
    for (int c = threadIdx.x; c < C; c += TH) {
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc = fmaf(p[s], stacked[((long long)b * S + s) * C + c], acc);
        fused[(long long)b * C + c] = acc;
    }
}

// dstacked[b,s,:] = w[b,s] * dfused[b,:];  dw[b,s] = <stacked[b,s,:], dfused[b,:]>;
// dscores[b,s] = w[b,s] * (dw[b,s] - sum_t w[b,t] dw[b,t])
__global__ void __launch_bounds__(TH)
attn_fuse_bwd_kernel(const float* __restrict__ stacked, const float* __restrict__ w, const float* __restrict__ dfused,
                     float* __restrict__ dstacked, float* __restrict__ dscores, int S, int C) {
    __shared__ float red[SMAX][TH / 32];
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float dot[SMAX];
    for (int s = 0; s < S; ++s) dot[s] = 0.f;
    for (int c = threadIdx.x; c < C; c += TH) {
        const float g = dfused[(long long)b * C + c];
        for (int s = 0; s < S; ++s) {
            const long long o = ((long long)b * S + s) * C + c;
            dot[s] = fmaf(stacked[o], g, dot[s]);
            dstacked[o] = w[b * S + s] * g;
        }
    }
    for (int s = 0; s < S; ++s) {
        const float v = lr::warp_sum(dot[s]);
        if (lane == 0) red[s][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float dw[SMAX], mean = 0.f;
        for (int s = 0; s < S; ++s) {
            float v = 0.f;
            for (int q = 0; q < TH / 32; ++q) v += red[s][q];
            dw[s] = v;
            mean = fmaf(w[b * S + s], v, mean);
        }
        for (int s = 0; s < S; ++s) dscores[b * S + s] = w[b * S + s] * (dw[s] - mean);
    }
}

}  // namespace fu

extern "C" int lr_attn_fuse_fwd(const float* stacked, const float* scores, float* weights, float* fused, int B, int S,
                                int C, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && S > 0 && S <= fu::SMAX && C > 0, "lr_attn_fuse_fwd: need 1 <= S <= 8, C > 0");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(stacked && scores && weights && fused, "lr_attn_fuse_fwd: null pointer");
    fu::attn_fuse_fwd_kernel<<<B, fu::TH, 0, stream>>>(stacked, scores, weights, fused, S, C);
    lr::count_launch();
    LR_CHECK_LAUNCH("attn_fuse_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_attn_fuse_bwd(const float* stacked, const float* weights, const float* dfused, float* dstacked,
                                float* dscores, int B, int S, int C, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && S > 0 && S <= fu::SMAX && C > 0, "lr_attn_fuse_bwd: need 1 <= S <= 8, C > 0");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(stacked && weights && dfused && dstacked && dscores, "lr_attn_fuse_bwd: null pointer");
    fu::attn_fuse_bwd_kernel<<<B, fu::TH, 0, stream>>>(stacked, weights, dfused, dstacked, dscores, S, C);
    lr::count_launch();
    LR_CHECK_LAUNCH("attn_fuse_bwd_kernel");
    return LR_OK;
}
