// Attention fusion of the late triple-fusion model (audio_cues_video/models/late_fusion_mobile.py:6-19):
//   stacked = stack(feats, dim=1) [B,S,C]; scores = attn(stacked) [B,S]; w = softmax(scores, dim=1);
//   fused[b,:] = sum_s w[b,s] * stacked[b,s,:]
// (the attn MLP itself is two lr_gemm calls on the [B*S, C] rows).  One CTA per clip; S <= 8.
#include "nn_common.cuh"

namespace fu {

constexpr int TH = 128;
constexpr int SMAX = 32;   // modalities (3) or pooled time steps (10, audio/models/lstm_resnet_attn_model.py:78-84)

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
    LR_CHECK_ARG(B >= 0 && S > 0 && S <= fu::SMAX && C > 0, "lr_attn_fuse_fwd: need 1 <= S <= 32, C > 0");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(stacked && scores && weights && fused, "lr_attn_fuse_fwd: null pointer");
    fu::attn_fuse_fwd_kernel<<<B, fu::TH, 0, stream>>>(stacked, scores, weights, fused, S, C);
    lr::count_launch();
    LR_CHECK_LAUNCH("attn_fuse_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_attn_fuse_bwd(const float* stacked, const float* weights, const float* dfused, float* dstacked,
                                float* dscores, int B, int S, int C, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && S > 0 && S <= fu::SMAX && C > 0, "lr_attn_fuse_bwd: need 1 <= S <= 32, C > 0");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(stacked && weights && dfused && dstacked && dscores, "lr_attn_fuse_bwd: null pointer");
    fu::attn_fuse_bwd_kernel<<<B, fu::TH, 0, stream>>>(stacked, weights, dfused, dstacked, dscores, S, C);
    lr::count_launch();
    LR_CHECK_LAUNCH("attn_fuse_bwd_kernel");
    return LR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Learnable scalar late fusion (audio_video/models/late_fusion.py:82,92, late_fusion_fast.py:34,58):
//   fused = alpha * a + (1 - alpha) * v
// and the channel-major flatten of a channels-last activation (x.view(B, -1) on an NCHW tensor,
// audio_video/models/middle_fusion.py:29): y[b, c*HW + p] = x[b, p, c].
namespace fu {

__global__ void __launch_bounds__(256)
alpha_fuse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ v, const float* __restrict__ alpha,
                      float* __restrict__ out, long long n) {
    const float al = *alpha;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        out[i] = al * a[i] + (1.f - al) * v[i];
}

// one block: da = alpha * dout, dv = (1 - alpha) * dout, dalpha += sum (a - v) * dout
__global__ void __launch_bounds__(256)
alpha_fuse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ v, const float* __restrict__ alpha,
                      const float* __restrict__ dout, float* __restrict__ da, float* __restrict__ dv,
                      float* __restrict__ dalpha, long long n) {
    __shared__ float red[8];
    const float al = *alpha;
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += 256) {
        const float g = dout[i];
        da[i] = al * g;
        dv[i] = (1.f - al) * g;
        s = fmaf(a[i] - v[i], g, s);
    }
    s = lr::warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        *dalpha += t;
    }
}

// to_nchw != 0: y[b*ldy + c*HW + p] = x[(b*HW + p)*C + c];  else the inverse (x[(b*HW + p)*C + c] = y[b*ldy + c*HW + p])
__global__ void __launch_bounds__(256)
flatten_kernel(float* __restrict__ x, float* __restrict__ y, long long ldy, int B, int HW, int C, int to_nchw) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    float* xb = x + (long long)b * HW * C;
    float* yb = y + (long long)b * ldy;
    if (to_nchw) {
        for (int i = ty; i < 32; i += 8)
            if (p0 + i < HW && c0 + tx < C) tile[i][tx] = xb[(long long)(p0 + i) * C + c0 + tx];
        __syncthreads();
        for (int i = ty; i < 32; i += 8)
            if (c0 + i < C && p0 + tx < HW) yb[(long long)(c0 + i) * HW + p0 + tx] = tile[tx][i];
    } else {
        for (int i = ty; i < 32; i += 8)
            if (c0 + i < C && p0 + tx < HW) tile[tx][i] = yb[(long long)(c0 + i) * HW + p0 + tx];
        __syncthreads();
        for (int i = ty; i < 32; i += 8)
            if (p0 + i < HW && c0 + tx < C) xb[(long long)(p0 + i) * C + c0 + tx] = tile[i][tx];
    }
}

}  // namespace fu

extern "C" int lr_alpha_fuse_fwd(const float* a, const float* v, const float* alpha, float* out, long long n,
                                 lr_stream_t stream) {
    LR_CHECK_ARG(n >= 0, "lr_alpha_fuse_fwd: negative size");
    if (n == 0) return LR_OK;
    LR_CHECK_ARG(a && v && alpha && out, "lr_alpha_fuse_fwd: null pointer");
    long long g = (n + 255) / 256;
    if (g > 1024) g = 1024;
    fu::alpha_fuse_fwd_kernel<<<(unsigned)g, 256, 0, stream>>>(a, v, alpha, out, n);
    lr::count_launch();
    LR_CHECK_LAUNCH("alpha_fuse_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_alpha_fuse_bwd(const float* a, const float* v, const float* alpha, const float* dout, float* da,
                                 float* dv, float* dalpha, long long n, lr_stream_t stream) {
    LR_CHECK_ARG(n >= 0, "lr_alpha_fuse_bwd: negative size");
    if (n == 0) return LR_OK;
    LR_CHECK_ARG(a && v && alpha && dout && da && dv && dalpha, "lr_alpha_fuse_bwd: null pointer");
    fu::alpha_fuse_bwd_kernel<<<1, 256, 0, stream>>>(a, v, alpha, dout, da, dv, dalpha, n);
    lr::count_launch();
    LR_CHECK_LAUNCH("alpha_fuse_bwd_kernel");
    return LR_OK;
}

extern "C" int lr_flatten_nchw(float* x_nhwc, float* y_nchw, long long ldy, int B, int HW, int C, int to_nchw,
                               lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && HW > 0 && C > 0 && ldy >= (long long)HW * C, "lr_flatten_nchw: bad shape");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(x_nhwc && y_nchw, "lr_flatten_nchw: null pointer");
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
    fu::flatten_kernel<<<grid, 256, 0, stream>>>(x_nhwc, y_nchw, ldy, B, HW, C, to_nchw);
    lr::count_launch();
    LR_CHECK_LAUNCH("flatten_kernel");
    return LR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// ShuffleNetV2 unit tail (torchvision shufflenetv2.py InvertedResidual.forward): out = channel_shuffle(cat(a, b), 2).
// In channels-last rows that is an interleave: out[r, 2c] = a[r, c], out[r, 2c + 1] = b[r, c]; a and b may be column
// slices of wider matrices (x.chunk(2, dim=1) leaves x1 in place), hence the row strides.
namespace fu {

__global__ void __launch_bounds__(256)
shuffle2_fwd_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b, long long ldb,
                    float* __restrict__ out, long long rows, int Ch) {
    const long long n = rows * Ch;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const long long r = i / Ch;
        const int c = (int)(i - r * Ch);
        reinterpret_cast<float2*>(out)[i] = make_float2(a[r * lda + c], b[r * ldb + c]);
    }
}

__global__ void __launch_bounds__(256)
shuffle2_bwd_kernel(const float* __restrict__ dout, float* __restrict__ da, long long lda, float* __restrict__ db,
                    long long ldb, long long rows, int Ch) {
    const long long n = rows * Ch;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const long long r = i / Ch;
        const int c = (int)(i - r * Ch);
        const float2 g = reinterpret_cast<const float2*>(dout)[i];
        da[r * lda + c] = g.x;
        db[r * ldb + c] = g.y;
    }
}

}  // namespace fu

extern "C" int lr_shuffle2_fwd(const float* a, long long lda, const float* b, long long ldb, float* out, long long rows,
                               int Ch, lr_stream_t stream) {
    LR_CHECK_ARG(rows >= 0 && Ch > 0 && lda >= Ch && ldb >= Ch, "lr_shuffle2_fwd: bad shape");
    if (rows == 0) return LR_OK;
    LR_CHECK_ARG(a && b && out, "lr_shuffle2_fwd: null pointer");
    long long g = (rows * Ch + 255) / 256;
    const long long cap = (long long)lr::sm_count() * 8;
    if (g > cap) g = cap;
    fu::shuffle2_fwd_kernel<<<(unsigned)g, 256, 0, stream>>>(a, lda, b, ldb, out, rows, Ch);
    lr::count_launch();
    LR_CHECK_LAUNCH("shuffle2_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_shuffle2_bwd(const float* dout, float* da, long long lda, float* db, long long ldb, long long rows, int Ch,
                               lr_stream_t stream) {
    LR_CHECK_ARG(rows >= 0 && Ch > 0 && lda >= Ch && ldb >= Ch, "lr_shuffle2_bwd: bad shape");
    if (rows == 0) return LR_OK;
    LR_CHECK_ARG(dout && da && db, "lr_shuffle2_bwd: null pointer");
    long long g = (rows * Ch + 255) / 256;
    const long long cap = (long long)lr::sm_count() * 8;
    if (g > cap) g = cap;
    fu::shuffle2_bwd_kernel<<<(unsigned)g, 256, 0, stream>>>(dout, da, lda, db, ldb, rows, Ch);
    lr::count_launch();
    LR_CHECK_LAUNCH("shuffle2_bwd_kernel");
    return LR_OK;
}
