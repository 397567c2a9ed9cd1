// Direct convolutions of the MobileNetV3 trunk that are NOT GEMM-shaped (channels-last, fp32):
//   stem   : 3x3 stride-2 conv 3 -> 16 reading the lip frames in the layout the caller has
//            (uint8 (B,T,H,W,3) straight from the .npy, or float (B,3,T,H,W) as the reference's
//            forward() receives it) -- the /255, permute(3,0,1,2) and TimeDistributed
//            permute/contiguous/view of video/data_utils/dataset_loader.py:90,96 and
//            audio_video/models/middle_fusion_fast.py:32-33 are folded into the addressing.
// (the depthwise convolutions live in dwconv.cu)
// Every forward kernel also emits the per-channel sum and sum of squares of its raw output
// (double atomics) so that train-mode BatchNorm needs no extra pass over the activation.
#include "nn_common.cuh"

namespace cv {

constexpr int TH = 256;

// ------------------------------------------------------------------------------------------ stem
struct StemIn {
    const void* x;
    int is_u8;                       // 1: uint8 values scaled by `scale`; 0: float
    long long sf, sc, sh, sw;        // element strides of frame / channel / row / column
    float scale;
    int F, H, W, Ho, Wo;
};
__device__ __forceinline__ float stem_load(const StemIn& in, int f, int c, int h, int w) {
    const long long off = (long long)f * in.sf + (long long)c * in.sc + (long long)h * in.sh + (long long)w * in.sw;
    return in.is_u8 ? lr::u8_scaled(static_cast<const unsigned char*>(in.x)[off], in.scale)
                    : static_cast<const float*>(in.x)[off] * in.scale;
}
// Frame index f = b*T + t maps to (b, t) strides through `sf` only when the (b, t) pair is
// addressable with one stride; for the float (B,3,T,H,W) layout that is not the case, so the
// caller passes T and the two strides and we resolve here.
struct StemIn2 {
    StemIn in; int T; long long sb, st;
};
__device__ __forceinline__ long long frame_base(const StemIn2& s, int f) {
    const int b = f / s.T, t = f - b * s.T;
    return (long long)b * s.sb + (long long)t * s.st;
}
__device__ __forceinline__ float stem_ld(const StemIn2& s, long long fb, int c, int h, int w) {
    const long long off = fb + (long long)c * s.in.sc + (long long)h * s.in.sh + (long long)w * s.in.sw;
    return s.in.is_u8 ? lr::u8_scaled(static_cast<const unsigned char*>(s.in.x)[off], s.in.scale)
                      : static_cast<const float*>(s.in.x)[off] * s.in.scale;
}

constexpr int SC = 16;               // stem output channels
constexpr int ST = 27;               // 3 in-channels * 3 * 3 taps

__global__ void __launch_bounds__(TH)
stem_fwd_kernel(const StemIn2 s, const float* __restrict__ w /*[16][3][3][3]*/, float* __restrict__ y,
                double* __restrict__ stats) {
    __shared__ float ws[ST][SC];      // [ci*9 + kh*3 + kw][co]
    __shared__ float ssum[TH / 32][SC], ssq[TH / 32][SC];      // one row per warp, added in warp order (no float atomics)
    for (int i = threadIdx.x; i < ST * SC; i += TH) { const int co = i / ST, tp = i - co * ST; ws[tp][co] = w[i]; }
    __syncthreads();
    const int Ho = s.in.Ho, Wo = s.in.Wo, H = s.in.H, W = s.in.W;
    const long long total = (long long)s.in.F * Ho * Wo;
    float lsum[SC], lsq[SC];
#pragma unroll
    for (int c = 0; c < SC; ++c) { lsum[c] = 0.f; lsq[c] = 0.f; }
    for (long long pix = (long long)blockIdx.x * TH + threadIdx.x; pix < total; pix += (long long)gridDim.x * TH) {
        const int wo = int(pix % Wo);
        const long long r = pix / Wo;
        const int ho = int(r % Ho), f = int(r / Ho);
        const long long fb = frame_base(s, f);
        float acc[SC];
#pragma unroll
        for (int c = 0; c < SC; ++c) acc[c] = 0.f;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hi = ho * 2 - 1 + kh;
                if (hi < 0 || hi >= H) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int wi = wo * 2 - 1 + kw;
                    if (wi < 0 || wi >= W) continue;
                    const float xv = stem_ld(s, fb, ci, hi, wi);
                    const float* wr = ws[ci * 9 + kh * 3 + kw];
#pragma unroll
                    for (int c = 0; c < SC; ++c) acc[c] = fmaf(xv, wr[c], acc[c]);
                }
            }
        float* o = y + pix * SC;
#pragma unroll
        for (int c = 0; c < SC; c += 4) nn::st4(o + c, make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]));
#pragma unroll
        for (int c = 0; c < SC; ++c) { lsum[c] += acc[c]; lsq[c] = fmaf(acc[c], acc[c], lsq[c]); }
    }
#pragma unroll
    for (int c = 0; c < SC; ++c) {
        const float a = lr::warp_sum(lsum[c]), b = lr::warp_sum(lsq[c]);
        if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5][c] = a; ssq[threadIdx.x >> 5][c] = b; }
    }
    __syncthreads();
    if (threadIdx.x < SC) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w_ = 0; w_ < TH / 32; ++w_) { a += ssum[w_][threadIdx.x]; b += ssq[w_][threadIdx.x]; }
        nn::atomic_add_double(stats + threadIdx.x, (double)a);
        nn::atomic_add_double(stats + SC + threadIdx.x, (double)b);
    }
}

// dW[co][tap] = sum_pix dy[pix][co] * x[pix, tap].  Thread (co, tap-slot) owns outputs, the block
// stages PT pixels of dy and of the im2col patch in shared memory; persistent over pixel tiles.
constexpr int PT = 64;
__global__ void __launch_bounds__(TH)
stem_wgrad_kernel(const StemIn2 s, const float* __restrict__ dy, float* __restrict__ dw /*[16][27]*/) {
    __shared__ float xs[PT][ST + 1];
    __shared__ float ds[PT][SC + 1];
    const int Ho = s.in.Ho, Wo = s.in.Wo, H = s.in.H, W = s.in.W;
    const long long total = (long long)s.in.F * Ho * Wo;
    // 432 outputs over 256 threads: thread t owns (co = t & 15, taps tp = t >> 4 and tp + 16 (< 27))
    const int co = threadIdx.x & 15, tp0 = threadIdx.x >> 4, tp1 = tp0 + 16;
    float a0 = 0.f, a1 = 0.f;
    for (long long base = (long long)blockIdx.x * PT; base < total; base += (long long)gridDim.x * PT) {
        const int np = (int)min((long long)PT, total - base);
        for (int i = threadIdx.x; i < PT * ST; i += TH) {
            const int pl = i / ST, tp = i - pl * ST;
            float v = 0.f;
            if (pl < np) {
                const long long pix = base + pl;
                const int wo = int(pix % Wo);
                const long long r = pix / Wo;
                const int ho = int(r % Ho), f = int(r / Ho);
                const int ci = tp / 9, kh = (tp - ci * 9) / 3, kw = tp - ci * 9 - kh * 3;
                const int hi = ho * 2 - 1 + kh, wi = wo * 2 - 1 + kw;
                if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = stem_ld(s, frame_base(s, f), ci, hi, wi);
            }
            xs[pl][tp] = v;
        }
        for (int i = threadIdx.x; i < PT * SC; i += TH) {
            const int pl = i >> 4, c = i & 15;
            ds[pl][c] = pl < np ? dy[(base + pl) * SC + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int pl = 0; pl < PT; ++pl) {
            const float d = ds[pl][co];
            a0 = fmaf(d, xs[pl][tp0], a0);
            if (tp1 < ST) a1 = fmaf(d, xs[pl][tp1], a1);
        }
        __syncthreads();
    }
    atomicAdd(&dw[co * ST + tp0], a0);
    if (tp1 < ST) atomicAdd(&dw[co * ST + tp1], a1);
}

}  // namespace cv

// ------------------------------------------------------------------------------------------ C ABI
static int make_stem(cv::StemIn2& s, const void* x, int is_u8, int B, int T, int H, int W, long long sb,
                     long long st, long long sc, long long sh, long long sw, float scale) {
    s.in.x = x; s.in.is_u8 = is_u8; s.in.sf = 0; s.in.sc = sc; s.in.sh = sh; s.in.sw = sw; s.in.scale = scale;
    s.in.F = B * T; s.in.H = H; s.in.W = W; s.in.Ho = (H + 2 - 3) / 2 + 1; s.in.Wo = (W + 2 - 3) / 2 + 1;
    s.T = T; s.sb = sb; s.st = st;
    return 0;
}

extern "C" int lr_stem_conv_fwd(const void* x, int is_u8, int B, int T, int H, int W, long long sb, long long st,
                                long long sc, long long sh, long long sw, float scale, const float* w, float* y,
                                double* stats, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && T > 0 && H > 0 && W > 0, "lr_stem_conv_fwd: bad shape");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(x && w && y && stats, "lr_stem_conv_fwd: null pointer");
    LR_CHECK_ALIGN(y);
    cv::StemIn2 s; make_stem(s, x, is_u8, B, T, H, W, sb, st, sc, sh, sw, scale);
    const long long total = (long long)s.in.F * s.in.Ho * s.in.Wo;
    const long long want = (total + cv::TH - 1) / cv::TH;
    const int grid = (int)(want < (long long)lr::sm_count() * 8 ? want : (long long)lr::sm_count() * 8);
    cv::stem_fwd_kernel<<<grid, cv::TH, 0, stream>>>(s, w, y, stats);
    lr::count_launch();
    LR_CHECK_LAUNCH("stem_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_stem_conv_wgrad(const void* x, int is_u8, int B, int T, int H, int W, long long sb, long long st,
                                  long long sc, long long sh, long long sw, float scale, const float* dy, float* dw,
                                  lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && T > 0 && H > 0 && W > 0, "lr_stem_conv_wgrad: bad shape");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(x && dy && dw, "lr_stem_conv_wgrad: null pointer");
    cv::StemIn2 s; make_stem(s, x, is_u8, B, T, H, W, sb, st, sc, sh, sw, scale);
    const long long total = (long long)s.in.F * s.in.Ho * s.in.Wo;
    const long long want = (total + cv::PT - 1) / cv::PT;
    const int grid = (int)(want < (long long)lr::sm_count() * 4 ? want : (long long)lr::sm_count() * 4);
    cv::stem_wgrad_kernel<<<grid, cv::TH, 0, stream>>>(s, dy, dw);
    lr::count_launch();
    LR_CHECK_LAUNCH("stem_wgrad_kernel");
    return LR_OK;
}
