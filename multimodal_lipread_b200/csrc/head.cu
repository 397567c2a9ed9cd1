// Small fused kernels at the two ends of the fusion model (fp32):
//   audio branch of MidFusionFast: Conv2d(1,16,3,padding=1) + ReLU + MaxPool2d(2) in one pass, writing the
//     (B, 16*40*58) row that audio_fc consumes (audio_video/models/middle_fusion_fast.py:8-13,28-30),
//     and its weight / bias gradient;
//   CrossEntropyLoss(mean) forward + gradient in one kernel (audio_video/train.py:129,65);
//   Adam step over the flat parameter buffer (audio_video/train.py:130,67; torch.optim.Adam defaults,
//     optional coupled L2 weight decay as audio/train.py:155 and video/train.py:207-211 use).
#include "nn_common.cuh"

namespace hd {

constexpr int TH = 256;
constexpr int AC = 16;      // audio conv channels

// out[b, c, ph, pw] = relu(max_{2x2} (conv(x)[b, c, 2ph+dy, 2pw+dx] + bias[c])); arg[...] = index of the max
__global__ void __launch_bounds__(TH)
audio_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out, long long ldo, unsigned char* __restrict__ arg, int B, int Hh, int Ww,
                      int Ph, int Pw) {
    __shared__ float ws[AC * 9], bs[AC];
    for (int i = threadIdx.x; i < AC * 9; i += TH) ws[i] = w[i];
    if (threadIdx.x < AC) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const long long per_clip = (long long)AC * Ph * Pw;
    const long long total = (long long)B * per_clip;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < total; i += (long long)gridDim.x * TH) {
        const int pw = int(i % Pw);
        long long r = i / Pw;
        const int ph = int(r % Ph); r /= Ph;
        const int c = int(r % AC), b = int(r / AC);
        const float* xb = x + (long long)b * Hh * Ww;
        // 4x4 input patch around the 2x2 window
        float patch[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int h = 2 * ph - 1 + a;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int q = 2 * pw - 1 + d;
                patch[a][d] = (h >= 0 && h < Hh && q >= 0 && q < Ww) ? xb[(long long)h * Ww + q] : 0.f;
            }
        }
        const float* wc = ws + c * 9;
        float best = -INFINITY; int bi = 0;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float v = bs[c];
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) v = fmaf(patch[dy + kh][dx + kw], wc[kh * 3 + kw], v);
                if (v > best) { best = v; bi = dy * 2 + dx; }
            }
        out[(long long)b * ldo + (i - (long long)b * per_clip)] = fmaxf(best, 0.f);
        arg[i] = (unsigned char)(best > 0.f ? bi : 4);       // 4 = gradient blocked by the ReLU
    }
}

// dW[c][kh][kw] += sum dA * x[b, h+kh-1, w+kw-1] at the arg-max position; db[c] += sum dA (where relu passed)
__global__ void __launch_bounds__(TH)
audio_conv_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dA, long long lda,
                      const unsigned char* __restrict__ arg, float* __restrict__ dw, float* __restrict__ db, int B,
                      int Hh, int Ww, int Ph, int Pw) {
    __shared__ float red[TH / 32][10];
    const int c = blockIdx.x % AC, b = blockIdx.x / AC;
    const float* xb = x + (long long)b * Hh * Ww;
    const long long per_clip = (long long)AC * Ph * Pw;
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.f;
    for (int i = threadIdx.x; i < Ph * Pw; i += TH) {
        const int ph = i / Pw, pw = i - ph * Pw;
        const long long e = (long long)c * Ph * Pw + i;
        const int a = arg[(long long)b * per_clip + e];
        if (a == 4) continue;
        const float g = dA[(long long)b * lda + e];
        const int h = 2 * ph + (a >> 1), q = 2 * pw + (a & 1);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int hh = h + kh - 1, qq = q + kw - 1;
                if (hh >= 0 && hh < Hh && qq >= 0 && qq < Ww) acc[kh * 3 + kw] = fmaf(g, xb[(long long)hh * Ww + qq], acc[kh * 3 + kw]);
            }
        acc[9] += g;
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const float v = lr::warp_sum(acc[i]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        float s = 0.f;
        for (int wq = 0; wq < TH / 32; ++wq) s += red[wq][threadIdx.x];
        if (threadIdx.x < 9) atomicAdd(&dw[c * 9 + threadIdx.x], s); else atomicAdd(&db[c], s);
    }
}

// One block, one warp per row at a time: loss += -log_softmax(logits)[label] * inv_n ; dlogits = (softmax - onehot) * inv_n.
// Every warp walks its rows in index order and the per-warp partial losses are added in warp order, so the loss is
// bit-reproducible (a batch is at most a few hundred rows of <= 40 logits: one block is plenty).
constexpr int CE_TH = 1024;
__global__ void __launch_bounds__(CE_TH)
ce_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, float* __restrict__ loss,
          float* __restrict__ dlogits, int* __restrict__ correct, int B, int C, float inv_n) {
    __shared__ float wloss[CE_TH / 32];
    __shared__ int wcorr[CE_TH / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float lsum = 0.f; int csum = 0;
    for (int row = warp; row < B; row += CE_TH / 32) {
        const float* l = logits + (long long)row * C;
        float mx = -INFINITY; int am = 0;
        for (int j = lane; j < C; j += 32) if (l[j] > mx) { mx = l[j]; am = j; }
        // arg-max with the lowest index on ties (torch.max semantics)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, am, o);
            if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
        }
        float se = 0.f;
        for (int j = lane; j < C; j += 32) se += expf(l[j] - mx);
        se = lr::warp_sum(se);
        const float lse = logf(se) + mx;
        const int y = (int)labels[row];
        if (dlogits)
            for (int j = lane; j < C; j += 32) dlogits[(long long)row * C + j] = (expf(l[j] - lse) - (j == y ? 1.f : 0.f)) * inv_n;
        lsum += (lse - l[y]) * inv_n;
        csum += (am == y) ? 1 : 0;
    }
    if (lane == 0) { wloss[warp] = lsum; wcorr[warp] = csum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f; int n = 0;
        for (int w = 0; w < CE_TH / 32; ++w) { t += wloss[w]; n += wcorr[w]; }
        *loss += t;                                  // accumulates onto the caller's (zeroed) scalar, as before
        if (correct) *correct += n;
    }
}

struct AdamState { float step; float bc1; float bc2_sqrt; float lr; };

// advance the step counter and refresh the bias corrections (device-side so that a captured CUDA graph
// replays correctly); lr < 0 keeps the current learning rate
__global__ void adam_tick_kernel(AdamState* st, float beta1, float beta2) {
    const float step = st->step + 1.f;
    st->step = step;
    st->bc1 = 1.f - powf(beta1, step);
    st->bc2_sqrt = sqrtf(1.f - powf(beta2, step));
}

__global__ void __launch_bounds__(TH)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            const AdamState* __restrict__ st, long long n, float beta1, float beta2, float eps, float wd,
            float grad_scale) {
    const float step_size = st->lr / st->bc1, bc2s = st->bc2_sqrt;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < n; i += (long long)gridDim.x * TH) {
        float gi = g[i] * grad_scale;
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = m[i] + (gi - m[i]) * (1.f - beta1);           // torch: exp_avg.lerp_(grad, 1 - beta1)
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2s + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

}  // namespace hd

extern "C" int lr_audio_conv_fwd(const float* x, const float* w, const float* bias, float* out, long long ldo,
                                 unsigned char* arg, int B, int H, int W, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "lr_audio_conv_fwd: bad shape");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(x && w && bias && out && arg, "lr_audio_conv_fwd: null pointer");
    const int Ph = H / 2, Pw = W / 2;
    const long long total = (long long)B * hd::AC * Ph * Pw;
    long long g = (total + hd::TH - 1) / hd::TH;
    const long long cap = (long long)lr::sm_count() * 16;
    if (g > cap) g = cap;
    hd::audio_conv_fwd_kernel<<<(unsigned)g, hd::TH, 0, stream>>>(x, w, bias, out, ldo, arg, B, H, W, Ph, Pw);
    lr::count_launch();
    LR_CHECK_LAUNCH("audio_conv_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_audio_conv_bwd(const float* x, const float* dA, long long lda, const unsigned char* arg, float* dw,
                                 float* db, int B, int H, int W, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "lr_audio_conv_bwd: bad shape");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(x && dA && arg && dw && db, "lr_audio_conv_bwd: null pointer");
    hd::audio_conv_bwd_kernel<<<B * hd::AC, hd::TH, 0, stream>>>(x, dA, lda, arg, dw, db, B, H, W, H / 2, W / 2);
    lr::count_launch();
    LR_CHECK_LAUNCH("audio_conv_bwd_kernel");
    return LR_OK;
}

extern "C" int lr_ce_loss(const float* logits, const long long* labels, float* loss, float* dlogits, int* correct,
                          int B, int C, float inv_n, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && C > 0, "lr_ce_loss: bad shape");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(logits && labels && loss, "lr_ce_loss: null pointer");
    hd::ce_kernel<<<1, hd::CE_TH, 0, stream>>>(logits, labels, loss, dlogits, correct, B, C, inv_n);
    lr::count_launch();
    LR_CHECK_LAUNCH("ce_kernel");
    return LR_OK;
}

extern "C" size_t lr_adam_state_bytes(void) { return sizeof(hd::AdamState); }

extern "C" int lr_adam_step(float* p, const float* g, float* m, float* v, void* state, long long n, float beta1,
                            float beta2, float eps, float weight_decay, float grad_scale, lr_stream_t stream) {
    LR_CHECK_ARG(n >= 0, "lr_adam_step: negative size");
    LR_CHECK_ARG(p && g && m && v && state, "lr_adam_step: null pointer");
    hd::adam_tick_kernel<<<1, 1, 0, stream>>>(static_cast<hd::AdamState*>(state), beta1, beta2);
    lr::count_launch();
    LR_CHECK_LAUNCH("adam_tick_kernel");
    if (n == 0) return LR_OK;
    long long gsz = (n + hd::TH - 1) / hd::TH;
    const long long cap = (long long)lr::sm_count() * 16;
    if (gsz > cap) gsz = cap;
    hd::adam_kernel<<<(unsigned)gsz, hd::TH, 0, stream>>>(p, g, m, v, static_cast<const hd::AdamState*>(state), n, beta1,
                                                         beta2, eps, weight_decay, grad_scale);
    lr::count_launch();
    LR_CHECK_LAUNCH("adam_kernel");
    return LR_OK;
}
