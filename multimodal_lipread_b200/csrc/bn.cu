// Channels-last [rows, C] kernels around the convolutions (fp32):
//   train-mode BatchNorm (+activation, +residual) forward / backward with the batch statistics taken from
//   the double sum / sum-of-squares the producing conv emitted (torch.nn.BatchNorm2d in training mode:
//   biased variance for normalisation, unbiased for running_var, momentum update, num_batches_tracked),
//   eval-mode BatchNorm (running statistics), per-frame average pooling, squeeze-excitation scale,
//   element-wise activation backward and bias-gradient column sums.
// torchvision MobileNetV3 call sites: Conv2dNormActivation, SqueezeExcitation, avgpool
// (audio_video/models/middle_fusion_fast.py:15-17,34).
#include "nn_common.cuh"

namespace bn {

constexpr int TH = 256;

struct Bn {
    long long rows; int C;
    const double* stats;            // [2C] sum, sumsq of x over rows (train) -- null in eval mode
    const float* gamma; const float* beta;
    float* running_mean; float* running_var; long long* nbt;
    float eps, momentum;
    int act, training;
};

// per-thread BN coefficients of the thread's 4 channels
struct Coef { float4 scale, shift, mean, invstd; };

// Mean and variance come from the double sums (the subtraction E[x^2] - E[x]^2 needs the precision); everything after
// that is float like the tensor itself: the per-thread prologue must stay cheap next to a main loop of a few rows
// (a double division or square root is a ~100-instruction sequence).
__device__ __forceinline__ void coef1(const Bn& b, int c, double inv_n, float& scale, float& shift, float& mean, float& invstd) {
    if (b.training) {
        const double m = b.stats[c] * inv_n;
        double var = b.stats[b.C + c] * inv_n - m * m;
        if (var < 0.0) var = 0.0;
        mean = (float)m;
        invstd = 1.0f / sqrtf((float)var + b.eps);
    } else {
        mean = b.running_mean[c];
        invstd = 1.0f / sqrtf(b.running_var[c] + b.eps);
    }
    scale = b.gamma[c] * invstd;
    shift = b.beta[c] - mean * scale;
}
__device__ __forceinline__ Coef coef4(const Bn& b, int c) {
    Coef k;
    const double inv_n = 1.0 / (double)b.rows;
    coef1(b, c, inv_n, k.scale.x, k.shift.x, k.mean.x, k.invstd.x);
    coef1(b, c + 1, inv_n, k.scale.y, k.shift.y, k.mean.y, k.invstd.y);
    coef1(b, c + 2, inv_n, k.scale.z, k.shift.z, k.mean.z, k.invstd.z);
    coef1(b, c + 3, inv_n, k.scale.w, k.shift.w, k.mean.w, k.invstd.w);
    return k;
}

// z = act(x * scale + shift) (+ residual).  Block (0,0) also performs the running-statistics update.
// ACT is a template parameter (round 2): with a run-time activation every element paid an indirect branch and the
// divergence bookkeeping around it (ncu: BRA / BRX / BSSY / BSYNC / FSETP / FSEL ~25 % of the executed instructions),
// every access a 64-bit multiply (IMAD 20 %), and a `break` inside the unrolled row loop pushed the in-flight vectors to
// local memory.  The rows of a thread are now walked with running pointers and a per-pass valid count.
template <typename T, bool RES, int ACT>
__global__ void __launch_bounds__(TH, 4)
bn_act_fwd_kernel(const Bn b, const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ z,
                  int rows_per_block, int res_pre, int gw) {
    const nn::CgMap map(b.C, blockIdx.y * gw, gw);
    if (b.training && b.running_mean && blockIdx.x == 0 && blockIdx.y == 0) {
        for (int c = threadIdx.x; c < b.C; c += blockDim.x) {
            const double n = (double)b.rows;
            const double m = b.stats[c] / n;
            double var = b.stats[b.C + c] / n - m * m;
            if (var < 0.0) var = 0.0;
            const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
            b.running_mean[c] = (1.f - b.momentum) * b.running_mean[c] + b.momentum * (float)m;
            b.running_var[c] = (1.f - b.momentum) * b.running_var[c] + b.momentum * (float)unb;
        }
        if (threadIdx.x == 0 && b.nbt) *b.nbt += 1;
    }
    if (!map.active) return;
    const int c = map.cg * 4;
    const Coef k = coef4(b, c);
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(b.rows, r0 + rows_per_block);
    // independent rows in flight per thread (memory-level parallelism): the same BYTES in flight in both storage
    // modes -- the raw vectors (float4 / 2 x bf16x2) stay in registers until used, so bf16 affords twice the rows
    // (with a residual stream the rows in flight stay at 4: RES is a template flag so that its array vanishes otherwise)
    constexpr int U = (sizeof(T) == 2 && !RES) ? 8 : 4;
    const long long step = map.rpp;
    const long long rs = step * b.C;                                  // elements between two rows of this thread
    long long r = r0 + map.rlane;
    const T* px = x + r * b.C + c;
    const T* pq = RES ? res + r * b.C + c : nullptr;
    T* pz = z + r * b.C + c;
    const bool pre = RES && res_pre;
    for (; r < r1; r += U * step, px += U * rs, pz += U * rs) {
        const int nv = (int)min((long long)U, (r1 - r + step - 1) / step);     // valid rows of this pass
        typename nn::Raw4<T>::type vr[U], qr[RES ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) if (u < nv) vr[u] = nn::ldraw(px + u * rs);
        if (RES) {
#pragma unroll
            for (int u = 0; u < U; ++u) if (u < nv) qr[u] = nn::ldraw(pq + u * rs);
            pq += U * rs;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u < nv) {
                const float4 v = nn::cvt4(vr[u]);
                float4 o = make_float4(fmaf(v.x, k.scale.x, k.shift.x), fmaf(v.y, k.scale.y, k.shift.y),
                                       fmaf(v.z, k.scale.z, k.shift.z), fmaf(v.w, k.scale.w, k.shift.w));
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (RES) q = nn::cvt4(qr[u]);
                if (pre) { o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }   // ResNet BasicBlock: act(bn(x) + identity)
                o.x = nn::act_fwd(o.x, ACT); o.y = nn::act_fwd(o.y, ACT); o.z = nn::act_fwd(o.z, ACT); o.w = nn::act_fwd(o.w, ACT);
                if (RES && !pre) { o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
                nn::st4(pz + u * rs, o);
            }
        }
    }
}

// dz masked by the activation derivative expressed through the forward OUTPUT z (pre-activation residual blocks,
// where u + residual is not recomputable from x alone)
__device__ __forceinline__ float4 mask_by_out(float4 g, const float4 z, int act) {
    g.x *= nn::act_grad_from_out(z.x, act); g.y *= nn::act_grad_from_out(z.y, act);
    g.z *= nn::act_grad_from_out(z.z, act); g.w *= nn::act_grad_from_out(z.w, act);
    return g;
}

// backward pass 1: dy = dz * act'(u), u = x*scale + shift;  sums[c] += dy, sums[C+c] += dy * xhat
// ZO: the derivative goes through the saved OUTPUT zo (ResNet pre-activation residual blocks): a third stream, so the
// rows in flight stay at 4 there; otherwise bf16 keeps 8 rows (the same bytes as fp32's 4) in flight.
template <typename T, bool ZO, int ACT>
__global__ void __launch_bounds__(TH, 3)
bn_act_bwd_reduce_kernel(const Bn b, const T* __restrict__ x, const T* __restrict__ dz,
                         const T* __restrict__ zo, double* __restrict__ sums, int rows_per_block, int gw) {
    extern __shared__ float sh[];                       // [2C] (only this block column's channels are touched)
    const int c_lo = blockIdx.y * gw * 4, c_hi = min(b.C, c_lo + gw * 4);
    for (int i = c_lo + threadIdx.x; i < c_hi; i += blockDim.x) { sh[i] = 0.f; sh[b.C + i] = 0.f; }
    __syncthreads();
    const nn::CgMap map(b.C, blockIdx.y * gw, gw);
    if (map.active) {
        const int c = map.cg * 4;
        const Coef k = coef4(b, c);
        const long long r0 = (long long)blockIdx.x * rows_per_block;
        const long long r1 = min(b.rows, r0 + rows_per_block);
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
        const unsigned long long keep = nn::l2_policy_evict_last();     // the apply pass re-reads x and dz right away
        constexpr int U = (sizeof(T) == 2 && !ZO) ? 8 : 4;
        constexpr int GA = ZO ? LR_ACT_NONE : ACT;                      // ZO: the mask comes from the forward output
        const long long step = map.rpp;
        const long long rs = step * b.C;
        long long r = r0 + map.rlane;
        const T* px = x + r * b.C + c;
        const T* pg = dz + r * b.C + c;
        const T* pzo = ZO ? zo + r * b.C + c : nullptr;
        for (; r < r1; r += U * step, px += U * rs, pg += U * rs) {
            const int nv = (int)min((long long)U, (r1 - r + step - 1) / step);
            typename nn::Raw4<T>::type vv[U], gg[U], zz[ZO ? U : 1];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (u < nv) {
                    vv[u] = nn::ldraw_hint(px + u * rs, keep); gg[u] = nn::ldraw_hint(pg + u * rs, keep);
                    if (ZO) zz[u] = nn::ldraw_hint(pzo + u * rs, keep);
                }
            if (ZO) pzo += U * rs;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (u < nv) {
                    const float4 v = nn::cvt4(vv[u]);
                    float4 g = nn::cvt4(gg[u]);
                    if (ZO) g = mask_by_out(g, nn::cvt4(zz[u]), ACT);
                    const float d0 = g.x * nn::act_grad(fmaf(v.x, k.scale.x, k.shift.x), GA);
                    const float d1 = g.y * nn::act_grad(fmaf(v.y, k.scale.y, k.shift.y), GA);
                    const float d2 = g.z * nn::act_grad(fmaf(v.z, k.scale.z, k.shift.z), GA);
                    const float d3 = g.w * nn::act_grad(fmaf(v.w, k.scale.w, k.shift.w), GA);
                    s1.x += d0; s1.y += d1; s1.z += d2; s1.w += d3;
                    s2.x = fmaf(d0, (v.x - k.mean.x) * k.invstd.x, s2.x);
                    s2.y = fmaf(d1, (v.y - k.mean.y) * k.invstd.y, s2.y);
                    s2.z = fmaf(d2, (v.z - k.mean.z) * k.invstd.z, s2.z);
                    s2.w = fmaf(d3, (v.w - k.mean.w) * k.invstd.w, s2.w);
                }
            }
        }
        atomicAdd(&sh[c], s1.x); atomicAdd(&sh[c + 1], s1.y); atomicAdd(&sh[c + 2], s1.z); atomicAdd(&sh[c + 3], s1.w);
        atomicAdd(&sh[b.C + c], s2.x); atomicAdd(&sh[b.C + c + 1], s2.y);
        atomicAdd(&sh[b.C + c + 2], s2.z); atomicAdd(&sh[b.C + c + 3], s2.w);
    }
    __syncthreads();
    for (int i = c_lo + threadIdx.x; i < c_hi; i += blockDim.x) {
        nn::atomic_add_double(sums + i, (double)sh[i]);
        nn::atomic_add_double(sums + b.C + i, (double)sh[b.C + i]);
    }
}

// backward pass 2: dx = gamma*invstd * (dy - mean(dy) - xhat * mean(dy*xhat))   (training)
//                  dx = gamma*invstd * dy                                        (eval)
// block (0,0) writes dgamma += sum(dy*xhat), dbeta += sum(dy).
template <typename T, bool ZO, int ACT>
__global__ void __launch_bounds__(TH, 3)
bn_act_bwd_apply_kernel(const Bn b, const T* __restrict__ x, const T* __restrict__ dz,
                        const T* __restrict__ zo, T* __restrict__ dres,
                        const double* __restrict__ sums, T* __restrict__ dx, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, int rows_per_block, int gw) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && dgamma) {
        for (int c = threadIdx.x; c < b.C; c += blockDim.x) {
            dbeta[c] += (float)sums[c];
            dgamma[c] += (float)sums[b.C + c];
        }
    }
    const nn::CgMap map(b.C, blockIdx.y * gw, gw);
    if (!map.active) return;
    const int c = map.cg * 4;
    // dx = scale * (dy - mean(dy) - xhat * mean(dy * xhat)) with xhat = (x - mean) * invstd, folded into per-channel
    // constants so that a thread keeps 20 of them instead of 24:  dx = scale * dy - ((x - mean) * ka + kd)
    float sc[4], sh[4], mu[4], ka[4], kd[4];
    {
        const Coef k = coef4(b, c);
        const double inv_n = b.training ? 1.0 / (double)b.rows : 0.0;
        const float scale[4] = {k.scale.x, k.scale.y, k.scale.z, k.scale.w}, shift[4] = {k.shift.x, k.shift.y, k.shift.z, k.shift.w};
        const float mean[4] = {k.mean.x, k.mean.y, k.mean.z, k.mean.w}, invstd[4] = {k.invstd.x, k.invstd.y, k.invstd.z, k.invstd.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float m1 = (float)(sums[c + j] * inv_n), m2 = (float)(sums[b.C + c + j] * inv_n);
            sc[j] = scale[j]; sh[j] = shift[j]; mu[j] = mean[j];
            ka[j] = scale[j] * invstd[j] * m2;                  // scale already holds gamma * invstd
            kd[j] = scale[j] * m1;
        }
    }
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(b.rows, r0 + rows_per_block);
    const unsigned long long drop = nn::l2_policy_evict_first();       // last use of x and dz
    constexpr int U = (sizeof(T) == 2 && !ZO) ? 8 : 4;
    constexpr int GA = ZO ? LR_ACT_NONE : ACT;
    const long long step = map.rpp;
    const long long rs = step * b.C;
    long long r = r0 + map.rlane;
    const T* px = x + r * b.C + c;
    const T* pg = dz + r * b.C + c;
    const T* pzo = ZO ? zo + r * b.C + c : nullptr;
    T* pdr = (ZO && dres) ? dres + r * b.C + c : nullptr;
    T* po = dx + r * b.C + c;
    for (; r < r1; r += U * step, px += U * rs, pg += U * rs, po += U * rs) {
        const int nv = (int)min((long long)U, (r1 - r + step - 1) / step);
        typename nn::Raw4<T>::type vv[U], gg[U], zz[ZO ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (u < nv) {
                vv[u] = nn::ldraw_hint(px + u * rs, drop); gg[u] = nn::ldraw_hint(pg + u * rs, drop);
                if (ZO) zz[u] = nn::ldraw_hint(pzo + u * rs, drop);
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u < nv) {
                const float4 v = nn::cvt4(vv[u]);
                float4 g = nn::cvt4(gg[u]);
                if (ZO) {
                    g = mask_by_out(g, nn::cvt4(zz[u]), ACT);
                    if (pdr) nn::st4(pdr + u * rs, g);                             // gradient of the identity branch
                }
                float4 o;
                o.x = sc[0] * (g.x * nn::act_grad(fmaf(v.x, sc[0], sh[0]), GA)) - fmaf(v.x - mu[0], ka[0], kd[0]);
                o.y = sc[1] * (g.y * nn::act_grad(fmaf(v.y, sc[1], sh[1]), GA)) - fmaf(v.y - mu[1], ka[1], kd[1]);
                o.z = sc[2] * (g.z * nn::act_grad(fmaf(v.z, sc[2], sh[2]), GA)) - fmaf(v.z - mu[2], ka[2], kd[2]);
                o.w = sc[3] * (g.w * nn::act_grad(fmaf(v.w, sc[3], sh[3]), GA)) - fmaf(v.w - mu[3], ka[3], kd[3]);
                nn::st4(po + u * rs, o);
            }
        }
        if (ZO) { pzo += U * rs; if (pdr) pdr += U * rs; }
    }
}

// ------------------------------------------------------------------------- per-frame pooling / SE
// mode 0: p[f,c]  = mean_hw a[f,hw,c]
// mode 1: p[f,c]  = sum_hw a[f,hw,c] * g[f,hw,c]        (ds of the SE scale)
// Deterministic: every thread sums its row lane in index order, the row lanes of a channel are then added in lane
// order from a shared-memory slab -- no atomics, so two runs on the same input are bit-identical (eval forwards,
// validate() and model selection depend on that).
template <typename T>
__global__ void __launch_bounds__(TH)
frame_reduce_kernel(const T* __restrict__ a, const T* __restrict__ g, float* __restrict__ p, int HW, int C,
                    int mode) {
    __shared__ __align__(16) float sh[TH * 4];          // [row lane][4 * groups of this pass]
    const long long f = blockIdx.x;
    const float sc = mode == 0 ? 1.f / (float)HW : 1.f;
    for (int cg0 = 0; cg0 < (C >> 2); cg0 += blockDim.x) {
        const nn::CgMap map(C, cg0);
        const int w = min(map.ncg - cg0, (int)blockDim.x);               // channel groups of this pass
        if (map.active) {
            const int c = map.cg * 4;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            // four rows in flight per thread (the loads go out together; the additions keep their index order), the
            // rows of a thread walked with running pointers
            constexpr int U = 4;
            const long long rs = (long long)map.rpp * C;
            const T* pa = a + (f * HW + map.rlane) * C + c;
            const T* pg = mode == 1 ? g + (f * HW + map.rlane) * C + c : nullptr;
            for (int r = map.rlane; r < HW; r += U * map.rpp, pa += U * rs) {
                const int nv = min(U, (HW - r + map.rpp - 1) / map.rpp);
                typename nn::Raw4<T>::type va[U], vg[U];
#pragma unroll
                for (int u = 0; u < U; ++u) if (u < nv) va[u] = nn::ldraw(pa + u * rs);
                if (mode == 1) {
#pragma unroll
                    for (int u = 0; u < U; ++u) if (u < nv) vg[u] = nn::ldraw(pg + u * rs);
                    pg += U * rs;
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (u < nv) {
                        float4 v = nn::cvt4(va[u]);
                        if (mode == 1) { const float4 q = nn::cvt4(vg[u]); v.x *= q.x; v.y *= q.y; v.z *= q.z; v.w *= q.w; }
                        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                    }
            }
            *reinterpret_cast<float4*>(&sh[(map.rlane * w + (map.cg - cg0)) * 4]) = s;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 4 * w; i += blockDim.x) {
            float t = 0.f;
            for (int rl = 0; rl < map.rpp; ++rl) t += sh[rl * 4 * w + i];
            p[f * C + cg0 * 4 + i] = t * sc;
        }
        __syncthreads();
    }
}

// out[f,hw,c] = (a ? a[f,hw,c] * s[f,c] : 0) + (dp ? dp[f,c] * inv_hw : 0)
//   SE forward            : a = activation, s = gate, dp = null
//   SE backward           : a = db, s = gate, dp = gradient of the pooled value
//   avg-pool backward     : a = null, dp = gradient of the pooled value
// (round 2: a block walks whole frames -- the first version divided a 64-bit row index by HW for every 4-channel vector
// and re-read the gate for every row; now the gate / pooled gradient of a frame sit in registers, four rows are in flight
// per thread, and the rows are walked with running pointers)
template <typename T, bool HAS_A, bool HAS_DP>
__global__ void __launch_bounds__(TH)
frame_scale_kernel(const T* __restrict__ a, const float* __restrict__ s, const float* __restrict__ dp,
                   T* __restrict__ out, int F, int HW, int C, float inv_hw, int gw) {
    const nn::CgMap map(C, blockIdx.y * gw, gw);
    if (!map.active) return;
    const int c = map.cg * 4;
    constexpr int U = 4;
    const long long rs = (long long)map.rpp * C;
    for (int f = blockIdx.x; f < F; f += gridDim.x) {
        float4 gate = make_float4(0.f, 0.f, 0.f, 0.f), add = gate;
        if (HAS_A) gate = nn::ld4(s + (long long)f * C + c);
        if (HAS_DP) add = nn::ld4(dp + (long long)f * C + c);
        const long long base = ((long long)f * HW + map.rlane) * C + c;
        const T* pa = HAS_A ? a + base : nullptr;
        T* po = out + base;
        for (int r = map.rlane; r < HW; r += U * map.rpp, po += U * rs) {
            const int nv = min(U, (HW - r + map.rpp - 1) / map.rpp);
            typename nn::Raw4<T>::type va[U];
            if (HAS_A) {
#pragma unroll
                for (int u = 0; u < U; ++u) if (u < nv) va[u] = nn::ldraw(pa + u * rs);
                pa += U * rs;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (u < nv) {
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (HAS_A) {
                        const float4 v = nn::cvt4(va[u]);
                        o.x = v.x * gate.x; o.y = v.y * gate.y; o.z = v.z * gate.z; o.w = v.w * gate.w;
                    }
                    if (HAS_DP) {                                  // the same fused operation as the first version
                        o.x = fmaf(add.x, inv_hw, o.x); o.y = fmaf(add.y, inv_hw, o.y);
                        o.z = fmaf(add.z, inv_hw, o.z); o.w = fmaf(add.w, inv_hw, o.w);
                    }
                    nn::st4(po + u * rs, o);
                }
        }
    }
}

// dy *= act'(y) with the derivative expressed through the OUTPUT (ReLU, hard-sigmoid), in place
template <typename T>
__global__ void __launch_bounds__(TH)
act_bwd_kernel(T* __restrict__ dy, const T* __restrict__ y, long long n, int act) {
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < n; i += (long long)gridDim.x * TH)
        nn::st1(dy + i, nn::ld1(dy + i) * nn::act_grad_from_out(nn::ld1(y + i), act));
}

// y = act(x) (stand-alone activation, e.g. the ReLU between nn.LSTM and the classifier in video/models/resnet_lstm.py:152)
__global__ void __launch_bounds__(TH)
act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int act) {
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < n; i += (long long)gridDim.x * TH)
        y[i] = nn::act_fwd(x[i], act);
}

// db[n] += sum_m dY[m, n]   (row stride ld)
template <typename T>
__global__ void __launch_bounds__(TH)
colsum_kernel(const T* __restrict__ dY, long long ld, long long M, int N, float* __restrict__ db,
              int rows_per_block) {
    // thread -> column (coalesced), y-lanes over rows
    const int cols = min(N, TH);
    const int rpp = TH / cols;
    const int rl = threadIdx.x / cols;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(M, r0 + rows_per_block);
    if (rl >= rpp) return;
    for (int n = blockIdx.y * cols + threadIdx.x % cols; n < N; n += gridDim.y * cols) {
        float s = 0.f;
        for (long long r = r0 + rl; r < r1; r += rpp) s += nn::ld1(dY + r * ld + n);
        atomicAdd(&db[n], s);
    }
}

__global__ void __launch_bounds__(TH)
copy2d_kernel(float* __restrict__ dst, long long ldd, const float* __restrict__ src, long long lds, int rows, int cols) {
    const long long n = (long long)rows * cols;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < n; i += (long long)gridDim.x * TH) {
        const long long r = i / cols; const int c = int(i - r * cols);
        dst[r * ldd + c] = src[r * lds + c];
    }
}

}  // namespace bn

// ------------------------------------------------------------------------------------------ C ABI
// grid of the streaming [rows, C] kernels: block columns of gw channel groups (grid.y) x row ranges (grid.x), about
// 3 blocks per SM in total (= what fits at once: one wave, each thread streams many rows so that the coefficient
// prologue and the statistic atomics are amortised), every block at least 4 passes of its row lanes deep
static int rows_per_block_for(long long rows, int C, dim3* grid, int* gw_out, int blocks_per_sm = 3) {
    const int ncg = C >> 2;
    const int gw = nn::cg_col_width(C);
    const int ncols = (ncg + gw - 1) / gw;
    const int rpp = bn::TH / gw;                                  // row lanes of a full-width column
    long long want = ((long long)lr::sm_count() * blocks_per_sm + ncols - 1) / ncols;
    long long rpb = (rows + want - 1) / want;
    if (rpb < 4LL * rpp) rpb = 4LL * rpp;
    rpb = ((rpb + rpp - 1) / rpp) * rpp;
    *grid = dim3((unsigned)((rows + rpb - 1) / rpb), (unsigned)ncols);
    *gw_out = gw;
    return (int)rpb;
}

static bn::Bn make_bn(long long rows, int C, const double* stats, const float* gamma, const float* beta, float* rm,
                      float* rv, long long* nbt, float eps, float momentum, int act, int training) {
    bn::Bn b;
    b.rows = rows; b.C = C; b.stats = stats; b.gamma = gamma; b.beta = beta; b.running_mean = rm; b.running_var = rv;
    b.nbt = nbt; b.eps = eps; b.momentum = momentum; b.act = act; b.training = training;
    return b;
}

#define LR_BN_CHECK(name)                                                                        \
    LR_CHECK_ARG(rows >= 0 && C > 0 && (C & 3) == 0, name ": need rows >= 0 and C %% 4 == 0");  \
    LR_CHECK_ARG(act >= LR_ACT_NONE && act <= LR_ACT_RELU6, name ": bad activation");        \
    LR_CHECK_ARG(training ? stats != nullptr : (running_mean && running_var), name ": missing statistics"); \
    if (rows == 0) return LR_OK

template <typename T>
static int bn_act_fwd_impl(const T* x, const double* stats, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, long long* num_batches_tracked, float eps,
                           float momentum, int act, int training, const T* residual, int res_pre, T* z,
                           long long rows, int C, lr_stream_t stream) {
    LR_BN_CHECK("lr_bn_act_fwd");
    LR_CHECK_ARG(x && gamma && beta && z, "lr_bn_act_fwd: null pointer");
    LR_CHECK_ALIGN(x); LR_CHECK_ALIGN(z); LR_CHECK_ALIGN(residual);
    dim3 grid; int gw; const int rpb = rows_per_block_for(rows, C, &grid, &gw);
    const bn::Bn b = make_bn(rows, C, stats, gamma, beta, running_mean, running_var, num_batches_tracked, eps, momentum, act, training);
#define LR_BN_FWD(ACT_)                                                                                         \
    do {                                                                                                        \
        if (residual) bn::bn_act_fwd_kernel<T, true, ACT_><<<grid, bn::TH, 0, stream>>>(b, x, residual, z, rpb, res_pre, gw); \
        else bn::bn_act_fwd_kernel<T, false, ACT_><<<grid, bn::TH, 0, stream>>>(b, x, residual, z, rpb, res_pre, gw);        \
    } while (0)
    switch (act) {
        case LR_ACT_RELU: LR_BN_FWD(LR_ACT_RELU); break;
        case LR_ACT_HSWISH: LR_BN_FWD(LR_ACT_HSWISH); break;
        case LR_ACT_HSIGMOID: LR_BN_FWD(LR_ACT_HSIGMOID); break;
        case LR_ACT_RELU6: LR_BN_FWD(LR_ACT_RELU6); break;
        default: LR_BN_FWD(LR_ACT_NONE); break;
    }
#undef LR_BN_FWD
    lr::count_launch();
    LR_CHECK_LAUNCH("bn_act_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_bn_act_fwd(const float* x, const double* stats, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, long long* num_batches_tracked, float eps,
                             float momentum, int act, int training, const float* residual, int res_pre, float* z,
                             long long rows, int C, lr_stream_t stream) {
    return bn_act_fwd_impl<float>(x, stats, gamma, beta, running_mean, running_var, num_batches_tracked, eps, momentum, act,
                                  training, residual, res_pre, z, rows, C, stream);
}
extern "C" int lr_bn_act_fwd_h(const void* x, const double* stats, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, long long* num_batches_tracked, float eps,
                               float momentum, int act, int training, const void* residual, int res_pre, void* z,
                               long long rows, int C, lr_stream_t stream) {
    return bn_act_fwd_impl<nn::bf16>(static_cast<const nn::bf16*>(x), stats, gamma, beta, running_mean, running_var,
                                     num_batches_tracked, eps, momentum, act, training,
                                     static_cast<const nn::bf16*>(residual), res_pre, static_cast<nn::bf16*>(z), rows, C, stream);
}

template <typename T>
static int bn_act_bwd_impl(const T* x, const double* stats, const float* gamma, const float* beta,
                           const float* running_mean, const float* running_var, float eps, int act, int training,
                           const T* dz, const T* z_out, T* dres, double* sums, T* dx, float* dgamma, float* dbeta,
                           long long rows, int C, lr_stream_t stream) {
    LR_BN_CHECK("lr_bn_act_bwd");
    LR_CHECK_ARG(x && gamma && beta && dz && sums && dx, "lr_bn_act_bwd: null pointer");
    LR_CHECK_ALIGN(x); LR_CHECK_ALIGN(dz); LR_CHECK_ALIGN(dx); LR_CHECK_ALIGN(z_out); LR_CHECK_ALIGN(dres);
    LR_CHECK_ARG(!z_out || act == LR_ACT_RELU || act == LR_ACT_RELU6 || act == LR_ACT_NONE,
                 "lr_bn_act_bwd: output-form derivative exists for ReLU / ReLU6 only");
    dim3 grid; int gw; const int rpb = rows_per_block_for(rows, C, &grid, &gw);
    const bn::Bn b = make_bn(rows, C, stats, gamma, beta, const_cast<float*>(running_mean),
                             const_cast<float*>(running_var), nullptr, eps, 0.f, act, training);
    const size_t smem = 2 * (size_t)C * sizeof(float);
#define LR_BN_BWD(ACT_)                                                                                         \
    do {                                                                                                        \
        if (z_out) bn::bn_act_bwd_reduce_kernel<T, true, ACT_><<<grid, bn::TH, smem, stream>>>(b, x, dz, z_out, sums, rpb, gw); \
        else bn::bn_act_bwd_reduce_kernel<T, false, ACT_><<<grid, bn::TH, smem, stream>>>(b, x, dz, z_out, sums, rpb, gw);      \
        lr::count_launch();                                                                                     \
        if (z_out) bn::bn_act_bwd_apply_kernel<T, true, ACT_><<<grid, bn::TH, 0, stream>>>(b, x, dz, z_out, dres, sums, dx, dgamma, dbeta, rpb, gw); \
        else bn::bn_act_bwd_apply_kernel<T, false, ACT_><<<grid, bn::TH, 0, stream>>>(b, x, dz, z_out, dres, sums, dx, dgamma, dbeta, rpb, gw);      \
    } while (0)
    switch (act) {
        case LR_ACT_RELU: LR_BN_BWD(LR_ACT_RELU); break;
        case LR_ACT_HSWISH: LR_BN_BWD(LR_ACT_HSWISH); break;
        case LR_ACT_HSIGMOID: LR_BN_BWD(LR_ACT_HSIGMOID); break;
        case LR_ACT_RELU6: LR_BN_BWD(LR_ACT_RELU6); break;
        default: LR_BN_BWD(LR_ACT_NONE); break;
    }
#undef LR_BN_BWD
    LR_CHECK_LAUNCH("bn_act_bwd_reduce_kernel");
    lr::count_launch();
    LR_CHECK_LAUNCH("bn_act_bwd_apply_kernel");
    return LR_OK;
}

extern "C" int lr_bn_act_bwd(const float* x, const double* stats, const float* gamma, const float* beta,
                             const float* running_mean, const float* running_var, float eps, int act, int training,
                             const float* dz, const float* z_out, float* dres, double* sums /*[2C], zeroed by the caller*/,
                             float* dx, float* dgamma, float* dbeta, long long rows, int C, lr_stream_t stream) {
    return bn_act_bwd_impl<float>(x, stats, gamma, beta, running_mean, running_var, eps, act, training, dz, z_out, dres, sums,
                                  dx, dgamma, dbeta, rows, C, stream);
}
extern "C" int lr_bn_act_bwd_h(const void* x, const double* stats, const float* gamma, const float* beta,
                               const float* running_mean, const float* running_var, float eps, int act, int training,
                               const void* dz, const void* z_out, void* dres, double* sums, void* dx, float* dgamma,
                               float* dbeta, long long rows, int C, lr_stream_t stream) {
    typedef nn::bf16 H;
    return bn_act_bwd_impl<H>(static_cast<const H*>(x), stats, gamma, beta, running_mean, running_var, eps, act, training,
                              static_cast<const H*>(dz), static_cast<const H*>(z_out), static_cast<H*>(dres), sums,
                              static_cast<H*>(dx), dgamma, dbeta, rows, C, stream);
}

template <typename T>
static int frame_reduce_impl(const T* a, const T* g, float* p, int F, int HW, int C, int mode, lr_stream_t stream) {
    LR_CHECK_ARG(F >= 0 && HW > 0 && C > 0 && (C & 3) == 0, "lr_frame_reduce: bad shape");
    LR_CHECK_ARG(mode == 0 || (mode == 1 && g), "lr_frame_reduce: bad mode");
    if (F == 0) return LR_OK;
    LR_CHECK_ARG(a && p, "lr_frame_reduce: null pointer");
    LR_CHECK_ALIGN(a); LR_CHECK_ALIGN(g);
    bn::frame_reduce_kernel<T><<<F, bn::TH, 0, stream>>>(a, g, p, HW, C, mode);
    lr::count_launch();
    LR_CHECK_LAUNCH("frame_reduce_kernel");
    return LR_OK;
}
extern "C" int lr_frame_reduce(const float* a, const float* g, float* p, int F, int HW, int C, int mode,
                               lr_stream_t stream) {
    return frame_reduce_impl<float>(a, g, p, F, HW, C, mode, stream);
}
extern "C" int lr_frame_reduce_h(const void* a, const void* g, float* p, int F, int HW, int C, int mode,
                                 lr_stream_t stream) {
    return frame_reduce_impl<nn::bf16>(static_cast<const nn::bf16*>(a), static_cast<const nn::bf16*>(g), p, F, HW, C, mode, stream);
}

template <typename T>
static int frame_scale_impl(const T* a, const float* s, const float* dp, T* out, int F, int HW, int C, lr_stream_t stream) {
    LR_CHECK_ARG(F >= 0 && HW > 0 && C > 0 && (C & 3) == 0, "lr_frame_scale: bad shape");
    LR_CHECK_ARG((a && s) || dp, "lr_frame_scale: nothing to do");
    if (F == 0) return LR_OK;
    LR_CHECK_ARG(out, "lr_frame_scale: null pointer");
    LR_CHECK_ALIGN(a); LR_CHECK_ALIGN(s); LR_CHECK_ALIGN(dp); LR_CHECK_ALIGN(out);
    const int gw = nn::cg_col_width(C);
    const int ncols = ((C >> 2) + gw - 1) / gw;
    int gx = (8 * lr::sm_count() + ncols - 1) / ncols;          // about eight blocks per SM in all, never more than frames
    if (gx > F) gx = F;
    const dim3 grid((unsigned)gx, (unsigned)ncols);
    const float inv_hw = 1.f / (float)HW;
    if (a && dp) bn::frame_scale_kernel<T, true, true><<<grid, bn::TH, 0, stream>>>(a, s, dp, out, F, HW, C, inv_hw, gw);
    else if (a) bn::frame_scale_kernel<T, true, false><<<grid, bn::TH, 0, stream>>>(a, s, dp, out, F, HW, C, inv_hw, gw);
    else bn::frame_scale_kernel<T, false, true><<<grid, bn::TH, 0, stream>>>(a, s, dp, out, F, HW, C, inv_hw, gw);
    lr::count_launch();
    LR_CHECK_LAUNCH("frame_scale_kernel");
    return LR_OK;
}
extern "C" int lr_frame_scale(const float* a, const float* s, const float* dp, float* out, int F, int HW, int C,
                              lr_stream_t stream) {
    return frame_scale_impl<float>(a, s, dp, out, F, HW, C, stream);
}
extern "C" int lr_frame_scale_h(const void* a, const float* s, const float* dp, void* out, int F, int HW, int C,
                                lr_stream_t stream) {
    return frame_scale_impl<nn::bf16>(static_cast<const nn::bf16*>(a), s, dp, static_cast<nn::bf16*>(out), F, HW, C, stream);
}

template <typename T>
static int act_bwd_impl(T* dy, const T* y, long long n, int act, lr_stream_t stream) {
    LR_CHECK_ARG(n >= 0, "lr_act_bwd: negative size");
    LR_CHECK_ARG(act == LR_ACT_RELU || act == LR_ACT_HSIGMOID || act == LR_ACT_RELU6 || act == LR_ACT_NONE, "lr_act_bwd: activation has no output-form derivative");
    if (n == 0 || act == LR_ACT_NONE) return LR_OK;
    LR_CHECK_ARG(dy && y, "lr_act_bwd: null pointer");
    long long g = (n + bn::TH - 1) / bn::TH;
    const long long cap = (long long)lr::sm_count() * 16;
    if (g > cap) g = cap;
    bn::act_bwd_kernel<T><<<(unsigned)g, bn::TH, 0, stream>>>(dy, y, n, act);
    lr::count_launch();
    LR_CHECK_LAUNCH("act_bwd_kernel");
    return LR_OK;
}
extern "C" int lr_act_bwd(float* dy, const float* y, long long n, int act, lr_stream_t stream) {
    return act_bwd_impl<float>(dy, y, n, act, stream);
}
extern "C" int lr_act_bwd_h(void* dy, const void* y, long long n, int act, lr_stream_t stream) {
    return act_bwd_impl<nn::bf16>(static_cast<nn::bf16*>(dy), static_cast<const nn::bf16*>(y), n, act, stream);
}

extern "C" int lr_act_fwd(const float* x, float* y, long long n, int act, lr_stream_t stream) {
    LR_CHECK_ARG(n >= 0, "lr_act_fwd: negative size");
    LR_CHECK_ARG(act >= LR_ACT_NONE && act <= LR_ACT_RELU6, "lr_act_fwd: bad activation");
    if (n == 0) return LR_OK;
    LR_CHECK_ARG(x && y, "lr_act_fwd: null pointer");
    long long g = (n + bn::TH - 1) / bn::TH;
    const long long cap = (long long)lr::sm_count() * 16;
    if (g > cap) g = cap;
    bn::act_fwd_kernel<<<(unsigned)g, bn::TH, 0, stream>>>(x, y, n, act);
    lr::count_launch();
    LR_CHECK_LAUNCH("act_fwd_kernel");
    return LR_OK;
}

template <typename T>
static int colsum_impl(const T* dY, long long ld, long long M, int N, float* db, lr_stream_t stream) {
    LR_CHECK_ARG(M >= 0 && N > 0, "lr_colsum: bad shape");
    if (M == 0) return LR_OK;
    LR_CHECK_ARG(dY && db, "lr_colsum: null pointer");
    const int cols = N < bn::TH ? N : bn::TH;
    const int rpp = bn::TH / cols;
    long long rpb = (M + (long long)lr::sm_count() * 2 - 1) / ((long long)lr::sm_count() * 2);
    rpb = ((rpb + rpp - 1) / rpp) * rpp;
    if (rpb < rpp) rpb = rpp;
    dim3 grid((unsigned)((M + rpb - 1) / rpb), 1);
    bn::colsum_kernel<T><<<grid, bn::TH, 0, stream>>>(dY, ld, M, N, db, (int)rpb);
    lr::count_launch();
    LR_CHECK_LAUNCH("colsum_kernel");
    return LR_OK;
}
extern "C" int lr_colsum(const float* dY, long long ld, long long M, int N, float* db, lr_stream_t stream) {
    return colsum_impl<float>(dY, ld, M, N, db, stream);
}
extern "C" int lr_colsum_h(const void* dY, long long ld, long long M, int N, float* db, lr_stream_t stream) {
    return colsum_impl<nn::bf16>(static_cast<const nn::bf16*>(dY), ld, M, N, db, stream);
}

// fp32 -> bf16 copy of a contiguous buffer (the bf16 shadow of the flat parameter buffer, refreshed once per step,
// and of re-laid-out weight matrices); n need not be a multiple of 4
namespace bn {
__global__ void __launch_bounds__(TH)
cast_bf16_kernel(const float* __restrict__ src, nn::bf16* __restrict__ dst, long long n) {
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < n4; i += (long long)gridDim.x * TH)
        nn::st4(dst + 4 * i, nn::ld4(src + 4 * i));
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[4 * n4 + threadIdx.x] = __float2bfloat16_rn(src[4 * n4 + threadIdx.x]);
}
}  // namespace bn
extern "C" int lr_cast_bf16(const float* src, void* dst, long long n, lr_stream_t stream) {
    LR_CHECK_ARG(n >= 0, "lr_cast_bf16: negative size");
    if (n == 0) return LR_OK;
    LR_CHECK_ARG(src && dst, "lr_cast_bf16: null pointer");
    LR_CHECK_ALIGN(src);
    LR_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & 7) == 0, "lr_cast_bf16: dst must be 8-byte aligned");
    long long g = ((n >> 2) + bn::TH - 1) / bn::TH;
    const long long cap = (long long)lr::sm_count() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    bn::cast_bf16_kernel<<<(unsigned)g, bn::TH, 0, stream>>>(src, static_cast<nn::bf16*>(dst), n);
    lr::count_launch();
    LR_CHECK_LAUNCH("cast_bf16_kernel");
    return LR_OK;
}

extern "C" int lr_copy2d(float* dst, long long ldd, const float* src, long long lds, int rows, int cols,
                         lr_stream_t stream) {
    LR_CHECK_ARG(rows >= 0 && cols >= 0, "lr_copy2d: negative shape");
    if (rows == 0 || cols == 0) return LR_OK;
    LR_CHECK_ARG(dst && src, "lr_copy2d: null pointer");
    long long g = ((long long)rows * cols + bn::TH - 1) / bn::TH;
    const long long cap = (long long)lr::sm_count() * 8;
    if (g > cap) g = cap;
    bn::copy2d_kernel<<<(unsigned)g, bn::TH, 0, stream>>>(dst, ldd, src, lds, rows, cols);
    lr::count_launch();
    LR_CHECK_LAUNCH("copy2d_kernel");
    return LR_OK;
}
