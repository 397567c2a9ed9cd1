// nn.MultiheadAttention over the time axis of a clip (T <= 64 frames, head width <= 128): the temporal heads of
// video/models/resnet_attn.py:23-35,95-111 and of the TransformerEncoder layers (resnet_trans.py:96-103).
// One CTA per (clip, head): the q / k / v tiles of that head ([T, d] each, read out of the packed in-projection
// [B*T, 3E]) live in shared memory; the probability matrix P [B, heads, T, T] is the only saved tensor.
// The work is tiny (B * heads CTAs of ~0.5 MFLOP) and latency-bound; it exists so that the step stays one CUDA graph
// of lipread_b200 kernels.  Scores / apply are separate entry points so that attention dropout (lr_dropout_fwd on P)
// can sit between them exactly where torch applies it.
#include "common.cuh"

namespace mha {

constexpr int TH = 128;
constexpr int MAXT = 64;
constexpr int MAXD = 128;
constexpr size_t SMEM_CEILING = (2 * MAXT * (MAXD + 1) + MAXT * (MAXT + 1) * 2) * sizeof(float);

struct Geo { int B, T, E, heads, d; long long ld; };

__device__ __forceinline__ void load_tile(float* dst, const float* src, long long ld, int T, int d, float scale) {
    for (int idx = threadIdx.x; idx < T * d; idx += TH) {
        const int i = idx / d, c = idx - i * d;
        dst[i * (d + 1) + c] = src[(long long)i * ld + c] * scale;
    }
}

// P[b,h,i,:] = softmax_j( (q_i * scale) . k_j )
__global__ void __launch_bounds__(TH) scores_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ P, Geo g, float scale) {
    extern __shared__ float sm[];
    const int b = blockIdx.x / g.heads, h = blockIdx.x % g.heads, T = g.T, d = g.d;
    float* q = sm; float* k = q + T * (d + 1); float* s = k + T * (d + 1);
    const float* base = qkv + (long long)b * T * g.ld + h * d;
    load_tile(q, base, g.ld, T, d, scale);
    load_tile(k, base + g.E, g.ld, T, d, 1.f);
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * T; idx += TH) {
        const int i = idx / T, j = idx - i * T;
        float acc = 0.f;
        for (int c = 0; c < d; ++c) acc = fmaf(q[i * (d + 1) + c], k[j * (d + 1) + c], acc);
        s[i * (T + 1) + j] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T; i += TH) {
        float* row = s + i * (T + 1);
        float m = row[0];
        for (int j = 1; j < T; ++j) m = fmaxf(m, row[j]);
        float sum = 0.f;
        for (int j = 0; j < T; ++j) { row[j] = expf(row[j] - m); sum += row[j]; }
        const float inv = 1.f / sum;
        float* out = P + ((long long)blockIdx.x * T + i) * T;
        for (int j = 0; j < T; ++j) out[j] = row[j] * inv;
    }
}

// O[b, i, h*d + c] = sum_j P[b,h,i,j] v[b, j, h*d + c]
__global__ void __launch_bounds__(TH) apply_fwd_kernel(const float* __restrict__ P, const float* __restrict__ qkv,
                                                       float* __restrict__ O, Geo g) {
    extern __shared__ float sm[];
    const int b = blockIdx.x / g.heads, h = blockIdx.x % g.heads, T = g.T, d = g.d;
    float* v = sm; float* p = v + T * (d + 1);
    load_tile(v, qkv + (long long)b * T * g.ld + 2 * g.E + h * d, g.ld, T, d, 1.f);
    for (int idx = threadIdx.x; idx < T * T; idx += TH) p[(idx / T) * (T + 1) + idx % T] = P[(long long)blockIdx.x * T * T + idx];
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * d; idx += TH) {
        const int i = idx / d, c = idx - i * d;
        float acc = 0.f;
        for (int j = 0; j < T; ++j) acc = fmaf(p[i * (T + 1) + j], v[j * (d + 1) + c], acc);
        O[((long long)b * T + i) * g.E + h * d + c] = acc;
    }
}

// dP[b,h,i,j] = sum_c dO[b,i,hd+c] v[b,j,hd+c];   dV[b,j,hd+c] = sum_i P[b,h,i,j] dO[b,i,hd+c]
__global__ void __launch_bounds__(TH) apply_bwd_kernel(const float* __restrict__ dO, const float* __restrict__ P,
                                                       const float* __restrict__ qkv, float* __restrict__ dP,
                                                       float* __restrict__ dqkv, Geo g) {
    extern __shared__ float sm[];
    const int b = blockIdx.x / g.heads, h = blockIdx.x % g.heads, T = g.T, d = g.d;
    float* v = sm; float* go = v + T * (d + 1); float* p = go + T * (d + 1);
    load_tile(v, qkv + (long long)b * T * g.ld + 2 * g.E + h * d, g.ld, T, d, 1.f);
    load_tile(go, dO + (long long)b * T * g.E + h * d, g.E, T, d, 1.f);
    for (int idx = threadIdx.x; idx < T * T; idx += TH) p[(idx / T) * (T + 1) + idx % T] = P[(long long)blockIdx.x * T * T + idx];
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * T; idx += TH) {
        const int i = idx / T, j = idx - i * T;
        float acc = 0.f;
        for (int c = 0; c < d; ++c) acc = fmaf(go[i * (d + 1) + c], v[j * (d + 1) + c], acc);
        dP[(long long)blockIdx.x * T * T + idx] = acc;
    }
    float* dv = dqkv + (long long)b * T * g.ld + 2 * g.E + h * d;
    for (int idx = threadIdx.x; idx < T * d; idx += TH) {
        const int j = idx / d, c = idx - j * d;
        float acc = 0.f;
        for (int i = 0; i < T; ++i) acc = fmaf(p[i * (T + 1) + j], go[i * (d + 1) + c], acc);
        dv[(long long)j * g.ld + c] = acc;
    }
}

// dS = P o (dP - rowsum(dP o P));  dq_i = scale * sum_j dS_ij k_j;  dk_j = scale * sum_i dS_ij q_i
__global__ void __launch_bounds__(TH) scores_bwd_kernel(const float* __restrict__ P, const float* __restrict__ dP,
                                                        const float* __restrict__ qkv, float* __restrict__ dqkv, Geo g,
                                                        float scale) {
    extern __shared__ float sm[];
    const int b = blockIdx.x / g.heads, h = blockIdx.x % g.heads, T = g.T, d = g.d;
    float* q = sm; float* k = q + T * (d + 1); float* ds = k + T * (d + 1);
    const float* base = qkv + (long long)b * T * g.ld + h * d;
    load_tile(q, base, g.ld, T, d, 1.f);
    load_tile(k, base + g.E, g.ld, T, d, 1.f);
    for (int i = threadIdx.x; i < T; i += TH) {
        const float* pr = P + ((long long)blockIdx.x * T + i) * T;
        const float* dr = dP + ((long long)blockIdx.x * T + i) * T;
        float dot = 0.f;
        for (int j = 0; j < T; ++j) dot = fmaf(pr[j], dr[j], dot);
        for (int j = 0; j < T; ++j) ds[i * (T + 1) + j] = pr[j] * (dr[j] - dot) * scale;
    }
    __syncthreads();
    float* dq = dqkv + (long long)b * T * g.ld + h * d;
    float* dk = dq + g.E;
    for (int idx = threadIdx.x; idx < T * d; idx += TH) {
        const int i = idx / d, c = idx - i * d;
        float aq = 0.f, ak = 0.f;
        for (int j = 0; j < T; ++j) {
            aq = fmaf(ds[i * (T + 1) + j], k[j * (d + 1) + c], aq);
            ak = fmaf(ds[j * (T + 1) + i], q[j * (d + 1) + c], ak);
        }
        dq[(long long)i * g.ld + c] = aq;
        dk[(long long)i * g.ld + c] = ak;
    }
}

static int geometry(Geo& g, int B, int T, int E, int heads, long long ld, const char* who) {
    if (!(B >= 0 && T >= 1 && T <= MAXT && heads >= 1 && E % heads == 0 && E / heads <= MAXD && ld >= 3LL * E))
        return lr::fail(LR_EINVAL, "%s: bad shape (B %d, T %d <= %d, E %d, heads %d, head width <= %d, ld %lld >= 3E)", who,
                        B, T, MAXT, E, heads, MAXD, ld);
    g = Geo{B, T, E, heads, E / heads, ld};
    return LR_OK;
}

template <class K>
static void allow_smem(K kernel) {      // fixed ceiling, set once per kernel (never lowered by a later, smaller launch)
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_CEILING);
}

static size_t smem_bytes(const Geo& g) {
    return ((size_t)2 * g.T * (g.d + 1) + (size_t)g.T * (g.T + 1)) * sizeof(float);
}

}  // namespace mha

extern "C" int lr_mha_scores_fwd(const float* qkv, long long ld, float* P, int B, int T, int E, int heads, lr_stream_t stream) {
    mha::Geo g;
    if (int rc = mha::geometry(g, B, T, E, heads, ld, "lr_mha_scores_fwd")) return rc;
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(qkv && P, "lr_mha_scores_fwd: null pointer");
    static const bool once = (mha::allow_smem(mha::scores_fwd_kernel), true);
    (void)once;
    mha::scores_fwd_kernel<<<B * heads, mha::TH, mha::smem_bytes(g), stream>>>(qkv, P, g, sqrtf(1.f / (float)g.d));
    lr::count_launch();
    LR_CHECK_LAUNCH("mha::scores_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_mha_apply_fwd(const float* P, const float* qkv, long long ld, float* O, int B, int T, int E, int heads,
                                lr_stream_t stream) {
    mha::Geo g;
    if (int rc = mha::geometry(g, B, T, E, heads, ld, "lr_mha_apply_fwd")) return rc;
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(P && qkv && O, "lr_mha_apply_fwd: null pointer");
    static const bool once = (mha::allow_smem(mha::apply_fwd_kernel), true);
    (void)once;
    mha::apply_fwd_kernel<<<B * heads, mha::TH, mha::smem_bytes(g), stream>>>(P, qkv, O, g);
    lr::count_launch();
    LR_CHECK_LAUNCH("mha::apply_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_mha_apply_bwd(const float* dO, const float* P, const float* qkv, long long ld, float* dP, float* dqkv,
                                int B, int T, int E, int heads, lr_stream_t stream) {
    mha::Geo g;
    if (int rc = mha::geometry(g, B, T, E, heads, ld, "lr_mha_apply_bwd")) return rc;
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(dO && P && qkv && dP && dqkv, "lr_mha_apply_bwd: null pointer");
    static const bool once = (mha::allow_smem(mha::apply_bwd_kernel), true);
    (void)once;
    mha::apply_bwd_kernel<<<B * heads, mha::TH, mha::smem_bytes(g), stream>>>(dO, P, qkv, dP, dqkv, g);
    lr::count_launch();
    LR_CHECK_LAUNCH("mha::apply_bwd_kernel");
    return LR_OK;
}

extern "C" int lr_mha_scores_bwd(const float* P, const float* dP, const float* qkv, long long ld, float* dqkv, int B, int T,
                                 int E, int heads, lr_stream_t stream) {
    mha::Geo g;
    if (int rc = mha::geometry(g, B, T, E, heads, ld, "lr_mha_scores_bwd")) return rc;
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(P && dP && qkv && dqkv, "lr_mha_scores_bwd: null pointer");
    static const bool once = (mha::allow_smem(mha::scores_bwd_kernel), true);
    (void)once;
    mha::scores_bwd_kernel<<<B * heads, mha::TH, mha::smem_bytes(g), stream>>>(P, dP, qkv, dqkv, g, sqrtf(1.f / (float)g.d));
    lr::count_launch();
    LR_CHECK_LAUNCH("mha::scores_bwd_kernel");
    return LR_OK;
}
