// General fp32 GEMM used for every dense contraction on the path that is not a depthwise / stem conv:
// 1x1 convolutions (channels-last activations make them plain GEMMs), SE / LSTM / classifier
// linears, and all their dgrad / wgrad passes.
//
//   C[M,N] (ldc) = epilogue( sum_k A(m,k) * B(k,n) )
//     a_trans = 0: A stored [M][K] (lda)        a_trans = 1: A stored [K][M] (lda)   (wgrad: dY^T)
//     b_trans = 0: B stored [N][K] (ldb)  "NT"  b_trans = 1: B stored [K][N] (ldb)   "NN"
//   epilogue: + bias[n]  -> act -> + R[m,n] (ldr; may alias C = accumulate) -> store
//             optional per-column sum / sum-of-squares of the stored value into double stats[2N]
//             (train-mode BatchNorm statistics come out of the producing conv for free)
//   ksplit > 1: the K range is split over gridDim.z and partial sums are atomically added into C
//             (C must already hold zeros or the value to accumulate onto); used for the long
//             reductions of wgrad (K = frames*pixels) and for audio_fc (K = 37120, M = batch).
#include "nn_common.cuh"

namespace gm {

constexpr int BM = 64, BN = 64, BK = 16, TH = 256, LD = 68;

struct P {
    const float* A; const float* B; float* C;
    int M, N, K;
    long long lda, ldb, ldc, ldr;
    const float* bias; const float* R;
    double* stats;
    int act, kchunk;
    float* part;            // ordered split-K: slice z stores its partial tile at part[z][M][N] instead of adding into C
};

// tile source contiguous along k: stored [rows][K]; thread -> (row = t>>2, 4 consecutive k)
__device__ __forceinline__ float4 load_kc(const float* __restrict__ S, long long ld, int rows, int K, int row0,
                                          int k0, bool vec_ok) {
    const int t = threadIdx.x, r = row0 + (t >> 2), k = k0 + (t & 3) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
        const float* p = S + (long long)r * ld + k;
        if (vec_ok && k + 3 < K) v = nn::ld4(p);
        else {
            if (k < K) v.x = p[0];
            if (k + 1 < K) v.y = p[1];
            if (k + 2 < K) v.z = p[2];
            if (k + 3 < K) v.w = p[3];
        }
    }
    return v;
}
__device__ __forceinline__ void store_kc(float (*T)[LD], float4 v) {
    const int t = threadIdx.x, r = t >> 2, k = (t & 3) * 4;
    T[k][r] = v.x; T[k + 1][r] = v.y; T[k + 2][r] = v.z; T[k + 3][r] = v.w;
}
// tile source contiguous along m/n: stored [K][cols]; thread -> (k = t>>4, 4 consecutive cols)
__device__ __forceinline__ float4 load_mc(const float* __restrict__ S, long long ld, int cols, int K, int col0,
                                          int k0, bool vec_ok) {
    const int t = threadIdx.x, k = k0 + (t >> 4), c = col0 + (t & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K) {
        const float* p = S + (long long)k * ld + c;
        if (vec_ok && c + 3 < cols) v = nn::ld4(p);
        else {
            if (c < cols) v.x = p[0];
            if (c + 1 < cols) v.y = p[1];
            if (c + 2 < cols) v.z = p[2];
            if (c + 3 < cols) v.w = p[3];
        }
    }
    return v;
}
__device__ __forceinline__ void store_mc(float (*T)[LD], float4 v) {
    const int t = threadIdx.x;
    *reinterpret_cast<float4*>(&T[t >> 4][(t & 15) * 4]) = v;
}

template <bool AT, bool BT>
__global__ void __launch_bounds__(TH) gemm_kernel(const P p) {
    __shared__ __align__(16) float As[BK][LD];
    __shared__ __align__(16) float Bs[BK][LD];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * p.kchunk;
    const int kend = min(p.K, kbeg + p.kchunk);
    const bool a_vec = ((p.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    const bool b_vec = ((p.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra, rb;
    if (kbeg < kend) {
        ra = AT ? load_mc(p.A, p.lda, p.M, kend, m0, kbeg, a_vec) : load_kc(p.A, p.lda, p.M, kend, m0, kbeg, a_vec);
        rb = BT ? load_mc(p.B, p.ldb, p.N, kend, n0, kbeg, b_vec) : load_kc(p.B, p.ldb, p.N, kend, n0, kbeg, b_vec);
    }
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        if (AT) store_mc(As, ra); else store_kc(As, ra);
        if (BT) store_mc(Bs, rb); else store_kc(Bs, rb);
        __syncthreads();
        if (k0 + BK < kend) {
            ra = AT ? load_mc(p.A, p.lda, p.M, kend, m0, k0 + BK, a_vec) : load_kc(p.A, p.lda, p.M, kend, m0, k0 + BK, a_vec);
            rb = BT ? load_mc(p.B, p.ldb, p.N, kend, n0, k0 + BK, b_vec) : load_kc(p.B, p.ldb, p.N, kend, n0, k0 + BK, b_vec);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---------------- epilogue
    const int n = n0 + tx * 4;
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.bias && blockIdx.z == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (n + j < p.N) bias[j] = p.bias[n + j];
    }
    float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};
    const bool c_vec = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (n + 3 < p.N);
    const bool r_vec = p.R && ((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.R) & 15) == 0) && (n + 3 < p.N);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
        float* crow = p.C + (long long)m * p.ldc + n;
        if (p.part) {                                      // ordered split-K: plain stores, reduced in slice order later
            float* prow = p.part + ((long long)blockIdx.z * p.M + m) * p.N + n;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < p.N) prow[j] = acc[i][j];
            continue;
        }
        if (gridDim.z > 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < p.N) atomicAdd(crow + j, v[j]);
            continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = nn::act_fwd(v[j], p.act);
        if (p.R) {
            const float* rrow = p.R + (long long)m * p.ldr + n;
            if (r_vec) { const float4 r = nn::ld4(rrow); v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w; }
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (n + j < p.N) v[j] += rrow[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { csum[j] += v[j]; csq[j] = fmaf(v[j], v[j], csq[j]); }
        if (c_vec) nn::st4(crow, make_float4(v[0], v[1], v[2], v[3]));
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < p.N) crow[j] = v[j];
        }
    }
    if (p.stats && gridDim.z == 1) {
        // reduce the 16 row-groups (ty) of each column through shared memory, then one double atomic
        // per column and block
        float (*red)[LD] = As;                       // [16][64] sums   (loop above ended with a barrier)
        float (*req)[LD] = Bs;                       // [16][64] squares
#pragma unroll
        for (int j = 0; j < 4; ++j) { red[ty][tx * 4 + j] = csum[j]; req[ty][tx * 4 + j] = csq[j]; }
        __syncthreads();
        if (t < BN && n0 + t < p.N) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) { s += red[r][t]; q += req[r][t]; }
            nn::atomic_add_double(p.stats + n0 + t, (double)s);
            nn::atomic_add_double(p.stats + p.N + n0 + t, (double)q);
        }
    }
}

// C[m,n] = act(sum_z part[z][m][n] + bias[n]), slices added in z order (deterministic split-K, second pass)
__global__ void __launch_bounds__(TH)
splitk_reduce_kernel(const float* __restrict__ part, int nz, float* __restrict__ C, long long ldc, int M, int N,
                     const float* __restrict__ bias, int act) {
    const long long total = (long long)M * N;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < total; i += (long long)gridDim.x * TH) {
        const int m = (int)(i / N), n = (int)(i - (long long)m * N);
        float s = 0.f;
        for (int z = 0; z < nz; ++z) s += part[(long long)z * total + i];
        if (bias) s += bias[n];
        C[(long long)m * ldc + n] = nn::act_fwd(s, act);
    }
}

}  // namespace gm

extern "C" int lr_gemm(const float* A, long long lda, int a_trans, const float* B, long long ldb, int b_trans,
                       float* C, long long ldc, int M, int N, int K, const float* bias, int act,
                       const float* R, long long ldr, double* stats, int ksplit, lr_stream_t stream) {
    LR_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "lr_gemm: negative dimension");
    if (M == 0 || N == 0) return LR_OK;
    LR_CHECK_ARG(A && B && C, "lr_gemm: null pointer");
    LR_CHECK_ARG(act >= LR_ACT_NONE && act <= LR_ACT_RELU6, "lr_gemm: bad activation %d", act);
    LR_CHECK_ARG(ksplit >= 1, "lr_gemm: ksplit must be >= 1");
    LR_CHECK_ARG(ksplit == 1 || (act == LR_ACT_NONE && !R && !stats),
                 "lr_gemm: split-K accumulates atomically and cannot fuse act / residual / stats");
    gm::P p;
    p.A = A; p.B = B; p.C = C; p.M = M; p.N = N; p.K = K;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldr = ldr; p.bias = bias; p.R = R; p.stats = stats; p.act = act;
    p.part = nullptr;
    int kchunk = (K + ksplit - 1) / ksplit;
    kchunk = ((kchunk + gm::BK - 1) / gm::BK) * gm::BK;
    if (kchunk == 0) kchunk = gm::BK;
    const int nz = K == 0 ? 1 : (K + kchunk - 1) / kchunk;
    p.kchunk = kchunk;
    const long long mt = (M + gm::BM - 1) / gm::BM, nt = (N + gm::BN - 1) / gm::BN;
    LR_CHECK_ARG(nt <= 65535 && nz <= 65535, "lr_gemm: N or split too large");
    dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)nz);
    if (a_trans) {
        if (b_trans) gm::gemm_kernel<true, true><<<grid, gm::TH, 0, stream>>>(p);
        else gm::gemm_kernel<true, false><<<grid, gm::TH, 0, stream>>>(p);
    } else {
        if (b_trans) gm::gemm_kernel<false, true><<<grid, gm::TH, 0, stream>>>(p);
        else gm::gemm_kernel<false, false><<<grid, gm::TH, 0, stream>>>(p);
    }
    lr::count_launch();
    LR_CHECK_LAUNCH("gemm_kernel");
    return LR_OK;
}

extern "C" size_t lr_gemm_splitk_workspace_bytes(int M, int N, int ksplit) {
    return (size_t)(ksplit < 1 ? 1 : ksplit) * (size_t)M * (size_t)N * sizeof(float);
}

extern "C" int lr_gemm_splitk(const float* A, long long lda, int a_trans, const float* B, long long ldb, int b_trans,
                              float* C, long long ldc, int M, int N, int K, const float* bias, int act, int ksplit,
                              float* ws, size_t ws_bytes, lr_stream_t stream) {
    LR_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "lr_gemm_splitk: negative dimension");
    if (M == 0 || N == 0) return LR_OK;
    LR_CHECK_ARG(A && B && C && ws, "lr_gemm_splitk: null pointer");
    LR_CHECK_ARG(act >= LR_ACT_NONE && act <= LR_ACT_RELU6, "lr_gemm_splitk: bad activation %d", act);
    LR_CHECK_ARG(ksplit >= 1, "lr_gemm_splitk: ksplit must be >= 1");
    if (ws_bytes < lr_gemm_splitk_workspace_bytes(M, N, ksplit))
        return lr::fail(LR_ENOSPC, "lr_gemm_splitk: workspace %zu < %zu bytes", ws_bytes,
                        lr_gemm_splitk_workspace_bytes(M, N, ksplit));
    gm::P p;
    p.A = A; p.B = B; p.C = C; p.M = M; p.N = N; p.K = K;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldr = 0; p.bias = nullptr; p.R = nullptr; p.stats = nullptr; p.act = LR_ACT_NONE;
    p.part = ws;
    int kchunk = (K + ksplit - 1) / ksplit;
    kchunk = ((kchunk + gm::BK - 1) / gm::BK) * gm::BK;
    if (kchunk == 0) kchunk = gm::BK;
    const int nz = K == 0 ? 1 : (K + kchunk - 1) / kchunk;      // <= ksplit: every slice z < nz writes its whole tile
    p.kchunk = kchunk;
    const long long mt = (M + gm::BM - 1) / gm::BM, nt = (N + gm::BN - 1) / gm::BN;
    LR_CHECK_ARG(nt <= 65535 && nz <= 65535, "lr_gemm_splitk: N or split too large");
    dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)nz);
    if (a_trans) {
        if (b_trans) gm::gemm_kernel<true, true><<<grid, gm::TH, 0, stream>>>(p);
        else gm::gemm_kernel<true, false><<<grid, gm::TH, 0, stream>>>(p);
    } else {
        if (b_trans) gm::gemm_kernel<false, true><<<grid, gm::TH, 0, stream>>>(p);
        else gm::gemm_kernel<false, false><<<grid, gm::TH, 0, stream>>>(p);
    }
    lr::count_launch();
    LR_CHECK_LAUNCH("gemm_kernel (ordered split-K)");
    long long g = ((long long)M * N + gm::TH - 1) / gm::TH;
    if (g > 4LL * lr::sm_count()) g = 4LL * lr::sm_count();
    gm::splitk_reduce_kernel<<<(unsigned)g, gm::TH, 0, stream>>>(ws, nz, C, ldc, M, N, bias, act);
    lr::count_launch();
    LR_CHECK_LAUNCH("splitk_reduce_kernel");
    return LR_OK;
}
