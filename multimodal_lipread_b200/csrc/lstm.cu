// nn.LSTM recurrence (one direction of one layer), forward and BPTT, fp32.
// Reference call sites: audio_video/models/middle_fusion_fast.py:18,35-36 (BiLSTM(576,128), out[:, -1]),
// audio_video/models/early_fusion.py:62-69,82-83, video/models/resnet_lstm.py:113-120.
// Semantics (torch.nn.LSTM): gate rows ordered i, f, g, o; i,f,o = sigmoid, g = tanh;
// c_t = f c_{t-1} + i g; h_t = o tanh(c_t); h_0 = c_0 = 0; the reverse direction walks t = T-1 .. 0 and
// stores its output at the same t.
//
// The input projection x_t W_ih^T + b_ih + b_hh for all t is ONE GEMM done by the caller (lr_gemm);
// these kernels only run the sequential part.  `nsteps` <= T lets the caller run a direction over a
// prefix of its walk only: a head that reads out[:, -1] needs the reverse direction's FIRST step
// (t = T-1) and nothing else.
#include "nn_common.cuh"

namespace ls {

constexpr int R = 4;          // batch rows per CTA (W_hh is re-read once per CTA and step)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct Fwd {
    const float* xproj; long long ldx;   // [B*T, 4H] pre-activations from the input (biases included)
    const float* bhh;                    // [4H] second bias (b_hh), added here; may be null
    const float* whh;                    // [4H, H]
    float* out; long long ldo;           // h_t -> out[(b*T+t)*ldo + 0..H)   (already offset to the direction's columns)
    float* gates;                        // [B,T,4H] activated gates (saved for backward), may be null
    float* cst;                          // [B,T,H]  cell states (saved), may be null
    float* hprev;                        // [B,T,H]  h_{prev step} (saved), may be null
    int B, T, H, nsteps, reverse;
};

__global__ void __launch_bounds__(512)
lstm_fwd_kernel(const Fwd p) {
    extern __shared__ __align__(16) float sm[];
    const int H = p.H, G = 4 * p.H;
    float* hs = sm;                       // [R][H]
    float* cs = hs + R * H;               // [R][H]
    float* gs = cs + R * H;               // [R][G]
    const int b0 = blockIdx.x * R;
    const int nr = min(R, p.B - b0);
    for (int i = threadIdx.x; i < 2 * R * H; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    for (int s = 0; s < p.nsteps; ++s) {
        const int t = p.reverse ? p.T - 1 - s : s;
        // gate pre-activations: gs[r][j] = xproj[b,t,j] + sum_k whh[j,k] * h[r][k]
        for (int j = threadIdx.x; j < G; j += blockDim.x) {
            float acc[R];
            const float bj = p.bhh ? p.bhh[j] : 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
                acc[r] = r < nr ? p.xproj[((long long)(b0 + r) * p.T + t) * p.ldx + j] + bj : 0.f;
            const float* wr = p.whh + (long long)j * H;
            for (int k = 0; k < H; k += 4) {
                const float4 w = nn::ld4(wr + k);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float4 h = *reinterpret_cast<const float4*>(hs + r * H + k);
                    acc[r] = fmaf(w.x, h.x, acc[r]); acc[r] = fmaf(w.y, h.y, acc[r]);
                    acc[r] = fmaf(w.z, h.z, acc[r]); acc[r] = fmaf(w.w, h.w, acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) gs[r * G + j] = acc[r];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nr * H; i += blockDim.x) {
            const int r = i / H, k = i - r * H;
            const float ig = sigmoidf_(gs[r * G + k]);
            const float fg = sigmoidf_(gs[r * G + H + k]);
            const float gg = tanhf(gs[r * G + 2 * H + k]);
            const float og = sigmoidf_(gs[r * G + 3 * H + k]);
            const float c = fg * cs[i] + ig * gg;
            const float h = og * tanhf(c);
            const long long row = (long long)(b0 + r) * p.T + t;
            if (p.gates) {
                float* g = p.gates + row * G;
                g[k] = ig; g[H + k] = fg; g[2 * H + k] = gg; g[3 * H + k] = og;
            }
            if (p.cst) p.cst[row * H + k] = c;
            if (p.hprev) p.hprev[row * H + k] = hs[i];
            p.out[row * p.ldo + k] = h;
            cs[i] = c;
            hs[i] = h;          // each (r,k) is owned by one thread; gs was fully consumed before the barrier below
        }
        __syncthreads();
    }
}

struct Bwd {
    const float* dout; long long ldo;     // external gradient of h: dout_step < 0: dout[(b*T+t)*ldo + k] for every t;
    int dout_step;                        // dout_step >= 0: only h at t == dout_step has one, dout[b*ldo + k]
    const float* gates; const float* cst; // saved by the forward
    const float* whh;
    float* dgates;                        // [B,T,4H] gradient of the gate pre-activations (rows of unvisited t untouched)
    int B, T, H, nsteps, reverse;
};

__global__ void __launch_bounds__(512)
lstm_bwd_kernel(const Bwd p) {
    extern __shared__ __align__(16) float sm[];
    const int H = p.H, G = 4 * p.H;
    float* dh = sm;                        // [R][H] gradient flowing into h_t from step t+1 (in walk order)
    float* dc = dh + R * H;                // [R][H]
    float* dg = dc + R * H;                // [R][G]
    float* part = dg + R * G;              // [R][G] partial sums of the W_hh^T product
    const int b0 = blockIdx.x * R;
    const int nr = min(R, p.B - b0);
    for (int i = threadIdx.x; i < 2 * R * H; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    for (int s = p.nsteps - 1; s >= 0; --s) {
        const int t = p.reverse ? p.T - 1 - s : s;
        const int tprev = p.reverse ? t + 1 : t - 1;          // the step walked just before t
        for (int i = threadIdx.x; i < nr * H; i += blockDim.x) {
            const int r = i / H, k = i - r * H;
            const long long row = (long long)(b0 + r) * p.T + t;
            const float* g = p.gates + row * G;
            const float ig = g[k], fg = g[H + k], gg = g[2 * H + k], og = g[3 * H + k];
            const float c = p.cst[row * H + k];
            const float cprev = s > 0 ? p.cst[((long long)(b0 + r) * p.T + tprev) * H + k] : 0.f;
            const float tc = tanhf(c);
            float dht = dh[i];
            if (p.dout) {
                if (p.dout_step < 0) dht += p.dout[row * p.ldo + k];
                else if (t == p.dout_step) dht += p.dout[(long long)(b0 + r) * p.ldo + k];
            }
            const float dct = dc[i] + dht * og * (1.f - tc * tc);
            const float d_i = dct * gg, d_f = dct * cprev, d_g = dct * ig, d_o = dht * tc;
            const float pi = d_i * ig * (1.f - ig), pf = d_f * fg * (1.f - fg);
            const float pg = d_g * (1.f - gg * gg), po = d_o * og * (1.f - og);
            float* o = p.dgates + row * G;
            o[k] = pi; o[H + k] = pf; o[2 * H + k] = pg; o[3 * H + k] = po;
            dg[r * G + k] = pi; dg[r * G + H + k] = pf; dg[r * G + 2 * H + k] = pg; dg[r * G + 3 * H + k] = po;
            dc[i] = dct * fg;
        }
        __syncthreads();
        if (s > 0) {
            // dh_prev[r][k] = sum_j dg[r][j] * whh[j,k]; thread (q, k) sums j in its quarter q
            for (int i = threadIdx.x; i < G; i += blockDim.x) {
                const int q = i / H, k = i - q * H;
                float acc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = 0.f;
                for (int j = q * H; j < (q + 1) * H; ++j) {
                    const float w = p.whh[(long long)j * H + k];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = fmaf(dg[r * G + j], w, acc[r]);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) part[r * G + i] = acc[r];
            }
            __syncthreads();
            for (int i = threadIdx.x; i < R * H; i += blockDim.x) {
                const int r = i / H, k = i - r * H;
                dh[i] = part[r * G + k] + part[r * G + H + k] + part[r * G + 2 * H + k] + part[r * G + 3 * H + k];
            }
            __syncthreads();
        }
    }
}

}  // namespace ls

extern "C" int lr_lstm_fwd(const float* xproj, long long ldx, const float* bhh, const float* whh, float* out, long long ldo,
                           float* gates, float* cst, float* hprev, int B, int T, int H, int nsteps, int reverse,
                           lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && T > 0 && H > 0 && (H & 3) == 0, "lr_lstm_fwd: bad shape (H %% 4 != 0?)");
    LR_CHECK_ARG(nsteps >= 0 && nsteps <= T, "lr_lstm_fwd: nsteps outside 0..T");
    if (B == 0 || nsteps == 0) return LR_OK;
    LR_CHECK_ARG(xproj && whh && out, "lr_lstm_fwd: null pointer");
    LR_CHECK_ALIGN(whh);
    ls::Fwd p;
    p.xproj = xproj; p.ldx = ldx; p.bhh = bhh; p.whh = whh; p.out = out; p.ldo = ldo; p.gates = gates; p.cst = cst; p.hprev = hprev;
    p.B = B; p.T = T; p.H = H; p.nsteps = nsteps; p.reverse = reverse;
    const size_t smem = (size_t)(2 * ls::R * H + ls::R * 4 * H) * sizeof(float);
    LR_CHECK_ARG(smem <= 200 * 1024, "lr_lstm_fwd: hidden size %d too large", H);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(ls::lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return lr::fail(LR_ECUDA, "lr_lstm_fwd smem: %s", cudaGetErrorString(e));
        configured = smem;
    }
    const int threads = 4 * H < 512 ? ((4 * H + 31) / 32) * 32 : 512;
    ls::lstm_fwd_kernel<<<(B + ls::R - 1) / ls::R, threads, smem, stream>>>(p);
    lr::count_launch();
    LR_CHECK_LAUNCH("lstm_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_lstm_bwd(const float* dout, long long ldo, int dout_step, const float* gates, const float* cst, const float* whh,
                           float* dgates, int B, int T, int H, int nsteps, int reverse, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && T > 0 && H > 0 && (H & 3) == 0, "lr_lstm_bwd: bad shape");
    LR_CHECK_ARG(nsteps >= 0 && nsteps <= T, "lr_lstm_bwd: nsteps outside 0..T");
    if (B == 0 || nsteps == 0) return LR_OK;
    LR_CHECK_ARG(gates && cst && whh && dgates, "lr_lstm_bwd: null pointer");
    ls::Bwd p;
    p.dout = dout; p.ldo = ldo; p.gates = gates; p.cst = cst; p.whh = whh; p.dgates = dgates;
    p.B = B; p.T = T; p.H = H; p.nsteps = nsteps; p.reverse = reverse; p.dout_step = dout_step;
    const size_t smem = (size_t)(2 * ls::R * H + 2 * ls::R * 4 * H) * sizeof(float);
    LR_CHECK_ARG(smem <= 200 * 1024, "lr_lstm_bwd: hidden size %d too large", H);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(ls::lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return lr::fail(LR_ECUDA, "lr_lstm_bwd smem: %s", cudaGetErrorString(e));
        configured = smem;
    }
    const int threads = 4 * H < 512 ? ((4 * H + 31) / 32) * 32 : 512;
    ls::lstm_bwd_kernel<<<(B + ls::R - 1) / ls::R, threads, smem, stream>>>(p);
    lr::count_launch();
    LR_CHECK_LAUNCH("lstm_bwd_kernel");
    return LR_OK;
}
