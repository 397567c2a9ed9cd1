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
#include <cooperative_groups.h>
#include <cstdlib>

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
            for (int k = 0; k < (s > 0 ? H : 0); k += 4) {    // h_{-1} = 0: the first step's product is exactly zero
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

// =====================================================================================================================
// Cluster version of the recurrence (thread-block clusters + distributed shared memory, sm_90+/sm_100a).
//
// The one-CTA-per-4-rows kernels above re-read the whole W_hh (4H x H floats: 256 KB at H = 128) from L2 on every step
// and are latency bound (12 us per step).  Here a cluster of NC = 8 CTAs owns BG = 32 batch rows; CTA c owns the
// hidden units [c*U, (c+1)*U), U = H / 8, i.e. the 4U gate rows {i,f,g,o} x U, whose W_hh slice (4U x H floats: 32 KB at
// H = 128, 128 KB at H = 256) stays in its shared memory for all T steps.  Per step a CTA computes the gates of its
// units for all 32 rows from h_{t-1} (a full [32][H] copy in its own shared memory), updates c / h for its units and
// pushes its slice of h_t into the h buffer of all 8 CTAs through distributed shared memory; one cluster barrier per
// step (the h buffer is double buffered).  The backward kernel does the transposed product the same way: every CTA
// forms the partial W_hh^T dgates sum over its gate rows for all H columns and scatters column slices to their
// owners (a reduce-scatter over DSMEM), two part buffers, one cluster barrier per step.
namespace lsc {

constexpr int NC = 8;          // CTAs per cluster (portable maximum)
// batch rows per cluster (template parameter BGT): 32, or 16 when that still fits one wave -- twice the clusters,
// half the rows (and gate activations) per thread and step, for the small batches where one cluster would walk alone
constexpr int TH = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(const void* smem_ptr, unsigned rank) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// thread t: unit u = t % U, row group rg = t / U (TH / U groups of RPT = BGT * U / TH rows)
template <int H, int BGT>
__global__ void __launch_bounds__(TH)
lstm_fwd_cluster_kernel(const ls::Fwd p) {
    constexpr int U = H / NC;                 // hidden units per CTA
    constexpr int NRG = TH / U;               // row groups
    constexpr int RPT = BGT / NRG;             // rows per thread
    constexpr int HP = H + 4;                 // h row pitch (floats): rows land on different banks
    static_assert(BGT % NRG == 0 && RPT >= 1, "bad thread mapping");
    extern __shared__ __align__(16) float sm[];
    float* Wt = sm;                           // [H][U][4]   Wt[k][u][g] = whh[(g*H + c*U + u)*H + k]
    float* hb = Wt + H * U * 4;               // [2][BGT][HP]
    const unsigned c = cluster_ctarank();
    const int b0 = (blockIdx.x / NC) * BGT;
    const int u = threadIdx.x % U, rg = threadIdx.x / U;
    for (int i = threadIdx.x; i < 4 * U * H; i += TH) {
        const int k = i % H, ju = i / H;                       // ju = g*U + uu: coalesced reads of the row
        const int g = ju / U, uu = ju - g * U;
        Wt[(k * U + uu) * 4 + g] = p.whh[(long long)(g * H + c * U + uu) * H + k];
    }
    for (int i = threadIdx.x; i < 2 * BGT * HP; i += TH) hb[i] = 0.f;
    float bias[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[g] = p.bhh ? p.bhh[g * H + c * U + u] : 0.f;
    float cst[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) cst[r] = 0.f;
    uint32_t remote[NC];
#pragma unroll
    for (int d = 0; d < NC; ++d) remote[d] = map_to_rank(hb, d);
    cluster_sync();                           // every CTA's buffers are initialised before anyone writes into them
    for (int s = 0; s < p.nsteps; ++s) {
        const int t = p.reverse ? p.T - 1 - s : s;
        const float* hcur = hb + (s & 1) * BGT * HP;
        float acc[RPT][4], xp[RPT][4];          // xp: issued now, consumed after the recurrent product (latency hidden)
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int b = b0 + rg * RPT + r;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                xp[r][g] = b < p.B ? p.xproj[((long long)b * p.T + t) * p.ldx + g * H + c * U + u] + bias[g] : 0.f;
                acc[r][g] = 0.f;
            }
        }
        if (s > 0) {                          // h_0 = 0: the first step has no recurrent term
#pragma unroll 2
            for (int k = 0; k < H; k += 4) {
                float4 w[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(Wt + ((k + kk) * U + u) * 4);
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const float4 h = *reinterpret_cast<const float4*>(hcur + (rg * RPT + r) * HP + k);
                    const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        acc[r][0] = fmaf(w[kk].x, hv[kk], acc[r][0]); acc[r][1] = fmaf(w[kk].y, hv[kk], acc[r][1]);
                        acc[r][2] = fmaf(w[kk].z, hv[kk], acc[r][2]); acc[r][3] = fmaf(w[kk].w, hv[kk], acc[r][3]);
                    }
                }
            }
        }
        const int nxt = ((s + 1) & 1) * BGT * HP;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int rl = rg * RPT + r, b = b0 + rl;
            const float ig = sigmoidf_(acc[r][0] + xp[r][0]), fg = sigmoidf_(acc[r][1] + xp[r][1]);
            const float gg = tanhf(acc[r][2] + xp[r][2]), og = sigmoidf_(acc[r][3] + xp[r][3]);
            const float cn = fg * cst[r] + ig * gg;
            const float hn = og * tanhf(cn);
            const int col = c * U + u;
            if (b < p.B) {
                const long long row = (long long)b * p.T + t;
                if (p.gates) { float* gp = p.gates + row * 4 * H; gp[col] = ig; gp[H + col] = fg; gp[2 * H + col] = gg; gp[3 * H + col] = og; }
                if (p.cst) p.cst[row * H + col] = cn;
                if (p.hprev) p.hprev[row * H + col] = hcur[rl * HP + col];
                p.out[row * p.ldo + col] = hn;
            }
            cst[r] = cn;
            const uint32_t off = (uint32_t)(nxt + rl * HP + col) * 4u;
#pragma unroll
            for (int d = 0; d < NC; ++d) st_cluster(remote[d] + off, hn);
        }
        cluster_sync();
    }
}

// BPTT with the same partition.  Per step: local dgates for the CTA's units, partial dh_prev[r][k] = sum over the
// CTA's 4U gate rows of dg[r][j] * whh[j][k] for ALL k, scattered to the owner of column k.
template <int H, int BGT>
__global__ void __launch_bounds__(TH)
lstm_bwd_cluster_kernel(const ls::Bwd p) {
    constexpr int U = H / NC;
    constexpr int NRG = TH / U;
    constexpr int RPT = BGT / NRG;
    constexpr int GP = 4 * U + 4;             // dg row pitch
    constexpr int KPT = H / 32;               // columns per thread in the transposed product (32 column lanes)
    constexpr int NRG2 = TH / 32;             // row groups of the transposed product
    constexpr int RPT2 = BGT / NRG2;
    extern __shared__ __align__(16) float sm[];
    float* Ws = sm;                           // [4U][H]    Ws[g*U+uu][k] = whh[(g*H + c*U + uu)*H + k]
    float* dg = Ws + 4 * U * H;               // [BGT][GP]   this step's gate gradients of the CTA's units
    float* part = dg + BGT * GP;               // [2][NC][BGT][U]  partial dh of MY units from every CTA
    const unsigned c = cluster_ctarank();
    const int b0 = (blockIdx.x / NC) * BGT;
    const int u = threadIdx.x % U, rg = threadIdx.x / U;
    for (int i = threadIdx.x; i < 4 * U * H; i += TH) {
        const int k = i % H, ju = i / H;
        const int g = ju / U, uu = ju - g * U;
        Ws[ju * H + k] = p.whh[(long long)(g * H + c * U + uu) * H + k];
    }
    for (int i = threadIdx.x; i < 2 * NC * BGT * U; i += TH) part[i] = 0.f;
    float dc[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) dc[r] = 0.f;
    uint32_t remote[NC];
#pragma unroll
    for (int d = 0; d < NC; ++d) remote[d] = map_to_rank(part, d);
    const int kl = threadIdx.x % 32, rg2 = threadIdx.x / 32;
    cluster_sync();
    for (int s = p.nsteps - 1; s >= 0; --s) {
        const int t = p.reverse ? p.T - 1 - s : s;
        const int tprev = p.reverse ? t + 1 : t - 1;
        const int par = s & 1;
        // ---- gate gradients of my units: dh = sum of the 8 partials scattered to me in the previous iteration
        const float* pin = part + par * NC * BGT * U;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int rl = rg * RPT + r, b = b0 + rl;
            const int col = c * U + u;
            float pi = 0.f, pf = 0.f, pg = 0.f, po = 0.f;
            if (b < p.B) {
                float dht = 0.f;
                if (s < p.nsteps - 1) {
#pragma unroll
                    for (int d = 0; d < NC; ++d) dht += pin[(d * BGT + rl) * U + u];
                }
                const long long row = (long long)b * p.T + t;
                if (p.dout) {
                    if (p.dout_step < 0) dht += p.dout[row * p.ldo + col];
                    else if (t == p.dout_step) dht += p.dout[(long long)b * p.ldo + col];
                }
                const float* gq = p.gates + row * 4 * H;
                const float ig = gq[col], fg = gq[H + col], gg = gq[2 * H + col], og = gq[3 * H + col];
                const float cc = p.cst[row * H + col];
                const float cprev = s > 0 ? p.cst[((long long)b * p.T + tprev) * H + col] : 0.f;
                const float tc = tanhf(cc);
                const float dct = dc[r] + dht * og * (1.f - tc * tc);
                pi = dct * gg * ig * (1.f - ig); pf = dct * cprev * fg * (1.f - fg);
                pg = dct * ig * (1.f - gg * gg); po = dht * tc * og * (1.f - og);
                float* o = p.dgates + row * 4 * H;
                o[col] = pi; o[H + col] = pf; o[2 * H + col] = pg; o[3 * H + col] = po;
                dc[r] = dct * fg;
            }
            float* d = dg + rl * GP;
            d[u] = pi; d[U + u] = pf; d[2 * U + u] = pg; d[3 * U + u] = po;
        }
        __syncthreads();
        if (s > 0) {
            // ---- partial[r][k] = sum_j dg[r][j] * Ws[j][k]; thread: columns kl + 32*i, rows rg2*RPT2 ..
            float acc[RPT2][KPT];
#pragma unroll
            for (int r = 0; r < RPT2; ++r)
#pragma unroll
                for (int i = 0; i < KPT; ++i) acc[r][i] = 0.f;
            for (int j = 0; j < 4 * U; j += 4) {
                float w[4][KPT];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int i = 0; i < KPT; ++i) w[jj][i] = Ws[(j + jj) * H + kl + 32 * i];
#pragma unroll
                for (int r = 0; r < RPT2; ++r) {
                    const float4 g4 = *reinterpret_cast<const float4*>(dg + (rg2 * RPT2 + r) * GP + j);
                    const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                        for (int i = 0; i < KPT; ++i) acc[r][i] = fmaf(gv[jj], w[jj][i], acc[r][i]);
                }
            }
            // scatter: column k belongs to CTA k / U; it lands in that CTA's part[(s-1)&1][src = c][r][k % U]
            const int npar = (s - 1) & 1;
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                const int k = kl + 32 * i;
                const int dst = k / U, ku = k - dst * U;
#pragma unroll
                for (int r = 0; r < RPT2; ++r) {
                    const int rl = rg2 * RPT2 + r;
                    st_cluster(remote[dst] + (uint32_t)(((npar * NC + (int)c) * BGT + rl) * U + ku) * 4u, acc[r][i]);
                }
            }
        }
        cluster_sync();
    }
}

template <typename Kern>
static int launch_cluster(Kern kern, int clusters, size_t smem, lr_stream_t stream, const void* arg, const char* name,
                          void (*launcher)(Kern, cudaLaunchConfig_t*, const void*)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "%s smem: %s", name, cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * NC));
    cfg.blockDim = dim3(TH);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    launcher(kern, &cfg, arg);
    return LR_OK;
}

template <int H, int BGT> static size_t fwd_smem() { return (size_t)(H * (H / NC) * 4 + 2 * BGT * (H + 4)) * sizeof(float); }
template <int H, int BGT> static size_t bwd_smem() {
    return (size_t)(4 * (H / NC) * H + BGT * (4 * (H / NC) + 4) + 2 * NC * BGT * (H / NC)) * sizeof(float);
}

template <int H, int BGT>
static int run_fwd(const ls::Fwd& p, lr_stream_t stream) {
    auto k = lstm_fwd_cluster_kernel<H, BGT>;
    return launch_cluster(k, (p.B + BGT - 1) / BGT, fwd_smem<H, BGT>(), stream, &p, "lstm_fwd_cluster_kernel",
                          +[](decltype(k) kk, cudaLaunchConfig_t* cfg, const void* a) {
                              cudaLaunchKernelEx(cfg, kk, *static_cast<const ls::Fwd*>(a));
                          });
}
template <int H, int BGT>
static int run_bwd(const ls::Bwd& p, lr_stream_t stream) {
    auto k = lstm_bwd_cluster_kernel<H, BGT>;
    return launch_cluster(k, (p.B + BGT - 1) / BGT, bwd_smem<H, BGT>(), stream, &p, "lstm_bwd_cluster_kernel",
                          +[](decltype(k) kk, cudaLaunchConfig_t* cfg, const void* a) {
                              cudaLaunchKernelEx(cfg, kk, *static_cast<const ls::Bwd*>(a));
                          });
}


// 16 rows per cluster while all clusters still run in one wave, else 32
static bool small_groups(int B) { return ((B + 15) / 16) * NC <= lr::sm_count() && getenv("LIPREAD_LSTM_BG32") == nullptr; }

template <int H> static int run_fwd_auto(const ls::Fwd& p, lr_stream_t stream) {
    return small_groups(p.B) ? run_fwd<H, 16>(p, stream) : run_fwd<H, 32>(p, stream);
}
template <int H> static int run_bwd_auto(const ls::Bwd& p, lr_stream_t stream) {
    return small_groups(p.B) ? run_bwd<H, 16>(p, stream) : run_bwd<H, 32>(p, stream);
}

}  // namespace lsc

// =====================================================================================================================
// Grid-cooperative version for hidden sizes whose W_hh (4H x H floats: 4 MB at H = 512) does not fit the shared
// memory of one 8-CTA cluster: NC = H / 8 CTAs per group of 32 batch rows (64 CTAs at H = 512: twice the SMs and half
// the per-step FMA chain of the first version's H / 16), CTA c owns 8 hidden units and keeps its W_hh slice (32 gate
// rows x H: 64 KB at H = 512) in shared memory for the whole walk.  h_t is exchanged through the
// `out` tensor itself (it is the next step's input and lives in L2), dgates through the `dgates` tensor; one
// cooperative grid barrier per step.  Launched with cudaLaunchAttributeCooperative (all CTAs co-resident).
namespace lsg {

constexpr int U = 8, BG = 32, TH = 256, NRG = TH / U, RPT = BG / NRG;    // 8 units x 32 row groups of 1 row: H/8 CTAs per group

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(TH)
lstm_fwd_coop_kernel(const ls::Fwd p) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    extern __shared__ __align__(16) float sm[];
    const int H = p.H, HP = H + 4, NC = H / U;
    float* Wt = sm;                                   // [H][U][4]
    float* hb = Wt + (size_t)H * U * 4;               // [BG][HP]  h_{t-1}
    const int c = blockIdx.x % NC, b0 = (blockIdx.x / NC) * BG;
    const int u = threadIdx.x % U, rg = threadIdx.x / U;
    for (int i = threadIdx.x; i < 4 * U * H; i += TH) {
        const int k = i % H, ju = i / H;
        const int g = ju / U, uu = ju - g * U;
        Wt[(k * U + uu) * 4 + g] = p.whh[(long long)(g * H + c * U + uu) * H + k];
    }
    float bias[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[g] = p.bhh ? p.bhh[g * H + c * U + u] : 0.f;
    float cst[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) cst[r] = 0.f;
    const int col = c * U + u;
    const bool vec_out = (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    for (int s = 0; s < p.nsteps; ++s) {
        const int t = p.reverse ? p.T - 1 - s : s;
        const int tprev = p.reverse ? t + 1 : t - 1;
        // h_{t-1}: the whole [BG][H] block from `out` (written by all CTAs of the group in the previous step)
        // (8 independent L2 loads in flight per thread: one at a time, this refill was most of the step's latency)
        for (int i0 = threadIdx.x; i0 < BG * (H / 4); i0 += 8 * TH) {
            float4 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = i0 + q * TH;
                v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < BG * (H / 4)) {
                    const int r = i / (H / 4), k4 = (i - r * (H / 4)) * 4;
                    const int b = b0 + r;
                    if (s > 0 && b < p.B) {
                        const float* src = p.out + ((long long)b * p.T + tprev) * p.ldo + k4;
                        v[q] = vec_out ? __ldcg(reinterpret_cast<const float4*>(src))
                                       : make_float4(__ldcg(src), __ldcg(src + 1), __ldcg(src + 2), __ldcg(src + 3));
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = i0 + q * TH;
                if (i < BG * (H / 4)) {
                    const int r = i / (H / 4), k4 = (i - r * (H / 4)) * 4;
                    *reinterpret_cast<float4*>(hb + r * HP + k4) = v[q];
                }
            }
        }
        float acc[RPT][4], xp[RPT][4];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int b = b0 + rg * RPT + r;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                xp[r][g] = b < p.B ? p.xproj[((long long)b * p.T + t) * p.ldx + g * H + col] + bias[g] : 0.f;
                acc[r][g] = 0.f;
            }
        }
        __syncthreads();
        if (s > 0) {
#pragma unroll 2
            for (int k = 0; k < H; k += 4) {
                float4 w[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(Wt + ((k + kk) * U + u) * 4);
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const float4 h = *reinterpret_cast<const float4*>(hb + (rg * RPT + r) * HP + k);
                    const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        acc[r][0] = fmaf(w[kk].x, hv[kk], acc[r][0]); acc[r][1] = fmaf(w[kk].y, hv[kk], acc[r][1]);
                        acc[r][2] = fmaf(w[kk].z, hv[kk], acc[r][2]); acc[r][3] = fmaf(w[kk].w, hv[kk], acc[r][3]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int rl = rg * RPT + r, b = b0 + rl;
            const float ig = sigmoidf_(acc[r][0] + xp[r][0]), fg = sigmoidf_(acc[r][1] + xp[r][1]);
            const float gg = tanhf(acc[r][2] + xp[r][2]), og = sigmoidf_(acc[r][3] + xp[r][3]);
            const float cn = fg * cst[r] + ig * gg;
            const float hn = og * tanhf(cn);
            if (b < p.B) {
                const long long row = (long long)b * p.T + t;
                if (p.gates) { float* gp = p.gates + row * 4 * H; gp[col] = ig; gp[H + col] = fg; gp[2 * H + col] = gg; gp[3 * H + col] = og; }
                if (p.cst) p.cst[row * H + col] = cn;
                if (p.hprev) p.hprev[row * H + col] = hb[rl * HP + col];
                p.out[row * p.ldo + col] = hn;
            }
            cst[r] = cn;
        }
        __threadfence();
        grid.sync();                                  // also protects hb against the next step's refill
    }
}

__global__ void __launch_bounds__(TH)
lstm_bwd_coop_kernel(const ls::Bwd p) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    extern __shared__ __align__(16) float sm[];
    const int H = p.H, G = 4 * H, NC = H / U;
    constexpr int JC = 256;                           // dgates columns staged per chunk
    constexpr int JP = JC + 4;
    float* Wc = sm;                                   // [4H][U]   Wc[j][u'] = whh[j*H + c*U + u']
    float* dgs = Wc + (size_t)G * U;                  // [BG][JP]  a chunk of this step's dgates (all units)
    float* dhs = dgs + BG * JP;                       // [BG][U]   dh of my units for the next visited step
    const int c = blockIdx.x % NC, b0 = (blockIdx.x / NC) * BG;
    const int u = threadIdx.x % U, rg = threadIdx.x / U;
    for (int i = threadIdx.x; i < G * U; i += TH) {
        const int uu = i % U, j = i / U;
        Wc[i] = p.whh[(long long)j * H + c * U + uu];
    }
    for (int i = threadIdx.x; i < BG * U; i += TH) dhs[i] = 0.f;
    float dc[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) dc[r] = 0.f;
    const int col = c * U + u;
    __syncthreads();
    for (int s = p.nsteps - 1; s >= 0; --s) {
        const int t = p.reverse ? p.T - 1 - s : s;
        const int tprev = p.reverse ? t + 1 : t - 1;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int rl = rg * RPT + r, b = b0 + rl;
            if (b < p.B) {
                float dht = dhs[rl * U + u];
                const long long row = (long long)b * p.T + t;
                if (p.dout) {
                    if (p.dout_step < 0) dht += p.dout[row * p.ldo + col];
                    else if (t == p.dout_step) dht += p.dout[(long long)b * p.ldo + col];
                }
                const float* gq = p.gates + row * G;
                const float ig = gq[col], fg = gq[H + col], gg = gq[2 * H + col], og = gq[3 * H + col];
                const float cc = p.cst[row * H + col];
                const float cprev = s > 0 ? p.cst[((long long)b * p.T + tprev) * H + col] : 0.f;
                const float tc = tanhf(cc);
                const float dct = dc[r] + dht * og * (1.f - tc * tc);
                float* o = p.dgates + row * G;
                o[col] = dct * gg * ig * (1.f - ig);
                o[H + col] = dct * cprev * fg * (1.f - fg);
                o[2 * H + col] = dct * ig * (1.f - gg * gg);
                o[3 * H + col] = dht * tc * og * (1.f - og);
                dc[r] = dct * fg;
            }
        }
        __threadfence();
        grid.sync();
        if (s > 0) {
            // dh_prev[r][u] = sum_j dgates[r, t, j] * whh[j][col] over ALL 4H gate rows, staged JC columns at a time
            float acc[RPT];
#pragma unroll
            for (int r = 0; r < RPT; ++r) acc[r] = 0.f;
            // chunks of JC columns are double buffered through registers: the next chunk's L2 loads are in flight
            // while the current one is multiplied
            constexpr int PF = BG * (JC / 4) / TH;                        // float4 per thread and chunk (8)
            float4 pre[PF];
            auto fetch = [&](int j0) {
#pragma unroll
                for (int q = 0; q < PF; ++q) {
                    const int i = threadIdx.x + q * TH;
                    const int r = i / (JC / 4), j4 = (i - r * (JC / 4)) * 4;
                    const int b = b0 + r;
                    pre[q] = b < p.B ? __ldcg(reinterpret_cast<const float4*>(p.dgates + ((long long)b * p.T + t) * G + j0 + j4))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            fetch(0);
            for (int j0 = 0; j0 < G; j0 += JC) {
                __syncthreads();
#pragma unroll
                for (int q = 0; q < PF; ++q) {
                    const int i = threadIdx.x + q * TH;
                    const int r = i / (JC / 4), j4 = (i - r * (JC / 4)) * 4;
                    *reinterpret_cast<float4*>(dgs + r * JP + j4) = pre[q];
                }
                __syncthreads();
                if (j0 + JC < G) fetch(j0 + JC);
#pragma unroll 4
                for (int j = 0; j < JC; j += 4) {
                    float w[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) w[jj] = Wc[(size_t)(j0 + j + jj) * U + u];
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        const float4 g4 = *reinterpret_cast<const float4*>(dgs + (rg * RPT + r) * JP + j);
                        acc[r] = fmaf(g4.x, w[0], acc[r]); acc[r] = fmaf(g4.y, w[1], acc[r]);
                        acc[r] = fmaf(g4.z, w[2], acc[r]); acc[r] = fmaf(g4.w, w[3], acc[r]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RPT; ++r) dhs[(rg * RPT + r) * U + u] = acc[r];      // read back by the same thread
        }
    }
}

static size_t fwd_smem(int H) { return ((size_t)H * U * 4 + (size_t)BG * (H + 4)) * sizeof(float); }
static size_t bwd_smem(int H) { return ((size_t)4 * H * U + (size_t)BG * (256 + 4) + (size_t)BG * U) * sizeof(float); }
static bool usable(int H, int B) {
    // shared memory fits and every CTA of the cooperative grid can be resident at once (one CTA per SM)
    return H % U == 0 && H >= 64 && fwd_smem(H) <= 220 * 1024 && bwd_smem(H) <= 220 * 1024 &&
           (long long)(H / U) * ((B + BG - 1) / BG) <= lr::sm_count();
}

template <typename Kern, typename Arg>
static int launch(Kern kern, int ctas, size_t smem, lr_stream_t stream, const Arg& arg, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "%s smem: %s", name, cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(TH);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, arg);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "%s launch: %s", name, cudaGetErrorString(e));
    return LR_OK;
}

}  // namespace lsg

// LIPREAD_LSTM=simple selects the one-CTA-per-4-rows kernels (debugging / A-B timing); read once.
static bool lr_lstm_use_cluster() {
    static const bool v = [] { const char* e = getenv("LIPREAD_LSTM"); return !(e && e[0] == 's'); }();
    return v;
}

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
    // more than one step of a long walk: the cluster kernel keeps W_hh in shared memory (H = 128 / 256)
    if (nsteps > 1 && (H == 128 || H == 256) && lr_lstm_use_cluster()) {
        const int rc = H == 128 ? lsc::run_fwd_auto<128>(p, stream) : lsc::run_fwd_auto<256>(p, stream);
        if (rc) return rc;
        lr::count_launch();
        LR_CHECK_LAUNCH("lstm_fwd_cluster_kernel");
        return LR_OK;
    }
    if (nsteps > 1 && H > 256 && lsg::usable(H, B) && lr_lstm_use_cluster()) {
        const int rc = lsg::launch(lsg::lstm_fwd_coop_kernel, (H / lsg::U) * ((B + lsg::BG - 1) / lsg::BG), lsg::fwd_smem(H),
                                   stream, p, "lstm_fwd_coop_kernel");
        if (rc) return rc;
        lr::count_launch();
        LR_CHECK_LAUNCH("lstm_fwd_coop_kernel");
        return LR_OK;
    }
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
    if (nsteps > 1 && (H == 128 || H == 256) && lr_lstm_use_cluster()) {
        const int rc = H == 128 ? lsc::run_bwd_auto<128>(p, stream) : lsc::run_bwd_auto<256>(p, stream);
        if (rc) return rc;
        lr::count_launch();
        LR_CHECK_LAUNCH("lstm_bwd_cluster_kernel");
        return LR_OK;
    }
    if (nsteps > 1 && H > 256 && lsg::usable(H, B) && lr_lstm_use_cluster()) {
        const int rc = lsg::launch(lsg::lstm_bwd_coop_kernel, (H / lsg::U) * ((B + lsg::BG - 1) / lsg::BG), lsg::bwd_smem(H),
                                   stream, p, "lstm_bwd_coop_kernel");
        if (rc) return rc;
        lr::count_launch();
        LR_CHECK_LAUNCH("lstm_bwd_coop_kernel");
        return LR_OK;
    }
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

