// nn.LSTM recurrence on the tensor cores (precision "bf16"): the same walk as lstm.cu (reference call sites
// audio_video/models/middle_fusion_fast.py:18,35-36, audio_video/models/early_fusion.py:62-69,82-83,
// video/models/resnet_lstm.py:113-120), with W_hh and h_{t-1} rounded to bfloat16 and the products h_{t-1} W_hh^T on
// tcgen05 (fp32 accumulation in TMEM).  Everything else -- the input projection, the gate activations, c_t, the saved
// gates / cell states for BPTT -- stays fp32.
//
// Partition (H = 128 / 256 / 512).  The 4H gate rows are dealt to NC = 1 / 4 / 16 CTAs of one thread-block cluster per
// group of 32 batch rows: a CTA owns UC = 128 / 64 / 32 hidden units, i.e. MT = 4 / 2 / 1 tiles of 128 gate rows
// (row g*32 + u of a tile = gate g in {i, f, g, o} of the tile's unit u), whose W_hh slice -- always 512 rows x H bf16
// ... always 128 KB -- is written ONCE into shared memory in the UMMA K-major / 128-byte-swizzle layout and stays
// there for the whole walk (the fp32 cluster kernels hold 64 KB slices; the cooperative H = 512 kernel re-read h
// through L2 and paid a grid barrier per step: 17 us per step).  Per step:
//     D_m[128 gate rows x 32 batch] = W_m[128 x H] . h_{t-1}^T[H x 32]      MT x H/16 tcgen05.mma (M 128, N 32, K 16),
//                                                                          one elected thread, commit per tile
//     16 warps: tcgen05.ld (warp w: rows 32 (w & 3).., batch columns 8 (w >> 2)..) -> a [128][33] shared tile ->
//     thread (unit, 2 batch rows) forms the four gates (+ input projection), c_t, h_t, stores out / gates / c / h_prev,
//     and h_t as bf16 goes -- in 16-byte chunks at their swizzled place -- into the NEXT step's h operand buffer of
//     every CTA of the cluster (st.shared::cluster); one cluster barrier per step (h is double buffered).
// BPTT (lstm_bwd_tc_kernel) keeps the same ownership: a CTA forms dgates for its units (fp32, stored), rounds them to
// bf16 as the B operand, multiplies by the TRANSPOSED slice W^T[H x its gate rows] (again 128 KB resident) to get its
// partial of dh_{t-1} for ALL H units, and scatters 32-unit slices to their owners through distributed shared memory,
// where they are added in fixed source order (deterministic).
#include <cuda_bf16.h>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace lt {

using namespace tcc;

constexpr int CW = 16;                       // compute warps
constexpr int TH = (CW + 1) * 32;            // + the MMA warp
constexpr int BG = 32;                       // batch rows per cluster = the MMA's N
constexpr int GP = 33;                       // pitch of the gate tile (floats)

struct Fwd {
    const float* xproj; long long ldx;
    const float* bhh;
    const float* whh;
    float* out; long long ldo;
    float* gates; float* cst; float* hprev;
    int B, T, H, nsteps, reverse;
    int dbg;                             // timing probes only (LIPREAD_LSTM_DBG bit mask; results are garbage when set)
};
struct Bwd {
    const float* dout; long long ldo;
    int dout_step;
    const float* gates; const float* cst;
    const float* whh;
    float* dgates;
    int B, T, H, nsteps, reverse;
};

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// fast, accurate-enough activations for the bf16 mode (relative error ~1e-6: far below the bf16 operand rounding)
__device__ __forceinline__ float sigm(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) { return 2.f * sigm(2.f * x) - 1.f; }

__device__ __forceinline__ uint4 pack8(const float4& a, const float4& b) {
    const uint2 lo = nn::pack4(a), hi = nn::pack4(b);
    return make_uint4(lo.x, lo.y, hi.x, hi.y);
}

template <int H> struct Geo {
    static constexpr int MT = H >= 256 ? 512 / H : 1;  // 128-row gate tiles per CTA (H = 128: one tile, 4 CTAs -- a single CTA
                                                      // with four tiles has 8 cells per thread and step and was slower)
    static constexpr int UC = 32 * MT;               // hidden units per CTA
    static constexpr int NC = H / UC;                // CTAs per cluster: 4, 4, 16
    static constexpr int KB = H / 64;                // 64-element k-blocks of the forward product
    static constexpr uint32_t A_BYTES = (uint32_t)MT * 128u * (uint32_t)H * 2u;    // the resident W_hh slice: 32 / 128 / 128 KB
    static constexpr uint32_t HB_BYTES = KB * 4096u;                 // one h operand buffer: [32 rows][H] bf16, swizzled
    static constexpr uint32_t G_BYTES = 128u * GP * 4u;              // 16 896
    static constexpr int HSP = UC + 8;                               // staging pitch (bf16 elements): 16-byte aligned rows
    static constexpr uint32_t HS_BYTES = 32u * HSP * 2u;
    static constexpr size_t FWD_SMEM = 1024 + A_BYTES + 2 * HB_BYTES + MT * G_BYTES + HS_BYTES;   // one gate tile per m
    // backward: B operand = dgates of the CTA's 128 MT gate rows [32 rows][128 MT] bf16 (2 MT k-blocks), receive buffer
    // for the dh partials [NC sources][32 batch][UC] fp32
    static constexpr int KBB = 2 * MT;
    static constexpr uint32_t DG_BYTES = KBB * 4096u;
    static constexpr uint32_t R_BYTES = (uint32_t)NC * 32u * UC * 4u;
    static constexpr size_t BWD_SMEM = 1024 + A_BYTES + DG_BYTES + R_BYTES;
};

// ---------------------------------------------------------------------------------------------------- forward
template <int H>
__global__ void __launch_bounds__(TH, 1)
lstm_fwd_tc_kernel(const Fwd p) {
    using G_ = Geo<H>;
    constexpr int MT = G_::MT, UC = G_::UC, NC = G_::NC, KB = G_::KB, HSP = G_::HSP;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned c = NC > 1 ? cluster_ctarank() : 0u;
    const int b0 = (blockIdx.x / NC) * BG;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sH0 = sA + G_::A_BYTES;
    uint8_t* const gA = gbase;
    float* const Gt = reinterpret_cast<float*>(gbase + G_::A_BYTES + 2 * G_::HB_BYTES);          // [MT][128][GP]
    __nv_bfloat16* const Hs = reinterpret_cast<__nv_bfloat16*>(gbase + G_::A_BYTES + 2 * G_::HB_BYTES + MT * G_::G_BYTES);

    // W_hh slice -> the A operand: tile m, k-block kb: [128 rows][128 B], 16-byte chunk j of row r at j ^ (r & 7)
    // (four chunks = eight 16-byte loads in flight per thread: the 256 KB fp32 slice is a latency chain otherwise)
    for (int idx0 = tid; idx0 < MT * 128 * (H / 8); idx0 += 4 * TH) {
        float4 lo[4], hi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * TH;
            if (idx < MT * 128 * (H / 8)) {
                const int r = idx / (H / 8), jc = idx - r * (H / 8);
                const int m = r >> 7, rr = r & 127, g = rr >> 5, ul = rr & 31;
                const float4* src = reinterpret_cast<const float4*>(p.whh + (long long)(g * H + (int)c * UC + 32 * m + ul) * H + 8 * jc);
                lo[u] = __ldg(src); hi[u] = __ldg(src + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * TH;
            if (idx < MT * 128 * (H / 8)) {
                const int r = idx / (H / 8), jc = idx - r * (H / 8);
                const int m = r >> 7, rr = r & 127;
                const int kb = jc >> 3, j = jc & 7;
                *reinterpret_cast<uint4*>(gA + (size_t)(m * KB + kb) * 16384 + rr * 128 + ((j ^ (rr & 7)) << 4)) = pack8(lo[u], hi[u]);
            }
        }
    }
    if (tid == 0) {
        for (int m = 0; m < MT; ++m) mbar_init(smem_u32(&bars[m]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // accumulators per tile.  Measured (profiles/r2_lstm_microbench.txt): the 32 MMAs of a step take 1.24 us; spreading
    // them over 4 independent accumulators changes nothing (137 -> 138 us at H = 512), i.e. the chain is bound by the
    // operand fetch of a 128 x 32 x 16 MMA (4 KB of W per instruction), not by the accumulation dependency: one it is
    constexpr int NA = 1;
    constexpr uint32_t TCOLS = MT * NA * 32 < 32 ? 32 : MT * NA * 32;
    if (warp == CW) tmem_alloc(smem_u32(&tmem_slot), TCOLS);
    fence_proxy_async();                                  // the A operand was written through the generic proxy
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_slot;
    if (NC > 1) cluster_sync();                           // every CTA of the cluster is resident before remote stores

    // compute-thread roles
    const int q = warp & 3, cq = warp >> 2;               // TMEM read: rows 32 q.., batch columns 8 cq..
    const int ul = lane, bl0 = 2 * warp;                   // cell update: unit ul of each tile, batch rows bl0, bl0 + 1
    float cstate[MT][2], hlast[MT][2], bias[MT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        cstate[m][0] = cstate[m][1] = 0.f; hlast[m][0] = hlast[m][1] = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) bias[m][g] = (warp < CW && p.bhh) ? p.bhh[g * H + (int)c * UC + 32 * m + ul] : 0.f;
    }

    // the input projection of a step is fetched one step ahead (it does not depend on the recurrence): the loads of step
    // s + 1 are issued after the cells of step s and land during the exchange, the barrier and the MMAs
    float xp[MT][4][2];
    auto fetch_xp = [&](int step) {
        const int tt = p.reverse ? p.T - 1 - step : step;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const int col = (int)c * UC + 32 * m + ul;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int b = b0 + bl0 + i;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    xp[m][g][i] = b < p.B ? __ldg(p.xproj + ((long long)b * p.T + tt) * p.ldx + g * H + col) : 0.f;
            }
        }
    };
    if (warp < CW) fetch_xp(0);

    for (int s = 0; s < p.nsteps; ++s) {
        const int t = p.reverse ? p.T - 1 - s : s;
        const uint32_t sHprev = sH0 + (uint32_t)((s & 1) ^ 1) * G_::HB_BYTES;     // h_{s-1}
        const uint32_t sHnext = sH0 + (uint32_t)(s & 1) * G_::HB_BYTES;           // h_s
        if (s > 0) {
            // h_{s-1} is complete in every CTA: the barrier was arrived at right after each thread's part of the exchange
            // (end of the previous step), so the output stores of that step and this step's fetches overlap its latency
            if (NC > 1 && !(p.dbg & 8)) cluster_wait(); else __syncthreads();
            if (warp == CW && lane == 0) {
                fence_proxy_async_all();
                fence_after();
                const uint32_t idesc = idesc_bf16(BG, false, false);
#pragma unroll 1
                for (int m = 0; m < MT; ++m) {
#pragma unroll 1
                    for (int kb = 0; kb < ((p.dbg & 2) ? 0 : KB); ++kb) {
                        const uint64_t ad = desc_k_sw128(sA + (uint32_t)(m * KB + kb) * 16384u);
                        const uint64_t bd = desc_k_sw128(sHprev + (uint32_t)kb * 4096u);
                        // NA independent accumulators per tile (k-block kb -> accumulator kb % NA), summed by the epilogue:
                        // a 128 x 32 x 16 MMA is so short that a single accumulation chain is latency bound
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_bf16(tmem + (uint32_t)(32 * (m * NA + (kb % NA))), ad + 2 * k, bd + 2 * k, idesc, (kb >= NA || k) ? 1u : 0u);
                    }
                    mma_commit(smem_u32(&bars[m]));        // tile m's cells run while the MMAs of tile m + 1 .. are in flight
                }
            }
        }
        const bool more = s + 1 < p.nsteps;
        float sv[MT][2][6];                                // i, f, g, o, c, h of this step's cells: stored after the exchange
        if (warp < CW) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int col = (int)c * UC + 32 * m + ul;
                float* const Gm = Gt + m * 128 * GP;
                if (s > 0) {
                    mbar_wait(smem_u32(&bars[m]), (uint32_t)(s - 1) & 1u);
                    fence_after();
                    float v[8];
                    tmem_ld8(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * m * NA + 8 * cq), v);
#pragma unroll
                    for (int a = 1; a < NA; ++a) {
                        float w[8];
                        tmem_ld8(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * (m * NA + a) + 8 * cq), w);
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] += w[i];
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) Gm[(32 * q + lane) * GP + 8 * cq + i] = v[i];
                    compute_sync();
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int bl = bl0 + i;
                    float a[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) a[g] = (s > 0 ? Gm[(32 * g + ul) * GP + bl] : 0.f) + xp[m][g][i] + bias[m][g];
                    const float ig = sigm(a[0]), fg = sigm(a[1]), gg = tanh_(a[2]), og = sigm(a[3]);
                    const float cn = fg * cstate[m][i] + ig * gg;
                    const float hn = og * tanh_(cn);
                    sv[m][i][0] = ig; sv[m][i][1] = fg; sv[m][i][2] = gg; sv[m][i][3] = og; sv[m][i][4] = cn; sv[m][i][5] = hn;
                    cstate[m][i] = cn;
                    if (more) {
                        if (NC > 1) {
                            Hs[bl * HSP + 32 * m + ul] = __float2bfloat16(hn);
                        } else {                           // one CTA: straight into the next step's operand buffer
                            const int kb = col >> 6, j = (col & 63) >> 3, e = col & 7;
                            *reinterpret_cast<__nv_bfloat16*>(gbase + (sHnext - base) + kb * 4096 + bl * 128 + ((j ^ (bl & 7)) << 4) + 2 * e) =
                                __float2bfloat16(hn);
                        }
                    }
                }
            }
            if (more) {
                if (NC > 1) {
                    // h_s (bf16) -> the next step's operand buffer of every CTA of the cluster, 16 bytes at a time:
                    // (chunk, destination) pairs dealt to all 512 threads, consecutive lanes = consecutive chunks of a row
                    compute_sync();                        // the staged h_s is complete
                    for (int idx = tid; idx < ((p.dbg & 1) ? 0 : 32 * (UC / 8) * NC); idx += CW * 32) {
                        const int ch = idx % (32 * (UC / 8));
                        const unsigned d = (unsigned)(idx / (32 * (UC / 8)));
                        const int bl = ch / (UC / 8), jc = ch - bl * (UC / 8);
                        const uint4 v = *reinterpret_cast<const uint4*>(Hs + bl * HSP + 8 * jc);
                        const int colg = (int)c * UC + 8 * jc, kb = colg >> 6, j = (colg & 63) >> 3;
                        const uint32_t dst = sHnext + (uint32_t)kb * 4096u + (uint32_t)bl * 128u + (uint32_t)((j ^ (bl & 7)) << 4);
                        st_cluster_v4(map_to_rank(dst, d), v);
                    }
                }
                fence_proxy_async_all();                   // generic-proxy writes -> visible to the tensor core's reads
            }
            fence_before();
        }
        if (more && NC > 1 && !(p.dbg & 8)) cluster_arrive();   // this thread's part of h_s is on its way
        if (warp < CW) {
            // ---- off the critical path: this step's outputs and the next step's input projection
            if (more && !(p.dbg & 16)) fetch_xp(s + 1);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int col = (int)c * UC + 32 * m + ul;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int b = b0 + bl0 + i;
                    if (b < p.B && !(p.dbg & 4)) {
                        const long long row = (long long)b * p.T + t;
                        if (p.gates) {
                            float* gp = p.gates + row * 4 * H;
                            gp[col] = sv[m][i][0]; gp[H + col] = sv[m][i][1]; gp[2 * H + col] = sv[m][i][2]; gp[3 * H + col] = sv[m][i][3];
                        }
                        if (p.cst) p.cst[row * H + col] = sv[m][i][4];
                        if (p.hprev) p.hprev[row * H + col] = hlast[m][i];
                        p.out[row * p.ldo + col] = sv[m][i][5];
                    }
                    hlast[m][i] = sv[m][i][5];
                }
            }
        }
    }
    if (NC > 1) cluster_sync(); else __syncthreads();      // nobody leaves while its shared memory may still be written
    if (warp == CW) {
        fence_after();
        tmem_dealloc(tmem, TCOLS);
    }
}

// ---------------------------------------------------------------------------------------------------- backward
template <int H>
__global__ void __launch_bounds__(TH, 1)
lstm_bwd_tc_kernel(const Bwd p) {
    using G_ = Geo<H>;
    constexpr int MT = G_::MT, UC = G_::UC, NC = G_::NC, KBB = G_::KBB;
    constexpr int KT = H / 128;                            // 128-row tiles of the transposed product (all H units)
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned c = NC > 1 ? cluster_ctarank() : 0u;
    const int b0 = (blockIdx.x / NC) * BG;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sDG = sA + G_::A_BYTES, sR = sDG + G_::DG_BYTES;
    uint8_t* const gA = gbase;
    uint8_t* const gDG = gbase + G_::A_BYTES;
    float* const R = reinterpret_cast<float*>(gbase + G_::A_BYTES + G_::DG_BYTES);                  // [NC][32][UC]

    // W^T slice -> the A operand: tile kt (units 128 kt ..), k-block kb over the CTA's local gate rows r = 128 m + 32 g + u:
    // A[k][r] = W[g H + c UC + 32 m + u][k].  Thread = (local gate row r, 8 consecutive k): one 32-byte read, eight
    // 2-byte scattered stores (setup only).
    for (int idx = tid; idx < MT * 128 * (H / 8); idx += TH) {
        const int r = idx / (H / 8), k8 = (idx - r * (H / 8)) * 8;
        const int m = r >> 7, rr = r & 127, g = rr >> 5, u = rr & 31;
        const float4* src = reinterpret_cast<const float4*>(p.whh + (long long)(g * H + (int)c * UC + 32 * m + u) * H + k8);
        const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
        const float w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        const int kb = r >> 6, j = (r & 63) >> 3, e = r & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k8 + i, kt = k >> 7, kr = k & 127;
            *reinterpret_cast<__nv_bfloat16*>(gA + (size_t)(kt * KBB + kb) * 16384 + kr * 128 + ((j ^ (kr & 7)) << 4) + 2 * e) =
                __float2bfloat16(w[i]);
        }
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr uint32_t TCOLS = KT * 32 < 32 ? 32 : KT * 32;
    if (warp == CW) tmem_alloc(smem_u32(&tmem_slot), TCOLS);
    fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_slot;
    if (NC > 1) cluster_sync();

    const int q = warp & 3, cq = warp >> 2;
    const int ul = lane, bl0 = 2 * warp;
    const int G4 = 4 * H;
    float dc[MT][2], dh[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m) { dc[m][0] = dc[m][1] = 0.f; dh[m][0] = dh[m][1] = 0.f; }

    // what a step reads from the forward pass (and the external gradient) does not depend on the recurrence: it is
    // fetched one step ahead, right after the previous step's dgates, and lands during that step's MMAs / exchange
    float sv[MT][2][7];                                     // i, f, g, o, c, c_prev, dout
    auto fetch = [&](int step) {
        const int tt = p.reverse ? p.T - 1 - step : step;
        const int tp = p.reverse ? tt + 1 : tt - 1;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const int col = (int)c * UC + 32 * m + ul;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int b = b0 + bl0 + i;
#pragma unroll
                for (int j = 0; j < 7; ++j) sv[m][i][j] = 0.f;
                if (b < p.B) {
                    const long long row = (long long)b * p.T + tt;
                    const float* g = p.gates + row * G4;
                    sv[m][i][0] = __ldg(g + col); sv[m][i][1] = __ldg(g + H + col);
                    sv[m][i][2] = __ldg(g + 2 * H + col); sv[m][i][3] = __ldg(g + 3 * H + col);
                    sv[m][i][4] = __ldg(p.cst + row * H + col);
                    if (step > 0) sv[m][i][5] = __ldg(p.cst + ((long long)b * p.T + tp) * H + col);
                    if (p.dout) {
                        if (p.dout_step < 0) sv[m][i][6] = __ldg(p.dout + row * p.ldo + col);
                        else if (tt == p.dout_step) sv[m][i][6] = __ldg(p.dout + (long long)b * p.ldo + col);
                    }
                }
            }
        }
    };
    if (warp < CW) fetch(p.nsteps - 1);
    if (NC > 1 && p.nsteps > 1) cluster_arrive();              // pairs with the first iteration's wait (the buffer starts free)

    for (int s = p.nsteps - 1; s >= 0; --s) {
        const int t = p.reverse ? p.T - 1 - s : s;
        float pdv[MT][2][4];
        if (warp < CW) {
            // ---- dgates of the CTA's units at step t: fp32 out, bf16 straight into the B operand (k-blocks of 64 local
            // gate rows r = 128 m + 32 g + u, [32 batch rows][128 B] swizzled)
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int col = (int)c * UC + 32 * m + ul;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int bl = bl0 + i, b = b0 + bl;
                    const float ig = sv[m][i][0], fg = sv[m][i][1], gg = sv[m][i][2], og = sv[m][i][3];
                    const float tc = tanh_(sv[m][i][4]), cprev = sv[m][i][5];
                    const float dht = dh[m][i] + sv[m][i][6];
                    const float dct = dc[m][i] + dht * og * (1.f - tc * tc);
                    const float pd[4] = {dct * gg * ig * (1.f - ig), dct * cprev * fg * (1.f - fg),
                                         dct * ig * (1.f - gg * gg), dht * tc * og * (1.f - og)};
#pragma unroll
                    for (int g = 0; g < 4; ++g) pdv[m][i][g] = pd[g];   // stored to HBM after the MMAs have been issued
                    dc[m][i] = dct * fg;                     // rows beyond B: all inputs are zero, everything stays zero
                    if (s > 0) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int r = 128 * m + 32 * g + ul, kb = r >> 6, j = (r & 63) >> 3, e = r & 7;
                            *reinterpret_cast<__nv_bfloat16*>(gDG + kb * 4096 + bl * 128 + ((j ^ (bl & 7)) << 4) + 2 * e) =
                                __float2bfloat16(pd[g]);
                        }
                    }
                }
            }
        }
        auto store_dgates = [&]() {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int col = (int)c * UC + 32 * m + ul;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int b = b0 + bl0 + i;
                    if (b < p.B) {
                        float* o = p.dgates + ((long long)b * p.T + t) * G4;
                        o[col] = pdv[m][i][0]; o[H + col] = pdv[m][i][1]; o[2 * H + col] = pdv[m][i][2]; o[3 * H + col] = pdv[m][i][3];
                    }
                }
            }
        };
        if (s == 0) {                                       // no step before the first one: dh_{-1} is not needed
            if (warp < CW) store_dgates();
            break;
        }
        if (warp < CW) {
            fence_proxy_async();
            fence_before();
        }
        // the B operand is complete (a CTA-local matter).  That every CTA has consumed the receive buffer of the previous
        // step is a cluster matter, but only the scatter below needs it: the barrier was ARRIVED at right after the sums
        // were read (end of the previous iteration / before the loop) and is waited for just before the scatter, so its
        // latency is off the MMA's critical path
        __syncthreads();
        if (warp < CW) {                                    // while the MMAs run: this step's dgates to HBM, next step's inputs
            store_dgates();
            fetch(s - 1);
        }
        if (warp == CW && lane == 0) {
            fence_proxy_async();
            fence_after();
            const uint32_t idesc = idesc_bf16(BG, false, false);
#pragma unroll 1
            for (int kt = 0; kt < KT; ++kt)
#pragma unroll 1
                for (int kb = 0; kb < KBB; ++kb) {
                    const uint64_t ad = desc_k_sw128(sA + (uint32_t)(kt * KBB + kb) * 16384u);
                    const uint64_t bd = desc_k_sw128(sDG + (uint32_t)kb * 4096u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) mma_bf16(tmem + (uint32_t)(32 * kt), ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                }
            mma_commit(smem_u32(&bars[0]));
        }
        if (NC > 1) cluster_wait();                            // every CTA's receive buffer is free (arrived: see above)
        if (warp < CW) {
            mbar_wait(smem_u32(&bars[0]), (uint32_t)(p.nsteps - 1 - s) & 1u);
            fence_after();
            // partial dh_{t-1}[k][b] over this CTA's gate rows, for ALL units k: row k of tile kt belongs to CTA k / UC
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                float v[8];
                tmem_ld8(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * kt + 8 * cq), v);
                const int k = 128 * kt + 32 * q + lane;
                const unsigned owner = (unsigned)(k / UC);
                const int ku = k - (int)owner * UC;
                const uint32_t dst0 = sR + (uint32_t)((((int)c * 32 + 8 * cq) * UC + ku) * 4);
                if (NC > 1) {
                    const uint32_t rdst = map_to_rank(dst0, owner);
#pragma unroll
                    for (int i = 0; i < 8; ++i) st_cluster_f32(rdst + (uint32_t)(i * UC * 4), v[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) R[(8 * cq + i) * UC + ku] = v[i];
                }
            }
            fence_before();
        }
        if (NC > 1) cluster_sync(); else __syncthreads();    // all partials have landed
        if (warp < CW) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float a = 0.f;
#pragma unroll 4
                    for (int src = 0; src < NC; ++src) a += R[(src * 32 + bl0 + i) * UC + 32 * m + ul];   // fixed order
                    dh[m][i] = a;
                }
        }
        if (NC > 1 && s > 1) cluster_arrive();                 // this CTA's receive buffer may be overwritten again (the
                                                               // next iteration waits for it unless it is the last one)
    }
    if (NC > 1) cluster_sync(); else __syncthreads();
    if (warp == CW) {
        fence_after();
        tmem_dealloc(tmem, TCOLS);
    }
}

template <int H, typename Kern, typename Arg>
static int launch(Kern kern, const Arg& arg, int B, size_t smem, lr_stream_t stream, const char* name) {
    constexpr int NC = Geo<H>::NC;
    cudaError_t e = lr::ensure_max_dynamic_smem(kern, (int)smem);
    if (e == cudaSuccess && NC > 8) {
        static bool allowed = false;                        // idempotent attribute: a repeated set is harmless
        if (!allowed) { e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); allowed = e == cudaSuccess; }
    }
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "%s attributes: %s", name, cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(((B + BG - 1) / BG) * NC));
    cfg.blockDim = dim3(TH);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, arg);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "%s launch: %s", name, cudaGetErrorString(e));
    lr::count_launch();
    return LR_OK;
}

}  // namespace lt

extern "C" int lr_lstm_fwd_tc(const float* xproj, long long ldx, const float* bhh, const float* whh, float* out, long long ldo,
                              float* gates, float* cst, float* hprev, int B, int T, int H, int nsteps, int reverse,
                              lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && T > 0 && (H == 128 || H == 256 || H == 512), "lr_lstm_fwd_tc: H must be 128, 256 or 512");
    LR_CHECK_ARG(nsteps >= 0 && nsteps <= T, "lr_lstm_fwd_tc: nsteps outside 0..T");
    if (B == 0 || nsteps == 0) return LR_OK;
    LR_CHECK_ARG(xproj && whh && out, "lr_lstm_fwd_tc: null pointer");
    LR_CHECK_ALIGN(whh);
    lt::Fwd p;
    p.xproj = xproj; p.ldx = ldx; p.bhh = bhh; p.whh = whh; p.out = out; p.ldo = ldo; p.gates = gates; p.cst = cst; p.hprev = hprev;
    p.B = B; p.T = T; p.H = H; p.nsteps = nsteps; p.reverse = reverse;
    static const int dbg = getenv("LIPREAD_LSTM_DBG") ? atoi(getenv("LIPREAD_LSTM_DBG")) : 0;
    p.dbg = dbg;
    int rc;
    if (H == 128) rc = lt::launch<128>(lt::lstm_fwd_tc_kernel<128>, p, B, lt::Geo<128>::FWD_SMEM, stream, "lstm_fwd_tc_kernel");
    else if (H == 256) rc = lt::launch<256>(lt::lstm_fwd_tc_kernel<256>, p, B, lt::Geo<256>::FWD_SMEM, stream, "lstm_fwd_tc_kernel");
    else rc = lt::launch<512>(lt::lstm_fwd_tc_kernel<512>, p, B, lt::Geo<512>::FWD_SMEM, stream, "lstm_fwd_tc_kernel");
    if (rc) return rc;
    LR_CHECK_LAUNCH("lstm_fwd_tc_kernel");
    return LR_OK;
}

extern "C" int lr_lstm_bwd_tc(const float* dout, long long ldo, int dout_step, const float* gates, const float* cst,
                              const float* whh, float* dgates, int B, int T, int H, int nsteps, int reverse, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && T > 0 && (H == 128 || H == 256 || H == 512), "lr_lstm_bwd_tc: H must be 128, 256 or 512");
    LR_CHECK_ARG(nsteps >= 0 && nsteps <= T, "lr_lstm_bwd_tc: nsteps outside 0..T");
    if (B == 0 || nsteps == 0) return LR_OK;
    LR_CHECK_ARG(gates && cst && whh && dgates, "lr_lstm_bwd_tc: null pointer");
    LR_CHECK_ALIGN(whh);
    lt::Bwd p;
    p.dout = dout; p.ldo = ldo; p.dout_step = dout_step; p.gates = gates; p.cst = cst; p.whh = whh; p.dgates = dgates;
    p.B = B; p.T = T; p.H = H; p.nsteps = nsteps; p.reverse = reverse;
    int rc;
    if (H == 128) rc = lt::launch<128>(lt::lstm_bwd_tc_kernel<128>, p, B, lt::Geo<128>::BWD_SMEM, stream, "lstm_bwd_tc_kernel");
    else if (H == 256) rc = lt::launch<256>(lt::lstm_bwd_tc_kernel<256>, p, B, lt::Geo<256>::BWD_SMEM, stream, "lstm_bwd_tc_kernel");
    else rc = lt::launch<512>(lt::lstm_bwd_tc_kernel<512>, p, B, lt::Geo<512>::BWD_SMEM, stream, "lstm_bwd_tc_kernel");
    if (rc) return rc;
    LR_CHECK_LAUNCH("lstm_bwd_tc_kernel");
    return LR_OK;
}
