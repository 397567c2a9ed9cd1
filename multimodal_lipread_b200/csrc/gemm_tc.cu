// Tensor-core GEMM for the trunk's 1x1 convolutions (sm_100a: TMA -> shared memory -> tcgen05.mma kind::tf32 /
// kind::f16 -> TMEM accumulator -> tcgen05.ld epilogue).  Two operand types share the kernel: fp32 storage with TF32
// products (lr_gemm_tf32) and bf16 storage with bf16 products (lr_gemm_bf16, precision "bf16": activations, their
// gradients and a bf16 shadow of the weights live in HBM as bfloat16; the accumulator is fp32 in both cases and the
// result leaves as fp32 or bf16).  A 128-byte swizzled smem row holds 32 floats or 64 bf16, so a pipeline stage is
// the same 16 KB + B tile in both modes and covers twice the reduce-depth in bf16.
//
//   C[M,N] (ldc) = epilogue( A[M,K] (lda, K-major) . B[N,K]^T (ldb, K-major) )        fp32 in HBM, TF32 products,
//                                                                                     fp32 accumulation in TMEM
// The channels-last activation matrix IS the K-major A operand and the torch conv weight [Cout,Cin] IS the
// K-major B operand, so both are fetched by TMA straight from where they live (128-byte swizzle, zero fill
// for the K / N / M tails) and never converted or repacked.  These GEMMs are HBM-bound (AI 6..40 FLOP/B):
// the point of the tensor pipe here is to make the math free so the kernel streams A in and C out.
//
// One CTA computes a 128 x BN tile (BN = N rounded up to 16, <= 256; wider N is split over gridDim.y):
//   warp 0      TMA producer (one elected lane), 128 x 32-float A box + BN x 32-float B box per stage
//   warp 1      TMEM allocation + MMA issue (one lane): 4 x tcgen05.mma (K = 8 each) per stage, tcgen05.commit
//               hands the stage back to the producer and finally signals the epilogue
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 16 columns at a time -> bias / activation -> a per-warp shared-memory
//               slab (transpose) -> contiguous row segments to HBM (+ residual), 512 bytes per store instruction;
//               per-column sum / sum of squares (train-mode BatchNorm statistics) read column-wise from the slab,
//               shared-memory float atomics, one double atomic per column and CTA
// Several CTAs are co-resident per SM (smem <= 2/SM for BN <= 96), so one CTA's epilogue overlaps another's loads.
//
// Tile batching (P::mt > 1, ksplit == 1): for the tall-skinny 1x1 convolutions (M up to 1.8 M rows, K and N <= 96: one or
// two k-blocks per tile) a CTA's life was a chain of latencies -- barrier init, TMEM allocation, descriptor fetch, one
// TMA round trip, one MMA, the epilogue, the store drain, 144 statistic atomics -- about 9 us for 20 KB of traffic, and
// 3 509 .. 14 036 such CTAs per launch.  With mt consecutive 128-row tiles per CTA the producer streams all their
// k-blocks through the same ring, each tile accumulates in its own TMEM columns and signals its own barrier, the
// epilogue warps drain tile t while the loads / MMAs of tile t+1.. are in flight, and the BatchNorm statistics leave
// once per CTA (mt times fewer atomics).
#include <cuda.h>
#include <cstring>
#include <cstdlib>

#include "nn_common.cuh"

namespace tc {

constexpr int BM = 128, MAX_STAGES = 8, THREADS = 192, MAX_MT = 8;
constexpr int A_STAGE_BYTES = BM * 128;        // 16 KB: 128 rows of one 128-byte swizzle span
// per operand type: elements per 128-byte span = reduce-elements per stage (K-major) = MN-elements per chunk (MN-major)
template <typename TO> struct Op { static constexpr int BK = 128 / (int)sizeof(TO); };
constexpr int EPI_PW = 64, EPI_PITCH = EPI_PW + 4;                    // epilogue panel width / slab pitch (floats)
constexpr int EPI_BYTES = 4 * 32 * EPI_PITCH * 4;                    // four 32-row slabs

struct P {
    int M, N, K, BN, stages;
    int sw;              // bytes per shared-memory operand row = swizzle span: 128, or 64 / 32 for thin K (K-major bf16 only)
    int mt;              // consecutive M tiles per CTA (1 unless ksplit == 1 and the output leaves through TMA)
    int tile_cols;       // TMEM columns per tile
    int kchunk;          // reduce-dimension elements handled by one CTA (multiple of BK); gridDim.z CTAs split K
    int atomic_out;      // 1: C += tile with red.global.add (split-K wgrad); bias / act / residual / stats unused
    int tma_out;         // 1: tiles leave through TMA (store, or reduce-add when accum_out) from swizzled 32x32 boxes
    int accum_out;       // tma_out only: C += tile (split-K, or an accumulating call whose R aliases C)
    void* C; long long ldc;          // fp32 or bf16 (the kernel's TC)
    const float* bias; const void* R; long long ldr;
    double* stats;
    int act;
#ifdef LR_TRACE
    long long* trace;    // [gridDim.x][32] clock64 stamps (debug builds only, -DLR_TRACE)
#endif
};
#ifdef LR_TRACE
#define LR_STAMP(i) do { p.trace[(long long)blockIdx.x * 32 + (i)] = clock64(); } while (0)
#else
#define LR_STAMP(i) do { } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows at a 128 B pitch, 8-row groups
// 1024 B apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// the same for a 64- or 32-byte row (thin K: K <= 32 / 16 bf16 elements): rows at a `sw`-byte pitch, 8-row groups
// 8 * sw bytes apart, layout type 4 = SWIZZLE_64B, 6 = SWIZZLE_32B (TMA writes it with the matching swizzle mode).
// A 128 x 16 operand then takes 4 KB of shared memory instead of a 16 KB slot that is three quarters zero fill, so
// four times as many tiles are in flight per SM.
__device__ __forceinline__ uint64_t make_desc_k_sw(uint32_t saddr, int sw) {
    const uint64_t lt = sw == 128 ? 2ull : (sw == 64 ? 4ull : 6ull);
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((8 * sw) >> 4) << 32) | (1ull << 46) | (lt << 61);
}
// MN-major operand.  For 32-bit (tf32) operands the only MN-major shared-memory layout the tensor core accepts
// is the 128-byte swizzle with a 32-byte base (layout type 1, CUTLASS Layout_MN_SW128_32B_Atom; TMA writes it
// with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): the tile is stored as chunks of 32 MN-contiguous floats; inside a
// chunk the 32 reduce-rows sit at a 128 B pitch, the swizzle atom is 4 rows (SBO = 512 B between 4-row groups),
// chunks are 4096 B apart (= LBO).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}
// MN-major 16-bit operand: the plain 128-byte swizzle (layout type 2; CUTLASS Layout_MN_SW128_Atom<bf16>, TMA
// CU_TENSOR_MAP_SWIZZLE_128B): chunks of 64 MN-contiguous elements; inside a chunk the 64 reduce-rows of a stage sit
// at a 128 B pitch, the swizzle atom is 8 rows (SBO = 1024 B between 8-row groups), chunks are 8192 B apart (= LBO).
__device__ __forceinline__ uint64_t make_desc_mn_sw128_b16(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16 with bf16 operands: D = F32, A = B = BF16 (format 1)
__device__ __forceinline__ uint32_t make_idesc_bf16(int bn, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
           ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// instruction descriptor, kind::tf32: D = F32, A = B = TF32, M = 128, N = bn; bit 15 / 16: A / B is MN-major
__device__ __forceinline__ uint32_t make_idesc_tf32(int bn, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
           ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <typename TO, typename TC, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const P p) {
    constexpr int BK = Op<TO>::BK;                         // 32 (tf32) / 64 (bf16): also the MN-chunk width
    constexpr bool H = sizeof(TO) == 2;
    constexpr uint32_t CHUNK_BYTES = (uint32_t)BK * 128u;  // one MN-major chunk: BK reduce-rows of 128 bytes
    static_assert(sizeof(TC) == 4 || sizeof(TC) == 2, "output is fp32 or bf16");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + MAX_MT];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_sum[4][256], s_sq[4][256], s_bias[256];   // statistics: one row per epilogue warp

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_first = blockIdx.x * p.mt * BM, n0 = blockIdx.y * p.BN;
    const int nt = min(p.mt, (p.M - m_first + BM - 1) / BM);            // tiles of this CTA (>= 1)
    const int kbeg = blockIdx.z * p.kchunk;
    const int kend = min(p.K, kbeg + p.kchunk);
    const int num_kb = (kend - kbeg + BK - 1) / BK;
    const int b_chunks = (p.BN + BK - 1) / BK;
    const uint32_t a_stage_bytes = (A_MN || B_MN) ? (uint32_t)A_STAGE_BYTES : (uint32_t)BM * (uint32_t)p.sw;
    const uint32_t b_stage_bytes = B_MN ? (uint32_t)b_chunks * CHUNK_BYTES : (uint32_t)p.BN * (uint32_t)((A_MN || B_MN) ? 128 : p.sw);
    const uint32_t stage_bytes = a_stage_bytes + b_stage_bytes;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;        // 1024 B alignment for SWIZZLE_128B
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[MAX_STAGES]);
    const uint32_t tmem_full = smem_u32(&bars[2 * MAX_STAGES]);
    uint32_t tmem_cols = 32;
    // the epilogue reads 32 columns at a time (fp32 out) or two such loads per 64-column box (bf16 out): tile_cols
    while (tmem_cols < (uint32_t)(p.mt * p.tile_cols)) tmem_cols <<= 1;

    if (threadIdx.x == 0) LR_STAMP(0);
    if (threadIdx.x == 32) {                               // the descriptors' first fetch overlaps the CTA's set-up
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        if (p.tma_out || sizeof(TC) == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    }
    for (int i = threadIdx.x; i < 256; i += THREADS) {
#pragma unroll
        for (int w = 0; w < 4; ++w) { s_sum[w][i] = 0.f; s_sq[w][i] = 0.f; }
        s_bias[i] = (p.bias && i < p.BN && blockIdx.y * p.BN + i < p.N) ? p.bias[blockIdx.y * p.BN + i] : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int t = 0; t < p.mt; ++t) mbar_init(tmem_full + 8 * t, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    if (threadIdx.x == 0) LR_STAMP(1);

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int t = 0; t < nt; ++t) {
                const int m0 = m_first + t * BM;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    if (it < 6) LR_STAMP(2 + it);
                    const uint32_t a_dst = tiles + s * stage_bytes, b_dst = a_dst + a_stage_bytes;
                    mbar_expect_tx(full0 + 8 * s, stage_bytes);
                    const int k0 = kbeg + kb * BK;
                    if (A_MN) {
#pragma unroll
                        for (int c = 0; c < BM / BK; ++c) tma_load_2d(a_dst + c * CHUNK_BYTES, &tmA, full0 + 8 * s, m0 + c * BK, k0);
                    } else {
                        tma_load_2d(a_dst, &tmA, full0 + 8 * s, k0, m0);
                    }
                    if (B_MN) {
                        for (int c = 0; c < b_chunks; ++c) tma_load_2d(b_dst + c * CHUNK_BYTES, &tmB, full0 + 8 * s, n0 + c * BK, k0);
                    } else {
                        tma_load_2d(b_dst, &tmB, full0 + 8 * s, k0, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = H ? make_idesc_bf16(p.BN, A_MN, B_MN) : make_idesc_tf32(p.BN, A_MN, B_MN);
            int it = 0;
            for (int t = 0; t < nt; ++t) {
                const uint32_t tmem_t = tmem_d + (uint32_t)(t * p.tile_cols);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(full0 + 8 * s, ph);
                    if (it < 6) LR_STAMP(8 + it);
                    fence_after();
                    const uint32_t a_src = tiles + s * stage_bytes, b_src = a_src + a_stage_bytes;
                    const bool thin = !(A_MN || B_MN) && p.sw != 128;                 // both operands K-major at a thin span
                    const uint64_t adesc = A_MN ? (H ? make_desc_mn_sw128_b16(a_src) : make_desc_mn_sw128(a_src))
                                                : (thin ? make_desc_k_sw(a_src, p.sw) : make_desc_k_sw128(a_src));
                    const uint64_t bdesc = B_MN ? (H ? make_desc_mn_sw128_b16(b_src) : make_desc_mn_sw128(b_src))
                                                : (thin ? make_desc_k_sw(b_src, p.sw) : make_desc_k_sw128(b_src));
                    const int nk = thin ? p.sw / 32 : 4;                              // MMAs (32 bytes of reduce depth each) per stage
                    // one MMA consumes 32 bytes of reduce-depth (8 tf32 / 16 bf16 elements): K-major: 32 bytes along the
                    // swizzled row (2 address units); MN-major: 8 / 16 reduce-rows of 128 bytes (64 / 128 address units)
                    constexpr int MN_STEP = H ? 128 : 64;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k >= nk) break;
                        if (H) mma_bf16(tmem_t, adesc + (A_MN ? MN_STEP : 2) * k, bdesc + (B_MN ? MN_STEP : 2) * k, idesc,
                                        (kb > 0 || k > 0) ? 1u : 0u);
                        else mma_tf32(tmem_t, adesc + (A_MN ? MN_STEP : 2) * k, bdesc + (B_MN ? MN_STEP : 2) * k, idesc,
                                      (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    mma_commit(empty0 + 8 * s);
                }
                mma_commit(tmem_full + 8 * t);
            }
            LR_STAMP(14);
        }
    } else {
        // ---------------- epilogue: warp w owns TMEM lanes 32*(w & 3) .. +31 == tile rows.
        // A thread reads ITS row from TMEM, so storing straight to HBM would write 32 different rows per
        // instruction (16 bytes of 32 cache lines: the store unit serialises them -- measured 3-5k cycles per 16
        // columns).  Instead each warp transposes through a private shared-memory slab (32 rows x 64 columns, pitch
        // 68 floats: conflict-free both ways) and then writes whole contiguous row segments, 512 bytes per
        // instruction.  The slabs reuse the pipeline buffers, which are idle once the last MMA has completed.
        const int q = warp & 3;
        float* const Cf = static_cast<float*>(p.C);                          // used by the fp32-output paths only
        const float* const Rf = static_cast<const float*>(p.R);
        const bool vec_c = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        const bool vec_r = p.R && ((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.R) & 15) == 0);
        // staging boxes: behind the pipeline ring when tiles are batched (the ring is still busy with later tiles),
        // else on top of it (idle once the only tile's last MMA has completed)
        const uint32_t epi0 = p.mt > 1 ? tiles + (uint32_t)p.stages * stage_bytes : tiles;
        int nbox = 0;                                                        // runs across tiles: two boxes in flight
        for (int t = 0; t < nt; ++t) {
        const int m0 = m_first + t * BM;
        const int row0 = m0 + q * 32;
        const int rv = max(0, min(32, p.M - row0));                         // valid rows of this warp's slab
        const uint32_t tmem_t = tmem_d + (uint32_t)(t * p.tile_cols);
        float* slab = reinterpret_cast<float*>(smem_raw + (tiles - smem_u32(smem_raw))) + q * 32 * EPI_PITCH;
        mbar_wait(tmem_full + 8 * t, 0);
        if (threadIdx.x == 64) LR_STAMP(15);
        fence_after();
        if constexpr (sizeof(TC) == 2) {
            // ---- bf16 output: 32 rows x 64 columns per box (again one 128-byte swizzled row per tile row).  A lane
            // owns its row: two 32-column TMEM loads -> bias / activation (+ residual, read straight from HBM: the
            // lane's 64 bf16 are one contiguous 128-byte line) -> round to bf16 -> eight 16-byte chunks at chunk
            // position j ^ (row & 7) -> TMA store (clips rows >= M, columns >= N).  Train-mode BatchNorm statistics are
            // taken from the ROUNDED values (what the consumer will normalise): lane L sums columns 2L, 2L+1 down the
            // 32 rows, one conflict-free 4-byte read per row.
            const nn::bf16* const Rh = static_cast<const nn::bf16*>(p.R);
            const uint32_t stg0 = epi0 + (uint32_t)q * 8192u;
            uint8_t* stg0_g = smem_raw + (stg0 - smem_u32(smem_raw));
            const int plain_out = (p.bias == nullptr && p.act == LR_ACT_NONE) ? 1 : 0;
            for (int c0 = 0; c0 < p.BN; c0 += 64, ++nbox) {
                const int n = n0 + c0;
                if (n >= p.N) break;                                         // warp-uniform
                const int ncol = min(64, p.BN - c0);                         // columns of this box: 16, 32, 48 or 64
                const uint32_t buf = (uint32_t)(nbox & 1) * 4096u;
                if (nbox >= 2) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }
                uint8_t* rowp = stg0_g + buf + lane * 128;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (32 * h >= ncol) break;                               // narrow tiles (N = 16 / 24 / 32): one half only
                    float v[32];
                    tmem_ld32(tmem_t + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + 32 * h), v);
                    if (plain_out == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bq = *reinterpret_cast<const float4*>(&s_bias[c0 + 32 * h + j]);
                            v[j] += bq.x; v[j + 1] += bq.y; v[j + 2] += bq.z; v[j + 3] += bq.w;
                        }
                        switch (p.act) {
                            case LR_ACT_RELU:
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                                break;
                            case LR_ACT_NONE: break;
                            default:
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = nn::act_fwd(v[j], p.act);
                        }
                    }
                    if (Rh && lane < rv) {
                        const nn::bf16* rrow = Rh + (long long)(row0 + lane) * p.ldr + n + 32 * h;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (n + 32 * h + 8 * j < p.N) {                  // N % 8 == 0: a chunk is wholly in or out
                                const uint4 u = *reinterpret_cast<const uint4*>(rrow + 8 * j);
                                const float4 a = nn::unpack4(make_uint2(u.x, u.y)), b = nn::unpack4(make_uint2(u.z, u.w));
                                v[8 * j] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += a.z; v[8 * j + 3] += a.w;
                                v[8 * j + 4] += b.x; v[8 * j + 5] += b.y; v[8 * j + 6] += b.z; v[8 * j + 7] += b.w;
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (32 * h + 8 * j < ncol) {                         // chunks beyond the tile are clipped by TMA anyway
                            const uint2 lo = nn::pack4(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]));
                            const uint2 hi = nn::pack4(make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]));
                            *reinterpret_cast<uint4*>(rowp + (((4 * h + j) ^ (lane & 7)) << 4)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && num_kb > 0) {
                    tma_store_2d(&tmC, stg0 + buf, n, row0);
                    bulk_commit();
                }
                if (p.stats) {
                    // column pair cp = columns 2cp, 2cp+1 = 4-byte word (cp & 3) of 16-byte chunk (cp >> 2).  A box of
                    // ncol <= 32 / <= 16 columns has only 16 / 8 pairs: the lanes then split the 32 rows 2 / 4 ways
                    // (row groups) and the groups are added by shuffles in fixed order.
                    const int pg = ncol > 32 ? 32 : (ncol > 16 ? 16 : 8), nrg = 32 / pg;
                    const int cp = lane & (pg - 1), rg = lane / pg;
                    const uint8_t* colp = stg0_g + buf + (cp & 3) * 4;
                    float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
#pragma unroll 4
                    for (int r = rg; r < rv; r += nrg) {
                        const unsigned w = *reinterpret_cast<const unsigned*>(colp + r * 128 + (((cp >> 2) ^ (r & 7)) << 4));
                        const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
                        a0 += x.x; b0 = fmaf(x.x, x.x, b0); a1 += x.y; b1 = fmaf(x.y, x.y, b1);
                    }
                    for (int o = pg; o < 32; o <<= 1) {
                        a0 += __shfl_xor_sync(0xffffffffu, a0, o); b0 += __shfl_xor_sync(0xffffffffu, b0, o);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, o); b1 += __shfl_xor_sync(0xffffffffu, b1, o);
                    }
                    if (rg == 0 && n + 2 * cp < p.N && 2 * cp < ncol) {      // N even; the slot is this lane's own
                        s_sum[q][c0 + 2 * cp] += a0; s_sq[q][c0 + 2 * cp] += b0;
                        s_sum[q][c0 + 2 * cp + 1] += a1; s_sq[q][c0 + 2 * cp + 1] += b1;
                    }
                }
            }
        } else
        if (p.tma_out) {
            // ---- TMA epilogue: 32 rows x 32 columns at a time through a 128-byte-swizzled box (row r, 16-byte chunk
            // c stored at chunk c ^ (r & 7): the row-per-lane writes and the column-per-lane statistic reads are both
            // bank-conflict free); one lane hands the box to TMA, which clips rows >= M / columns >= N and either
            // stores or reduce-adds it.  Two boxes per warp are in flight.
            const uint32_t stg0 = epi0 + (uint32_t)q * 8192u;
            uint8_t* stg0_g = smem_raw + (stg0 - smem_u32(smem_raw));
            const bool do_out = num_kb > 0 || !p.accum_out;
            const int plain_out = (p.atomic_out || (p.bias == nullptr && p.act == LR_ACT_NONE)) ? 1 : 0;
            for (int c0 = 0; c0 < p.BN; c0 += 32, ++nbox) {
                const int n = n0 + c0;
                if (n >= p.N) break;                                         // warp-uniform
                const uint32_t buf = (uint32_t)(nbox & 1) * 4096u;
                if (threadIdx.x == 64 && nbox == 1) LR_STAMP(19);
                if (nbox >= 2) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }
                uint8_t* rowp = stg0_g + buf + lane * 128;
                {
                    float v[32];
                    tmem_ld32(tmem_t + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    if (threadIdx.x == 64 && nbox == 1) LR_STAMP(20);
                    if (plain_out == 0) {                                    // bias and / or activation (uniform branch)
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bq = *reinterpret_cast<const float4*>(&s_bias[c0 + j]);
                            v[j] += bq.x; v[j + 1] += bq.y; v[j + 2] += bq.z; v[j + 3] += bq.w;
                        }
                        switch (p.act) {                                      // hoisted: one branch per box, not per element
                            case LR_ACT_RELU:
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                                break;
                            case LR_ACT_NONE: break;
                            default:
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = nn::act_fwd(v[j], p.act);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) =
                            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if (threadIdx.x == 64 && nbox == 1) LR_STAMP(21);
                fence_proxy_async();
                if (threadIdx.x == 64 && nbox == 1) LR_STAMP(22);
                __syncwarp();
                if (lane == 0 && do_out) {
                    if (p.accum_out) tma_reduce_add_2d(&tmC, stg0 + buf, n, row0);
                    else tma_store_2d(&tmC, stg0 + buf, n, row0);
                    bulk_commit();
                }
                if (threadIdx.x == 64 && nbox == 1) LR_STAMP(23);
                if (p.stats && n + lane < p.N) {
                    const uint8_t* colp = stg0_g + buf + (lane & 3) * 4;
                    float a = 0.f, b = 0.f;
                    if (rv == 32) {
                        // the 8 rows of a swizzle period hit 8 different 16-byte chunks: their offsets are loop constants
                        float a1 = 0.f, b1 = 0.f, a2 = 0.f, b2 = 0.f, a3 = 0.f, b3 = 0.f;
#pragma unroll
                        for (int r = 0; r < 32; r += 4) {
                            const float x0 = *reinterpret_cast<const float*>(colp + r * 128 + (((lane >> 2) ^ (r & 7)) << 4));
                            const float x1 = *reinterpret_cast<const float*>(colp + (r + 1) * 128 + (((lane >> 2) ^ ((r + 1) & 7)) << 4));
                            const float x2 = *reinterpret_cast<const float*>(colp + (r + 2) * 128 + (((lane >> 2) ^ ((r + 2) & 7)) << 4));
                            const float x3 = *reinterpret_cast<const float*>(colp + (r + 3) * 128 + (((lane >> 2) ^ ((r + 3) & 7)) << 4));
                            a += x0; b = fmaf(x0, x0, b); a1 += x1; b1 = fmaf(x1, x1, b1);
                            a2 += x2; b2 = fmaf(x2, x2, b2); a3 += x3; b3 = fmaf(x3, x3, b3);
                        }
                        a = (a + a1) + (a2 + a3); b = (b + b1) + (b2 + b3);
                    } else {
                        for (int r = 0; r < rv; ++r) {
                            const float x = *reinterpret_cast<const float*>(colp + r * 128 + (((lane >> 2) ^ (r & 7)) << 4));
                            a += x; b = fmaf(x, x, b);
                        }
                    }
                    s_sum[q][c0 + lane] += a;                                // warp q's own slot (this lane's): no atomics,
                    s_sq[q][c0 + lane] += b;                                 // the four warps are added in fixed order below
                }
                if (threadIdx.x == 64 && nbox == 1) LR_STAMP(24);
            }
        } else
        for (int p0 = 0; p0 < p.BN; p0 += EPI_PW) {
            const int nbase = n0 + p0;
            if (nbase >= p.N) break;                                         // warp-uniform
            const int pwv = min(min(EPI_PW, p.BN - p0), p.N - nbase);        // valid columns of this panel
            // ---- TMEM -> registers (bias, activation) -> slab[row = lane][col]
            for (int c0 = 0; c0 < pwv; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_t + ((uint32_t)(q * 32) << 16) + (uint32_t)(p0 + c0), v);
                if (!p.atomic_out) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = nn::act_fwd(v[j] + s_bias[p0 + c0 + j], p.act);
                }
                float* dst = slab + lane * EPI_PITCH + c0;
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            __syncwarp();
            // ---- slab -> HBM in contiguous row segments (+ residual); the final value goes back to the slab for the statistics
            if (num_kb > 0 || !p.atomic_out) {
                if (vec_c && (pwv & 3) == 0 && (!p.R || vec_r)) {
                    const int nv = pwv >> 2, total = rv * nv;
                    for (int idx = lane; idx < total; idx += 32) {
                        const int r = idx / nv, c4 = (idx - r * nv) * 4;
                        float4 x = *reinterpret_cast<const float4*>(slab + r * EPI_PITCH + c4);
                        float* cp = Cf + (long long)(row0 + r) * p.ldc + nbase + c4;
                        if (p.atomic_out) { atomicAdd(reinterpret_cast<float4*>(cp), x); continue; }
                        if (p.R) {
                            const float4 t = nn::ld4(Rf + (long long)(row0 + r) * p.ldr + nbase + c4);
                            x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
                            if (p.stats) *reinterpret_cast<float4*>(slab + r * EPI_PITCH + c4) = x;
                        }
                        nn::st4(cp, x);
                    }
                } else {
                    const int total = rv * pwv;
                    for (int idx = lane; idx < total; idx += 32) {
                        const int r = idx / pwv, c = idx - r * pwv;
                        float x = slab[r * EPI_PITCH + c];
                        float* cp = Cf + (long long)(row0 + r) * p.ldc + nbase + c;
                        if (p.atomic_out) { atomicAdd(cp, x); continue; }
                        if (p.R) {
                            x += Rf[(long long)(row0 + r) * p.ldr + nbase + c];
                            if (p.stats) slab[r * EPI_PITCH + c] = x;
                        }
                        *cp = x;
                    }
                }
            }
            // ---- per-column sum / sum of squares over this warp's valid rows (train-mode BatchNorm statistics)
            if (p.stats) {
                __syncwarp();
                for (int c = lane; c < pwv; c += 32) {
                    float a = 0.f, b = 0.f;
                    for (int r = 0; r < rv; ++r) { const float x = slab[r * EPI_PITCH + c]; a += x; b = fmaf(x, x, b); }
                    s_sum[q][p0 + c] = a;
                    s_sq[q][p0 + c] = b;
                }
            }
            __syncwarp();
        }
        }   // tiles of this CTA
        if (p.tma_out || sizeof(TC) == 2) {
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
        }
        fence_before();
        if (threadIdx.x == 64) LR_STAMP(16);
        if (p.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");            // the four epilogue warps only
            const int col = threadIdx.x - 64;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cc = col + 128 * h;
                if (cc < p.BN && n0 + cc < p.N) {
                    // the CTA's 128 rows in fixed warp order; across CTAs the double atomics add 24-bit partials into
                    // 53-bit sums, which is exact (order-independent) unless the partials span more than 2^29
                    const float cs = (s_sum[0][cc] + s_sum[1][cc]) + (s_sum[2][cc] + s_sum[3][cc]);
                    const float cq = (s_sq[0][cc] + s_sq[1][cc]) + (s_sq[2][cc] + s_sq[3][cc]);
                    nn::atomic_add_double(p.stats + n0 + cc, (double)cs);
                    nn::atomic_add_double(p.stats + p.N + n0 + cc, (double)cq);
                }
            }
        }
    }
    if (threadIdx.x == 64) LR_STAMP(17);
    __syncthreads();
    if (threadIdx.x == 0) LR_STAMP(18);
    if (warp == 1) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// row-major matrix [rows][cols] (ld elements, fp32 or bf16) -> tensor map with a {box_cols, box_rows} box whose
// inner extent is one 128-byte span, 128 B swizzle (the 32-byte-atom flavour for MN-major fp32 operands)
static int make_map(CUtensorMap* map, const void* base, bool is_bf16, long long rows, long long cols, long long ld,
                    int box_cols, int box_rows, bool atom32 = false, int sw = 128) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return lr::fail(LR_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const size_t es = is_bf16 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                            : (sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B)),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return lr::fail(LR_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return LR_OK;
}

template <typename TO, typename TC>
static int launch(const void* A, long long lda, int a_trans, const void* B, long long ldb, int b_trans, void* C,
                  long long ldc, int M, int N, int K, const float* bias, int act, const void* R, long long ldr,
                  double* stats, int ksplit, lr_stream_t stream, const char* name) {
    constexpr bool H = sizeof(TO) == 2, HC = sizeof(TC) == 2;
    constexpr int BK = Op<TO>::BK;
    const int ntile = (N + 255) / 256;
    int bn = (N + ntile - 1) / ntile;
    bn = (bn + 15) / 16 * 16;
    if (ntile > 1) bn = HC ? (bn + 63) / 64 * 64 : (bn + 31) / 32 * 32;   // output boxes must not straddle two CTAs' tiles
    P p;
    p.M = M; p.N = N; p.K = K; p.BN = bn; p.C = C; p.ldc = ldc; p.bias = bias; p.R = R; p.ldr = ldr; p.stats = stats; p.act = act;
    int kchunk = (K + ksplit - 1) / ksplit;
    kchunk = (kchunk + BK - 1) / BK * BK;
    const int nz = (K + kchunk - 1) / kchunk;
    p.kchunk = kchunk;
    p.atomic_out = ksplit > 1 ? 1 : 0;
#ifdef LR_TRACE
    extern long long* g_lr_trace;
    p.trace = g_lr_trace;
#endif
    const int num_kb = kchunk / BK;
    // Output path: TMA store / reduce-add whenever C is TMA-addressable (16-byte aligned rows) and the residual
    // is absent or IS C (then the call accumulates: reduce-add); otherwise the shared-memory slab path (fp32 out only).
    const bool c_tma_ok = ((ldc * (long long)sizeof(TC)) & 15) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
    const bool r_is_c = R == C && ldr == ldc;
    if (HC) {
        if (!c_tma_ok) return lr::fail(LR_EALIGN, "%s: a bf16 result needs 16-byte aligned rows (ldc %% 8 == 0)", name);
        if (R && (((ldr * 2) & 15) != 0 || (reinterpret_cast<uintptr_t>(R) & 15) != 0 || (N & 7) != 0))
            return lr::fail(LR_EALIGN, "%s: a bf16 residual needs 16-byte aligned rows and N %% 8 == 0", name);
        if (ksplit > 1) return lr::fail(LR_EINVAL, "%s: split-K accumulates in fp32 (use an fp32 result)", name);
        p.tma_out = 1; p.accum_out = 0;
    } else {
        p.tma_out = (c_tma_ok && (R == nullptr || (r_is_c && !stats && act == LR_ACT_NONE && !bias))) ? 1 : 0;
        p.accum_out = (p.atomic_out || (R != nullptr && r_is_c)) ? 1 : 0;
    }
    // tile batching: tall single-pass problems whose result leaves through TMA (own staging boxes).  mt tiles per CTA
    // while the grid still has >= ~3 CTAs per SM slot and the tiles' accumulators fit 256 TMEM columns (two CTAs / SM)
    p.tile_cols = HC ? (bn + 31) / 32 * 32 : (bn + 31) / 32 * 32;
    p.mt = 1;
    {
        static const int mt_env = getenv("LIPREAD_GEMM_MT") ? atoi(getenv("LIPREAD_GEMM_MT")) : 0;
        const long long tiles_m = (M + BM - 1) / BM, n_tiles = (N + bn - 1) / bn;
        if (nz == 1 && (HC || p.tma_out) && mt_env != 1 && num_kb == 1) {
            // measured on B200 (tools/microbench.py gemm, profiles/r2_gemm_tile_batching.txt): single-k-block tiles gain
            // from batching -- up to 8 per CTA for N <= 32 (1.8 M x 16 x 32: 124 -> 73 us), 2 for wider tiles
            // (449 152 x 72 x 16: 53 -> 49 us) -- while two-k-block tiles (K = 72 .. 96) and grids below ~3 CTAs per
            // SM slot do not
            const int cap = p.tile_cols <= 32 ? MAX_MT : 2;
            int mt = 1;
            while (mt < cap && 2 * mt * p.tile_cols <= 256 && tiles_m * n_tiles / (2 * mt) >= 3LL * lr::sm_count()) mt *= 2;
            if (mt_env > 1) mt = mt_env;
            while (mt > 1 && mt * p.tile_cols > 512) mt /= 2;
            p.mt = mt;
        }
    }
    // thin K (K-major bf16 operands, one k-block, K <= 32): 32- / 64-byte operand rows instead of the 128-byte span
    static const bool thin_env = !(getenv("LIPREAD_GEMM_THIN") && getenv("LIPREAD_GEMM_THIN")[0] == '0');
    p.sw = 128;
    if (thin_env && H && !a_trans && !b_trans && nz == 1 && num_kb == 1 && K <= 32) p.sw = K <= 16 ? 32 : 64;
    const int total_kb = num_kb * p.mt;
    static const int st_env = getenv("LIPREAD_GEMM_STAGES") ? atoi(getenv("LIPREAD_GEMM_STAGES")) : 0;
    // batched tiles: two ring slots of the 128-byte span (smem per CTA stays small enough for three CTAs per SM: more
    // loads in flight); thin spans are 4 / 2 times smaller, so every tile of the CTA gets its own slot
    const int dflt_stages = p.mt > 1 ? (p.sw == 128 ? 2 : (p.sw == 64 ? 4 : 8)) : 4;
    int max_stages = st_env >= 1 && st_env <= MAX_STAGES ? st_env : dflt_stages;
    {   // multi-k-block problems with more CTAs than SMs: cap the ring so that two CTAs share an SM (one tile per CTA, so
        // one CTA's epilogue runs under the other's main loop; config 5: 13.45 -> 13.32 ms, config 2: 7.13 -> 7.07 ms)
        static const bool occ_off = getenv("LIPREAD_GEMM_2CTA") && getenv("LIPREAD_GEMM_2CTA")[0] == '0';
        if (!occ_off && p.mt == 1 && p.sw == 128) {
            const size_t bst = b_trans ? (size_t)((bn + BK - 1) / BK) * (size_t)BK * 128 : (size_t)bn * 128;
            int cap = (int)((110 * 1024) / (A_STAGE_BYTES + bst));
            if (cap < 2) cap = 2;
            const long long ctas = (long long)((M + BM - 1) / BM) * ((N + bn - 1) / bn) * nz;
            if (ctas > lr::sm_count() && max_stages > cap) max_stages = cap;
        }
    }
    p.stages = total_kb < max_stages ? total_kb : max_stages;
    CUtensorMap ma, mb, mc;
    if (p.tma_out) {
        int rcc = make_map(&mc, C, HC, M, N, ldc, HC ? 64 : 32, 32);
        if (rcc) return rcc;
    } else {
        memset(&mc, 0, sizeof(mc));
    }
    const int bk_box = p.sw == 128 ? BK : p.sw / (int)sizeof(TO);      // elements per operand row of the span
    int rc = a_trans ? make_map(&ma, A, H, K, M, lda, BK, BK, !H) : make_map(&ma, A, H, M, K, lda, bk_box, BM, false, p.sw);
    if (rc) return rc;
    rc = b_trans ? make_map(&mb, B, H, K, N, ldb, BK, BK, !H) : make_map(&mb, B, H, N, K, ldb, bk_box, bn, false, p.sw);
    if (rc) return rc;
    const size_t b_stage = b_trans ? (size_t)((bn + BK - 1) / BK) * (size_t)BK * 128 : (size_t)bn * (size_t)p.sw;
    const size_t a_stage = (a_trans || b_trans) ? (size_t)A_STAGE_BYTES : (size_t)BM * (size_t)p.sw;
    size_t smem = (size_t)p.stages * (a_stage + b_stage);
    if (p.mt > 1) smem += 32768;                                // staging boxes of the epilogue behind the ring
    else if (smem < (size_t)EPI_BYTES) smem = EPI_BYTES;        // one tile: the epilogue slabs reuse the pipeline buffers
    smem += 1024;
    {
        cudaError_t e = lr::ensure_max_dynamic_smem(gemm_tc_kernel<TO, TC, false, false>, 200 * 1024);
        if (e == cudaSuccess) e = lr::ensure_max_dynamic_smem(gemm_tc_kernel<TO, TC, false, true>, 200 * 1024);
        if (e == cudaSuccess) e = lr::ensure_max_dynamic_smem(gemm_tc_kernel<TO, TC, true, true>, 200 * 1024);
        if (e == cudaSuccess) e = lr::ensure_max_dynamic_smem(gemm_tc_kernel<TO, TC, true, false>, 200 * 1024);
        if (e != cudaSuccess) return lr::fail(LR_ECUDA, "%s smem: %s", name, cudaGetErrorString(e));
    }
    dim3 grid((unsigned)(((M + BM - 1) / BM + p.mt - 1) / p.mt), (unsigned)((N + bn - 1) / bn), (unsigned)nz);
    if (a_trans) {
        if (b_trans) gemm_tc_kernel<TO, TC, true, true><<<grid, THREADS, smem, stream>>>(ma, mb, mc, p);
        else gemm_tc_kernel<TO, TC, true, false><<<grid, THREADS, smem, stream>>>(ma, mb, mc, p);
    } else {
        if (b_trans) gemm_tc_kernel<TO, TC, false, true><<<grid, THREADS, smem, stream>>>(ma, mb, mc, p);
        else gemm_tc_kernel<TO, TC, false, false><<<grid, THREADS, smem, stream>>>(ma, mb, mc, p);
    }
    lr::count_launch();
    LR_CHECK_LAUNCH("gemm_tc_kernel");
    return LR_OK;
}

}  // namespace tc

extern "C" int lr_gemm_tf32(const float* A, long long lda, int a_trans, const float* B, long long ldb, int b_trans,
                            float* C, long long ldc, int M, int N, int K, const float* bias, int act, const float* R,
                            long long ldr, double* stats, int ksplit, lr_stream_t stream) {
    LR_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "lr_gemm_tf32: bad dimension");
    if (M == 0 || N == 0) return LR_OK;
    LR_CHECK_ARG(A && B && C, "lr_gemm_tf32: null pointer");
    LR_CHECK_ARG(act >= LR_ACT_NONE && act <= LR_ACT_RELU6, "lr_gemm_tf32: bad activation %d", act);
    LR_CHECK_ARG((lda & 3) == 0 && (ldb & 3) == 0, "lr_gemm_tf32: lda / ldb must be multiples of 4 floats (TMA 16-byte stride)");
    LR_CHECK_ARG(ksplit >= 1, "lr_gemm_tf32: ksplit must be >= 1");
    LR_CHECK_ARG(ksplit == 1 || (act == LR_ACT_NONE && !R && !stats && !bias),
                 "lr_gemm_tf32: split-K accumulates atomically and cannot fuse bias / act / residual / stats");
    LR_CHECK_ALIGN(A); LR_CHECK_ALIGN(B);
    return tc::launch<float, float>(A, lda, a_trans, B, ldb, b_trans, C, ldc, M, N, K, bias, act, R, ldr, stats, ksplit,
                                    stream, "lr_gemm_tf32");
}

extern "C" int lr_gemm_bf16(const void* A, long long lda, int a_trans, const void* B, long long ldb, int b_trans,
                            void* C, long long ldc, int c_bf16, int M, int N, int K, const float* bias, int act,
                            const void* R, long long ldr, double* stats, int ksplit, lr_stream_t stream) {
    LR_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "lr_gemm_bf16: bad dimension");
    if (M == 0 || N == 0) return LR_OK;
    LR_CHECK_ARG(A && B && C, "lr_gemm_bf16: null pointer");
    LR_CHECK_ARG(act >= LR_ACT_NONE && act <= LR_ACT_RELU6, "lr_gemm_bf16: bad activation %d", act);
    LR_CHECK_ARG((lda & 7) == 0 && (ldb & 7) == 0, "lr_gemm_bf16: lda / ldb must be multiples of 8 bf16 (TMA 16-byte stride)");
    LR_CHECK_ARG(ksplit >= 1, "lr_gemm_bf16: ksplit must be >= 1");
    LR_CHECK_ARG(ksplit == 1 || (act == LR_ACT_NONE && !R && !stats && !bias),
                 "lr_gemm_bf16: split-K accumulates atomically and cannot fuse bias / act / residual / stats");
    LR_CHECK_ALIGN(A); LR_CHECK_ALIGN(B);
    if (c_bf16)
        return tc::launch<nn::bf16, nn::bf16>(A, lda, a_trans, B, ldb, b_trans, C, ldc, M, N, K, bias, act, R, ldr, stats,
                                              ksplit, stream, "lr_gemm_bf16");
    return tc::launch<nn::bf16, float>(A, lda, a_trans, B, ldb, b_trans, C, ldc, M, N, K, bias, act, R, ldr, stats, ksplit,
                                       stream, "lr_gemm_bf16");
}
