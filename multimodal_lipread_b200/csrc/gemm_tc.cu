// Tensor-core GEMM for the trunk's 1x1 convolutions (sm_100a: TMA -> shared memory -> tcgen05.mma kind::tf32
// -> TMEM accumulator -> tcgen05.ld epilogue).
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
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 16 columns at a time -> bias / activation / residual -> row-major
//               float4 stores; per-column sum / sum of squares (train-mode BatchNorm statistics) reduced with a
//               transpose-reduce shuffle network, shared-memory float atomics, one double atomic per column/CTA
// Several CTAs are co-resident per SM (smem <= 2/SM for BN <= 96), so one CTA's epilogue overlaps another's loads.
#include <cuda.h>

#include "nn_common.cuh"

namespace tc {

constexpr int BM = 128, BK = 32, MAX_STAGES = 4, THREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 4;     // 16 KB

struct P {
    int M, N, K, BN, stages;
    int kchunk;          // reduce-dimension elements handled by one CTA (multiple of BK); gridDim.z CTAs split K
    int atomic_out;      // 1: C += tile with red.global.add (split-K wgrad); bias / act / residual / stats unused
    float* C; long long ldc;
    const float* bias; const float* R; long long ldr;
    double* stats;
    int act;
};

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
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows at a 128 B pitch, 8-row groups
// 1024 B apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
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

// column sums of a 32 (lanes) x 16 (values) tile: afterwards every lane holds the full 32-lane sum of column
// ((lane>>1) & 15) -- 16 shuffles instead of 80.
__device__ __forceinline__ float colsum16(const float* v, int lane) {
    float a[8];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = hi ? v[i] : v[i + 8];
            const float keep = hi ? v[i + 8] : v[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    float b[4];
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = hi ? a[i] : a[i + 4];
            const float keep = hi ? a[i + 4] : a[i];
            b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    float c[2];
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = hi ? b[i] : b[i + 2];
            const float keep = hi ? b[i + 2] : b[i];
            c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    float d;
    {
        const bool hi = lane & 2;
        const float send = hi ? c[0] : c[1];
        const float keep = hi ? c[1] : c[0];
        d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return d + __shfl_xor_sync(0xffffffffu, d, 1);
}
// column index held by `lane` after colsum16
__device__ __forceinline__ int colsum16_col(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const P p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 1];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_sum[256], s_sq[256];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * p.BN;
    const int kbeg = blockIdx.z * p.kchunk;
    const int kend = min(p.K, kbeg + p.kchunk);
    const int num_kb = (kend - kbeg + BK - 1) / BK;
    const int b_chunks = (p.BN + 31) / 32;
    const uint32_t b_stage_bytes = B_MN ? (uint32_t)b_chunks * 4096u : (uint32_t)p.BN * BK * 4;
    const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;        // 1024 B alignment for SWIZZLE_128B
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[MAX_STAGES]);
    const uint32_t tmem_full = smem_u32(&bars[2 * MAX_STAGES]);
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)p.BN) tmem_cols <<= 1;

    for (int i = threadIdx.x; i < 256; i += THREADS) { s_sum[i] = 0.f; s_sq[i] = 0.f; }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tmem_full, 1);
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

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                const uint32_t a_dst = tiles + s * stage_bytes, b_dst = a_dst + A_STAGE_BYTES;
                mbar_expect_tx(full0 + 8 * s, stage_bytes);
                const int k0 = kbeg + kb * BK;
                if (A_MN) {
#pragma unroll
                    for (int c = 0; c < BM / 32; ++c) tma_load_2d(a_dst + c * 4096, &tmA, full0 + 8 * s, m0 + c * 32, k0);
                } else {
                    tma_load_2d(a_dst, &tmA, full0 + 8 * s, k0, m0);
                }
                if (B_MN) {
                    for (int c = 0; c < b_chunks; ++c) tma_load_2d(b_dst + c * 4096, &tmB, full0 + 8 * s, n0 + c * 32, k0);
                } else {
                    tma_load_2d(b_dst, &tmB, full0 + 8 * s, k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(p.BN, A_MN, B_MN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                fence_after();
                const uint32_t a_src = tiles + s * stage_bytes, b_src = a_src + A_STAGE_BYTES;
                const uint64_t adesc = A_MN ? make_desc_mn_sw128(a_src) : make_desc_k_sw128(a_src);
                const uint64_t bdesc = B_MN ? make_desc_mn_sw128(b_src) : make_desc_k_sw128(b_src);
                // one MMA consumes 8 reduce-elements: K-major: 32 bytes along the swizzled row (2 address units);
                // MN-major: one 8-row group = 1024 bytes (64 address units)
#pragma unroll
                for (int k = 0; k < BK / 8; ++k)
                    mma_tf32(tmem_d, adesc + (A_MN ? 64 : 2) * k, bdesc + (B_MN ? 64 : 2) * k, idesc,
                             (kb > 0 || k > 0) ? 1u : 0u);
                mma_commit(empty0 + 8 * s);
            }
            mma_commit(tmem_full);
        }
    } else {
        // ---------------- epilogue: warp w owns TMEM lanes 32*(w & 3) .. +31 == tile rows
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < p.M;
        mbar_wait(tmem_full, 0);
        fence_after();
        const bool vec_c = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        const bool vec_r = p.R && ((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.R) & 15) == 0);
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
            const int n = n0 + c0;
            if (n >= p.N) break;                         // warp-uniform
            float v[16];
            tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            const bool full16 = n + 15 < p.N;
            if (p.atomic_out) {
                if (row_ok && num_kb > 0) {
                    float* cr = p.C + (long long)row * p.ldc + n;
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (n + j < p.N) atomicAdd(cr + j, v[j]);
                }
                continue;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x = v[j];
                if (p.bias && n + j < p.N) x += p.bias[n + j];
                v[j] = nn::act_fwd(x, p.act);
            }
            if (p.R && row_ok) {
                const float* rr = p.R + (long long)row * p.ldr + n;
                if (vec_r && full16) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) { const float4 t = nn::ld4(rr + j); v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w; }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (n + j < p.N) v[j] += rr[j];
                }
            }
            if (row_ok) {
                float* cr = p.C + (long long)row * p.ldc + n;
                if (vec_c && full16) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) nn::st4(cr + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (n + j < p.N) cr[j] = v[j];
                }
            }
            if (p.stats) {
                float sq[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { if (!row_ok) v[j] = 0.f; sq[j] = v[j] * v[j]; }
                const float cs = colsum16(v, lane), cq = colsum16(sq, lane);
                if ((lane & 1) == 0) {
                    const int col = c0 + colsum16_col(lane);
                    atomicAdd(&s_sum[col], cs);
                    atomicAdd(&s_sq[col], cq);
                }
            }
        }
        fence_before();
        if (p.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");            // the four epilogue warps only
            const int col = threadIdx.x - 64;
            if (col < p.BN && n0 + col < p.N) {
                nn::atomic_add_double(p.stats + n0 + col, (double)s_sum[col]);
                nn::atomic_add_double(p.stats + p.N + n0 + col, (double)s_sq[col]);
            }
            if (col + 128 < p.BN && n0 + col + 128 < p.N) {
                nn::atomic_add_double(p.stats + n0 + col + 128, (double)s_sum[col + 128]);
                nn::atomic_add_double(p.stats + p.N + n0 + col + 128, (double)s_sq[col + 128]);
            }
        }
    }
    __syncthreads();
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

// row-major fp32 matrix [rows][cols] (ld floats) -> tensor map with a {32 floats, box_rows} box, 128 B swizzle
static int make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows,
                    bool mn_major = false) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return lr::fail(LR_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return lr::fail(LR_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
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
    const int ntile = (N + 255) / 256;
    int bn = (N + ntile - 1) / ntile;
    bn = (bn + 15) / 16 * 16;
    tc::P p;
    p.M = M; p.N = N; p.K = K; p.BN = bn; p.C = C; p.ldc = ldc; p.bias = bias; p.R = R; p.ldr = ldr; p.stats = stats; p.act = act;
    int kchunk = (K + ksplit - 1) / ksplit;
    kchunk = (kchunk + tc::BK - 1) / tc::BK * tc::BK;
    const int nz = (K + kchunk - 1) / kchunk;
    p.kchunk = kchunk;
    p.atomic_out = ksplit > 1 ? 1 : 0;
    const int num_kb = kchunk / tc::BK;
    p.stages = num_kb < tc::MAX_STAGES ? num_kb : tc::MAX_STAGES;
    CUtensorMap ma, mb;
    int rc = a_trans ? tc::make_map(&ma, A, K, M, lda, 32, true) : tc::make_map(&ma, A, M, K, lda, tc::BM);
    if (rc) return rc;
    rc = b_trans ? tc::make_map(&mb, B, K, N, ldb, 32, true) : tc::make_map(&mb, B, N, K, ldb, bn);
    if (rc) return rc;
    const size_t b_stage = b_trans ? (size_t)((bn + 31) / 32) * 4096 : (size_t)bn * tc::BK * 4;
    const size_t smem = (size_t)p.stages * (tc::A_STAGE_BYTES + b_stage) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::gemm_tf32_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tf32_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tf32_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::gemm_tf32_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return lr::fail(LR_ECUDA, "lr_gemm_tf32 smem: %s", cudaGetErrorString(e));
        configured = true;
    }
    dim3 grid((unsigned)((M + tc::BM - 1) / tc::BM), (unsigned)((N + bn - 1) / bn), (unsigned)nz);
    if (a_trans) {
        if (b_trans) tc::gemm_tf32_kernel<true, true><<<grid, tc::THREADS, smem, stream>>>(ma, mb, p);
        else tc::gemm_tf32_kernel<true, false><<<grid, tc::THREADS, smem, stream>>>(ma, mb, p);
    } else {
        if (b_trans) tc::gemm_tf32_kernel<false, true><<<grid, tc::THREADS, smem, stream>>>(ma, mb, p);
        else tc::gemm_tf32_kernel<false, false><<<grid, tc::THREADS, smem, stream>>>(ma, mb, p);
    }
    lr::count_launch();
    LR_CHECK_LAUNCH("gemm_tf32_kernel");
    return LR_OK;
}
