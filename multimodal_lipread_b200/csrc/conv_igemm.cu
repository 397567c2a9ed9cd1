// Implicit-GEMM 3x3 / stride 1 / pad 1 convolution on bf16 channels-last activations (sm_100a): the ResNet-18 trunk's
// thirteen 3x3 s1 convolutions (torchvision BasicBlock conv1 / conv2: video/models/resnet_lstm.py:79-110,
// audio/models/resnet_model.py:12-17), their input gradients and their weight gradients -- with NO patch matrix.
//
//   forward   y[p, n]   = sum_{tap, c} x[p + off(tap), c] * Wt[n, tap*Cin + c]          (off = (r-1, s-1), zero outside)
//   dgrad     dx[p, c]  = sum_{tap, n} dy[p - off(tap), n] * Wd[c, tap*Cout + n]         (same kernel, flip = 1)
//   wgrad     dW[n, tap*Cin + c] += sum_p dy[p, n] * x[p + off(tap), c]                  (second kernel)
//
// The A operand of tap (r, s) is the SAME activation tile shifted by (r-1, s-1).  It is fetched by a 4-D TMA tile of
// the activation seen as (C, W, H, F): box (64 channels, W, bh rows, bf frames) at coordinates (c0, s-1, h0+r-1, f0);
// the part of the box that falls outside the image is zero filled by the TMA unit -- the convolution's padding -- and
// the box lands in shared memory as [pixels][64 channels] rows of 128 bytes with the 128-byte swizzle, i.e. exactly
// the K-major (forward / dgrad) or MN-major (wgrad: pixels are the reduce dimension) operand tcgen05.mma reads.
// One CTA owns up to 256 output pixels (two M = 128 accumulators in TMEM) x BN output channels and walks the
// 9 * Cin/64 (tap, channel-chunk) K-blocks through a 4-stage TMA -> tcgen05.mma(kind::f16, fp32 accumulate) pipeline:
//   warp 0      TMA producer (one lane)          warp 1      TMEM allocation + MMA issue (one lane)
//   warps 2..5  epilogue: tcgen05.ld -> (+ residual) -> bf16 -> swizzled staging -> coalesced 16-byte stores, and the
//               per-channel sum / sum of squares of the rounded values (train-mode BatchNorm statistics)
//
// Stride 2 (the first convolution of layer2 / 3 / 4; forward and wgrad; s2 = 1): the tile grid is the OUTPUT grid
// (H, W) = (Hi / 2, Wi / 2) and input pixel (2 ho + r - 1, 2 wo + s - 1) has FIXED row / column parities for a given tap, so
// the input is mapped as the 5-D tensor ((pw, C), Wi / 2, ph, Hi / 2, F) and the tap's operand is again one dense box:
// (64 channels at parity pw, W, 1, bh, bf) at ((pw) * Cin + c0, -1 or 0, ph, h0 - 1 or h0, f0).  Same bytes, same shared
// memory layout, same zero fill at the borders -- no patch matrix for these layers either.  (Their input gradient is
// the stride-1 kernel on the zero-stuffed dy, see lr_zero_stuff2_h.)
#include <cstdlib>

#include "tc_common.cuh"

namespace ig {

using namespace tcc;

constexpr int THREADS = 192, MAX_STAGES = 4;

struct FP {
    int F, H, W, Cin, N;        // N: output channels of this call (Cout forward, Cin of the conv for dgrad)
    int bh, bf, tpf;            // box (64, W, bh, bf); tpf = tiles per frame when bf == 1 (else 0: a tile = bf whole frames)
    int rows, nm;               // pixels per tile (<= 256), M = 128 accumulators per tile (1 or 2)
    int BN, stages, flip, s2;
    nn::bf16* y; const nn::bf16* R; double* stats;
};

__global__ void __launch_bounds__(THREADS)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const FP p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_sum[4][128], s_sq[4][128];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, n0 = blockIdx.y * p.BN;
    int f0, h0, valid;
    if (p.tpf) { f0 = tile / p.tpf; h0 = (tile - f0 * p.tpf) * p.bh; valid = min(p.bh, p.H - h0) * p.W; }
    else { f0 = tile * p.bf; h0 = 0; valid = min(p.bf, p.F - f0) * p.H * p.W; }
    const long long base_row = ((long long)f0 * p.H + h0) * p.W;
    const int cchunks = p.Cin >> 6, num_kb = 9 * cchunks;
    const uint32_t a_bytes = (uint32_t)p.nm * 16384u, stage_bytes = a_bytes + (uint32_t)p.BN * 128u;
    const uint32_t tx_bytes = (uint32_t)p.rows * 128u + (uint32_t)p.BN * 128u;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[MAX_STAGES]), tmem_full = smem_u32(&bars[2 * MAX_STAGES]);
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(p.nm * p.BN)) tmem_cols <<= 1;

    if (threadIdx.x == 32) {                               // the descriptors' first fetch overlaps the CTA's set-up
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    }
    for (int i = threadIdx.x; i < 4 * 128; i += THREADS) { (&s_sum[0][0])[i] = 0.f; (&s_sq[0][0])[i] = 0.f; }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % p.stages;
                mbar_wait(empty0 + 8 * s, (((uint32_t)(kb / p.stages)) & 1u) ^ 1u);
                const int tap = kb / cchunks, cc = kb - tap * cchunks;
                const int r = tap / 3, q = tap - r * 3;
                const int dr = p.flip ? 1 - r : r - 1, dq = p.flip ? 1 - q : q - 1;
                const uint32_t a_dst = tiles + s * stage_bytes, b_dst = a_dst + a_bytes;
                mbar_expect_tx(full0 + 8 * s, tx_bytes);
                if (!p.s2) tma_load_4d(a_dst, &tmX, full0 + 8 * s, cc * 64, dq, h0 + dr, f0);
                else tma_load_5d(a_dst, &tmX, full0 + 8 * s, (dq & 1) * p.Cin + cc * 64, dq < 0 ? -1 : 0, dr & 1,
                                 h0 + (dr < 0 ? -1 : 0), f0);
                tma_load_2d(b_dst, &tmW, full0 + 8 * s, tap * p.Cin + cc * 64, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16(p.BN, false, false);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % p.stages;
                mbar_wait(full0 + 8 * s, ((uint32_t)(kb / p.stages)) & 1u);
                fence_after();
                const uint32_t a_src = tiles + s * stage_bytes, b_src = a_src + a_bytes;
                const uint64_t bdesc = desc_k_sw128(b_src);
                for (int mt = 0; mt < p.nm; ++mt) {
                    const uint64_t adesc = desc_k_sw128(a_src + (uint32_t)mt * 16384u);
#pragma unroll
                    for (int k = 0; k < 4; ++k)                       // 4 x 16 channels of this 64-channel chunk
                        mma_bf16(tmem_d + (uint32_t)(mt * p.BN), adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                mma_commit(empty0 + 8 * s);
            }
            mma_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        fence_after();
        uint8_t* slab = smem_raw + (tiles - smem_u32(smem_raw)) + q * 4096;         // the pipeline buffers are idle now
        for (int mt = 0; mt < p.nm; ++mt) {
            const int row_l = mt * 128 + q * 32;                                     // first tile row of this warp's slab
            const int rv = max(0, min(32, valid - row_l));
            for (int c0 = 0; c0 < p.BN; c0 += 64) {
                uint8_t* rowp = slab + lane * 128;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float v[32];
                    tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * p.BN + c0 + 32 * h), v);
                    if (p.R && lane < rv) {
                        const nn::bf16* rrow = p.R + (base_row + row_l + lane) * p.N + n0 + c0 + 32 * h;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 u = *reinterpret_cast<const uint4*>(rrow + 8 * j);
                            const float4 a = nn::unpack4(make_uint2(u.x, u.y)), b = nn::unpack4(make_uint2(u.z, u.w));
                            v[8 * j] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += a.z; v[8 * j + 3] += a.w;
                            v[8 * j + 4] += b.x; v[8 * j + 5] += b.y; v[8 * j + 6] += b.z; v[8 * j + 7] += b.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint2 lo = nn::pack4(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]));
                        const uint2 hi = nn::pack4(make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]));
                        *reinterpret_cast<uint4*>(rowp + (((4 * h + j) ^ (lane & 7)) << 4)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
                __syncwarp();
                if (p.stats) {                                            // lane L: columns 2L, 2L+1 of the ROUNDED values
                    const uint8_t* colp = slab + (lane & 3) * 4;
                    float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
                    for (int r = 0; r < rv; ++r) {
                        const unsigned w = *reinterpret_cast<const unsigned*>(colp + r * 128 + (((lane >> 2) ^ (r & 7)) << 4));
                        const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
                        a0 += x.x; b0 = fmaf(x.x, x.x, b0); a1 += x.y; b1 = fmaf(x.y, x.y, b1);
                    }
                    s_sum[q][c0 + 2 * lane] += a0; s_sq[q][c0 + 2 * lane] += b0;          // slots owned by this lane
                    s_sum[q][c0 + 2 * lane + 1] += a1; s_sq[q][c0 + 2 * lane + 1] += b1;
                }
                // coalesced stores: 4 rows x 8 chunks of 16 bytes per instruction
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + (lane >> 3), ch = lane & 7;
                    if (rr < rv) {
                        const uint4 u = *reinterpret_cast<const uint4*>(slab + rr * 128 + ((ch ^ (rr & 7)) << 4));
                        *reinterpret_cast<uint4*>(p.y + (base_row + row_l + rr) * p.N + n0 + c0 + ch * 8) = u;
                    }
                }
                __syncwarp();
            }
        }
        fence_before();
        if (p.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");            // the four epilogue warps only
            const int col = threadIdx.x - 64;
            if (col < p.BN) {
                const float cs = (s_sum[0][col] + s_sum[1][col]) + (s_sum[2][col] + s_sum[3][col]);
                const float cq = (s_sq[0][col] + s_sq[1][col]) + (s_sq[2][col] + s_sq[3][col]);
                nn::atomic_add_double(p.stats + n0 + col, (double)cs);
                nn::atomic_add_double(p.stats + p.N + n0 + col, (double)cq);
            }
        }
    }
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_d, tmem_cols); }
}

// ------------------------------------------------------------------------------------------------------- wgrad
struct WP {
    int F, H, W, Cin, Cout;
    int bh, bf, tpf;            // pixel block = box (64, W, bh, bf)
    int kr, krp;                // pixels per block, padded to a multiple of 16 (pad rows of the x tile stay zero)
    int nblocks, per_cta;       // pixel blocks in all, per CTA (grid.z splits them)
    int nch;                    // 64-column chunks of this N tile (<= 4): chunk q -> tap q / (Cin/64), channels (q % ..) * 64
    int stages, s2;
};

__global__ void __launch_bounds__(THREADS)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ CUtensorMap tmC, const WP p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 1];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;                      // output channels (rows of dW)
    const int cchunks = p.Cin >> 6;
    const int q0 = blockIdx.y * 4;                        // first 64-column chunk of this N tile
    const int nch = min(4, 9 * cchunks - q0);
    const int blk_lo = blockIdx.z * p.per_cta, blk_hi = min(p.nblocks, blk_lo + p.per_cta);
    const int nblk = max(0, blk_hi - blk_lo);
    const uint32_t chunk_bytes = (uint32_t)p.krp * 128u;
    const uint32_t stage_bytes = (uint32_t)(2 + nch) * chunk_bytes;
    const uint32_t tx_bytes = (uint32_t)(2 + nch) * (uint32_t)p.kr * 128u;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[MAX_STAGES]), tmem_full = smem_u32(&bars[2 * MAX_STAGES]);
    const int BN = nch * 64;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)BN) tmem_cols <<= 1;

    // rows kr .. krp of every chunk are never written by TMA: they must read as zeros
    {
        uint4* z = reinterpret_cast<uint4*>(smem_raw + (tiles - smem_u32(smem_raw)));
        const int n16 = (int)((uint32_t)p.stages * stage_bytes >> 4);
        for (int i = threadIdx.x; i < n16; i += THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nblk; ++i) {
                const int s = i % p.stages;
                mbar_wait(empty0 + 8 * s, (((uint32_t)(i / p.stages)) & 1u) ^ 1u);
                const int blk = blk_lo + i;
                int f0, h0;
                if (p.tpf) { f0 = blk / p.tpf; h0 = (blk - f0 * p.tpf) * p.bh; } else { f0 = blk * p.bf; h0 = 0; }
                const uint32_t a_dst = tiles + s * stage_bytes, b_dst = a_dst + 2u * chunk_bytes;
                mbar_expect_tx(full0 + 8 * s, tx_bytes);
                // dy through the SAME box shape as x (not as a run of rows): in a ragged block the pixels past the
                // frame are zero filled instead of being the next frame's first rows, which would meet valid x rows
                tma_load_4d(a_dst, &tmDy, full0 + 8 * s, m0, 0, h0, f0);
                tma_load_4d(a_dst + chunk_bytes, &tmDy, full0 + 8 * s, m0 + 64, 0, h0, f0);     // (zero filled past Cout)
                for (int j = 0; j < nch; ++j) {
                    const int qq = q0 + j, tap = qq / cchunks, cc = qq - tap * cchunks;
                    const int r = tap / 3, t = tap - r * 3;
                    if (!p.s2) tma_load_4d(b_dst + (uint32_t)j * chunk_bytes, &tmX, full0 + 8 * s, cc * 64, t - 1, h0 + r - 1, f0);
                    else tma_load_5d(b_dst + (uint32_t)j * chunk_bytes, &tmX, full0 + 8 * s, ((t - 1) & 1) * p.Cin + cc * 64,
                                     t == 0 ? -1 : 0, (r - 1) & 1, h0 + (r == 0 ? -1 : 0), f0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16(BN, true, true);
            const int k16 = p.krp >> 4;
            for (int i = 0; i < nblk; ++i) {
                const int s = i % p.stages;
                mbar_wait(full0 + 8 * s, ((uint32_t)(i / p.stages)) & 1u);
                fence_after();
                const uint32_t a_src = tiles + s * stage_bytes, b_src = a_src + 2u * chunk_bytes;
                const uint64_t adesc = desc_mn_sw128_b16(a_src, chunk_bytes), bdesc = desc_mn_sw128_b16(b_src, chunk_bytes);
                for (int k = 0; k < k16; ++k)                          // 16 pixels = 16 rows of 128 bytes = 128 address units
                    mma_bf16(tmem_d, adesc + 128ull * k, bdesc + 128ull * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
                mma_commit(empty0 + 8 * s);
            }
            mma_commit(tmem_full);
        }
    } else if (nblk > 0) {
        // fp32 tile -> swizzled 32 x 32 boxes -> TMA reduce-add into dW[Cout][9*Cin] (clips rows >= Cout, columns past the end)
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        fence_after();
        const uint32_t stg0 = tiles + (uint32_t)q * 8192u;
        uint8_t* stg0_g = smem_raw + (stg0 - smem_u32(smem_raw));
        int nbox = 0;
        for (int c0 = 0; c0 < BN; c0 += 32, ++nbox) {
            const uint32_t buf = (uint32_t)(nbox & 1) * 4096u;
            if (nbox >= 2) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }
            float v[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            uint8_t* rowp = stg0_g + buf + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { tma_reduce_add_2d(&tmC, stg0 + buf, q0 * 64 + c0, m0 + q * 32); bulk_commit(); }
        }
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        fence_before();
    }
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_d, tmem_cols); }
}

static int x_map(CUtensorMap* m, const void* x, int F, int H, int W, int C, int bh, int bf) {
    const long long dims[4] = {C, W, H, F};
    const long long strides[3] = {(long long)C * 2, (long long)W * C * 2, (long long)H * W * C * 2};
    const int box[4] = {64, W, bh, bf};
    return make_map(m, x, true, 4, dims, strides, box);
}
// input of a stride-2 window: ((pw, C), Wi / 2, ph, Hi / 2, F); box = (64, Wo, 1, bh, bf) on the OUTPUT grid (Hi, Wi even)
static int x_map_s2(CUtensorMap* m, const void* x, int F, int Hi, int Wi, int C, int bh, int bf) {
    const long long dims[5] = {2LL * C, Wi / 2, 2, Hi / 2, F};
    const long long strides[4] = {2LL * C * 2, (long long)Wi * C * 2, 2LL * Wi * C * 2, (long long)Hi * Wi * C * 2};
    const int box[5] = {64, Wi / 2, 1, bh, bf};
    return make_map(m, x, true, 5, dims, strides, box);
}

}  // namespace ig

#define LR_IG_CHECK(name)                                                                                     \
    LR_CHECK_ARG(F >= 0 && H > 0 && W > 0 && W <= 128, name ": bad image shape (W <= 128)");                  \
    if (F == 0) return LR_OK

// H, W: the OUTPUT grid (= the input grid for stride 1; the input is 2H x 2W for s2)
static int conv3x3_impl(const void* x, const void* wt, void* y, const void* R, double* stats, int F, int H, int W,
                        int Cin, int N, int flip, int s2, lr_stream_t stream) {
    LR_IG_CHECK("lr_conv3x3_bf16");
    LR_CHECK_ARG(Cin > 0 && (Cin & 63) == 0 && N > 0 && (N & 63) == 0, "lr_conv3x3_bf16: channels must be multiples of 64");
    LR_CHECK_ARG(x && wt && y, "lr_conv3x3_bf16: null pointer");
    LR_CHECK_ALIGN(x); LR_CHECK_ALIGN(wt); LR_CHECK_ALIGN(y); LR_CHECK_ALIGN(R);
    ig::FP p;
    p.F = F; p.H = H; p.W = W; p.Cin = Cin; p.N = N; p.flip = flip ? 1 : 0; p.s2 = s2;
    p.y = static_cast<nn::bf16*>(y); p.R = static_cast<const nn::bf16*>(R); p.stats = stats;
    int tiles;
    if (H * W <= 256) {
        p.bh = H; p.bf = 256 / (H * W); if (p.bf > F) p.bf = F; if (p.bf > 256) p.bf = 256;
        p.tpf = 0; tiles = (F + p.bf - 1) / p.bf;
    } else {
        p.bf = 1; p.bh = 256 / W; p.tpf = (H + p.bh - 1) / p.bh; tiles = F * p.tpf;
    }
    p.rows = W * p.bh * p.bf; p.nm = (p.rows + 127) / 128;
    p.BN = (N % 128 == 0) ? 128 : 64;
    const int num_kb = 9 * (Cin / 64);
    const size_t stage = (size_t)p.nm * 16384 + (size_t)p.BN * 128;
    int stages = (int)((200 * 1024 - 1024) / stage);
    if (stages > ig::MAX_STAGES) stages = ig::MAX_STAGES;
    {   // The kernel is one tile per CTA: load ramp, main loop and epilogue of a CTA are serial.  With more CTAs than SMs a
        // ring that lets TWO CTAs share an SM (one's epilogue under the other's main loop) beats a deeper ring for one:
        // measured (profiles/r2_conv_ring_depth.txt) 22x22x64: 131 -> 84 us, 11x11x128: 74 -> 53, 6x6x256: 57 -> 43;
        // the 3x3x512 layers have fewer CTAs than SMs and keep the deep ring (51 us; 63 with two stages).
        static const int cap_env = getenv("LIPREAD_CONV_STAGES") ? atoi(getenv("LIPREAD_CONV_STAGES")) : 0;
        int cap = cap_env;
        if (cap < 1 && (long long)tiles * (N / p.BN) > lr::sm_count()) {
            cap = (int)((113 * 1024 - 1024) / stage);
            if (cap < 2) cap = 2;
        }
        if (cap >= 1 && stages > cap) stages = cap;
    }
    if (stages > num_kb) stages = num_kb;
    p.stages = stages;
    CUtensorMap mx, mw;
    int rc = s2 ? ig::x_map_s2(&mx, x, F, 2 * H, 2 * W, Cin, p.bh, p.bf) : ig::x_map(&mx, x, F, H, W, Cin, p.bh, p.bf);
    if (rc) return rc;
    {
        const long long dims[2] = {9LL * Cin, N};
        const long long strides[1] = {9LL * Cin * 2};
        const int box[2] = {64, p.BN};
        rc = tcc::make_map(&mw, wt, true, 2, dims, strides, box);
        if (rc) return rc;
    }
    size_t smem = (size_t)stages * stage;
    if (smem < 16384) smem = 16384;                      // the epilogue's four 4 KB staging slabs
    smem += 1024;
    cudaError_t e = lr::ensure_max_dynamic_smem(ig::conv3x3_kernel, 200 * 1024);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "lr_conv3x3_bf16 smem: %s", cudaGetErrorString(e));
    dim3 grid((unsigned)tiles, (unsigned)(N / p.BN));
    ig::conv3x3_kernel<<<grid, ig::THREADS, smem, stream>>>(mx, mw, p);
    lr::count_launch();
    LR_CHECK_LAUNCH("conv3x3_kernel");
    return LR_OK;
}

extern "C" int lr_conv3x3_bf16(const void* x, const void* wt, void* y, const void* R, double* stats, int F, int H, int W,
                               int Cin, int N, int flip, lr_stream_t stream) {
    return conv3x3_impl(x, wt, y, R, stats, F, H, W, Cin, N, flip, 0, stream);
}
extern "C" int lr_conv3x3s2_bf16(const void* x, const void* wt, void* y, double* stats, int F, int Hi, int Wi, int Cin,
                                 int N, lr_stream_t stream) {
    LR_CHECK_ARG(Hi > 0 && Wi > 0 && (Hi & 1) == 0 && (Wi & 1) == 0, "lr_conv3x3s2_bf16: the input height and width must be even");
    return conv3x3_impl(x, wt, y, nullptr, stats, F, Hi / 2, Wi / 2, Cin, N, 0, 1, stream);
}

static int conv3x3_wgrad_impl(const void* dy, const void* x, float* dwp, int F, int H, int W, int Cin, int Cout, int s2,
                              lr_stream_t stream) {
    LR_IG_CHECK("lr_conv3x3_wgrad_bf16");
    LR_CHECK_ARG(Cin > 0 && (Cin & 63) == 0 && Cout > 0 && (Cout & 7) == 0, "lr_conv3x3_wgrad_bf16: Cin %% 64, Cout %% 8");
    LR_CHECK_ARG(dy && x && dwp, "lr_conv3x3_wgrad_bf16: null pointer");
    LR_CHECK_ALIGN(dy); LR_CHECK_ALIGN(x); LR_CHECK_ALIGN(dwp);
    ig::WP p;
    p.F = F; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.s2 = s2;
    if (H * W <= 128) {
        p.bh = H; p.bf = 128 / (H * W); if (p.bf > F) p.bf = F;
        p.tpf = 0; p.nblocks = (F + p.bf - 1) / p.bf;
    } else {
        p.bf = 1; p.bh = 128 / W; p.tpf = (H + p.bh - 1) / p.bh; p.nblocks = F * p.tpf;
    }
    p.kr = W * p.bh * p.bf; p.krp = (p.kr + 15) / 16 * 16;
    const int mtiles = (Cout + 127) / 128, ntiles = (9 * (Cin / 64) + 3) / 4;
    // split the pixel blocks so that about two waves of CTAs exist, each with at least 4 blocks to pipeline
    int splits = (2 * lr::sm_count() + mtiles * ntiles - 1) / (mtiles * ntiles);
    if (splits > (p.nblocks + 3) / 4) splits = (p.nblocks + 3) / 4;
    if (splits < 1) splits = 1;
    p.per_cta = (p.nblocks + splits - 1) / splits;
    splits = (p.nblocks + p.per_cta - 1) / p.per_cta;
    const size_t stage = (size_t)6 * p.krp * 128;        // sized for a full 4-chunk N tile
    int stages = (int)((200 * 1024 - 1024) / stage);
    if (stages > ig::MAX_STAGES) stages = ig::MAX_STAGES;
    {   // as in the forward kernel: a ring that lets two CTAs share an SM (config 2: 7.13 -> 7.04 ms with it)
        static const int cap_env = getenv("LIPREAD_WGRAD_STAGES") ? atoi(getenv("LIPREAD_WGRAD_STAGES")) : 0;
        int cap = cap_env;
        if (cap < 1 && (long long)mtiles * ntiles * splits > lr::sm_count()) {
            cap = (int)((113 * 1024 - 1024) / stage);
            if (cap < 1) cap = 1;
        }
        if (cap >= 1 && stages > cap) stages = cap;
    }
    if (stages > p.per_cta) stages = p.per_cta;
    if (stages < 1) stages = 1;
    p.stages = stages;
    p.nch = 4;
    CUtensorMap md, mx, mc;
    int rc = ig::x_map(&md, dy, F, H, W, Cout, p.bh, p.bf);
    if (rc) return rc;
    rc = s2 ? ig::x_map_s2(&mx, x, F, 2 * H, 2 * W, Cin, p.bh, p.bf) : ig::x_map(&mx, x, F, H, W, Cin, p.bh, p.bf);
    if (rc) return rc;
    {
        const long long dims[2] = {9LL * Cin, Cout};
        const long long strides[1] = {9LL * Cin * 4};
        const int box[2] = {32, 32};
        rc = tcc::make_map(&mc, dwp, false, 2, dims, strides, box);
        if (rc) return rc;
    }
    size_t smem = (size_t)stages * stage;
    if (smem < 32768) smem = 32768;                      // four warps x two 4 KB fp32 boxes
    smem += 1024;
    cudaError_t e = lr::ensure_max_dynamic_smem(ig::conv3x3_wgrad_kernel, 200 * 1024);
    if (e != cudaSuccess) return lr::fail(LR_ECUDA, "lr_conv3x3_wgrad_bf16 smem: %s", cudaGetErrorString(e));
    dim3 grid((unsigned)mtiles, (unsigned)ntiles, (unsigned)splits);
    ig::conv3x3_wgrad_kernel<<<grid, ig::THREADS, smem, stream>>>(md, mx, mc, p);
    lr::count_launch();
    LR_CHECK_LAUNCH("conv3x3_wgrad_kernel");
    return LR_OK;
}

extern "C" int lr_conv3x3_wgrad_bf16(const void* dy, const void* x, float* dwp, int F, int H, int W, int Cin, int Cout,
                                     lr_stream_t stream) {
    return conv3x3_wgrad_impl(dy, x, dwp, F, H, W, Cin, Cout, 0, stream);
}
extern "C" int lr_conv3x3s2_wgrad_bf16(const void* dy, const void* x, float* dwp, int F, int Hi, int Wi, int Cin, int Cout,
                                       lr_stream_t stream) {
    LR_CHECK_ARG(Hi > 0 && Wi > 0 && (Hi & 1) == 0 && (Wi & 1) == 0, "lr_conv3x3s2_wgrad_bf16: the input height and width must be even");
    return conv3x3_wgrad_impl(dy, x, dwp, F, Hi / 2, Wi / 2, Cin, Cout, 1, stream);
}
