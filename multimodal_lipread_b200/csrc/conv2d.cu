// Dense (groups == 1) k x k convolutions as GEMMs on an explicit patch matrix, plus the pooling / dropout
// kernels of the ResNet-18 trunk and the early-fusion audio encoder (channels-last fp32).
//
//   forward   col = im2col(x)            [F*Ho*Wo, Cin*kh*kw]   (column order (c, r, s) == torch's weight layout,
//             y   = col . W^T            lr_gemm_tf32 (tcgen05) with W[Cout][Cin*kh*kw] used as it lies in HBM)
//   wgrad     dW += dy^T . col           (same col, kept from the forward)
//   dgrad     colT = im2col_T(dy)        [F*H*W, Cout*kh*kw]: the gather form of the transposed convolution
//             dx   = colT . Wt^T         with Wt[Cin][Cout*kh*kw] = W transposed over its first two dims
// Reference call sites: torchvision resnet18 (video/models/resnet_lstm.py:79-110, audio/models/resnet_model.py:12-17,
// audio_cues_video/models/late_fusion_mobile.py:57-66), AudioEncoder (audio_video/models/early_fusion.py:21-35),
// torchvision mobilenet_v2 stem (audio_cues_video/models/late_fusion_mobile.py:33-40).
#include "nn_common.cuh"

namespace c2 {

constexpr int TH = 256;

struct Src {
    const void* x;
    int is_u8; float scale;
    int T; long long sb, st, sc, sh, sw;     // element (f, c, h, w) at (f/T)*sb + (f%T)*st + c*sc + h*sh + w*sw
    int Hs, Ws, C;                           // source frame geometry
};
struct Geo {
    int Hd, Wd;                              // destination rows run over (f, hd, wd)
    int kh, kw, stride, pad, transposed;     // pad: rows (and columns unless pad_w >= 0)
    int pad_w;                               // column padding of the tap-major kernel (Conv1d as a 1 x k window)
    int K, ldk;                              // K = C*kh*kw valid columns, ldk = row pitch (multiple of 4), tail zeroed
};

__device__ __forceinline__ float src_ld(const Src& s, long long fb, int c, int h, int w) {
    const long long off = fb + (long long)c * s.sc + (long long)h * s.sh + (long long)w * s.sw;
    return s.is_u8 ? lr::u8_scaled(static_cast<const unsigned char*>(s.x)[off], s.scale)
                   : static_cast<const float*>(s.x)[off] * s.scale;
}

// one thread -> 4 consecutive columns of one row (coalesced float4 store; the gathers hit L1)
template <typename T>
__global__ void __launch_bounds__(TH)
im2col_kernel(const Src s, const Geo g, T* __restrict__ col, long long rows) {
    const int q4 = g.ldk >> 2;
    const long long total = rows * q4;
    const int kk = g.kh * g.kw;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < total; i += (long long)gridDim.x * TH) {
        const long long row = i / q4;
        const int k0 = int(i - row * q4) * 4;
        const int wd = int(row % g.Wd);
        const long long t = row / g.Wd;
        const int hd = int(t % g.Hd), f = int(t / g.Hd);
        const long long fb = (long long)(f / s.T) * s.sb + (long long)(f % s.T) * s.st;
        int c = k0 / kk, rs = k0 - c * kk;
        int r = rs / g.kw, q = rs - r * g.kw;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float val = 0.f;
            if (k0 + j < g.K) {
                int hs, ws; bool ok;
                if (!g.transposed) {
                    hs = hd * g.stride - g.pad + r; ws = wd * g.stride - g.pad + q;
                    ok = hs >= 0 && hs < s.Hs && ws >= 0 && ws < s.Ws;
                } else {
                    const int hn = hd + g.pad - r, wn = wd + g.pad - q;
                    hs = hn / g.stride; ws = wn / g.stride;
                    ok = hn >= 0 && wn >= 0 && hs * g.stride == hn && ws * g.stride == wn && hs < s.Hs && ws < s.Ws;
                }
                if (ok) val = src_ld(s, fb, c, hs, ws);
            }
            v[j] = val;
            if (++q == g.kw) { q = 0; if (++r == g.kh) { r = 0; ++c; } }
        }
        nn::st4(col + row * g.ldk + k0, make_float4(v[0], v[1], v[2], v[3]));
    }
}

// Fast path of the ResNet stem on raw uint8 lip frames: 3 channels, 7x7 window, stride 2, padding 3, row pitch 152
// (147 valid columns, torch order k = c*49 + r*7 + q).  One block per (frame, output row): the 3 x 7 source rows the
// row needs are staged once in shared memory as floats (zero outside the image), then thread = (16-byte chunk j of the
// patch row, pixel lane) copies 8 consecutive columns per pixel from offsets it computed ONCE -- no per-element index
// arithmetic, 16-byte stores, 19 consecutive chunks per pixel = one contiguous 304-byte row.  (The generic kernel
// spent a chain of 64-bit divisions per 4 columns: 1 080 us for the 1.8 M-pixel stem of the benchmark; this one is
// bound by the 546 MB it writes.)
constexpr int S7_W = 96;                                   // staged row pitch (floats) >= 2*(Wd-1) + 7, Wd <= 45
template <typename T>
__global__ void __launch_bounds__(TH)
im2col_u8c3_7x7s2_kernel(const Src s, const Geo g, T* __restrict__ col, int rows_fh) {
    __shared__ float S[3 * 7 * S7_W];
    constexpr int NCH = 152 / 8;                           // 19 chunks of 8 columns
    const int j = threadIdx.x % NCH, pl = threadIdx.x / NCH;       // pixel lanes: 13 (247 of 256 threads work)
    int off[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = 8 * j + e;
        const int c = k / 49, rs = k - c * 49, r = rs / 7, q = rs - r * 7;
        off[e] = k < 147 ? (c * 7 + r) * S7_W + q : -1;
    }
    const unsigned char* xs = static_cast<const unsigned char*>(s.x);
    for (int fh = blockIdx.x; fh < rows_fh; fh += gridDim.x) {
        const int f = fh / g.Hd, hd = fh - f * g.Hd;
        const long long fb = (long long)(f / s.T) * s.sb + (long long)(f % s.T) * s.st;
        __syncthreads();                                   // the previous row's readers are done
        for (int i = threadIdx.x; i < 3 * 7 * S7_W; i += TH) {
            const int cr = i / S7_W, xw = i - cr * S7_W, c = cr / 7, r = cr - c * 7;
            const int hs = 2 * hd - 3 + r, ws = xw - 3;
            float v = 0.f;
            if (hs >= 0 && hs < s.Hs && ws >= 0 && ws < s.Ws)
                v = lr::u8_scaled(xs[fb + (long long)c * s.sc + (long long)hs * s.sh + (long long)ws * s.sw], s.scale);
            S[i] = v;
        }
        __syncthreads();
        if (pl < TH / NCH) {
            T* const orow = col + (long long)fh * g.Wd * 152 + 8 * j;
            for (int wd = pl; wd < g.Wd; wd += TH / NCH) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = off[e] >= 0 ? S[off[e] + 2 * wd] : 0.f;
                nn::st4(orow + (long long)wd * 152, make_float4(v[0], v[1], v[2], v[3]));
                nn::st4(orow + (long long)wd * 152 + 4, make_float4(v[4], v[5], v[6], v[7]));
            }
        }
    }
}

// Fast path of the stems on raw uint8 lip frames (3 channels, 3x3 window, any stride / padding): one thread per
// output pixel reads its 3 x 3 x 3 byte patch (L1 serves the overlap with the neighbours) and writes its whole
// 28-float row (7 coalesced float4 stores) -- no per-element index arithmetic.
template <typename T>
__global__ void __launch_bounds__(TH)
im2col_u8c3_3x3_kernel(const Src s, const Geo g, T* __restrict__ col, long long rows) {
    const unsigned char* xb = static_cast<const unsigned char*>(s.x);
    for (long long row = (long long)blockIdx.x * TH + threadIdx.x; row < rows; row += (long long)gridDim.x * TH) {
        const int wd = int(row % g.Wd);
        const long long t = row / g.Wd;
        const int hd = int(t % g.Hd), f = int(t / g.Hd);
        const long long fb = (long long)(f / s.T) * s.sb + (long long)(f % s.T) * s.st;
        float v[28];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int hs = hd * g.stride - g.pad + r;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int ws = wd * g.stride - g.pad + q;
                const bool ok = hs >= 0 && hs < s.Hs && ws >= 0 && ws < s.Ws;
                const unsigned char* px = xb + fb + (long long)hs * s.sh + (long long)ws * s.sw;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c * 9 + r * 3 + q] = ok ? lr::u8_scaled(px[c * s.sc], s.scale) : 0.f;
            }
        }
        v[27] = 0.f;
        T* o = col + row * g.ldk;
#pragma unroll
        for (int j = 0; j < 28; j += 4) nn::st4(o + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        for (int j = 28; j < g.ldk; j += 4) nn::st4(o + j, make_float4(0.f, 0.f, 0.f, 0.f));   // bf16 rows are padded to 32
    }
}

// Tap-major patch matrix of a channels-last float activation with C % 4 == 0:
//   col[(f,hd,wd)][(r*kw + s)*C + c] = x[f, hs, ws, c]   (0 outside the image)
// A thread moves one float4 of channels: the gather is a shifted, fully coalesced copy (both sides), which is what
// lets the ResNet trunk's patch matrices be written at HBM speed.  transposed: the dgrad operand (see lr_im2col).
template <typename T>
__global__ void __launch_bounds__(TH)
im2col_tap_kernel(const T* __restrict__ x, const Geo g, int Hs, int Ws, int C, T* __restrict__ col, long long rows) {
    // one thread = one (row, 4-channel group): the pixel decode is done once and reused for all kh*kw taps
    const int c4n = C >> 2;
    const long long total = rows * c4n;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < total; i += (long long)gridDim.x * TH) {
        const long long row = i / c4n;
        const int c = int(i - row * c4n) * 4;
        const int wd = int(row % g.Wd);
        const long long t = row / g.Wd;
        const int hd = int(t % g.Hd), f = int(t / g.Hd);
        const T* xf = x + (long long)f * Hs * Ws * C + c;
        T* o = col + row * g.ldk + c;
        for (int r = 0; r < g.kh; ++r) {
            int hs; bool okh;
            if (!g.transposed) { hs = hd * g.stride - g.pad + r; okh = hs >= 0 && hs < Hs; }
            else { const int hn = hd + g.pad - r; hs = hn / g.stride; okh = hn >= 0 && hs * g.stride == hn && hs < Hs; }
            for (int q = 0; q < g.kw; ++q) {
                int ws; bool ok;
                if (!g.transposed) { ws = wd * g.stride - g.pad_w + q; ok = okh && ws >= 0 && ws < Ws; }
                else { const int wn = wd + g.pad_w - q; ws = wn / g.stride; ok = okh && wn >= 0 && ws * g.stride == wn && ws < Ws; }
                const float4 v = ok ? nn::ld4(xf + ((long long)hs * Ws + ws) * C) : make_float4(0.f, 0.f, 0.f, 0.f);
                nn::st4(o + (long long)(r * g.kw + q) * C, v);
            }
        }
    }
}

// The shapes the ResNet trunks use (3x3 and 1x1 windows) with 16-byte accesses (8 bf16 / 4 fp32 channels per thread),
// the window unrolled at compile time and ALL loads of a thread issued before its stores: the generic kernel above ran
// the stride-2 patch matrices of the benchmark at 1.0-1.5 TB/s (one 8-byte load in flight per thread).
template <typename T, int KH, int KW, bool TR>
__global__ void __launch_bounds__(TH)
im2col_tap_vec_kernel(const T* __restrict__ x, const Geo g, int Hs, int Ws, int C, T* __restrict__ col, long long rows) {
    constexpr int V = 16 / (int)sizeof(T);
    const int cvn = C / V;
    const long long total = rows * cvn;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < total; i += (long long)gridDim.x * TH) {
        const long long row = i / cvn;
        const int c = int(i - row * cvn) * V;
        const int wd = int(row % g.Wd);
        const long long t = row / g.Wd;
        const int hd = int(t % g.Hd), f = int(t / g.Hd);
        const T* xf = x + (long long)f * Hs * Ws * C + c;
        uint4 v[KH * KW];
#pragma unroll
        for (int r = 0; r < KH; ++r) {
            int hs; bool okh;
            if (!TR) { hs = hd * g.stride - g.pad + r; okh = hs >= 0 && hs < Hs; }
            else { const int hn = hd + g.pad - r; hs = hn / g.stride; okh = hn >= 0 && hs * g.stride == hn && hs < Hs; }
#pragma unroll
            for (int q = 0; q < KW; ++q) {
                int ws; bool ok;
                if (!TR) { ws = wd * g.stride - g.pad_w + q; ok = okh && ws >= 0 && ws < Ws; }
                else { const int wn = wd + g.pad_w - q; ws = wn / g.stride; ok = okh && wn >= 0 && ws * g.stride == wn && ws < Ws; }
                v[r * KW + q] = ok ? __ldg(reinterpret_cast<const uint4*>(xf + ((long long)hs * Ws + ws) * C)) : make_uint4(0u, 0u, 0u, 0u);
            }
        }
        T* o = col + row * g.ldk + c;
#pragma unroll
        for (int j = 0; j < KH * KW; ++j) *reinterpret_cast<uint4*>(o + (long long)j * C) = v[j];
    }
}

// Weight layouts of the tap-major patch matrix (kk = kh*kw taps):
//   mode 0: wp[k][rs][c]  = w[k][c][rs]    forward / wgrad operand   [Cout][kk*Cin]
//   mode 1: wp[c][rs][k]  = w[k][c][rs]    dgrad operand             [Cin][kk*Cout]
//   mode 2: w[k][c][rs]   = wp[k][rs][c]   weight gradient back to torch's layout (overwrites w)
template <typename TD>
__global__ void __launch_bounds__(TH)
weight_tap_kernel(const float* __restrict__ src, TD* __restrict__ dst, int Cout, int Cin, int kk, int mode) {
    const long long n = (long long)Cout * Cin * kk;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < n; i += (long long)gridDim.x * TH) {
        // i enumerates the DESTINATION contiguously
        if (mode == 0) {
            const int c = int(i % Cin); const long long t = i / Cin; const int rs = int(t % kk), k = int(t / kk);
            nn::st1(dst + i, src[((long long)k * Cin + c) * kk + rs]);
        } else if (mode == 1) {
            const int k = int(i % Cout); const long long t = i / Cout; const int rs = int(t % kk), c = int(t / kk);
            nn::st1(dst + i, src[((long long)k * Cin + c) * kk + rs]);
        } else {
            const int rs = int(i % kk); const long long t = i / kk; const int c = int(t % Cin), k = int(t / Cin);
            nn::st1(dst + i, src[((long long)k * kk + rs) * Cin + c]);
        }
    }
}

// Every tap-major bf16 operand of a model in ONE launch (modes 0 and 1 of weight_tap_kernel): the weights only change in
// the optimizer step, so all re-layouts of a training step run up front from a device table -- 38 launches of ~9 us
// each in the ResNet-18 plan become one.  blockIdx.y = table entry.
struct WTapEntry { const float* src; nn::bf16* dst; int Cout, Cin, kk, mode; };
static_assert(sizeof(WTapEntry) == 32, "the host packs an entry as four 64-bit words");
__global__ void __launch_bounds__(TH)
weight_tap_batch_kernel(const WTapEntry* __restrict__ tab) {
    const WTapEntry e = tab[blockIdx.y];
    const int n = e.Cout * e.Cin * e.kk;
    for (int i = blockIdx.x * TH + threadIdx.x; i < n; i += gridDim.x * TH) {
        int k, c, rs;
        if (e.mode == 0) { c = i % e.Cin; const int t = i / e.Cin; rs = t % e.kk; k = t / e.kk; }
        else { k = i % e.Cout; const int t = i / e.Cout; rs = t % e.kk; c = t / e.kk; }
        e.dst[i] = __float2bfloat16_rn(e.src[(k * e.Cin + c) * e.kk + rs]);
    }
}

// wt[c][k][rs] = w[k][c][rs]  (row pitch of wt = ldt >= Cout*kk, tail zeroed by the caller once)
__global__ void __launch_bounds__(TH)
weight_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int Cin, int kk, long long ldt) {
    const long long n = (long long)Cout * Cin * kk;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < n; i += (long long)gridDim.x * TH) {
        const int rs = int(i % kk);
        const long long t = i / kk;
        const int k = int(t % Cout), c = int(t / Cout);
        wt[(long long)c * ldt + (long long)k * kk + rs] = w[((long long)k * Cin + c) * kk + rs];
    }
}

// ------------------------------------------------------------------------------------------ max pooling
// y[f,ho,wo,c] = max over the window (padding = -inf), arg = r*k + s of the FIRST maximum in scan order
// (torch.nn.MaxPool2d tie rule: `val > maxval`).
template <typename T>
__global__ void __launch_bounds__(TH)
maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, unsigned char* __restrict__ arg, int F, int H,
                   int W, int C, int k, int stride, int pad, int Ho, int Wo) {
    const int c4n = C >> 2;
    const long long total = (long long)F * Ho * Wo * c4n;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < total; i += (long long)gridDim.x * TH) {
        const int c = int(i % c4n) * 4;
        long long t = i / c4n;
        const int wo = int(t % Wo); t /= Wo;
        const int ho = int(t % Ho), f = int(t / Ho);
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        unsigned char a[4] = {0, 0, 0, 0};
        for (int r = 0; r < k; ++r) {
            const int h = ho * stride - pad + r;
            if (h < 0 || h >= H) continue;
            for (int s = 0; s < k; ++s) {
                const int w = wo * stride - pad + s;
                if (w < 0 || w >= W) continue;
                const float4 v = nn::ld4(x + (((long long)f * H + h) * W + w) * C + c);
                const unsigned char code = (unsigned char)(r * k + s);
                if (v.x > m[0] || v.x != v.x) { m[0] = v.x; a[0] = code; }
                if (v.y > m[1] || v.y != v.y) { m[1] = v.y; a[1] = code; }
                if (v.z > m[2] || v.z != v.z) { m[2] = v.z; a[2] = code; }
                if (v.w > m[3] || v.w != v.w) { m[3] = v.w; a[3] = code; }
            }
        }
        const long long o = (((long long)f * Ho + ho) * Wo + wo) * C + c;
        nn::st4(y + o, make_float4(m[0], m[1], m[2], m[3]));
        *reinterpret_cast<uchar4*>(arg + o) = make_uchar4(a[0], a[1], a[2], a[3]);
    }
}

// dx[f,h,w,c] = sum of dy over the windows whose saved arg-max is (h, w)   (gather form: no atomics)
// One block per (frame, input row): the window rows that can hold it are block constants, a thread owns 4 channels of
// one pixel of the row (all index arithmetic is 32-bit and per row, not a chain of 64-bit divisions per element).
template <typename T>
__global__ void __launch_bounds__(TH)
maxpool_bwd_kernel(const T* __restrict__ dy, const unsigned char* __restrict__ arg, T* __restrict__ dx, int F,
                   int H, int W, int C, int k, int stride, int pad, int Ho, int Wo) {
    const int c4n = C >> 2;
    const int items = W * c4n;
    for (int fh = blockIdx.x; fh < F * H; fh += gridDim.x) {
        const int f = fh / H, h = fh - f * H;
        const int ho_lo = max(0, (h + pad - k + stride) / stride), ho_hi = min(Ho - 1, (h + pad) / stride);
        const long long obase = (long long)f * Ho * Wo * C;
        T* const drow = dx + (long long)fh * W * C;
        for (int it = threadIdx.x; it < items; it += TH) {
            const int w = it / c4n, c = (it - w * c4n) * 4;
            const int wo_lo = max(0, (w + pad - k + stride) / stride), wo_hi = min(Wo - 1, (w + pad) / stride);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int ho = ho_lo; ho <= ho_hi; ++ho) {
                const int r = h + pad - ho * stride;
                if (r < 0 || r >= k) continue;
                for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                    const int sft = w + pad - wo * stride;
                    if (sft < 0 || sft >= k) continue;
                    const long long o = obase + (long long)(ho * Wo + wo) * C + c;
                    const uchar4 a = *reinterpret_cast<const uchar4*>(arg + o);
                    const float4 g = nn::ld4(dy + o);
                    const unsigned char code = (unsigned char)(r * k + sft);
                    if (a.x == code) acc.x += g.x;
                    if (a.y == code) acc.y += g.y;
                    if (a.z == code) acc.z += g.z;
                    if (a.w == code) acc.w += g.w;
                }
            }
            nn::st4(drow + (long long)w * C + c, acc);
        }
    }
}

// ---- the ResNet stem pool (MaxPool2d(3, 2, 1)) with 16-byte accesses: a thread owns V channels of one pixel (V = 4
// fp32 / 8 bf16), a block walks (frame, row) pairs so the index arithmetic is 32-bit and per row, and the window is
// unrolled at compile time.  The generic kernels above moved 8 bytes per thread behind a chain of 64-bit divisions:
// 201 / 382 us for the 928 x 44 x 44 x 64 map of the benchmark (0.23 / 0.13 of HBM).
template <typename T> struct PoolVec;
template <> struct PoolVec<float> {
    static constexpr int V = 4;
    typedef uchar4 Arg;
    __device__ static void load(const float* p, float (&v)[4]) { const float4 t = nn::ld4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ static void store(float* p, const float (&v)[4]) { nn::st4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <> struct PoolVec<nn::bf16> {
    static constexpr int V = 8;
    typedef uint2 Arg;
    __device__ static void load(const nn::bf16* p, float (&v)[8]) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const float4 a = nn::unpack4(make_uint2(t.x, t.y)), b = nn::unpack4(make_uint2(t.z, t.w));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ static void store(nn::bf16* p, const float (&v)[8]) {
        const uint2 a = nn::pack4(make_float4(v[0], v[1], v[2], v[3])), b = nn::pack4(make_float4(v[4], v[5], v[6], v[7]));
        *reinterpret_cast<uint4*>(p) = make_uint4(a.x, a.y, b.x, b.y);
    }
};

template <typename T, int K, int S, int P>
__global__ void __launch_bounds__(TH)
maxpool_fwd_vec_kernel(const T* __restrict__ x, T* __restrict__ y, unsigned char* __restrict__ arg, int F, int H,
                       int W, int C, int Ho, int Wo) {
    constexpr int V = PoolVec<T>::V;
    const int cvn = C / V, items = Wo * cvn;
    for (int fo = blockIdx.x; fo < F * Ho; fo += gridDim.x) {
        const int f = fo / Ho, ho = fo - f * Ho;
        const T* const xf = x + (long long)f * H * W * C;
        for (int it = threadIdx.x; it < items; it += TH) {
            const int wo = it / cvn, c = (it - wo * cvn) * V;
            float m[V];
            unsigned a[V];
#pragma unroll
            for (int i = 0; i < V; ++i) { m[i] = -INFINITY; a[i] = 0; }
#pragma unroll
            for (int r = 0; r < K; ++r) {
                const int h = ho * S - P + r;
                if (h < 0 || h >= H) continue;
#pragma unroll
                for (int q = 0; q < K; ++q) {
                    const int w = wo * S - P + q;
                    if (w < 0 || w >= W) continue;
                    float v[V];
                    PoolVec<T>::load(xf + ((long long)h * W + w) * C + c, v);
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        if (v[i] > m[i] || v[i] != v[i]) { m[i] = v[i]; a[i] = (unsigned)(r * K + q); }
                }
            }
            const long long o = ((long long)fo * Wo + wo) * C + c;
            PoolVec<T>::store(y + o, m);
            unsigned pk[V / 4];
#pragma unroll
            for (int i = 0; i < V / 4; ++i) pk[i] = a[4 * i] | (a[4 * i + 1] << 8) | (a[4 * i + 2] << 16) | (a[4 * i + 3] << 24);
            if (V == 4) *reinterpret_cast<unsigned*>(arg + o) = pk[0];
            else *reinterpret_cast<uint2*>(arg + o) = make_uint2(pk[0], pk[V / 4 - 1]);
        }
    }
}

template <typename T, int K, int S, int P>
__global__ void __launch_bounds__(TH)
maxpool_bwd_vec_kernel(const T* __restrict__ dy, const unsigned char* __restrict__ arg, T* __restrict__ dx, int F,
                       int H, int W, int C, int Ho, int Wo) {
    constexpr int V = PoolVec<T>::V;
    constexpr int NW = (K + S - 1) / S;                 // windows that can cover a pixel, per axis
    const int cvn = C / V, items = W * cvn;
    for (int fh = blockIdx.x; fh < F * H; fh += gridDim.x) {
        const int f = fh / H, h = fh - f * H;
        const long long obase = (long long)f * Ho * Wo * C;
        T* const drow = dx + (long long)fh * W * C;
        const int ho_hi = (h + P) / S;
        for (int it = threadIdx.x; it < items; it += TH) {
            const int w = it / cvn, c = (it - w * cvn) * V;
            const int wo_hi = (w + P) / S;
            float acc[V];
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
            for (int dh = NW - 1; dh >= 0; --dh) {              // ascending ho, wo: the generic kernel's summation order
                const int ho = ho_hi - dh, r = h + P - ho * S;
                if (ho < 0 || ho >= Ho || r >= K) continue;
#pragma unroll
                for (int dw_ = NW - 1; dw_ >= 0; --dw_) {
                    const int wo = wo_hi - dw_, q = w + P - wo * S;
                    if (wo < 0 || wo >= Wo || q >= K) continue;
                    const long long o = obase + (long long)(ho * Wo + wo) * C + c;
                    float g[V];
                    PoolVec<T>::load(dy + o, g);
                    unsigned a[V];
                    if (V == 4) { const uchar4 t = *reinterpret_cast<const uchar4*>(arg + o); a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; }
                    else {
                        const uint2 t = *reinterpret_cast<const uint2*>(arg + o);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { a[i] = (t.x >> (8 * i)) & 255u; a[(V - 4) + i] = (t.y >> (8 * i)) & 255u; }
                    }
                    const unsigned code = (unsigned)(r * K + q);
#pragma unroll
                    for (int i = 0; i < V; ++i) if (a[i] == code) acc[i] += g[i];
                }
            }
            PoolVec<T>::store(drow + (long long)w * C + c, acc);
        }
    }
}

// ------------------------------------------------------------------------------------------ zero stuffing
// up[f, 2 ho, 2 wo, :] = dy[f, ho, wo, :], zero elsewhere, on the INPUT grid [F, Hi, Wi, C] of a stride-2 convolution.
// The input gradient of a 3x3 / stride-2 / pad-1 convolution is the stride-1 correlation of this map with the mirrored
// taps -- i.e. exactly what the implicit-GEMM kernel computes with flip = 1 -- so the transposed patch matrix
// ([input pixels][9 Cout], three quarters of it zeros: 1 GB for the first stride-2 block of the benchmark) is never
// written.  One 16-byte store per thread (8 bf16 channels).
__global__ void __launch_bounds__(TH)
zero_stuff2_kernel(const uint4* __restrict__ dy, uint4* __restrict__ up, int F, int Ho, int Wo, int Hi, int Wi, int c8n) {
    const long long total = (long long)F * Hi * Wi * c8n;
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < total; i += (long long)gridDim.x * TH) {
        const int c = (int)(i % c8n);
        long long t = i / c8n;
        const int w = (int)(t % Wi); t /= Wi;
        const int h = (int)(t % Hi), f = (int)(t / Hi);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (!((h | w) & 1) && (h >> 1) < Ho && (w >> 1) < Wo)
            v = dy[(((long long)f * Ho + (h >> 1)) * Wo + (w >> 1)) * c8n + c];
        up[i] = v;
    }
}

// ------------------------------------------------------------------------------------------ dropout
// Counter-based generator: keep(i) = hash(seed, step, i) >= p.  `step` lives on the device and is advanced by
// lr_rng_tick once per training step, so a captured CUDA graph draws a fresh mask on every replay.
__device__ __forceinline__ uint32_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((x ^ (x >> 31)) >> 32);
}
__global__ void __launch_bounds__(TH)
dropout_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned char* __restrict__ mask, long long n,
                   float p, unsigned long long seed, const long long* __restrict__ step) {
    const unsigned long long base = seed * 0xD1342543DE82EF95ull + (unsigned long long)(*step) * 0xA0761D6478BD642Full;
    const float inv_keep = 1.f / (1.f - p);
    const uint32_t thr = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < n; i += (long long)gridDim.x * TH) {
        const bool keep = mix64(base + (unsigned long long)i) >= thr;
        mask[i] = keep ? 1 : 0;
        y[i] = keep ? x[i] * inv_keep : 0.f;
    }
}
__global__ void __launch_bounds__(TH)
dropout_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ mask, float* __restrict__ dx,
                   long long n, float inv_keep) {
    for (long long i = (long long)blockIdx.x * TH + threadIdx.x; i < n; i += (long long)gridDim.x * TH)
        dx[i] = mask[i] ? dy[i] * inv_keep : 0.f;
}
__global__ void rng_tick_kernel(long long* step) { *step += 1; }

static unsigned grid_for(long long work) {
    long long g = (work + TH - 1) / TH;
    const long long cap = (long long)lr::sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace c2

template <typename T>
static int im2col_impl(const void* x, int is_u8, float scale, int F, int T_, long long sb, long long st, long long sc,
                       long long sh, long long sw, int Hs, int Ws, int C, int kh, int kw, int stride, int pad,
                       int transposed, int Hd, int Wd, T* col, long long ldk, lr_stream_t stream) {
    LR_CHECK_ARG(F >= 0 && T_ > 0 && Hs > 0 && Ws > 0 && C > 0 && Hd > 0 && Wd > 0, "lr_im2col: bad shape");
    LR_CHECK_ARG(kh > 0 && kw > 0 && stride > 0 && pad >= 0, "lr_im2col: bad window");
    const long long K = (long long)C * kh * kw;
    LR_CHECK_ARG(ldk >= K && (ldk & 3) == 0, "lr_im2col: ldk must be >= C*kh*kw and a multiple of 4");
    if (F == 0) return LR_OK;
    LR_CHECK_ARG(x && col, "lr_im2col: null pointer");
    LR_CHECK_ALIGN(col);
    c2::Src s;
    s.x = x; s.is_u8 = is_u8; s.scale = scale; s.T = T_; s.sb = sb; s.st = st; s.sc = sc; s.sh = sh; s.sw = sw;
    s.Hs = Hs; s.Ws = Ws; s.C = C;
    c2::Geo g;
    g.Hd = Hd; g.Wd = Wd; g.kh = kh; g.kw = kw; g.stride = stride; g.pad = pad; g.transposed = transposed;
    g.pad_w = pad;
    g.K = (int)K; g.ldk = (int)ldk;
    const long long rows = (long long)F * Hd * Wd;
    if (is_u8 && C == 3 && kh == 7 && kw == 7 && stride == 2 && pad == 3 && !transposed && ldk == 152 &&
        2 * (Wd - 1) + 7 <= c2::S7_W) {
        const long long fh = (long long)F * Hd;
        const int grid = (int)(fh < 16LL * lr::sm_count() ? fh : 16LL * lr::sm_count());
        c2::im2col_u8c3_7x7s2_kernel<T><<<grid, c2::TH, 0, stream>>>(s, g, col, (int)fh);
    } else if (is_u8 && C == 3 && kh == 3 && kw == 3 && !transposed && ldk >= 28 && ldk <= 32)
        c2::im2col_u8c3_3x3_kernel<T><<<c2::grid_for(rows), c2::TH, 0, stream>>>(s, g, col, rows);
    else
        c2::im2col_kernel<T><<<c2::grid_for(rows * (ldk >> 2)), c2::TH, 0, stream>>>(s, g, col, rows);
    lr::count_launch();
    LR_CHECK_LAUNCH("im2col_kernel");
    return LR_OK;
}
extern "C" int lr_im2col(const void* x, int is_u8, float scale, int F, int T, long long sb, long long st, long long sc,
                         long long sh, long long sw, int Hs, int Ws, int C, int kh, int kw, int stride, int pad,
                         int transposed, int Hd, int Wd, float* col, long long ldk, lr_stream_t stream) {
    return im2col_impl<float>(x, is_u8, scale, F, T, sb, st, sc, sh, sw, Hs, Ws, C, kh, kw, stride, pad, transposed, Hd, Wd,
                              col, ldk, stream);
}
/* bf16 patch matrix (precision "bf16"): the source is still the caller's uint8 / float frames */
extern "C" int lr_im2col_h(const void* x, int is_u8, float scale, int F, int T, long long sb, long long st, long long sc,
                           long long sh, long long sw, int Hs, int Ws, int C, int kh, int kw, int stride, int pad,
                           int transposed, int Hd, int Wd, void* col, long long ldk, lr_stream_t stream) {
    return im2col_impl<nn::bf16>(x, is_u8, scale, F, T, sb, st, sc, sh, sw, Hs, Ws, C, kh, kw, stride, pad, transposed, Hd, Wd,
                                 static_cast<nn::bf16*>(col), ldk, stream);
}

template <typename T>
static int im2col_tap_impl(const T* x, int F, int Hs, int Ws, int C, int kh, int kw, int stride, int pad_h,
                           int pad_w, int transposed, int Hd, int Wd, T* col, lr_stream_t stream) {
    const int pad = pad_h;
    LR_CHECK_ARG(F >= 0 && Hs > 0 && Ws > 0 && C > 0 && (C & 3) == 0 && Hd > 0 && Wd > 0, "lr_im2col_tap: bad shape (C %% 4 != 0?)");
    LR_CHECK_ARG(kh > 0 && kw > 0 && stride > 0 && pad_h >= 0 && pad_w >= 0, "lr_im2col_tap: bad window");
    if (F == 0) return LR_OK;
    LR_CHECK_ARG(x && col, "lr_im2col_tap: null pointer");
    LR_CHECK_ALIGN(x); LR_CHECK_ALIGN(col);
    c2::Geo g;
    g.Hd = Hd; g.Wd = Wd; g.kh = kh; g.kw = kw; g.stride = stride; g.pad = pad; g.transposed = transposed;
    g.pad_w = pad_w;
    g.K = C * kh * kw; g.ldk = g.K;
    const long long rows = (long long)F * Hd * Wd;
    constexpr int V = 16 / (int)sizeof(T);
    const bool vec = C % V == 0 && ((kh == 3 && kw == 3) || (kh == 1 && kw == 1));
    const unsigned grid = c2::grid_for(rows * (C / (vec ? V : 4)));
    if (vec && kh == 3 && !transposed) c2::im2col_tap_vec_kernel<T, 3, 3, false><<<grid, c2::TH, 0, stream>>>(x, g, Hs, Ws, C, col, rows);
    else if (vec && kh == 3) c2::im2col_tap_vec_kernel<T, 3, 3, true><<<grid, c2::TH, 0, stream>>>(x, g, Hs, Ws, C, col, rows);
    else if (vec && !transposed) c2::im2col_tap_vec_kernel<T, 1, 1, false><<<grid, c2::TH, 0, stream>>>(x, g, Hs, Ws, C, col, rows);
    else if (vec) c2::im2col_tap_vec_kernel<T, 1, 1, true><<<grid, c2::TH, 0, stream>>>(x, g, Hs, Ws, C, col, rows);
    else c2::im2col_tap_kernel<T><<<grid, c2::TH, 0, stream>>>(x, g, Hs, Ws, C, col, rows);
    lr::count_launch();
    LR_CHECK_LAUNCH("im2col_tap_kernel");
    return LR_OK;
}
extern "C" int lr_im2col_tap(const float* x, int F, int Hs, int Ws, int C, int kh, int kw, int stride, int pad_h,
                             int pad_w, int transposed, int Hd, int Wd, float* col, lr_stream_t stream) {
    return im2col_tap_impl<float>(x, F, Hs, Ws, C, kh, kw, stride, pad_h, pad_w, transposed, Hd, Wd, col, stream);
}
extern "C" int lr_im2col_tap_h(const void* x, int F, int Hs, int Ws, int C, int kh, int kw, int stride, int pad_h,
                               int pad_w, int transposed, int Hd, int Wd, void* col, lr_stream_t stream) {
    return im2col_tap_impl<nn::bf16>(static_cast<const nn::bf16*>(x), F, Hs, Ws, C, kh, kw, stride, pad_h, pad_w, transposed,
                                     Hd, Wd, static_cast<nn::bf16*>(col), stream);
}

extern "C" int lr_weight_tap(const float* src, float* dst, int Cout, int Cin, int kk, int mode, lr_stream_t stream) {
    LR_CHECK_ARG(Cout > 0 && Cin > 0 && kk > 0 && mode >= 0 && mode <= 2, "lr_weight_tap: bad argument");
    LR_CHECK_ARG(src && dst, "lr_weight_tap: null pointer");
    c2::weight_tap_kernel<float><<<c2::grid_for((long long)Cout * Cin * kk), c2::TH, 0, stream>>>(src, dst, Cout, Cin, kk, mode);
    lr::count_launch();
    LR_CHECK_LAUNCH("weight_tap_kernel");
    return LR_OK;
}
/* the same re-layout straight into the bf16 operand of the tensor-core convolutions (modes 0 and 1) */
extern "C" int lr_weight_tap_h(const float* src, void* dst, int Cout, int Cin, int kk, int mode, lr_stream_t stream) {
    LR_CHECK_ARG(Cout > 0 && Cin > 0 && kk > 0 && (mode == 0 || mode == 1), "lr_weight_tap_h: bad argument (modes 0 and 1)");
    LR_CHECK_ARG(src && dst, "lr_weight_tap_h: null pointer");
    c2::weight_tap_kernel<nn::bf16><<<c2::grid_for((long long)Cout * Cin * kk), c2::TH, 0, stream>>>(
        src, static_cast<nn::bf16*>(dst), Cout, Cin, kk, mode);
    lr::count_launch();
    LR_CHECK_LAUNCH("weight_tap_kernel");
    return LR_OK;
}

extern "C" int lr_weight_tap_batch_h(const void* table, int n, long long max_elems, lr_stream_t stream) {
    LR_CHECK_ARG(n >= 0 && max_elems >= 0 && max_elems < (1LL << 31), "lr_weight_tap_batch_h: bad argument");
    if (n == 0 || max_elems == 0) return LR_OK;
    LR_CHECK_ARG(table, "lr_weight_tap_batch_h: null pointer");
    LR_CHECK_ALIGN(table);
    long long bx = (max_elems + 4LL * c2::TH - 1) / (4LL * c2::TH);          // about four elements per thread for the largest
    if (bx > 256) bx = 256;
    c2::weight_tap_batch_kernel<<<dim3((unsigned)bx, (unsigned)n), c2::TH, 0, stream>>>(static_cast<const c2::WTapEntry*>(table));
    lr::count_launch();
    LR_CHECK_LAUNCH("weight_tap_batch_kernel");
    return LR_OK;
}

extern "C" int lr_weight_transpose(const float* w, float* wt, int Cout, int Cin, int kk, long long ldt,
                                   lr_stream_t stream) {
    LR_CHECK_ARG(Cout > 0 && Cin > 0 && kk > 0 && ldt >= (long long)Cout * kk, "lr_weight_transpose: bad shape");
    LR_CHECK_ARG(w && wt, "lr_weight_transpose: null pointer");
    c2::weight_transpose_kernel<<<c2::grid_for((long long)Cout * Cin * kk), c2::TH, 0, stream>>>(w, wt, Cout, Cin, kk, ldt);
    lr::count_launch();
    LR_CHECK_LAUNCH("weight_transpose_kernel");
    return LR_OK;
}

template <typename T>
static int maxpool_fwd_impl(const T* x, T* y, unsigned char* arg, int F, int H, int W, int C, int k,
                            int stride, int pad, lr_stream_t stream) {
    LR_CHECK_ARG(F >= 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0, "lr_maxpool_fwd: bad shape (C %% 4 != 0?)");
    LR_CHECK_ARG(k > 0 && k <= 15 && stride > 0 && pad >= 0 && 2 * pad <= k, "lr_maxpool_fwd: bad window");
    if (F == 0) return LR_OK;
    LR_CHECK_ARG(x && y && arg, "lr_maxpool_fwd: null pointer");
    LR_CHECK_ALIGN(x); LR_CHECK_ALIGN(y);
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    LR_CHECK_ARG(Ho > 0 && Wo > 0, "lr_maxpool_fwd: window larger than the input");
    if (k == 3 && stride == 2 && pad == 1 && C % c2::PoolVec<T>::V == 0) {
        const long long fo = (long long)F * Ho;
        const int grid = (int)(fo < 32LL * lr::sm_count() ? fo : 32LL * lr::sm_count());
        c2::maxpool_fwd_vec_kernel<T, 3, 2, 1><<<grid, c2::TH, 0, stream>>>(x, y, arg, F, H, W, C, Ho, Wo);
        lr::count_launch();
        LR_CHECK_LAUNCH("maxpool_fwd_vec_kernel");
        return LR_OK;
    }
    c2::maxpool_fwd_kernel<T><<<c2::grid_for((long long)F * Ho * Wo * (C >> 2)), c2::TH, 0, stream>>>(
        x, y, arg, F, H, W, C, k, stride, pad, Ho, Wo);
    lr::count_launch();
    LR_CHECK_LAUNCH("maxpool_fwd_kernel");
    return LR_OK;
}
extern "C" int lr_maxpool_fwd(const float* x, float* y, unsigned char* arg, int F, int H, int W, int C, int k,
                              int stride, int pad, lr_stream_t stream) {
    return maxpool_fwd_impl<float>(x, y, arg, F, H, W, C, k, stride, pad, stream);
}
extern "C" int lr_maxpool_fwd_h(const void* x, void* y, unsigned char* arg, int F, int H, int W, int C, int k,
                                int stride, int pad, lr_stream_t stream) {
    return maxpool_fwd_impl<nn::bf16>(static_cast<const nn::bf16*>(x), static_cast<nn::bf16*>(y), arg, F, H, W, C, k, stride, pad, stream);
}

template <typename T>
static int maxpool_bwd_impl(const T* dy, const unsigned char* arg, T* dx, int F, int H, int W, int C, int k,
                            int stride, int pad, lr_stream_t stream) {
    LR_CHECK_ARG(F >= 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0, "lr_maxpool_bwd: bad shape");
    LR_CHECK_ARG(k > 0 && k <= 15 && stride > 0 && pad >= 0 && 2 * pad <= k, "lr_maxpool_bwd: bad window");
    if (F == 0) return LR_OK;
    LR_CHECK_ARG(dy && arg && dx, "lr_maxpool_bwd: null pointer");
    LR_CHECK_ALIGN(dy); LR_CHECK_ALIGN(dx);
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    const long long fh = (long long)F * H;
    const int grid = (int)(fh < 64LL * lr::sm_count() ? fh : 64LL * lr::sm_count());
    if (k == 3 && stride == 2 && pad == 1 && C % c2::PoolVec<T>::V == 0) {
        c2::maxpool_bwd_vec_kernel<T, 3, 2, 1><<<grid, c2::TH, 0, stream>>>(dy, arg, dx, F, H, W, C, Ho, Wo);
        lr::count_launch();
        LR_CHECK_LAUNCH("maxpool_bwd_vec_kernel");
        return LR_OK;
    }
    c2::maxpool_bwd_kernel<T><<<grid, c2::TH, 0, stream>>>(dy, arg, dx, F, H, W, C, k, stride, pad, Ho, Wo);
    lr::count_launch();
    LR_CHECK_LAUNCH("maxpool_bwd_kernel");
    return LR_OK;
}
extern "C" int lr_maxpool_bwd(const float* dy, const unsigned char* arg, float* dx, int F, int H, int W, int C, int k,
                              int stride, int pad, lr_stream_t stream) {
    return maxpool_bwd_impl<float>(dy, arg, dx, F, H, W, C, k, stride, pad, stream);
}
extern "C" int lr_maxpool_bwd_h(const void* dy, const unsigned char* arg, void* dx, int F, int H, int W, int C, int k,
                                int stride, int pad, lr_stream_t stream) {
    return maxpool_bwd_impl<nn::bf16>(static_cast<const nn::bf16*>(dy), arg, static_cast<nn::bf16*>(dx), F, H, W, C, k, stride, pad, stream);
}

extern "C" int lr_zero_stuff2_h(const void* dy, void* up, int F, int Ho, int Wo, int Hi, int Wi, int C, lr_stream_t stream) {
    LR_CHECK_ARG(F >= 0 && Ho > 0 && Wo > 0 && Hi >= 2 * Ho - 1 && Wi >= 2 * Wo - 1 && C > 0 && (C & 7) == 0,
                 "lr_zero_stuff2_h: bad shape (C %% 8, Hi >= 2 Ho - 1)");
    if (F == 0) return LR_OK;
    LR_CHECK_ARG(dy && up, "lr_zero_stuff2_h: null pointer");
    LR_CHECK_ALIGN(dy); LR_CHECK_ALIGN(up);
    c2::zero_stuff2_kernel<<<c2::grid_for((long long)F * Hi * Wi * (C >> 3)), c2::TH, 0, stream>>>(
        static_cast<const uint4*>(dy), static_cast<uint4*>(up), F, Ho, Wo, Hi, Wi, C >> 3);
    lr::count_launch();
    LR_CHECK_LAUNCH("zero_stuff2_kernel");
    return LR_OK;
}

extern "C" int lr_dropout_fwd(const float* x, float* y, unsigned char* mask, long long n, float p,
                              unsigned long long seed, const long long* step, lr_stream_t stream) {
    LR_CHECK_ARG(n >= 0 && p >= 0.f && p < 1.f, "lr_dropout_fwd: need n >= 0 and 0 <= p < 1");
    if (n == 0) return LR_OK;
    LR_CHECK_ARG(x && y && mask && step, "lr_dropout_fwd: null pointer");
    c2::dropout_fwd_kernel<<<c2::grid_for(n), c2::TH, 0, stream>>>(x, y, mask, n, p, seed, step);
    lr::count_launch();
    LR_CHECK_LAUNCH("dropout_fwd_kernel");
    return LR_OK;
}

extern "C" int lr_dropout_bwd(const float* dy, const unsigned char* mask, float* dx, long long n, float p,
                              lr_stream_t stream) {
    LR_CHECK_ARG(n >= 0 && p >= 0.f && p < 1.f, "lr_dropout_bwd: need n >= 0 and 0 <= p < 1");
    if (n == 0) return LR_OK;
    LR_CHECK_ARG(dy && mask && dx, "lr_dropout_bwd: null pointer");
    c2::dropout_bwd_kernel<<<c2::grid_for(n), c2::TH, 0, stream>>>(dy, mask, dx, n, 1.f / (1.f - p));
    lr::count_launch();
    LR_CHECK_LAUNCH("dropout_bwd_kernel");
    return LR_OK;
}

extern "C" int lr_rng_tick(long long* step, lr_stream_t stream) {
    LR_CHECK_ARG(step, "lr_rng_tick: null pointer");
    c2::rng_tick_kernel<<<1, 1, 0, stream>>>(step);
    lr::count_launch();
    LR_CHECK_LAUNCH("rng_tick_kernel");
    return LR_OK;
}
