// Depthwise k x k convolutions of the MobileNet trunks (k in {3,5}, stride in {1,2}, pad k/2, channels-last fp32),
// forward / dgrad / wgrad: shared-memory tiled and register-blocked along the image row.
//
// A CTA owns a chunk of 32 channels and UB "units" = (frame, band of rows).  The patch every unit needs (halo
// included, zero filled outside the image) is staged once in shared memory with coalesced 128-byte segments, so
// HBM sees each activation once per kernel.  A thread owns ONE channel (its k*k weights live in registers) and
// computes WS adjacent pixels of one row at a time: each input value fetched from shared memory feeds up to
// min(WS, k) outputs, which lifts the kernel off the shared-memory bandwidth bound (25 reads per output for a 5x5)
// that a one-output-per-thread mapping hits.  Reads are conflict-free (32 lanes = 32 consecutive channels).
// The forward kernel also emits the per-channel sum / sum of squares (train-mode BatchNorm statistics).
#include "nn_common.cuh"
#include "dwconv_small.cuh"

namespace dw {

constexpr int TH = 256;
constexpr int SMEM_BUDGET = 32 * 1024;     // per staging buffer; every kernel double-buffers
constexpr int MAX_UB = 32;

struct Geo {
    int F, H, W, C, Ho, Wo, k, stride, pad;
    int CC;                // channels per chunk (32, or 16 when C == 16)
    int RB, nb, UB;        // rows per band, bands per frame, units per CTA
    int rows_t, cols_t;    // staged tile (per unit), cols_t includes the right padding the row blocking may touch
    int lo_c;              // dgrad: first staged dy column (<= 0)
    int nseg;              // row segments of WS pixels
    long long units;       // F * nb
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 16 : 0;                       // src-size 0: the 16 destination bytes are zero filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int MAX_PIX = 512;      // pixels of one unit's staged patch (SMEM_BUDGET / 64 B = 512 at most)

// tab[rc] = (r << 16) | col for the rc-th pixel of a unit's patch: built once per CTA so that the staging loops
// run without integer divisions (they were the bottleneck: 5 divisions per 16-byte request)
__device__ __forceinline__ void build_pixel_table(const Geo& d, int* tab) {
    for (int rc = threadIdx.x; rc < d.rows_t * d.cols_t; rc += TH) tab[rc] = ((rc / d.cols_t) << 16) | (rc % d.cols_t);
}

// The integer side of the main loops was the bottleneck of these kernels (ncu, round 2: IMAD / IADD3 / ISETP 45 % of the
// executed instructions, FFMA 7 %: three integer divisions per row segment and one more per unit): the decomposition of
// a flat item index it -> (unit, row, segment) does not depend on the round, so it is tabulated once per CTA, and the
// (frame, first row) of the round's units once per round.
constexpr int MAX_ITEMS = 1024;
__device__ __forceinline__ void build_item_table(const Geo& d, int* itab) {
    const int per_unit = d.RB * d.nseg;
    for (int it = threadIdx.x; it < d.UB * per_unit && it < MAX_ITEMS; it += TH) {
        const int ul = it / per_unit, rs = it - ul * per_unit, ro = rs / d.nseg;
        itab[it] = (ul << 20) | (ro << 10) | (rs - ro * d.nseg);
    }
}
// utab[ul] = (frame, first row of the band) of unit u0 + ul; frame = -1 past the last unit
__device__ __forceinline__ void build_unit_table(const Geo& d, int u0, int2* utab) {
    if ((int)threadIdx.x < d.UB) {
        const int u = u0 + (int)threadIdx.x;
        const int f = u / d.nb;
        utab[threadIdx.x] = u < (int)d.units ? make_int2(f, (u - f * d.nb) * d.RB) : make_int2(-1, 0);
    }
}

// Asynchronously stage tile[ul][r][col][c] = src[f, row0(band) + r, col0 + col, chunk*CC + c] (zero outside
// [0,SH) x [0,SW) and past the last unit) for the UB units starting at u0; cp.async keeps many 16-byte requests
// in flight per thread without a register round trip, so a round's loads overlap the previous round's math.
// (T = float: 4 channels per 16-byte request; T = bf16: 8 channels; the tile keeps the storage type)
template <typename T>
__device__ __forceinline__ void stage_async(const Geo& d, const T* __restrict__ src, T* tile, const int* tab,
                                            int u0, int chunk, int SH, int SW, int band_rows, int row_shift,
                                            int row_off, int col0) {
    constexpr int V = 16 / (int)sizeof(T);                         // channels per 16-byte request
    const int g4 = d.CC / V, lanes = TH / g4;
    const int g = threadIdx.x % g4, p = threadIdx.x / g4;
    const int npix = d.rows_t * d.cols_t;
    const int c = chunk * d.CC + g * V;
    const bool cvalid = c < d.C;
    for (int ul = 0; ul < d.UB; ++ul) {
        const int u = u0 + ul;
        const int f = u / d.nb, b = u - f * d.nb;
        const int row0 = (b * band_rows + row_off) >> row_shift;
        const bool uvalid = cvalid && u < (int)d.units;
        const T* base = src + (((long long)f * SH + row0) * SW + col0) * d.C + c;
        T* dst = tile + ((long long)ul * npix * g4 + g) * V;
        for (int rc = p; rc < npix; rc += lanes) {
            const int e = tab[rc], r = e >> 16, col = e & 0xffff;
            const int row = row0 + r, cc = col0 + col;
            const bool valid = uvalid && row >= 0 && row < SH && cc >= 0 && cc < SW;
            cp_async16(dst + (long long)rc * g4 * V, valid ? base + ((long long)r * SW + col) * d.C : src, valid);
        }
    }
}

template <int K>
__device__ __forceinline__ void load_weights(const float* __restrict__ w, int c, bool ok, float* wr) {
#pragma unroll
    for (int t = 0; t < K * K; ++t) wr[t] = ok ? w[c * K * K + t] : 0.f;
}

// ------------------------------------------------------------------------------------------ forward
template <typename T, int K, int S, int WS, int CCT>
__global__ void __launch_bounds__(TH)
fwd_kernel(const Geo d, const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
           double* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    __shared__ float ssum[TH], ssq[TH];                 // one slot per thread: [row lane][channel], added in lane order
    __shared__ int tab[MAX_PIX];
    __shared__ int itab[MAX_ITEMS];
    __shared__ int2 utab[MAX_UB];
    build_pixel_table(d, tab);
    build_item_table(d, itab);
    __syncthreads();
    const int cl = threadIdx.x % CCT, rl = threadIdx.x / CCT, nrl = TH / CCT;
    const int chunk = blockIdx.y, c = chunk * CCT + cl;
    const bool ok = c < d.C && rl < nrl;               // (CC = 24: the last 16 threads have no row lane)
    const int tile_floats = d.UB * d.rows_t * d.cols_t * CCT;
    const int rounds = (int)((d.units + d.UB - 1) / d.UB);
    float wr[K * K];
    load_weights<K>(w, c, ok, wr);
    constexpr int NX = (WS - 1) * S + K;               // input columns one row segment touches
    float ls = 0.f, lq = 0.f;
    int buf = 0;
    if ((int)blockIdx.x < rounds) stage_async(d, x, smem, tab, blockIdx.x * d.UB, chunk, d.H, d.W, d.RB * S, 0, -d.pad, -d.pad);
    cp_async_commit();
    for (int rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
        const int nxt = rd + gridDim.x;
        if (nxt < rounds) stage_async(d, x, smem + (buf ^ 1) * tile_floats, tab, nxt * d.UB, chunk, d.H, d.W, d.RB * S, 0, -d.pad, -d.pad);
        cp_async_commit();
        build_unit_table(d, rd * d.UB, utab);
        cp_async_wait<1>();                             // this round's tile has landed (the prefetch may be in flight)
        __syncthreads();
        const T* tile = smem + buf * tile_floats;
        const int per_unit = d.RB * d.nseg;
        for (int it = rl; it < d.UB * per_unit && ok; it += nrl) {       // flat over (unit, row, segment): balanced lanes
          {
            const int e = itab[it], ul = e >> 20, ro = (e >> 10) & 1023, seg = e & 1023;
            const int2 uf = utab[ul];
            if (uf.x < 0) break;
            const int f = uf.x, hob = uf.y;
            const int ho = hob + ro, wo0 = seg * WS;
            if (ho >= d.Ho) continue;
            float acc[WS];
#pragma unroll
            for (int j = 0; j < WS; ++j) acc[j] = 0.f;
            const T* base = tile + ((ul * d.rows_t + ro * S) * d.cols_t + wo0 * S) * CCT + cl;
#pragma unroll
            for (int kh = 0; kh < K; ++kh) {
                float xr[NX];
#pragma unroll
                for (int i = 0; i < NX; ++i) xr[i] = nn::ld1(base + (kh * d.cols_t + i) * CCT);
#pragma unroll
                for (int j = 0; j < WS; ++j)
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) acc[j] = fmaf(xr[j * S + kw], wr[kh * K + kw], acc[j]);
            }
            T* o = y + (((long long)f * d.Ho + ho) * d.Wo + wo0) * d.C + c;
#pragma unroll
            for (int j = 0; j < WS; ++j)
                if (wo0 + j < d.Wo) {
                    nn::st1(o + (long long)j * d.C, acc[j]);
                    const float v = sizeof(T) == 2 ? __bfloat162float(__float2bfloat16_rn(acc[j])) : acc[j];   // the STORED value
                    ls += v; lq = fmaf(v, v, lq);
                }
        }
        }
        __syncthreads();                                // everyone is done with `buf` before it is refilled
        buf ^= 1;
    }
    if (stats) {
        // fixed-order block reduction (no float atomics: the forward is bit-reproducible); across blocks the double
        // atomics add 24-bit partials into 53-bit sums, exact unless the partials span more than 2^29
        ssum[threadIdx.x] = ok ? ls : 0.f; ssq[threadIdx.x] = ok ? lq : 0.f;
        __syncthreads();
        if (threadIdx.x < CCT && chunk * CCT + threadIdx.x < d.C) {
            float a = 0.f, b = 0.f;
            for (int l = 0; l < nrl; ++l) { a += ssum[l * CCT + threadIdx.x]; b += ssq[l * CCT + threadIdx.x]; }
            nn::atomic_add_double(stats + chunk * CCT + threadIdx.x, (double)a);
            nn::atomic_add_double(stats + d.C + chunk * CCT + threadIdx.x, (double)b);
        }
    }
}

// ------------------------------------------------------------------------------------------ wgrad
// dw[c,kh,kw] += sum_{f,ho,wo} dy[f,ho,wo,c] * x[f, ho*s-pad+kh, wo*s-pad+kw, c]; persistent, double buffered.
template <typename T>
__device__ __forceinline__ void stage_dy_async(const Geo& d, const T* __restrict__ dy, T* gt, int u0, int chunk,
                                               int gw) {
    constexpr int V = 16 / (int)sizeof(T);
    const int g4 = d.CC / V, lanes = TH / g4;
    const int g = threadIdx.x % g4, p = threadIdx.x / g4;
    const int cc = chunk * d.CC + g * V;
    const int npix = d.RB * gw;
    for (int ul = 0; ul < d.UB; ++ul) {
        const int u = u0 + ul;
        const int f = u / d.nb, ho0 = (u - f * d.nb) * d.RB;
        const bool uvalid = cc < d.C && u < (int)d.units;
        T* dst = gt + ((long long)ul * npix * g4 + g) * V;
        for (int rw = p; rw < npix; rw += lanes) {
            const int ro = rw / gw, wo = rw - ro * gw, ho = ho0 + ro;
            const bool valid = uvalid && ho < d.Ho && wo < d.Wo;
            cp_async16(dst + (long long)rw * g4 * V, valid ? dy + (((long long)f * d.Ho + ho) * d.Wo + wo) * d.C + cc : dy, valid);
        }
    }
}

template <typename T, int K, int S, int WS, int CCT>
__global__ void __launch_bounds__(TH)
wgrad_kernel(const Geo d, const T* __restrict__ dy, const T* __restrict__ x, float* __restrict__ dwt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    __shared__ float red[K * K][32];
    __shared__ int tab[MAX_PIX];
    __shared__ int itab[MAX_ITEMS];
    __shared__ int2 utab[MAX_UB];
    build_pixel_table(d, tab);
    build_item_table(d, itab);
    const int cl = threadIdx.x % CCT, rl = threadIdx.x / CCT, nrl = TH / CCT;
    const int chunk = blockIdx.y, c = chunk * CCT + cl;
    const bool ok = c < d.C && rl < nrl;               // (CC = 24: the last 16 threads have no row lane)
    const int gw = d.nseg * WS;                                           // padded dy row length
    const int x_floats = d.UB * d.rows_t * d.cols_t * CCT, g_floats = d.UB * d.RB * gw * CCT;
    const int buf_floats = x_floats + g_floats;
    const int rounds = (int)((d.units + d.UB - 1) / d.UB);
    for (int i = threadIdx.x; i < K * K * 32; i += TH) (&red[0][0])[i] = 0.f;
    __syncthreads();
    float acc[K * K];
#pragma unroll
    for (int t = 0; t < K * K; ++t) acc[t] = 0.f;
    constexpr int NX = (WS - 1) * S + K;
    int buf = 0;
    if ((int)blockIdx.x < rounds) {
        stage_async(d, x, smem, tab, blockIdx.x * d.UB, chunk, d.H, d.W, d.RB * S, 0, -d.pad, -d.pad);
        stage_dy_async(d, dy, smem + x_floats, blockIdx.x * d.UB, chunk, gw);
    }
    cp_async_commit();
    for (int rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
        const int nxt = rd + gridDim.x;
        if (nxt < rounds) {
            T* nb_ = smem + (buf ^ 1) * buf_floats;
            stage_async(d, x, nb_, tab, nxt * d.UB, chunk, d.H, d.W, d.RB * S, 0, -d.pad, -d.pad);
            stage_dy_async(d, dy, nb_ + x_floats, nxt * d.UB, chunk, gw);
        }
        cp_async_commit();
        build_unit_table(d, rd * d.UB, utab);
        cp_async_wait<1>();
        __syncthreads();
        const T* tile = smem + buf * buf_floats;
        const T* gt = tile + x_floats;
        if (ok) {
            const int per_unit = d.RB * d.nseg;
            for (int it = rl; it < d.UB * per_unit; it += nrl) {
              {
                const int e = itab[it], ul = e >> 20, ro = (e >> 10) & 1023, seg = e & 1023;
                if (utab[ul].x < 0) break;
                const int wo0 = seg * WS;
                float g[WS];
                const T* gb = gt + ((ul * d.RB + ro) * gw + wo0) * CCT + cl;
#pragma unroll
                for (int j = 0; j < WS; ++j) g[j] = nn::ld1(gb + j * CCT);
                const T* base = tile + ((ul * d.rows_t + ro * S) * d.cols_t + wo0 * S) * CCT + cl;
#pragma unroll
                for (int kh = 0; kh < K; ++kh) {
                    float xr[NX];
#pragma unroll
                    for (int i = 0; i < NX; ++i) xr[i] = nn::ld1(base + (kh * d.cols_t + i) * CCT);
#pragma unroll
                    for (int kw = 0; kw < K; ++kw)
#pragma unroll
                        for (int j = 0; j < WS; ++j) acc[kh * K + kw] = fmaf(g[j], xr[j * S + kw], acc[kh * K + kw]);
                }
              }
            }
        }
        __syncthreads();
        buf ^= 1;
    }
    __syncthreads();
    if (ok) {
#pragma unroll
        for (int t = 0; t < K * K; ++t) atomicAdd(&red[t][cl], acc[t]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K * CCT; i += TH) {
        const int t = i / CCT, j = i % CCT, cc = chunk * CCT + j;
        if (cc < d.C) atomicAdd(&dwt[cc * K * K + t], red[t][j]);
    }
}

// ------------------------------------------------------------------------------------------ dgrad
// dx[f,hi,wi,c] = sum_{kh,kw : (hi+pad-kh) % s == 0, (wi+pad-kw) % s == 0} dy[f,(hi+pad-kh)/s,(wi+pad-kw)/s,c] * w[c,kh,kw]
// units are (frame, band of RB INPUT rows); a thread produces WS adjacent dx pixels of one row (WS even when s == 2,
// so the column parities are compile-time).
template <int S>
__host__ __device__ constexpr int floor_div(int a) { return S == 1 ? a : (a >= 0 ? a / 2 : -((-a + 1) / 2)); }

template <typename T, int K, int S, int WS, int CCT>
__global__ void __launch_bounds__(TH)
dgrad_kernel(const Geo d, const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    static_assert(S == 1 || (WS % 2) == 0, "stride-2 dgrad needs an even row segment");
    constexpr int SH_ = S == 2 ? 1 : 0, P = K / 2;
    __shared__ int tab[MAX_PIX];
    __shared__ int itab[MAX_ITEMS];
    __shared__ int2 utab[MAX_UB];
    build_pixel_table(d, tab);
    build_item_table(d, itab);
    __syncthreads();
    const int cl = threadIdx.x % CCT, rl = threadIdx.x / CCT, nrl = TH / CCT;
    const int chunk = blockIdx.y, c = chunk * CCT + cl;
    const bool ok = c < d.C && rl < nrl;               // (CC = 24: the last 16 threads have no row lane)
    const int tile_floats = d.UB * d.rows_t * d.cols_t * CCT;
    const int rounds = (int)((d.units + d.UB - 1) / d.UB);
    float wr[K * K];
    load_weights<K>(w, c, ok, wr);
    constexpr int BASE_OFF = floor_div<S>(P - (K - 1));
    constexpr int NG = floor_div<S>(WS - 1 + P) - BASE_OFF + 1;          // dy columns one row segment touches
    int buf = 0;
    // staged dy rows of a band start at floor((hi0 + P - (K-1)) / S), columns at lo_c = floor((P - (K-1)) / S)
    if ((int)blockIdx.x < rounds) stage_async(d, dy, smem, tab, blockIdx.x * d.UB, chunk, d.Ho, d.Wo, d.RB, SH_, P - (K - 1), d.lo_c);
    cp_async_commit();
    for (int rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
        const int nxt = rd + gridDim.x;
        if (nxt < rounds) stage_async(d, dy, smem + (buf ^ 1) * tile_floats, tab, nxt * d.UB, chunk, d.Ho, d.Wo, d.RB, SH_, P - (K - 1), d.lo_c);
        cp_async_commit();
        build_unit_table(d, rd * d.UB, utab);
        cp_async_wait<1>();
        __syncthreads();
        const T* tile = smem + buf * tile_floats;
        const int per_unit = d.RB * d.nseg;
        for (int it = rl; it < d.UB * per_unit && ok; it += nrl) {
          {
            const int e = itab[it], ul = e >> 20, ri = (e >> 10) & 1023, seg = e & 1023;
            const int2 uf = utab[ul];
            if (uf.x < 0) break;
            const int f = uf.x, hi0 = uf.y;
            const int hi = hi0 + ri, wi0 = seg * WS;
            if (hi >= d.H) continue;
            const int lo_r = (hi0 + P - (K - 1)) >> SH_;
            const int tc0 = wi0 >> SH_;                                   // == first dy column of the segment - lo_c
            float acc[WS];
#pragma unroll
            for (int j = 0; j < WS; ++j) acc[j] = 0.f;
#pragma unroll
            for (int kh = 0; kh < K; ++kh) {
                const int hn = hi + P - kh;
                if (S == 2 && (hn & 1)) continue;
                const int tr = (hn >> SH_) - lo_r;
                const T* row = tile + ((ul * d.rows_t + tr) * d.cols_t + tc0) * CCT + cl;
                float gr[NG];
#pragma unroll
                for (int i = 0; i < NG; ++i) gr[i] = nn::ld1(row + i * CCT);
#pragma unroll
                for (int j = 0; j < WS; ++j)
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) {
                        // wi0 is a multiple of WS (even when S == 2): the parity of (wi + P - kw) is that of (j + P - kw)
                        if (S == 2 && ((j + P - kw) & 1)) continue;
                        acc[j] = fmaf(gr[floor_div<S>(j + P - kw) - BASE_OFF], wr[kh * K + kw], acc[j]);
                    }
            }
            T* o = dx + (((long long)f * d.H + hi) * d.W + wi0) * d.C + c;
#pragma unroll
            for (int j = 0; j < WS; ++j) if (wi0 + j < d.W) nn::st1(o + (long long)j * d.C, acc[j]);
          }
        }
        __syncthreads();
        buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------ host side
// row segment (adjacent outputs per thread and row).  (12-wide segments -- a whole 11-wide row per item -- were measured:
// 22x22x72 s2 forward 53 -> 51 us, dgrad 63 -> 59, wgrad 68 -> 72: not worth three more instantiations per kernel.)
static int pick_ws(int width, bool even_needed) {
    if (!even_needed && width <= 3) return 3;
    if (width <= 6 || width == 11 || width == 12) return 6;
    return 8;
}

static Geo make_geo(int F, int H, int W, int C, int k, int stride, int mode /*0 fwd, 1 wgrad, 2 dgrad*/, int ws) {
    Geo d;
    d.F = F; d.H = H; d.W = W; d.C = C; d.k = k; d.stride = stride; d.pad = k / 2;
    d.Ho = (H + 2 * d.pad - k) / stride + 1; d.Wo = (W + 2 * d.pad - k) / stride + 1;
    // channels per chunk: 32 lanes of a warp = 32 channels; 24 when that tiles C exactly and 32 does not (C = 72: three
    // full chunks instead of two full ones and one at a quarter)
    d.CC = C == 16 ? 16 : ((C % 32) != 0 && (C % 24) == 0 ? 24 : 32);
    d.lo_c = 0;
    const int width = mode == 2 ? W : d.Wo;
    d.nseg = (width + ws - 1) / ws;
    int RB = mode == 2 ? H : d.Ho;
    for (;;) {
        size_t bytes;
        if (mode == 2) {
            const int lo = stride == 2 ? floor_div<2>(d.pad - (k - 1)) : d.pad - (k - 1);
            const int hi = stride == 2 ? floor_div<2>(d.nseg * ws - 1 + d.pad) : d.nseg * ws - 1 + d.pad;
            d.rows_t = (stride == 2 ? (RB - 1 + k - 1) / 2 : RB - 1 + k - 1) + 2;
            d.lo_c = lo; d.cols_t = hi - lo + 1;
            bytes = (size_t)d.rows_t * d.cols_t * d.CC * 4;
        } else {
            d.rows_t = (RB - 1) * stride + k; d.cols_t = (d.nseg * ws - 1) * stride + k;
            bytes = (size_t)d.rows_t * d.cols_t * d.CC * 4 + (mode == 1 ? (size_t)RB * d.nseg * ws * d.CC * 4 : 0);
        }
        if (bytes <= (size_t)SMEM_BUDGET || RB == 1) {
            d.RB = RB;
            int ub = (int)((size_t)SMEM_BUDGET / bytes);
            if (ub < 1) ub = 1;
            if (ub > MAX_UB) ub = MAX_UB;
            d.UB = ub;
            break;
        }
        RB = (RB + 1) / 2;
    }
    d.nb = ((mode == 2 ? H : d.Ho) + d.RB - 1) / d.RB;
    d.units = (long long)F * d.nb;
    if (d.units < d.UB) d.UB = (int)d.units;
    while (d.UB > 1 && d.UB * d.RB * d.nseg > MAX_ITEMS) --d.UB;        // the item table of a round (build_item_table)
    return d;
}
// es: bytes per staged element (4 = fp32, 2 = bf16).  The tile GEOMETRY (make_geo) is the same for both storage
// types -- it is sized for fp32 -- so the bf16 kernels simply use half the shared memory.
static size_t smem_bytes(const Geo& d, int mode, int ws, int es) {
    size_t b = (size_t)d.UB * d.rows_t * d.cols_t * d.CC * es;
    if (mode == 1) b += (size_t)d.UB * d.RB * d.nseg * ws * d.CC * es;
    return 2 * b;
}
// persistent grid: enough CTAs to fill the GPU a few times over, never more than there are rounds
static dim3 persistent_grid(const Geo& d, int ctas_per_sm) {
    const int chunks = (d.C + d.CC - 1) / d.CC;
    long long gx = (d.units + d.UB - 1) / d.UB;
    const long long cap = ((long long)lr::sm_count() * ctas_per_sm + chunks - 1) / chunks;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    return dim3((unsigned)gx, chunks);
}

// The opt-in dynamic shared-memory limit is raised ONCE per kernel instantiation to a fixed ceiling, never to the
// size of the current launch: a later, smaller launch must not lower the limit under launches already captured in
// a CUDA graph (tools that re-launch graph kernel nodes one by one -- ncu -- use the function's current limit).
constexpr int SMEM_CEILING = 160 * 1024;

// CC (channels per chunk: 16 or 32) is a template parameter of the kernels: every shared-memory tile offset of the main
// loops is then a compile-time multiple of it, i.e. an immediate of the load instead of an IMAD + 64-bit add per tap
#define DW_LAUNCH_CC(KERNEL, K_, S_, WS_, CC_, ...)                                      \
    do {                                                                                 \
        err = lr::ensure_max_dynamic_smem(KERNEL<T, K_, S_, WS_, CC_>, SMEM_CEILING);    \
        if (err == cudaSuccess && smem > (size_t)SMEM_CEILING) err = cudaErrorInvalidValue; \
        if (err == cudaSuccess) KERNEL<T, K_, S_, WS_, CC_><<<grid, TH, smem, stream>>>(__VA_ARGS__); \
    } while (0)
#define DW_LAUNCH(KERNEL, K_, S_, WS_, ...)                                              \
    do {                                                                                 \
        if (d.CC == 16) DW_LAUNCH_CC(KERNEL, K_, S_, WS_, 16, __VA_ARGS__);              \
        else if (d.CC == 24) DW_LAUNCH_CC(KERNEL, K_, S_, WS_, 24, __VA_ARGS__);         \
        else DW_LAUNCH_CC(KERNEL, K_, S_, WS_, 32, __VA_ARGS__);                         \
    } while (0)
#define DW_DISPATCH_WS(KERNEL, K_, S_, ...)                                              \
    switch (ws) {                                                                        \
        case 3: DW_LAUNCH(KERNEL, K_, S_, 3, __VA_ARGS__); break;                        \
        case 6: DW_LAUNCH(KERNEL, K_, S_, 6, __VA_ARGS__); break;                        \
        default: DW_LAUNCH(KERNEL, K_, S_, 8, __VA_ARGS__); break;                       \
    }
#define DW_DISPATCH_WS_EVEN(KERNEL, K_, S_, ...)                                         \
    switch (ws) {                                                                        \
        case 6: DW_LAUNCH(KERNEL, K_, S_, 6, __VA_ARGS__); break;                        \
        default: DW_LAUNCH(KERNEL, K_, S_, 8, __VA_ARGS__); break;                       \
    }

}  // namespace dw

#define LR_DW_CHECK(name)                                                                              \
    LR_CHECK_ARG(F >= 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0, name ": bad shape (C %% 4 != 0?)"); \
    LR_CHECK_ARG(k == 3 || k == 5, name ": kernel size %d not in {3,5}", k);                          \
    LR_CHECK_ARG(stride == 1 || stride == 2, name ": stride %d not in {1,2}", stride);                \
    if (F == 0) return LR_OK

template <typename T>
static int dwconv_fwd_impl(const T* x, const float* w, T* y, double* stats, int F, int H, int W, int C,
                           int k, int stride, lr_stream_t stream) {
    using namespace dw;
    LR_DW_CHECK("lr_dwconv_fwd");
    LR_CHECK_ARG(x && w && y, "lr_dwconv_fwd: null pointer");
    LR_CHECK_ARG(sizeof(T) == 4 || (C & 7) == 0, "lr_dwconv_fwd: bf16 storage needs C %% 8 == 0");
    LR_CHECK_ALIGN(x); LR_CHECK_ALIGN(y);
    if (dws::dispatch<T>(0, x, w, y, stats, F, H, W, C, k, stride, stream)) {
        lr::count_launch();
        LR_CHECK_LAUNCH("dws::fwd_kernel");
        return LR_OK;
    }
    const int Wo = (W + 2 * (k / 2) - k) / stride + 1;
    const int ws = pick_ws(Wo, false);
    const Geo d = make_geo(F, H, W, C, k, stride, 0, ws);
    LR_CHECK_ARG(d.rows_t * d.cols_t <= MAX_PIX, "lr_dwconv: image row of %d pixels is too wide for the staged tile", W);
    const size_t smem = smem_bytes(d, 0, ws, (int)sizeof(T));
    const dim3 grid = persistent_grid(d, 4);
    cudaError_t err = cudaSuccess;
    if (k == 3 && stride == 1) { DW_DISPATCH_WS(fwd_kernel, 3, 1, d, x, w, y, stats) }
    else if (k == 3) { DW_DISPATCH_WS(fwd_kernel, 3, 2, d, x, w, y, stats) }
    else if (stride == 1) { DW_DISPATCH_WS(fwd_kernel, 5, 1, d, x, w, y, stats) }
    else { DW_DISPATCH_WS(fwd_kernel, 5, 2, d, x, w, y, stats) }
    if (err != cudaSuccess) return lr::fail(LR_ECUDA, "lr_dwconv_fwd smem: %s", cudaGetErrorString(err));
    lr::count_launch();
    LR_CHECK_LAUNCH("dw::fwd_kernel");
    return LR_OK;
}
extern "C" int lr_dwconv_fwd(const float* x, const float* w, float* y, double* stats, int F, int H, int W, int C,
                             int k, int stride, lr_stream_t stream) {
    return dwconv_fwd_impl<float>(x, w, y, stats, F, H, W, C, k, stride, stream);
}
extern "C" int lr_dwconv_fwd_h(const void* x, const float* w, void* y, double* stats, int F, int H, int W, int C,
                               int k, int stride, lr_stream_t stream) {
    return dwconv_fwd_impl<nn::bf16>(static_cast<const nn::bf16*>(x), w, static_cast<nn::bf16*>(y), stats, F, H, W, C, k, stride, stream);
}

template <typename T>
static int dwconv_dgrad_impl(const T* dy, const float* w, T* dx, int F, int H, int W, int C, int k,
                             int stride, lr_stream_t stream) {
    using namespace dw;
    LR_DW_CHECK("lr_dwconv_dgrad");
    LR_CHECK_ARG(dy && w && dx, "lr_dwconv_dgrad: null pointer");
    LR_CHECK_ARG(sizeof(T) == 4 || (C & 7) == 0, "lr_dwconv_dgrad: bf16 storage needs C %% 8 == 0");
    LR_CHECK_ALIGN(dy); LR_CHECK_ALIGN(dx);
    if (dws::dispatch<T>(1, dy, w, dx, nullptr, F, H, W, C, k, stride, stream)) {
        lr::count_launch();
        LR_CHECK_LAUNCH("dws::dgrad_kernel");
        return LR_OK;
    }
    const int ws = pick_ws(W, stride == 2);
    const Geo d = make_geo(F, H, W, C, k, stride, 2, ws);
    LR_CHECK_ARG(d.rows_t * d.cols_t <= MAX_PIX, "lr_dwconv: image row of %d pixels is too wide for the staged tile", W);
    const size_t smem = smem_bytes(d, 2, ws, (int)sizeof(T));
    const dim3 grid = persistent_grid(d, 4);
    cudaError_t err = cudaSuccess;
    if (k == 3 && stride == 1) { DW_DISPATCH_WS(dgrad_kernel, 3, 1, d, dy, w, dx) }
    else if (k == 3) { DW_DISPATCH_WS_EVEN(dgrad_kernel, 3, 2, d, dy, w, dx) }
    else if (stride == 1) { DW_DISPATCH_WS(dgrad_kernel, 5, 1, d, dy, w, dx) }
    else { DW_DISPATCH_WS_EVEN(dgrad_kernel, 5, 2, d, dy, w, dx) }
    if (err != cudaSuccess) return lr::fail(LR_ECUDA, "lr_dwconv_dgrad smem: %s", cudaGetErrorString(err));
    lr::count_launch();
    LR_CHECK_LAUNCH("dw::dgrad_kernel");
    return LR_OK;
}
extern "C" int lr_dwconv_dgrad(const float* dy, const float* w, float* dx, int F, int H, int W, int C, int k,
                               int stride, lr_stream_t stream) {
    return dwconv_dgrad_impl<float>(dy, w, dx, F, H, W, C, k, stride, stream);
}
extern "C" int lr_dwconv_dgrad_h(const void* dy, const float* w, void* dx, int F, int H, int W, int C, int k,
                                 int stride, lr_stream_t stream) {
    return dwconv_dgrad_impl<nn::bf16>(static_cast<const nn::bf16*>(dy), w, static_cast<nn::bf16*>(dx), F, H, W, C, k, stride, stream);
}

template <typename T>
static int dwconv_wgrad_impl(const T* dy, const T* x, float* dwt, int F, int H, int W, int C, int k,
                             int stride, lr_stream_t stream) {
    using namespace dw;
    LR_DW_CHECK("lr_dwconv_wgrad");
    LR_CHECK_ARG(dy && x && dwt, "lr_dwconv_wgrad: null pointer");
    LR_CHECK_ARG(sizeof(T) == 4 || (C & 7) == 0, "lr_dwconv_wgrad: bf16 storage needs C %% 8 == 0");
    LR_CHECK_ALIGN(dy); LR_CHECK_ALIGN(x);
    if (dws::dispatch<T>(2, dy, x, dwt, nullptr, F, H, W, C, k, stride, stream)) {
        lr::count_launch();
        LR_CHECK_LAUNCH("dws::wgrad_kernel");
        return LR_OK;
    }
    const int Wo = (W + 2 * (k / 2) - k) / stride + 1;
    const int ws = pick_ws(Wo, false);
    const Geo d = make_geo(F, H, W, C, k, stride, 1, ws);
    LR_CHECK_ARG(d.rows_t * d.cols_t <= MAX_PIX, "lr_dwconv: image row of %d pixels is too wide for the staged tile", W);
    const size_t smem = smem_bytes(d, 1, ws, (int)sizeof(T));
    const dim3 grid = persistent_grid(d, 4);
    cudaError_t err = cudaSuccess;
    if (k == 3 && stride == 1) { DW_DISPATCH_WS(wgrad_kernel, 3, 1, d, dy, x, dwt) }
    else if (k == 3) { DW_DISPATCH_WS(wgrad_kernel, 3, 2, d, dy, x, dwt) }
    else if (stride == 1) { DW_DISPATCH_WS(wgrad_kernel, 5, 1, d, dy, x, dwt) }
    else { DW_DISPATCH_WS(wgrad_kernel, 5, 2, d, dy, x, dwt) }
    if (err != cudaSuccess) return lr::fail(LR_ECUDA, "lr_dwconv_wgrad smem: %s", cudaGetErrorString(err));
    lr::count_launch();
    LR_CHECK_LAUNCH("dw::wgrad_kernel");
    return LR_OK;
}
extern "C" int lr_dwconv_wgrad(const float* dy, const float* x, float* dwt, int F, int H, int W, int C, int k,
                               int stride, lr_stream_t stream) {
    return dwconv_wgrad_impl<float>(dy, x, dwt, F, H, W, C, k, stride, stream);
}
extern "C" int lr_dwconv_wgrad_h(const void* dy, const void* x, float* dwt, int F, int H, int W, int C, int k,
                                 int stride, lr_stream_t stream) {
    return dwconv_wgrad_impl<nn::bf16>(static_cast<const nn::bf16*>(dy), static_cast<const nn::bf16*>(x), dwt, F, H, W, C, k, stride, stream);
}
