// K1: batched framed-STFT -> mel filterbank -> log -> per-utterance normalisation -> crop, one kernel.
// Replaces audio/utils/audio_processor.py:48-52,60-64 + audio/data_utils/dataset.py:52 (reference:
// torch.stft on the CPU, one clip at a time inside DataLoader workers).
//
// v1 layout (v0 ran one 512-thread CTA per SM through four block-wide barriers per 63-frame chunk and sat at 0.096
// of HBM: latency / issue bound at 16 resident warps that all waited on the same barriers):
//   * a CTA of 8 warps owns one clip at a time (persistent over clips), two CTAs per SM;
//   * each WARP takes groups of 4 consecutive frames through the FFT pipeline on its own 6.8 KB of shared memory,
//     with __syncwarp only: waveform staging -> windowed 8-point DFTs (lane = residue r, window and twiddles in
//     registers) -> 25-point DFTs (lane = frame x k2, exactly 32 tasks) -> even/odd split + power (lane = bin).
//     The FFT buffer is reused in place for every stage (ordering argued at each step below);
//   * the eight warps' groups of one round are 32 consecutive frames: after a block barrier the banded mel filters
//     + log run with lane = frame and a warp-uniform mel (exact tap counts, no divergence, weights broadcast);
//   * the 880 samples of a warp's NEXT group are fetched with coalesced 16-byte loads into registers while the
//     current group is in its power / mel stages (HBM latency hidden per warp), then stored to shared memory: HBM
//     sees a waveform once (the 240-sample overlap of neighbouring groups hits L1 / L2);
//   * the 80 x 126 log-mel tile stays in shared memory until the clip statistics are known; only the cropped
//     result (80 x n_out) is written.
// Per clip: 80 000 B in, 80 * 117 * 4 = 37 440 B out.  What binds: fp32 issue slots (about 13 k thread-instructions
// per frame), not HBM -- see DESIGN.md section 4.
#include "common.cuh"
#include "logmel_core.cuh"

namespace lm {

constexpr int WARPS = 8, THREADS = WARPS * 32;
constexpr int GROUPS_PER_WARP = GROUPS / WARPS;           // 4
// A warp's buffer, in floats: [0, 1600) the FFT buffer of its four frames (800 float2), whose first 804 floats later
// hold the four power rows (stride 201); [804, 1684) the staged samples of the group.  The buffers are 1700 floats
// apart: 4 (mod 32), which with the odd row stride puts the 32 power rows of a round on 32 different banks.
constexpr int WAV_OFS = GF * PLD;                         // 804
constexpr int WBUF = 1700;
static_assert(GROUPS % WARPS == 0 && WARPS * GF == 32, "a round of groups is 32 consecutive frames");
static_assert(WAV_OFS % 4 == 0 && WBUF % 4 == 0 && WAV_OFS + GSAMP <= WBUF && WBUF % 32 == 4, "staging / bank layout");
static_assert(WAV_OFS >= 2 * NHALF * (GF - 2) + 2 * NHALF - HOP * (GF - 1), "stage A in place: see the kernel");

struct Smem {
    float Zb[WARPS][WBUF];                                // per-warp frame pipeline buffers             54 400 B
    float L[NMEL][LLD];                                   // log-mel tile of the clip                    40 640 B
    Plan plan;                                            // tables                                       9 776 B
    float red[WARPS];
    float bcast[2];
};

__device__ __forceinline__ float block_sum(float v, float* red, float* out_slot) {
    v = lr::warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        float s = lane < THREADS / 32 ? red[lane] : 0.f;
        s = lr::warp_sum(s);
        if (lane == 0) *out_slot = s;
    }
    __syncthreads();
    return *out_slot;
}

// the 880 padded samples of group g of one clip -> 7 float4 per lane (lane owns float4 number lane + 32 q).
// Groups 1..30 lie inside the signal: plain 16-byte loads; groups 0 and 31 touch the reflect padding.
__device__ __forceinline__ void fetch_group(const float* __restrict__ x, int g, int lane, float4 (&pre)[7]) {
    const int base = GF * HOP * g;                         // padded position of the group's first sample
    if (g >= 1 && g < GROUPS - 1) {
        const float4* src = reinterpret_cast<const float4*>(x + (base - PAD));
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const int i = lane + 32 * q;
            pre[q] = i < GSAMP / 4 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const int p = base + 4 * (lane + 32 * q);
            pre[q] = lane + 32 * q < GSAMP / 4
                         ? make_float4(padded_sample(x, p), padded_sample(x, p + 1), padded_sample(x, p + 2), padded_sample(x, p + 3))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// two banded filters over NQ tap quads each: all loads first, then two independent FMA chains
template <int NQ>
__device__ __forceinline__ void mel_pair(const float* __restrict__ pa, const float* __restrict__ pb,
                                         const float4* __restrict__ wa, const float4* __restrict__ wb,
                                         float& acc_a, float& acc_b) {
    float xa[4 * NQ], xb[4 * NQ];
    float4 ua[NQ], ub[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) { ua[q] = wa[q]; ub[q] = wb[q]; }
#pragma unroll
    for (int j = 0; j < 4 * NQ; ++j) { xa[j] = pa[j]; xb[j] = pb[j]; }
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        a = fmaf(xa[4 * q], ua[q].x, a);     b = fmaf(xb[4 * q], ub[q].x, b);
        a = fmaf(xa[4 * q + 1], ua[q].y, a); b = fmaf(xb[4 * q + 1], ub[q].y, b);
        a = fmaf(xa[4 * q + 2], ua[q].z, a); b = fmaf(xb[4 * q + 2], ub[q].z, b);
        a = fmaf(xa[4 * q + 3], ua[q].w, a); b = fmaf(xb[4 * q + 3], ub[q].w, b);
    }
    acc_a = a; acc_b = b;
}

__global__ void __launch_bounds__(THREADS, 2)
logmel_kernel(const float* __restrict__ wav, const Plan* __restrict__ gplan, float* __restrict__ out,
              int B, int n_out, int mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    {   // tables -> shared memory once per CTA (the CTA is persistent)
        const int4* src = reinterpret_cast<const int4*>(gplan);
        int4* dst = reinterpret_cast<int4*>(&S.plan);
        for (int i = tid; i < int(sizeof(Plan) / sizeof(int4)); i += THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const Plan& T = S.plan;
    float* const Fw = S.Zb[warp];
    float2* const Zw = reinterpret_cast<float2*>(Fw);        // FFT buffer; its front later holds the power rows
    float* const wavbuf = Fw + WAV_OFS;                      // staged samples

    float4 pre[7];
    int clip = blockIdx.x;
    if (clip < B) fetch_group(wav + size_t(clip) * NSAMP, warp, lane, pre);

    for (; clip < B; clip += gridDim.x) {
#pragma unroll 1
        for (int gi = 0; gi < GROUPS_PER_WARP; ++gi) {
            const int g = warp + WARPS * gi;
            // ---- staging: registers -> the buffer's sample area (disjoint from the FFT buffer and the power rows)
#pragma unroll
            for (int q = 0; q < 7; ++q)
                if (lane + 32 * q < GSAMP / 4) reinterpret_cast<float4*>(wavbuf)[lane + 32 * q] = pre[q];
            __syncwarp();

            // ---- stage A, one frame per round, lane = residue r (25 of 32 lanes).  Frame f reads floats
            // [804 + 160 f, +400) of the buffer and writes Y to [400 f, +400): a round's stores never touch samples
            // that a LATER round still reads (f = 2 ends at 1200, frame 3 starts at 1284), and inside a round every
            // lane has loaded before any lane stores (the __syncwarp).  The previous round's stage D, which read the
            // power rows that Y overwrites, ended with a block barrier.
            {
                const int r = lane < 25 ? lane : 24;
                float w16[16];
                float2 tw7[7];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 ww = *reinterpret_cast<const float2*>(&T.win[2 * (r + 25 * j)]);
                    w16[2 * j] = ww.x; w16[2 * j + 1] = ww.y;
                }
#pragma unroll
                for (int k2 = 1; k2 < 8; ++k2) tw7[k2 - 1] = T.tw200[k2][r];
#pragma unroll 1
                for (int f = 0; f < GF; ++f) {
                    float2 v[8];
                    const float2* xf = reinterpret_cast<const float2*>(wavbuf + HOP * f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = xf[r + 25 * j];
                    __syncwarp();
                    stage_a_regs(v, w16, tw7);
                    if (lane < 25) {
                        float2* Yf = Zw + NHALF * f;
#pragma unroll
                        for (int k2 = 0; k2 < 8; ++k2) Yf[k2 * 25 + r] = v[k2];
                    }
                }
            }
            __syncwarp();

            // ---- stage B: 25-point DFTs in place, lane = (frame, k2): 32 tasks.  Loads: the 8 lanes of a frame are
            // 50 words apart, the frames of a half-warp 400 words: 16 different bank pairs.  Stores: 8 consecutive float2.
            {
                float2* Zf = Zw + NHALF * (lane >> 3);
                const int k2 = lane & 7;
                float2 y[25], z[25];
                stage_b_load(Zf, k2, y);
                dft25(y, z);
                __syncwarp();
                stage_b_store(Zf, k2, z);
            }
            __syncwarp();

            // ---- the warp's next group (possibly of the CTA's next clip) starts its way from HBM now and lands in
            // registers while stages C and D run
            {
                int nclip = clip, ng = g + WARPS;
                if (gi == GROUPS_PER_WARP - 1) { nclip = clip + gridDim.x; ng = warp; }
                if (nclip < B) fetch_group(wav + size_t(nclip) * NSAMP, ng, lane, pre);
            }

            // ---- stage C: even/odd split + power, lane = bin k (k = lane + 32 j <= 100) for the four frames.  All
            // powers are formed in registers first: the power rows (4 x 201 floats) overwrite the front of the FFT buffer.
            {
                float2 pw[4][GF];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = lane + 32 * j;
                    if (k <= 100) {
                        const float2 tw = T.tw400[k];
#pragma unroll
                        for (int f = 0; f < GF; ++f) {
                            const float2* Zf = Zw + NHALF * f;
                            pw[j][f] = stage_c_pair(Zf[k], Zf[k == 0 ? 0 : NHALF - k], tw);
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = lane + 32 * j;
                    if (k <= 100) {
#pragma unroll
                        for (int f = 0; f < GF; ++f) {
                            Fw[PLD * f + k] = pw[j][f].x;
                            Fw[PLD * f + NHALF - k] = pw[j][f].y;   // k == 100: the same bin, the same value
                        }
                    }
                }
            }
            __syncwarp();

            // ---- stage D for the round's 32 frames (frame 32 gi + lane; its power row lives in warp lane / 4's buffer):
            // banded mel filters + log with a warp-uniform mel m = warp + 8 s, i.e. exact tap counts and broadcast weights.
            // Same summation order as stage_d().
            __syncthreads();                               // the eight warps' power rows are complete
            {
                const float* Pf = S.Zb[lane >> 2] + PLD * (lane & 3);
                const int t = 32 * gi + lane;
                // two filters (m, m + 8: close tap counts) per iteration = two independent FMA chains; the longer one
                // sets the trip count, the shorter one adds exact zeros (zero weights, finite in-row power values)
#pragma unroll 1
                for (int ma = warp; ma < NMEL; ma += 2 * WARPS) {
                    const int mb = ma + WARPS;
                    const float* pa = Pf + T.mel_lo[ma];
                    const float* pb = Pf + T.mel_lo[mb];
                    const float4* wa = reinterpret_cast<const float4*>(T.mel_w[ma]);
                    const float4* wb = reinterpret_cast<const float4*>(T.mel_w[mb]);
                    const int nq = max(T.mel_nq[ma], T.mel_nq[mb]);
                    float acc_a, acc_b;
                    switch (nq) {                              // warp-uniform; each case is straight-line code
                        case 1: mel_pair<1>(pa, pb, wa, wb, acc_a, acc_b); break;
                        case 2: mel_pair<2>(pa, pb, wa, wb, acc_a, acc_b); break;
                        case 3: mel_pair<3>(pa, pb, wa, wb, acc_a, acc_b); break;
                        default: mel_pair<4>(pa, pb, wa, wb, acc_a, acc_b); break;
                    }
                    if (t < NFRAMES) { S.L[ma][t] = log_eps(acc_a); S.L[mb][t] = log_eps(acc_b); }
                }
            }
            __syncthreads();                               // the power rows may be overwritten (next round's stage A)
        }
        // ---- statistics over all 80 x 126 values (two-pass, shifted by L[0][0] so that a constant tile -- a silent
        // clip -- gives exactly zero deviations), normalise, crop, write.  Thread = (mel row warp + 8 a, frames lane + 32 b).
        float* o = out + size_t(clip) * NMEL * n_out;
        if (mode == LR_LOGMEL_RAW) {
#pragma unroll 1
            for (int m = warp; m < NMEL; m += WARPS)
                for (int t = lane; t < NFRAMES; t += 32) o[m * NFRAMES + t] = S.L[m][t];
        } else {
            const float pivot = S.L[0][0];
            float v[NMEL / WARPS][4];
            float s = 0.f;
#pragma unroll
            for (int a = 0; a < NMEL / WARPS; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int t = lane + 32 * b;
                    v[a][b] = t < NFRAMES ? S.L[warp + WARPS * a][t] - pivot : 0.f;
                    s += v[a][b];
                }
            const float mean_d = block_sum(s, S.red, &S.bcast[0]) * (1.0f / float(NMEL * NFRAMES));
            float q = 0.f;
#pragma unroll
            for (int a = 0; a < NMEL / WARPS; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float d = v[a][b] - mean_d;
                    v[a][b] = d;
                    if (lane + 32 * b < NFRAMES) q = fmaf(d, d, q);
                }
            const float var = block_sum(q, S.red, &S.bcast[1]) * (1.0f / float(NMEL * NFRAMES - 1));
            const float inv = 1.0f / (sqrtf(var) + 1e-9f);
#pragma unroll
            for (int a = 0; a < NMEL / WARPS; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int t = lane + 32 * b;
                    if (t < n_out) o[(warp + WARPS * a) * n_out + t] = v[a][b] * inv;
                }
        }
        __syncthreads();   // L is rewritten by the next clip
    }
}

// ---- plan construction (one tiny launch per process/device) ----------------------------------
__global__ void logmel_plan_kernel(const float* __restrict__ window, const float* __restrict__ fb,
                                   Plan* __restrict__ plan) {
    __shared__ double s_norm;
    const int tid = threadIdx.x;
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < NFFT; ++i) s += double(window[i]) * double(window[i]);
        s_norm = 0.5 / sqrt(s);
        plan->status = 0;
    }
    __syncthreads();
    for (int i = tid; i < NFFT; i += blockDim.x) plan->win[i] = float(double(window[i]) * s_norm);
    for (int i = tid; i < 8 * 25; i += blockDim.x) {
        const int k2 = i / 25, r = i - k2 * 25;
        double sn, cs;
        sincospi(-2.0 * double(r * k2) / 200.0, &sn, &cs);
        plan->tw200[k2][r] = make_float2(float(cs), float(sn));
    }
    for (int k = tid; k <= 100; k += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * double(k) / 400.0, &sn, &cs);
        plan->tw400[k] = make_float2(float(cs), float(sn));
    }
    for (int m = tid; m < NMEL; m += blockDim.x)
        if (!plan_mel(fb, m, &plan->mel_lo[m], &plan->mel_nq[m], plan->mel_w[m])) plan->status = 1;
}

// ---- normalize_spectrogram on its own ---------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
normalize_kernel(const float* __restrict__ x, float* __restrict__ out, int n) {
    __shared__ float red[THREADS / 32];
    __shared__ float bc[2];
    const float* r = x + size_t(blockIdx.x) * n;
    float* o = out + size_t(blockIdx.x) * n;
    const float pivot = r[0];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += THREADS) s += r[i] - pivot;
    const float mean_d = block_sum(s, red, &bc[0]) / float(n);
    float q = 0.f;
    for (int i = threadIdx.x; i < n; i += THREADS) { const float d = (r[i] - pivot) - mean_d; q = fmaf(d, d, q); }
    const float var = block_sum(q, red, &bc[1]) / float(n - 1);
    const float inv = 1.0f / (sqrtf(var) + 1e-9f);
    for (int i = threadIdx.x; i < n; i += THREADS) o[i] = ((r[i] - pivot) - mean_d) * inv;
}

}  // namespace lm

extern "C" size_t lr_logmel_plan_bytes(void) { return (sizeof(lm::Plan) + 15) & ~size_t(15); }

extern "C" int lr_logmel_plan_init(const float* window, const float* fb, void* plan, size_t plan_bytes,
                                   lr_stream_t stream) {
    LR_CHECK_ARG(window && fb && plan, "lr_logmel_plan_init: null pointer");
    LR_CHECK_ALIGN(plan);
    if (plan_bytes < lr_logmel_plan_bytes())
        return lr::fail(LR_ENOSPC, "lr_logmel_plan_init: plan buffer %zu < %zu bytes", plan_bytes,
                        lr_logmel_plan_bytes());
    lm::logmel_plan_kernel<<<1, 128, 0, stream>>>(window, fb, static_cast<lm::Plan*>(plan));
    lr::count_launch();
    LR_CHECK_LAUNCH("logmel_plan_kernel");
    return LR_OK;
}

extern "C" int lr_logmel_fwd(const float* wav, const void* plan, float* out, int B, int n_out, int mode,
                             lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0, "lr_logmel_fwd: negative batch");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(wav && plan && out, "lr_logmel_fwd: null pointer");
    LR_CHECK_ARG(mode == LR_LOGMEL_FRONTEND || mode == LR_LOGMEL_RAW, "lr_logmel_fwd: bad mode %d", mode);
    LR_CHECK_ARG(n_out >= 1 && n_out <= lm::NFRAMES, "lr_logmel_fwd: n_out %d outside 1..126", n_out);
    LR_CHECK_ARG(mode != LR_LOGMEL_RAW || n_out == lm::NFRAMES, "lr_logmel_fwd: raw mode needs n_out == 126");
    LR_CHECK_ALIGN(wav); LR_CHECK_ALIGN(plan); LR_CHECK_ALIGN(out);
    static const int smem = int(sizeof(lm::Smem));
    const cudaError_t attr = lr::ensure_max_dynamic_smem(lm::logmel_kernel, smem);
    if (attr != cudaSuccess) return lr::fail(LR_ECUDA, "logmel smem attribute: %s", cudaGetErrorString(attr));
    const int grid = B < 2 * lr::sm_count() ? B : 2 * lr::sm_count();        // two resident CTAs per SM
    lm::logmel_kernel<<<grid, lm::THREADS, smem, stream>>>(wav, static_cast<const lm::Plan*>(plan), out, B,
                                                           n_out, mode);
    lr::count_launch();
    LR_CHECK_LAUNCH("logmel_kernel");
    return LR_OK;
}

extern "C" int lr_normalize_fwd(const float* x, float* out, int B, int n, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && n >= 2, "lr_normalize_fwd: need B >= 0, n >= 2");
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(x && out, "lr_normalize_fwd: null pointer");
    lm::normalize_kernel<<<B, lm::THREADS, 0, stream>>>(x, out, n);
    lr::count_launch();
    LR_CHECK_LAUNCH("normalize_kernel");
    return LR_OK;
}
