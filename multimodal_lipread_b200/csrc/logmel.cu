// K1: batched framed-STFT -> mel filterbank -> log -> per-utterance normalisation -> crop, one kernel.
// Replaces audio/utils/audio_processor.py:48-52,60-64 + audio/data_utils/dataset.py:52 (reference:
// torch.stft on the CPU, one clip at a time inside DataLoader workers).
//
// One persistent CTA per SM loops over clips.  A clip's 126 frames go through shared memory in two
// chunks of 63; the 80x126 log-mel tile stays in shared memory until the clip statistics are known,
// so HBM sees each waveform once (80 000 B in) and only the cropped result (80*117*4 B out).
#include "common.cuh"
#include "logmel_core.cuh"

namespace lm {

constexpr int THREADS = 512;
constexpr int PLD = 203;                                  // power row stride (201 bins, odd-ish padding)

struct Smem {
    Plan plan;                                            // tables                              9 776 B
    float2 Z[CHUNK][NHALF];                               // stage A/B buffer (in place)       100 800 B
    float P[CHUNK][PLD];                                  // power spectra                      51 156 B
    float L[NMEL][LLD];                                   // log-mel tile of the clip           40 640 B
    float red[THREADS / 32];
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

__global__ void __launch_bounds__(THREADS, 1)
logmel_kernel(const float* __restrict__ wav, const Plan* __restrict__ gplan, float* __restrict__ out,
              int B, int n_out, int mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;

    {   // tables -> shared memory once per CTA (the CTA is persistent)
        const int4* src = reinterpret_cast<const int4*>(gplan);
        int4* dst = reinterpret_cast<int4*>(&S.plan);
        for (int i = tid; i < int(sizeof(Plan) / sizeof(int4)); i += THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const float* win = S.plan.win;
    const float2* tw200 = &S.plan.tw200[0][0];
    const float2* tw400 = S.plan.tw400;

    for (int clip = blockIdx.x; clip < B; clip += gridDim.x) {
        const float* x = wav + size_t(clip) * NSAMP;

        for (int chunk = 0; chunk < NFRAMES / CHUNK; ++chunk) {
            const int f0 = chunk * CHUNK;
            // A: windowed load + 8-point DFTs + outer twiddles
            for (int task = tid; task < CHUNK * 25; task += THREADS) {
                const int f = task / 25, r = task - f * 25;
                stage_a(x, f0 + f, r, win, tw200, S.Z[f]);
            }
            __syncthreads();
            // B: 25-point DFTs, in place.  The 8 tasks of a frame sit in one warp, so a warp-level
            // barrier between the loads and the stores is enough.
            {
                const int f = tid >> 3, k2 = tid & 7;
                float2 y[25], z[25];
                if (f < CHUNK) { stage_b_load(S.Z[f], k2, y); dft25(y, z); }
                __syncwarp();
                if (f < CHUNK) stage_b_store(S.Z[f], k2, z);
            }
            __syncthreads();
            // C: even/odd split + power
            for (int task = tid; task < CHUNK * 101; task += THREADS) {
                const int f = task / 101, k = task - f * 101;
                stage_c(S.Z[f], k, tw400, S.P[f]);
            }
            __syncthreads();
            // D: mel filters + log
            for (int task = tid; task < CHUNK * NMEL; task += THREADS) {
                const int f = task / NMEL, m = task - f * NMEL;
                S.L[m][f0 + f] = stage_d(S.P[f], m, S.plan.mel_lo, S.plan.mel_n, &S.plan.mel_w[0][0]);
            }
            __syncthreads();
        }

        float* o = out + size_t(clip) * NMEL * n_out;
        if (mode == LR_LOGMEL_RAW) {
            for (int i = tid; i < NMEL * NFRAMES; i += THREADS) {
                const int m = i / NFRAMES, t = i - m * NFRAMES;
                o[i] = S.L[m][t];
            }
        } else {
            // statistics over all 80 x 126 values, two-pass, shifted by L[0][0] so that a constant
            // tile (silent clip) gives exactly zero deviations.
            const float pivot = S.L[0][0];
            float s = 0.f;
            for (int i = tid; i < NMEL * NFRAMES; i += THREADS) {
                const int m = i / NFRAMES, t = i - m * NFRAMES;
                s += S.L[m][t] - pivot;
            }
            const float mean_d = block_sum(s, S.red, &S.bcast[0]) * (1.0f / float(NMEL * NFRAMES));
            float q = 0.f;
            for (int i = tid; i < NMEL * NFRAMES; i += THREADS) {
                const int m = i / NFRAMES, t = i - m * NFRAMES;
                const float d = (S.L[m][t] - pivot) - mean_d;
                q = fmaf(d, d, q);
            }
            const float var = block_sum(q, S.red, &S.bcast[1]) * (1.0f / float(NMEL * NFRAMES - 1));
            const float inv = 1.0f / (sqrtf(var) + 1e-9f);
            for (int i = tid; i < NMEL * n_out; i += THREADS) {
                const int m = i / n_out, t = i - m * n_out;
                o[i] = ((S.L[m][t] - pivot) - mean_d) * inv;
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
    for (int m = tid; m < NMEL; m += blockDim.x) {
        int lo = -1, hi = -1;
        for (int k = 0; k < NBINS; ++k)
            if (fb[k * NMEL + m] != 0.f) { if (lo < 0) lo = k; hi = k; }
        if (lo < 0) { lo = 0; hi = -1; }
        int n = hi - lo + 1;
        if (n > MAXTAPS) { plan->status = 1; n = MAXTAPS; }
        plan->mel_lo[m] = lo;
        plan->mel_n[m] = n;
        for (int j = 0; j < MAXTAPS; ++j) plan->mel_w[j][m] = (j < n) ? fb[(lo + j) * NMEL + m] : 0.f;
    }
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
    const int grid = B < lr::sm_count() ? B : lr::sm_count();
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
