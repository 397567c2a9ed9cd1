// PCM ingestion: packed, ragged int16 PCM clips -> the fixed-length fp32 waveform batch the log-mel kernel reads.
// Byte work, HBM-bound: 2*ch bytes in, 4 bytes out per sample; one pass, vector loads/stores where alignment allows.
// Semantics: audio/utils/audio_processor.py:29 (integer PCM -> float, NOT rescaled), :37 (mean over channels),
// :40-44 (truncate to target_samples or right zero-pad).
#include "common.cuh"

namespace ing {

constexpr int TH = 256;
constexpr int PER = 4;     // output samples per thread (one float4 store)

__global__ void __launch_bounds__(TH) pcm_ingest_kernel(const short* __restrict__ pcm, const long long* __restrict__ offset,
                                                        const int* __restrict__ n_frames, const int* __restrict__ channels,
                                                        float scale, float* __restrict__ wav, int target) {
    const int b = blockIdx.y;
    const long long off = offset[b];
    const int n = min(n_frames[b], target);                     // truncate (:40-41)
    const int ch = channels ? channels[b] : 1;
    const short* src = pcm + off;
    float* dst = wav + (long long)b * target;
    const int i0 = (blockIdx.x * TH + threadIdx.x) * PER;
    if (i0 >= target) return;
    float v[PER];
    if (ch == 1 && (off & 3) == 0 && i0 + PER <= n) {            // 8-byte aligned mono run
        const short4 s = *reinterpret_cast<const short4*>(src + i0);
        v[0] = (float)s.x * scale; v[1] = (float)s.y * scale; v[2] = (float)s.z * scale; v[3] = (float)s.w * scale;
    } else {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = i0 + j;
            float acc = 0.f;                                     // right zero-pad (:42-44)
            if (i < n) {
                const short* p = src + (long long)i * ch;
                acc = (float)p[0] * scale;
                for (int c = 1; c < ch; ++c) acc += (float)p[c] * scale;
                if (ch > 1) acc = __fdiv_rn(acc, (float)ch);     // samples.mean(dim=0) (:37)
            }
            v[j] = acc;
        }
    }
    if (i0 + PER <= target && (target & 3) == 0) {
        *reinterpret_cast<float4*>(dst + i0) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        for (int j = 0; j < PER && i0 + j < target; ++j) dst[i0 + j] = v[j];
    }
}

}  // namespace ing

extern "C" int lr_pcm_ingest(const short* pcm, const long long* offset, const int* n_frames, const int* channels,
                             float scale, float* wav, int B, int target, lr_stream_t stream) {
    LR_CHECK_ARG(B >= 0 && target > 0, "lr_pcm_ingest: bad shape (B %d, target %d)", B, target);
    if (B == 0) return LR_OK;
    LR_CHECK_ARG(B <= 65535, "lr_pcm_ingest: at most 65535 clips per call");
    LR_CHECK_ARG(pcm && offset && n_frames && wav, "lr_pcm_ingest: null pointer");
    LR_CHECK_ALIGN(pcm);
    LR_CHECK_ALIGN(wav);
    dim3 grid((unsigned)((target + ing::TH * ing::PER - 1) / (ing::TH * ing::PER)), (unsigned)B);
    ing::pcm_ingest_kernel<<<grid, ing::TH, 0, stream>>>(pcm, offset, n_frames, channels, scale, wav, target);
    lr::count_launch();
    LR_CHECK_LAUNCH("pcm_ingest_kernel");
    return LR_OK;
}
