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

// audio_cues_video/data_utils/dataset.py:256-258 divides a clip by 255 only `if arr.max() > 1.0`: a uint8 clip whose
// every pixel is 0 or 1 stays 0.0 / 1.0.  The stem kernels always divide uint8 by 255, so such a clip is rewritten
// as 0 / 255 here (one CTA per clip: max over the clip, then the rewrite if max <= 1) and comes out identical.
__global__ void __launch_bounds__(TH) u8_unit_clip_kernel(unsigned char* __restrict__ frames, long long clip_bytes) {
    __shared__ unsigned int smax;
    if (threadIdx.x == 0) smax = 0u;
    __syncthreads();
    unsigned char* clip = frames + (long long)blockIdx.x * clip_bytes;
    const long long n16 = clip_bytes / 16;
    unsigned int m = 0u;
    const uint4* v = reinterpret_cast<const uint4*>(clip);
    for (long long i = threadIdx.x; i < n16 && m <= 1u; i += TH) {
        const uint4 q = v[i];
        const unsigned int o = q.x | q.y | q.z | q.w;           // any byte > 1 sets a bit above bit 0 of some byte
        if (o & 0xFEFEFEFEu) m = 2u; else if (o) m = 1u;
    }
    for (long long i = n16 * 16 + threadIdx.x; i < clip_bytes; i += TH) m = max(m, (unsigned int)clip[i]);
    atomicMax(&smax, m);
    __syncthreads();
    if (smax != 1u) return;                                     // max > 1: divided as usual; max == 0: all zeros either way
    for (long long i = threadIdx.x; i < clip_bytes; i += TH) clip[i] = clip[i] ? 255 : 0;
}

}  // namespace ing

extern "C" int lr_u8_unit_clips(unsigned char* frames, int n_clips, long long clip_bytes, lr_stream_t stream) {
    LR_CHECK_ARG(n_clips >= 0 && clip_bytes > 0 && clip_bytes % 16 == 0, "lr_u8_unit_clips: clip_bytes must be a positive multiple of 16");
    if (n_clips == 0) return LR_OK;
    LR_CHECK_ARG(frames, "lr_u8_unit_clips: null pointer");
    LR_CHECK_ALIGN(frames);
    ing::u8_unit_clip_kernel<<<n_clips, ing::TH, 0, stream>>>(frames, clip_bytes);
    lr::count_launch();
    LR_CHECK_LAUNCH("u8_unit_clip_kernel");
    return LR_OK;
}

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
