/* lipread_b200 -- C ABI of the B200-native audio-visual hot path (liblipread_b200.so).
 *
 * The reference (Aswath25S/multimodal_lipread) is pure Python and has no FFI of its own
 * (SURVEY.md 8(b)); every entry point below names the reference call site whose arithmetic it
 * replaces.  The Python host side (multimodal_lipread_b200/_lib.py, ops.py) binds these symbols
 * with ctypes and exposes them as torch.library custom ops under the reference's own
 * nn.Module / AudioProcessor surface.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.
 *   - Every pointer is DEVICE memory owned by the caller (including workspaces / plans).
 *   - Functions only enqueue work on `stream`; they never allocate, synchronise or throw.
 *   - Return LR_OK (0) or a negative LR_E* code; lr_last_error() gives a thread-local message.
 *   - Tensors are contiguous; base pointers must be 16-byte aligned (LR_EALIGN otherwise).
 *   - Safe to call under CUDA-graph stream capture.
 */
#ifndef LIPREAD_B200_H
#define LIPREAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* lr_stream_t; /* == cudaStream_t */

enum {
    LR_OK = 0,
    LR_EINVAL = -1,  /* bad shape / argument */
    LR_EALIGN = -2,  /* pointer not 16-byte aligned */
    LR_ECUDA = -3,   /* CUDA launch error (message holds cudaGetErrorString) */
    LR_ENOSPC = -4   /* caller-provided workspace / plan buffer too small */
};

#define LR_ABI_VERSION 1

int lr_version(void);
const char* lr_last_error(void);
/* Number of kernels this library has launched in the calling process (all threads). */
unsigned long long lr_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * K1  log-mel frontend
 *   replaces  audio/utils/audio_processor.py:48-52  AudioProcessor.compute_melspectrogram
 *             audio/utils/audio_processor.py:60-64  AudioProcessor.normalize_spectrogram
 *             audio/data_utils/dataset.py:52, audio_video/data_utils/dataset_av.py:62  crop [:80,:117]
 *   constants fixed by audio/utils/audio_processor.py:9-21: 16 kHz, n_fft = win = 400, hop 160,
 *   80 mels, 20 000 samples -> 126 frames, reflect padding, power 2, window-normalised.
 *
 * The plan holds what torchaudio keeps as module buffers (window, mel filterbank) in the form
 * the kernel wants plus FFT twiddles; build it once per device with lr_logmel_plan_init from the
 * SAME fp32 `window[400]` and `fb[201*80]` tensors the reference's transform owns.
 * ------------------------------------------------------------------------------------------ */
#define LR_LOGMEL_SAMPLES 20000
#define LR_LOGMEL_FRAMES 126
#define LR_LOGMEL_MELS 80

size_t lr_logmel_plan_bytes(void); /* pure host function */
int lr_logmel_plan_init(const float* window /*[400]*/, const float* fb /*[201,80] row-major*/,
                        void* plan, size_t plan_bytes, lr_stream_t stream);

/* mode LR_LOGMEL_FRONTEND: out[B,80,n_out] = ((L - mean)/(std + 1e-9))[:, :n_out], statistics over
 *                          all 80x126 values of the clip (unbiased std), 1 <= n_out <= 126.
 * mode LR_LOGMEL_RAW:      out[B,80,126]   = L = ln(mel + 1e-9)   (n_out must be 126). */
enum { LR_LOGMEL_FRONTEND = 0, LR_LOGMEL_RAW = 1 };
int lr_logmel_fwd(const float* wav /*[B,20000]*/, const void* plan, float* out, int B, int n_out,
                  int mode, lr_stream_t stream);

/* normalize_spectrogram alone on [B, n] rows: (x - mean)/(std_unbiased + 1e-9). */
int lr_normalize_fwd(const float* x, float* out, int B, int n, lr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LIPREAD_B200_H */
