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
 *   - Every pointer is DEVICE memory owned by the caller (including workspaces / plans), except in the
 *     lr_host_* readers of the input path, which take host pointers and make no CUDA calls.
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

#define LR_ABI_VERSION 2

int lr_version(void);
const char* lr_last_error(void);
/* Number of kernels this library has launched in the calling process (all threads). */
unsigned long long lr_launch_count(void);
/* Zero `bytes` bytes of device memory on `stream` (cudaMemsetAsync: per-step clearing of the BatchNorm
 * statistic arena and of the flat gradient buffer; a memset node under graph capture). */
int lr_memset(void* ptr, size_t bytes, lr_stream_t stream);

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

/* PCM ingestion in front of lr_logmel_fwd: B ragged clips of interleaved int16 PCM, packed in one buffer, become
 * the fixed-length fp32 batch wav[B, target].  Clip b starts at pcm[offset[b]] and holds n_frames[b] sample frames
 * of channels[b] channels (channels == NULL: mono).
 *   wav[b, i] = i < min(n_frames[b], target) ? mean_c(scale * pcm[offset[b] + i*ch + c]) : 0
 * replaces  audio/utils/audio_processor.py:29   integer PCM samples -> float, not rescaled (scale = 1)
 *           audio/utils/audio_processor.py:37   samples.mean(dim=0) over channels (scale = 1/32768 for the
 *                                               torchaudio.load branch, :31)
 *           audio/utils/audio_processor.py:40-44 truncate to target_samples / right zero-pad
 * Decoding the container (m4a -> PCM, pydub / ffmpeg, :26-28) stays on the host. */
int lr_pcm_ingest(const short* pcm, const long long* offset /*[B]*/, const int* n_frames /*[B]*/,
                  const int* channels /*[B] or NULL*/, float scale, float* wav /*[B,target]*/, int B, int target,
                  lr_stream_t stream);

/* audio_cues_video/data_utils/dataset.py:256-258 scales a lip clip by 1/255 only `if arr.max() > 1.0`.  The stem
 * kernels always divide uint8 frames by 255; this pass rewrites, in place, every clip of frames[n_clips][clip_bytes]
 * whose maximum is exactly 1 as 0 / 255, so that it comes out as the 0.0 / 1.0 the reference feeds its model. */
int lr_u8_unit_clips(unsigned char* frames, int n_clips, long long clip_bytes, lr_stream_t stream);

/* Host-side batch readers of the input path (the ONLY entry points that take host pointers; no CUDA calls).
 * A whole batch of `.npy` files is read by `n_threads` native threads straight into a ring slot of (pinned) host
 * memory, from where one async copy takes it to the device; no arithmetic on the host.
 *   replaces  video/data_utils/dataset_loader.py:90 / audio_video/data_utils/dataset_av.py:70   np.load per clip in
 *             DataLoader workers (the astype(float32) / 255 and the permute are folded into the stem kernels)
 * lr_host_read_npy_u8: file i must hold a C-order uint8 array of exactly shape[0..ndim); its payload lands at
 *   dst + i * prod(shape).
 * lr_host_read_npy_pcm16: file i holds little-endian int16 PCM, (n,) or (n, channels) interleaved; at most
 *   target_frames sample frames land at dst + i * cap and meta[0][i] = i * cap (offset), meta[1][i] = frames kept,
 *   meta[2][i] = channels (meta is [3][n]) -- the arguments of lr_pcm_ingest.
 * Errors (LR_EINVAL, message names the first offending file): missing file, wrong dtype / order / shape, truncated. */
int lr_host_read_npy_u8(const char* const* paths, int n, unsigned char* dst, const long long* shape, int ndim,
                        int n_threads);
int lr_host_read_npy_pcm16(const char* const* paths, int n, short* dst, long long cap, int target_frames,
                           long long* meta, int n_threads);

/* normalize_spectrogram alone on [B, n] rows: (x - mean)/(std_unbiased + 1e-9). */
int lr_normalize_fwd(const float* x, float* out, int B, int n, lr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Trunk / head building blocks.  Activations are channels-last fp32: a (F, H, W, C) frame batch is a
 * row-major [rows = F*H*W, C] matrix, so the 1x1 convolutions of torchvision's MobileNetV3 / the SE and
 * classifier linears / nn.LSTM's projections are all lr_gemm calls on it.
 * ------------------------------------------------------------------------------------------ */
enum { LR_ACT_NONE = 0, LR_ACT_RELU = 1, LR_ACT_HSWISH = 2, LR_ACT_HSIGMOID = 3, LR_ACT_RELU6 = 4 };

/* C[M,N] = act(A.B + bias) + R, optional per-column sum / sum-of-squares into double stats[2N]
 * (train-mode BatchNorm statistics of a conv output), optional split-K with atomic accumulation.
 *   a_trans 0: A is [M][K] (lda);  1: A is [K][M] (lda)     b_trans 0: B is [N][K] (ldb);  1: B is [K][N] (ldb)
 * replaces: nn.Conv2d(k=1) / nn.Linear forward, dgrad and wgrad (cuDNN / cuBLAS in the reference), e.g.
 * audio_video/models/middle_fusion_fast.py:13,20-25 and torchvision mobilenetv3 Conv2dNormActivation. */
int lr_gemm(const float* A, long long lda, int a_trans, const float* B, long long ldb, int b_trans, float* C,
            long long ldc, int M, int N, int K, const float* bias, int act, const float* R, long long ldr,
            double* stats, int ksplit, lr_stream_t stream);

/* Deterministic split-K for FORWARD GEMMs with a long reduction and few output tiles (audio_fc: K = 37120, M = batch,
 * audio_video/models/middle_fusion_fast.py:13,30): slice z of the reduction stores its partial tile in ws[z][M][N] and a
 * second kernel adds the slices in z order, then bias and activation -- no atomics, so two runs are bit-identical.
 * ws: caller-allocated, lr_gemm_splitk_workspace_bytes(M, N, ksplit) bytes. */
size_t lr_gemm_splitk_workspace_bytes(int M, int N, int ksplit);
int lr_gemm_splitk(const float* A, long long lda, int a_trans, const float* B, long long ldb, int b_trans, float* C,
                   long long ldc, int M, int N, int K, const float* bias, int act, int ksplit, float* ws,
                   size_t ws_bytes, lr_stream_t stream);

/* Tensor-core variant of lr_gemm (same layout flags and epilogue): TF32 products with fp32 accumulation
 * (tcgen05.mma kind::tf32; operands fetched by TMA from the fp32 matrices where they live, K-major or MN-major
 * with the 128-byte swizzle; accumulator in TMEM).  lda and ldb must be multiples of 4 floats, A and B 16-byte
 * aligned.  ksplit > 1 splits the reduction over CTAs and adds the partial tiles into C atomically (wgrad).
 * Used for the 1x1 convolutions of the trunk: forward (NT), dgrad (NN) and wgrad (TN). */
int lr_gemm_tf32(const float* A, long long lda, int a_trans, const float* B, long long ldb, int b_trans, float* C,
                 long long ldc, int M, int N, int K, const float* bias, int act, const float* R, long long ldr,
                 double* stats, int ksplit, lr_stream_t stream);

/* MobileNetV3 stem: Conv2d(3,16,3,stride=2,padding=1,bias=False) on frames addressed in the caller's own
 * layout: element (b, t, c, h, w) at x[b*sb + t*st + c*sc + h*sh + w*sw] (uint8 if is_u8 else float), times
 * `scale` (1/255 for uint8 .npy frames).  Folds video/data_utils/dataset_loader.py:90,96 (/255, permute) and
 * the TimeDistributed permute/contiguous/view of audio_video/models/middle_fusion_fast.py:32-33.
 * y: [B*T, Ho, Wo, 16] raw conv output; stats: double[32] accumulated (caller zeroes). */
int lr_stem_conv_fwd(const void* x, int is_u8, int B, int T, int H, int W, long long sb, long long st,
                     long long sc, long long sh, long long sw, float scale, const float* w /*[16,3,3,3]*/,
                     float* y, double* stats, lr_stream_t stream);
int lr_stem_conv_wgrad(const void* x, int is_u8, int B, int T, int H, int W, long long sb, long long st,
                       long long sc, long long sh, long long sw, float scale, const float* dy,
                       float* dw /*[16,3,3,3], accumulated*/, lr_stream_t stream);

/* Depthwise k x k convolution (k in {3,5}, stride in {1,2}, padding k/2, no bias), x: [F,H,W,C],
 * w: [C,1,k,k] (torch layout), y: [F,Ho,Wo,C]; stats as above.  dw is accumulated into. */
int lr_dwconv_fwd(const float* x, const float* w, float* y, double* stats, int F, int H, int W, int C, int k,
                  int stride, lr_stream_t stream);
int lr_dwconv_dgrad(const float* dy, const float* w, float* dx, int F, int H, int W, int C, int k, int stride,
                    lr_stream_t stream);
int lr_dwconv_wgrad(const float* dy, const float* x, float* dw, int F, int H, int W, int C, int k, int stride,
                    lr_stream_t stream);

/* nn.BatchNorm2d (+ activation, + residual add) on [rows, C].  training != 0: batch statistics from
 * `stats` (double[2C] sum / sum of squares over the rows), running_mean / running_var / num_batches_tracked
 * updated exactly as torch does (unbiased variance, momentum); training == 0: running statistics.
 * res_pre == 0: z = act(bn(x)) + residual   (torchvision InvertedResidual: the projection has no activation)
 * res_pre != 0: z = act(bn(x) + residual)   (torchvision BasicBlock: out += identity; out = relu(out)) */
int lr_bn_act_fwd(const float* x, const double* stats, const float* gamma, const float* beta, float* running_mean,
                  float* running_var, long long* num_batches_tracked, float eps, float momentum, int act,
                  int training, const float* residual, int res_pre, float* z, long long rows, int C,
                  lr_stream_t stream);
/* dx from dz (gradient of z); dgamma / dbeta accumulated into; sums: double[2C] scratch zeroed by the caller.
 * z_out == NULL: the activation derivative is evaluated at bn(x) and the residual branch's gradient is dz itself.
 * z_out != NULL (res_pre blocks, ReLU / ReLU6 only): the derivative is taken through the forward output z_out;
 * the masked gradient dz * act'(z) -- which is also the identity branch's gradient -- is written to dres if given. */
int lr_bn_act_bwd(const float* x, const double* stats, const float* gamma, const float* beta,
                  const float* running_mean, const float* running_var, float eps, int act, int training,
                  const float* dz, const float* z_out, float* dres, double* sums, float* dx, float* dgamma,
                  float* dbeta, long long rows, int C, lr_stream_t stream);

/* Per-frame reductions over the HW pixels of [F, HW, C]:  mode 0: p = mean(a)  (AdaptiveAvgPool2d(1), SE squeeze);
 * mode 1: p = sum(a * g)  (gradient of the SE gate). */
int lr_frame_reduce(const float* a, const float* g, float* p, int F, int HW, int C, int mode, lr_stream_t stream);
/* out[f,hw,c] = a[f,hw,c] * s[f,c] + dp[f,c] / HW   (either term may be absent: a == NULL or dp == NULL):
 * SE excitation forward, SE backward, average-pool backward. */
int lr_frame_scale(const float* a, const float* s, const float* dp, float* out, int F, int HW, int C,
                   lr_stream_t stream);
/* Linear layers on at most 32 rows (heads: one row per clip), output columns spread over CTAs, fp32 FMA in a fixed order:
 *   lr_linear_small_fwd  : y[m, n] = act(b[n] + sum_k x[m*ldx + k] * w[n*K + k])             (nn.Linear [+ activation])
 *   lr_linear_small_dgrad: dx[m, k] = sum_n dy[m*ldy + n] * w[n*K + k] (+ r[m*ldr + k])     (its input gradient)
 * M <= 32; K and the row pitches multiples of 4 (forward: K <= 1600), dgrad: N multiple of 4 (<= 1088), K multiple of 32.
 * Replaces the tile-GEMM launches of the classifier MLPs (audio_video/models/middle_fusion_fast.py:20-25,38-39) and of
 * the single reverse LSTM step, which were latency chains of one or two CTAs. */
int lr_linear_small_fwd(const float* x, long long ldx, const float* w, const float* b, float* y, long long ldy, int M,
                        int N, int K, int act, lr_stream_t stream);
int lr_linear_small_dgrad(const float* dy, long long ldy, const float* w, float* dx, long long ldx, const float* r,
                          long long ldr, int M, int N, int K, lr_stream_t stream);

/* Squeeze-Excitation gate on the pooled per-frame vectors p [F, C] in ONE launch per direction (torchvision
 * SqueezeExcitation: fc1 -> activation -> fc2 -> scale_activation, the MobileNetV3 trunk the reference builds in
 * audio_video/models/middle_fusion_fast.py:15-17 and early_fusion.py:58-60):
 *   forward : h1 = act1(p . w1^T + b1) [F, Cs],  s = act2(h1 . w2^T + b2) [F, C]      (w1 [Cs, C], w2 [C, Cs], fp32 FMA,
 *             fixed summation order: bit-reproducible)
 *   backward: ds [F, C] holds dL/ds on entry and dz2 = ds * act2'(s) on exit; dz1 = (dz2 . w2) * act1'(h1) [F, Cs];
 *             dp = dz1 . w1 [F, C].  The weight gradients are the caller's (dz2^T h1, dz1^T p and their column sums).
 * act1 / act2: LR_ACT_NONE, RELU, RELU6 or HSIGMOID (derivatives taken through the outputs).  C, Cs multiples of 4. */
int lr_se_fc_fwd(const float* p, const float* w1, const float* b1, const float* w2, const float* b2, float* h1,
                 float* s, int F, int C, int Cs, int act1, int act2, lr_stream_t stream);
int lr_se_fc_bwd(float* ds, const float* s, const float* h1, const float* w1, const float* w2, float* dz1, float* dp,
                 int F, int C, int Cs, int act1, int act2, lr_stream_t stream);
/* y = act(x) element-wise (x == y allowed). */
int lr_act_fwd(const float* x, float* y, long long n, int act, lr_stream_t stream);
/* dy *= act'(.) expressed through the activation OUTPUT y (ReLU, ReLU6, hard-sigmoid), in place. */
int lr_act_bwd(float* dy, const float* y, long long n, int act, lr_stream_t stream);
/* db[n] += sum_m dY[m*ld + n]  (bias gradients) */
int lr_colsum(const float* dY, long long ld, long long M, int N, float* db, lr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense k x k convolutions (groups == 1) as GEMMs on an explicit patch matrix: torchvision resnet18
 * (video/models/resnet_lstm.py:79-110, audio/models/resnet_model.py:12-17), AudioEncoder
 * (audio_video/models/early_fusion.py:21-35), the mobilenet_v2 stem.
 *
 * lr_im2col gathers col[(f,hd,wd)][c*kh*kw + r*kw + s] (row pitch ldk, tail columns zeroed) from a source whose
 * element (f, c, h, w) lies at x[(f/T)*sb + (f%T)*st + c*sc + h*sh + w*sw] (uint8 if is_u8, times `scale`), so the
 * lip frames are read in the caller's layout exactly as lr_stem_conv_fwd does.
 *   transposed == 0 (forward / wgrad operand): source pixel (hd*stride - pad + r, wd*stride - pad + s);
 *                    the column order equals torch's weight layout [Cout][Cin][kh][kw], so y = col . W^T.
 *   transposed != 0 (dgrad operand, source = dy): source pixel ((hd + pad - r)/stride, (wd + pad - s)/stride) when
 *                    both divide exactly; dx = colT . Wt^T with Wt from lr_weight_transpose.
 * Out-of-range taps contribute 0. */
int lr_im2col(const void* x, int is_u8, float scale, int F, int T, long long sb, long long st, long long sc,
              long long sh, long long sw, int Hs, int Ws, int C, int kh, int kw, int stride, int pad, int transposed,
              int Hd, int Wd, float* col, long long ldk, lr_stream_t stream);
/* Tap-major variant for channels-last float activations with C % 4 == 0 (every ResNet convolution but the stem):
 * col[(f,hd,wd)][(r*kw + s)*C + c], a shifted float4 copy that is coalesced on both sides.  The matching weight
 * layouts come from lr_weight_tap: mode 0: wp[k][rs][c] = w[k][c][rs] (forward / wgrad operand), mode 1:
 * wp[c][rs][k] = w[k][c][rs] (dgrad operand), mode 2: w[k][c][rs] = wp[k][rs][c] (weight gradient back to torch's
 * layout).  pad_h / pad_w are separate so that nn.Conv1d(k, padding=p) over time (video/models/cnn.py:35-42) runs as a
 * 1 x k window (kh = 1, pad_h = 0, pad_w = p) on a [B, 1, T, C] map. */
int lr_im2col_tap(const float* x, int F, int Hs, int Ws, int C, int kh, int kw, int stride, int pad_h, int pad_w,
                  int transposed, int Hd, int Wd, float* col, lr_stream_t stream);
int lr_weight_tap(const float* src, float* dst, int Cout, int Cin, int kk, int mode, lr_stream_t stream);
/* modes 0 and 1 straight into the bf16 operand of the tensor-core convolutions (precision "bf16") */
int lr_weight_tap_h(const float* src, void* dst, int Cout, int Cin, int kk, int mode, lr_stream_t stream);
/* All tap-major bf16 operands of a model in one launch: `table` = n device entries of four 64-bit words
 * {src (const float*), dst (bf16*), Cout | Cin << 32, kk | mode << 32}, modes 0 / 1; max_elems = the largest Cout*Cin*kk. */
int lr_weight_tap_batch_h(const void* table, int n, long long max_elems, lr_stream_t stream);
/* wt[c][k*kk + rs] (row pitch ldt) = w[k][c][rs]: the dgrad weight of a dense convolution. */
int lr_weight_transpose(const float* w, float* wt, int Cout, int Cin, int kk, long long ldt, lr_stream_t stream);

/* nn.MaxPool2d(k, stride, pad) on [F,H,W,C] -> [F,Ho,Wo,C]; arg: uint8 window position of the first maximum
 * (torch's tie rule), saved for the backward.  The backward is a gather (no atomics). */
int lr_maxpool_fwd(const float* x, float* y, unsigned char* arg, int F, int H, int W, int C, int k, int stride,
                   int pad, lr_stream_t stream);
int lr_maxpool_bwd(const float* dy, const unsigned char* arg, float* dx, int F, int H, int W, int C, int k,
                   int stride, int pad, lr_stream_t stream);

/* nn.Dropout(p) in training mode: y = keep ? x / (1 - p) : 0 with keep drawn from a counter-based generator keyed by
 * (seed, *step, element index); *step is a device counter advanced once per train step by lr_rng_tick, so a
 * replayed CUDA graph draws fresh masks.  (The reference draws from torch's Philox stream; masks are statistically,
 * not bitwise, equivalent -- parity tests run with p = 0) */
int lr_dropout_fwd(const float* x, float* y, unsigned char* mask, long long n, float p, unsigned long long seed,
                   const long long* step, lr_stream_t stream);
int lr_dropout_bwd(const float* dy, const unsigned char* mask, float* dx, long long n, float p, lr_stream_t stream);
int lr_rng_tick(long long* step, lr_stream_t stream);

/* One direction of one nn.LSTM layer over `nsteps` <= T steps of its walk (forward: t = 0.., reverse: t = T-1..).
 * xproj: [B*T, 4H] = x W_ih^T + b_ih + b_hh (one lr_gemm); out rows (b*T+t) with stride ldo; gates / cst / hprev
 * ([B,T,4H], [B,T,H], [B,T,H]) are saved for the backward pass when non-NULL. */
int lr_lstm_fwd(const float* xproj, long long ldx, const float* bhh /*[4H] or NULL, added to xproj*/,
                const float* whh, float* out, long long ldo, float* gates, float* cst, float* hprev, int B, int T,
                int H, int nsteps, int reverse, lr_stream_t stream);
/* BPTT: dgates[B,T,4H] (gradient of the gate pre-activations) for the visited steps.  External gradient of
 * the outputs: dout_step < 0: dout[(b*T+t)*ldo + k] for every t; dout_step >= 0: only h at t == dout_step
 * receives dout[b*ldo + k] (a head that reads out[:, -1]). */
int lr_lstm_bwd(const float* dout, long long ldo, int dout_step, const float* gates, const float* cst, const float* whh,
                float* dgates, int B, int T, int H, int nsteps, int reverse, lr_stream_t stream);

/* The same recurrence / BPTT on the tensor cores (csrc/lstm_tc.cu, precision "bf16"): W_hh and h_{t-1} (forward),
 * W_hh^T and dgates (backward) are rounded to bfloat16 and multiplied by tcgen05.mma with fp32 accumulation; the input
 * projection, gate activations, cell states and every saved / returned tensor stay fp32.  H must be 128, 256 or 512
 * (1 / 4 / 16 CTAs of one cluster per 32 batch rows keep a 128 KB bf16 slice of W_hh resident in shared memory);
 * arguments and semantics as lr_lstm_fwd / lr_lstm_bwd.  Replaces nn.LSTM under torch.autocast(bfloat16) at the
 * reference's call sites (video/models/resnet_lstm.py:113-120, audio_video/models/early_fusion.py:62-69,
 * audio_video/models/middle_fusion_fast.py:18,35-36). */
int lr_lstm_fwd_tc(const float* xproj, long long ldx, const float* bhh, const float* whh, float* out, long long ldo,
                   float* gates, float* cst, float* hprev, int B, int T, int H, int nsteps, int reverse,
                   lr_stream_t stream);
int lr_lstm_bwd_tc(const float* dout, long long ldo, int dout_step, const float* gates, const float* cst, const float* whh,
                   float* dgates, int B, int T, int H, int nsteps, int reverse, lr_stream_t stream);

/* dst[r*ldd + c] = src[r*lds + c]  (strided row gather, e.g. out[:, -1] of a sequence into the fusion row) */
int lr_copy2d(float* dst, long long ldd, const float* src, long long lds, int rows, int cols, lr_stream_t stream);

/* MidFusionFast audio branch: Conv2d(1,16,3,padding=1) + ReLU + MaxPool2d(2) fused
 * (audio_video/models/middle_fusion_fast.py:8-12,28-29): x [B,H,W] -> out rows of 16*(H/2)*(W/2) (stride ldo),
 * arg: uint8 arg-max code per output (saved for the backward). */
int lr_audio_conv_fwd(const float* x, const float* w, const float* bias, float* out, long long ldo,
                      unsigned char* arg, int B, int H, int W, lr_stream_t stream);
int lr_audio_conv_bwd(const float* x, const float* dA, long long lda, const unsigned char* arg, float* dw, float* db,
                      int B, int H, int W, lr_stream_t stream);

/* ShuffleNetV2 unit tail: out[rows, 2*Ch] = channel_shuffle(cat(a, b), groups = 2), i.e. out[r, 2c] = a[r*lda + c],
 * out[r, 2c+1] = b[r*ldb + c]  (a / b may be column slices: x.chunk(2, dim=1) leaves x1 where it is); the backward
 * de-interleaves dout into da / db (written, same strides).
 * replaces: torchvision.models.shufflenetv2 InvertedResidual.forward / channel_shuffle as used by
 * video/models/shufflenet_lstm.py:37-55. */
int lr_shuffle2_fwd(const float* a, long long lda, const float* b, long long ldb, float* out, long long rows, int Ch,
                    lr_stream_t stream);
int lr_shuffle2_bwd(const float* dout, float* da, long long lda, float* db, long long ldb, long long rows, int Ch,
                    lr_stream_t stream);

/* nn.MultiheadAttention(embed_dim E, heads, batch_first) over the T time steps of each clip, around its in- and
 * out-projections (two lr_gemm calls): qkv is the packed in-projection [B*T, 3E] (row stride ld; q | k | v, head h in
 * columns h*d .. (h+1)*d of each section, d = E / heads), P [B, heads, T, T] the attention weights (saved for the
 * backward), O [B*T, E] the concatenated heads.  T <= 64, d <= 128.
 *   scores_fwd: P = softmax((q / sqrt(d)) k^T)            apply_fwd: O = P v     (nn.Dropout on P sits between them)
 *   apply_bwd:  dP = dO v^T, dqkv[v section] = P^T dO     scores_bwd: dqkv[q | k sections] through the softmax
 * replaces: video/models/resnet_attn.py:23-35,103 (TemporalAttention), the self-attention of
 * nn.TransformerEncoderLayer (video/models/resnet_trans.py:96-103). */
int lr_mha_scores_fwd(const float* qkv, long long ld, float* P, int B, int T, int E, int heads, lr_stream_t stream);
int lr_mha_apply_fwd(const float* P, const float* qkv, long long ld, float* O, int B, int T, int E, int heads,
                     lr_stream_t stream);
int lr_mha_apply_bwd(const float* dO, const float* P, const float* qkv, long long ld, float* dP, float* dqkv, int B,
                     int T, int E, int heads, lr_stream_t stream);
int lr_mha_scores_bwd(const float* P, const float* dP, const float* qkv, long long ld, float* dqkv, int B, int T, int E,
                      int heads, lr_stream_t stream);

/* nn.LayerNorm(D) over rows [rows, D] with the residual add of a post-norm nn.TransformerEncoderLayer fused in:
 *   s = a + b (b == NULL: s = a);  y = (s - mean(s)) / sqrt(var(s) + eps) * gamma + beta.   D <= 1024.
 * s [rows, D] and stats [rows, 2] = (mean, rstd) are saved for the backward, which writes ds (the gradient of both
 * a and b) and ACCUMULATES dgamma / dbeta.
 * replaces: norm1 / norm2 of nn.TransformerEncoderLayer (video/models/resnet_trans.py:96-103,
 * audio/models/lstm_resnet_trans_model.py:52-58). */
int lr_layernorm_fwd(const float* a, const float* b, const float* gamma, const float* beta, float eps, float* y, float* s,
                     float* stats, int rows, int D, lr_stream_t stream);
int lr_layernorm_bwd(const float* dy, const float* s, const float* stats, const float* gamma, float* ds, float* dgamma,
                     float* dbeta, int rows, int D, lr_stream_t stream);
/* out[f, t, c] = x[f, c] + r[t, c] (r == NULL: plain repeat): fc_out.unsqueeze(1).repeat(1, T, 1) + pe[:, :T]
 * (audio/models/lstm_resnet_trans_model.py:91-94, lstm_resnet_attn_model.py:78). */
int lr_add_bcast(const float* x, const float* r, float* out, int F, int T, int C, lr_stream_t stream);

/* AttentionFusion of the late triple-fusion model (audio_cues_video/models/late_fusion_mobile.py:6-19) around its
 * attn MLP (two lr_gemm calls): weights = softmax(scores[B,S], dim=1); fused[b,:] = sum_s weights[b,s] stacked[b,s,:].
 * Backward: dstacked = weights * dfused (the MLP's own backward then accumulates onto it), dscores through the softmax.
 * Also the additive attention pooling over S time steps of audio/models/lstm_resnet_attn_model.py:5-14.  S <= 32. */
int lr_attn_fuse_fwd(const float* stacked, const float* scores, float* weights, float* fused, int B, int S, int C,
                     lr_stream_t stream);
int lr_attn_fuse_bwd(const float* stacked, const float* weights, const float* dfused, float* dstacked, float* dscores,
                     int B, int S, int C, lr_stream_t stream);

/* Learnable scalar late fusion (audio_video/models/late_fusion.py:82,92; late_fusion_fast.py:34,58):
 * out = alpha * a + (1 - alpha) * v on n floats; backward: da, dv and dalpha += sum (a - v) * dout. */
int lr_alpha_fuse_fwd(const float* a, const float* v, const float* alpha, float* out, long long n, lr_stream_t stream);
int lr_alpha_fuse_bwd(const float* a, const float* v, const float* alpha, const float* dout, float* da, float* dv,
                      float* dalpha, long long n, lr_stream_t stream);
/* Channel-major flatten of a channels-last activation -- `x.view(B, -1)` on an NCHW tensor
 * (audio_video/models/middle_fusion.py:29): to_nchw != 0: y[b*ldy + c*HW + p] = x[(b*HW + p)*C + c]; to_nchw == 0:
 * the inverse copy (gradient path).  Neither pointer is const: the direction flag picks the destination. */
int lr_flatten_nchw(float* x_nhwc, float* y_nchw, long long ldy, int B, int HW, int C, int to_nchw,
                    lr_stream_t stream);

/* nn.CrossEntropyLoss(mean) forward + gradient (audio_video/train.py:129,65): loss += mean CE (caller zeroes),
 * dlogits = (softmax - onehot) * inv_n (may be NULL), correct += #(argmax == label) (may be NULL). */
int lr_ce_loss(const float* logits, const long long* labels, float* loss, float* dlogits, int* correct, int B, int C,
               float inv_n, lr_stream_t stream);

/* torch.optim.Adam step on flat buffers (audio_video/train.py:130,67).  state: device struct
 * {float step, bc1, bc2_sqrt, lr} (lr_adam_state_bytes()); the step counter advances on the device so the
 * call can be replayed from a CUDA graph.  g is multiplied by grad_scale first (1/world after the allreduce);
 * weight_decay is the coupled L2 form torch.optim.Adam uses. */
size_t lr_adam_state_bytes(void);
int lr_adam_step(float* p, const float* g, float* m, float* v, void* state, long long n, float beta1, float beta2,
                 float eps, float weight_decay, float grad_scale, lr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * bf16 storage (precision "bf16": the north star's "bf16 on 1 B200").  The big per-pixel tensors -- activations and
 * their gradients [rows, C] -- live in HBM as bfloat16 (void* below), which halves the bytes of every HBM-bound kernel;
 * arithmetic inside the kernels stays fp32, BatchNorm statistics stay double sums, parameters / gradients of
 * parameters / Adam state / per-frame vectors (SE gates, pooled features, LSTM) stay fp32.  Every `_h` entry point has
 * the argument meaning of its fp32 namesake above; C must be a multiple of 4 (depthwise: 8).
 * ------------------------------------------------------------------------------------------ */

/* Tensor-core GEMM on bf16 operands (tcgen05.mma kind::f16, fp32 accumulate in TMEM): A and B are bf16 (K-major or
 * MN-major, fetched by TMA with the 128-byte swizzle; lda / ldb multiples of 8), C is bf16 (c_bf16 = 1: activations,
 * dgrad; ldc multiple of 8; R, if given, is bf16 too) or fp32 (c_bf16 = 0: weight gradients -- split-K through TMA
 * reduce-add --, LSTM input projections).  bias fp32.  Same flags and epilogue as lr_gemm_tf32; the BatchNorm statistics
 * are those of the ROUNDED bf16 values. */
int lr_gemm_bf16(const void* A, long long lda, int a_trans, const void* B, long long ldb, int b_trans, void* C,
                 long long ldc, int c_bf16, int M, int N, int K, const float* bias, int act, const void* R,
                 long long ldr, double* stats, int ksplit, lr_stream_t stream);
/* dst[i] = bf16(src[i]): the bf16 shadow of the flat fp32 parameter buffer (refreshed once per step) */
int lr_cast_bf16(const float* src, void* dst, long long n, lr_stream_t stream);

int lr_bn_act_fwd_h(const void* x, const double* stats, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* num_batches_tracked, float eps, float momentum, int act,
                    int training, const void* residual, int res_pre, void* z, long long rows, int C,
                    lr_stream_t stream);
int lr_bn_act_bwd_h(const void* x, const double* stats, const float* gamma, const float* beta,
                    const float* running_mean, const float* running_var, float eps, int act, int training,
                    const void* dz, const void* z_out, void* dres, double* sums, void* dx, float* dgamma,
                    float* dbeta, long long rows, int C, lr_stream_t stream);
int lr_frame_reduce_h(const void* a, const void* g, float* p, int F, int HW, int C, int mode, lr_stream_t stream);
int lr_frame_scale_h(const void* a, const float* s, const float* dp, void* out, int F, int HW, int C,
                     lr_stream_t stream);
int lr_act_bwd_h(void* dy, const void* y, long long n, int act, lr_stream_t stream);
int lr_colsum_h(const void* dY, long long ld, long long M, int N, float* db, lr_stream_t stream);
int lr_dwconv_fwd_h(const void* x, const float* w, void* y, double* stats, int F, int H, int W, int C, int k,
                    int stride, lr_stream_t stream);
int lr_dwconv_dgrad_h(const void* dy, const float* w, void* dx, int F, int H, int W, int C, int k, int stride,
                      lr_stream_t stream);
int lr_dwconv_wgrad_h(const void* dy, const void* x, float* dwt, int F, int H, int W, int C, int k, int stride,
                      lr_stream_t stream);
int lr_im2col_h(const void* x, int is_u8, float scale, int F, int T, long long sb, long long st, long long sc,
                long long sh, long long sw, int Hs, int Ws, int C, int kh, int kw, int stride, int pad,
                int transposed, int Hd, int Wd, void* col, long long ldk, lr_stream_t stream);
int lr_im2col_tap_h(const void* x, int F, int Hs, int Ws, int C, int kh, int kw, int stride, int pad_h, int pad_w,
                    int transposed, int Hd, int Wd, void* col, lr_stream_t stream);
int lr_maxpool_fwd_h(const void* x, void* y, unsigned char* arg, int F, int H, int W, int C, int k, int stride,
                     int pad, lr_stream_t stream);
int lr_maxpool_bwd_h(const void* dy, const unsigned char* arg, void* dx, int F, int H, int W, int C, int k,
                     int stride, int pad, lr_stream_t stream);

/* Implicit-GEMM 3x3 / stride 1 / pad 1 convolution on bf16 channels-last activations (tcgen05.mma kind::f16, fp32
 * accumulate in TMEM; the shifted activation tile of every tap is a 4-D TMA box whose out-of-image part is zero filled:
 * no patch matrix).  Cin, N multiples of 64, W <= 128.
 *   lr_conv3x3_bf16        y[F*H*W, N] = conv(x[F,H,W,Cin], wt[N][9*Cin]) (+ R), wt tap-major (lr_weight_tap mode 0);
 *                          flip = 1: the input gradient, dx = conv(dy, wt[Cin_conv][9*Cout] (lr_weight_tap mode 1)) with
 *                          the taps mirrored; stats: double [2N] sums of the rounded outputs (train-mode BatchNorm).
 *   lr_conv3x3_wgrad_bf16  dwp[Cout][9*Cin] (fp32, tap-major) += dy^T * shifted(x), split over pixel blocks, TMA reduce-add.
 * replaces: torchvision BasicBlock conv1 / conv2 of the ResNet-18 trunks (video/models/resnet_lstm.py:79-110,
 * audio_video/models/ef_cnn_lstm_resnet.py:62-64,86, audio/models/resnet_model.py:12-17), forward, dgrad and wgrad. */
/* up[f, 2 ho, 2 wo, :] = dy[f, ho, wo, :], zero elsewhere, on the input grid [F, Hi, Wi, C] (bf16, C % 8 == 0): the
 * input gradient of a 3x3 / stride-2 / pad-1 convolution (torchvision BasicBlock conv1 of layer2 / 3 / 4,
 * video/models/resnet_lstm.py:79-110) is then lr_conv3x3_bf16(up, mirrored taps, flip = 1) -- no transposed patch matrix. */
int lr_zero_stuff2_h(const void* dy, void* up, int F, int Ho, int Wo, int Hi, int Wi, int C, lr_stream_t stream);
int lr_conv3x3_bf16(const void* x, const void* wt, void* y, const void* R, double* stats, int F, int H, int W, int Cin,
                    int N, int flip, lr_stream_t stream);
int lr_conv3x3_wgrad_bf16(const void* dy, const void* x, float* dwp, int F, int H, int W, int Cin, int Cout,
                          lr_stream_t stream);
/* The same for 3x3 / STRIDE 2 / pad 1 (torchvision BasicBlock conv1 of layer2 / 3 / 4): x [F, Hi, Wi, Cin] with Hi, Wi even,
 * y / dy on the [F, Hi/2, Wi/2] grid.  The input is mapped as the 5-D tensor ((column parity, C), Wi/2, row parity, Hi/2, F),
 * in which every tap of the stride-2 window is a dense box: forward and weight gradient without a patch matrix. */
int lr_conv3x3s2_bf16(const void* x, const void* wt, void* y, double* stats, int F, int Hi, int Wi, int Cin, int N,
                      lr_stream_t stream);
int lr_conv3x3s2_wgrad_bf16(const void* dy, const void* x, float* dwp, int F, int Hi, int Wi, int Cin, int Cout,
                            lr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LIPREAD_B200_H */
