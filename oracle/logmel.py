"""float64 numpy restatement of the reference log-mel frontend.  TEST INFRASTRUCTURE ONLY.

Follows (reference file:line)
  audio/utils/audio_processor.py:9-21   constants, MelSpectrogram(normalized=True)
  audio/utils/audio_processor.py:40-44  truncate / right zero-pad to 20 000 samples
  audio/utils/audio_processor.py:48-52  mel = MelSpectrogram(audio); log(mel + 1e-9)
  audio/utils/audio_processor.py:60-64  (spec - mean) / (std + 1e-9), unbiased std
  audio/data_utils/dataset.py:52        crop [:80, :117] AFTER normalisation
and the published algorithm of the pinned third-party wheel torchaudio==2.6.0
(requirements.txt:89), which is not vendored under /root/reference:
  functional.spectrogram        reflect pad n_fft//2, periodic Hann, onesided rFFT,
                                normalized=True -> divide by sqrt(sum(window^2)), power=2
  functional.melscale_fbanks    HTK mel scale, triangular filters, norm=None
Pinned against the reference's own output by tests/golden/make_golden.py
(tests/test_oracle_golden.py checks it).
"""
import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_MELS = 80
N_FREQS = N_FFT // 2 + 1          # 201
TARGET_SAMPLES = int(1.25 * SAMPLE_RATE)   # 20000  (audio_processor.py:14)
N_FRAMES = 1 + TARGET_SAMPLES // HOP       # 126
N_OUT = 117                                 # av_config.yaml:6 audio_input_size


def hann_window(n=N_FFT):
    """torch.hann_window(n, periodic=True) in float64."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def _hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + f / 700.0)


def _mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (m / 2595.0) - 1.0)


def melscale_fbanks(n_freqs=N_FREQS, f_min=0.0, f_max=SAMPLE_RATE / 2.0, n_mels=N_MELS,
                    sample_rate=SAMPLE_RATE):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') -> (n_freqs, n_mels)."""
    all_freqs = np.linspace(0.0, sample_rate // 2, n_freqs)
    m_pts = np.linspace(_hz_to_mel_htk(f_min), _hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = _mel_to_hz_htk(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]                       # (n_mels+1,)
    slopes = f_pts[None, :] - all_freqs[:, None]          # (n_freqs, n_mels+2)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))


def pad_or_truncate(x, n=TARGET_SAMPLES):
    """audio_processor.py:40-44."""
    x = np.asarray(x)
    if x.shape[-1] > n:
        return x[..., :n]
    if x.shape[-1] < n:
        pad = [(0, 0)] * (x.ndim - 1) + [(0, n - x.shape[-1])]
        return np.pad(x, pad)
    return x


def frames(x):
    """Reflect-pad by n_fft//2 and cut hop-strided frames: (..., S) -> (..., T, n_fft)."""
    x = np.asarray(x, dtype=np.float64)
    half = N_FFT // 2
    xp = np.concatenate([x[..., half:0:-1], x, x[..., -2:-half - 2:-1]], axis=-1)
    T = 1 + x.shape[-1] // HOP
    idx = HOP * np.arange(T)[:, None] + np.arange(N_FFT)[None, :]
    return xp[..., idx]


def log_mel(x, window=None, fb=None):
    """compute_melspectrogram: (..., 20000) -> (..., 80, 126) float64."""
    w = hann_window() if window is None else np.asarray(window, dtype=np.float64)
    fbm = melscale_fbanks() if fb is None else np.asarray(fb, dtype=np.float64)
    fr = frames(x) * w
    spec = np.fft.rfft(fr, axis=-1) / np.sqrt(np.sum(w * w))
    power = spec.real ** 2 + spec.imag ** 2               # (..., T, 201)
    mel = power @ fbm                                     # (..., T, 80)
    return np.log(np.swapaxes(mel, -1, -2) + 1e-9)


def normalize(spec):
    """normalize_spectrogram over ALL values of one clip (last two axes), unbiased std."""
    mean = spec.mean(axis=(-2, -1), keepdims=True)
    std = spec.std(axis=(-2, -1), keepdims=True, ddof=1)
    return (spec - mean) / (std + 1e-9)


def logmel_frontend(x, window=None, fb=None, n_out=N_OUT):
    """Whole frontend as the datasets apply it: log-mel -> normalise -> crop [:80, :n_out]."""
    return normalize(log_mel(x, window, fb))[..., :N_MELS, :n_out]
