"""B200 log-mel frontend behind the reference's AudioProcessor surface.

Mirrors audio/utils/audio_processor.py:8-64 (same constructor arguments, same method names and
meaning) but batched and on the GPU: the reference runs torchaudio's MelSpectrogram per clip on the
CPU inside DataLoader workers (audio_video/data_utils/dataset_av.py:58-66); here a whole batch of
raw waveforms goes through ONE CUDA kernel (csrc/logmel.cu).  File decoding (`load_audio`: pydub /
ffmpeg) is outside the hot path and stays with the caller; `pad_or_truncate` keeps its semantics.
"""
import math

import torch

from . import ops


def hann_window(n_fft=400, device=None):
    return torch.hann_window(n_fft, periodic=True, dtype=torch.float32, device=device)


def melscale_fbanks(n_freqs=201, f_min=0.0, f_max=8000.0, n_mels=80, sample_rate=16000):
    """HTK triangular filterbank, norm=None -- the arithmetic (and fp32 rounding) of the
    `mel_scale.fb` buffer owned by the reference's torchaudio transform (audio_processor.py:15-21)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).contiguous()


class AudioProcessor:
    def __init__(self, sample_rate=16000, n_mels=80, n_fft=400, hop_length=160, target_duration=1.25,
                 device="cuda"):
        if (sample_rate, n_mels, n_fft, hop_length, int(target_duration * sample_rate)) != (16000, 80, 400, 160, 20000):
            raise ValueError("the B200 log-mel kernel is specialised for the reference's constants "
                             "(16 kHz, 80 mels, n_fft 400, hop 160, 1.25 s)")
        self.sample_rate, self.n_mels, self.n_fft, self.hop_length = sample_rate, n_mels, n_fft, hop_length
        self.target_samples = int(target_duration * sample_rate)
        self.device = torch.device(device)
        self.window = hann_window(n_fft).to(self.device)
        self.fb = melscale_fbanks(n_fft // 2 + 1, 0.0, sample_rate / 2.0, n_mels, sample_rate).to(self.device)
        self.plan = ops.logmel_plan(self.window, self.fb)

    # audio_processor.py:40-44
    def pad_or_truncate(self, audio):
        n = audio.shape[-1]
        if n > self.target_samples:
            return audio[..., :self.target_samples]
        if n < self.target_samples:
            return torch.nn.functional.pad(audio, (0, self.target_samples - n))
        return audio

    def _batched(self, audio):
        x = self.pad_or_truncate(audio.to(self.device, torch.float32))
        return (x[None] if x.dim() == 1 else x).contiguous(), x.dim() == 1

    # audio_processor.py:48-52  (accepts (S,) like the reference, or a batch (B, S))
    def compute_melspectrogram(self, audio):
        x, single = self._batched(audio)
        out = ops.logmel(x, self.plan, 126, 1)
        return out[0] if single else out

    # audio_processor.py:60-64  (statistics per clip: over everything but a leading batch dim)
    def normalize_spectrogram(self, spec):
        if spec.dim() == 2:
            return ops.normalize(spec.contiguous()[None])[0]
        return ops.normalize(spec.contiguous())

    def frontend(self, audio, n_out=117):
        """Fused compute_melspectrogram -> normalize_spectrogram -> [:80, :n_out] (dataset_av.py:58-66)."""
        x, single = self._batched(audio)
        out = ops.logmel(x, self.plan, n_out, 0)
        return out[0] if single else out
