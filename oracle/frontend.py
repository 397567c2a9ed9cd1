"""fp32 torch/torchaudio port of the reference audio + video frontends.  TEST INFRASTRUCTURE ONLY.

This is the call sequence the reference itself executes on the CPU, restated without the
file decoding around it (pydub / librosa are not on the path's arithmetic):
  audio/utils/audio_processor.py:15-21   torchaudio.transforms.MelSpectrogram(16 kHz, 400, 160, 80, normalized=True)
  audio/utils/audio_processor.py:48-52   compute_melspectrogram
  audio/utils/audio_processor.py:60-64   normalize_spectrogram
  audio_video/data_utils/dataset_av.py:58-66   process -> normalise -> crop [:80, :117] -> float
  audio_video/data_utils/dataset_av.py:70-71   uint8 (T,H,W,C) -> float32 / 255 -> permute(3,0,1,2)
It is (a) the fp32 parity target of the CUDA log-mel kernel and (b) the CPU arm that
bench.py times as ``cpu_baseline`` / ``--impl reference`` ("kind": "port").
"""
import numpy as np
import torch
import torchaudio


class AudioProcessorPort:
    def __init__(self, sample_rate=16000, n_mels=80, n_fft=400, hop_length=160, target_duration=1.25):
        self.sample_rate = sample_rate
        self.n_mels = n_mels
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.target_samples = int(target_duration * sample_rate)
        self.mel_transform = torchaudio.transforms.MelSpectrogram(
            sample_rate=sample_rate, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels, normalized=True)

    @property
    def window(self):
        return self.mel_transform.spectrogram.window

    @property
    def fb(self):
        return self.mel_transform.mel_scale.fb

    def pad_or_truncate(self, audio):
        if audio.size(0) > self.target_samples:
            audio = audio[:self.target_samples]
        elif audio.size(0) < self.target_samples:
            audio = torch.nn.functional.pad(audio, (0, self.target_samples - audio.size(0)))
        return audio

    def compute_melspectrogram(self, audio):
        return torch.log(self.mel_transform(audio) + 1e-9)

    def normalize_spectrogram(self, spec):
        return (spec - spec.mean()) / (spec.std() + 1e-9)

    def clip_frontend(self, audio, n_out=117):
        """One clip, exactly as dataset_av.py:58-66 (reference semantics: a per-clip call)."""
        spec = self.normalize_spectrogram(self.compute_melspectrogram(audio))
        return spec[:80, :n_out].float()

    def batch_frontend_loop(self, wav, n_out=117):
        """(B, S) -> (B, 80, n_out) by the reference's per-clip loop + default collate."""
        return torch.stack([self.clip_frontend(w, n_out) for w in wav])

    def batch_frontend_batched(self, wav, n_out=117):
        """Same transform applied to the whole batch at once (statistics still per clip)."""
        spec = self.compute_melspectrogram(wav)
        mean = spec.mean(dim=(1, 2), keepdim=True)
        std = spec.std(dim=(1, 2), keepdim=True)
        return ((spec - mean) / (std + 1e-9))[:, :80, :n_out].contiguous()


def lips_u8_to_model_input(lips_u8):
    """(B, T, H, W, 3) uint8 -> (B, 3, T, H, W) float32, dataset_av.py:70-71 + default collate."""
    if isinstance(lips_u8, np.ndarray):
        lips_u8 = torch.from_numpy(lips_u8)
    return (lips_u8.to(torch.float32) / 255.0).permute(0, 4, 1, 2, 3).contiguous()
