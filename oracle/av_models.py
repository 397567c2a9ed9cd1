"""torch / torchvision restatement of the AV fusion models on the hot path.  TEST INFRASTRUCTURE ONLY.

The reference's models are thin compositions of third-party modules that are not vendored
under /root/reference (torchvision==0.21.0 ``mobilenet_v3_small``, torch==2.6.0 ``nn.LSTM`` /
``nn.Linear`` / ``nn.Conv2d``; requirements.txt:88,90).  This file restates the composition --
same sub-module names (so ``state_dict`` keys match), same construction order (so a seeded
init draws the same random numbers), same forward arithmetic -- with ``weights=None`` because
there is no network for the ImageNet checkpoint:

  MidFusionFastOracle      audio_video/models/middle_fusion_fast.py:5-39
  EarlyFusionMobileNetOracle  audio_video/models/early_fusion.py:14-110  (dropout p passed in)

``tests/golden/make_golden.py`` imports the *real* reference modules in the build container and
records their outputs; ``tests/test_oracle_golden.py`` pins these restatements to them.
"""
import torch
import torch.nn as nn
from torchvision.models import mobilenet_v3_small


class DictConfig:
    """Anything with .get('dotted.key', default) is a valid reference config (config/config.py:41-61)."""

    def __init__(self, d=None):
        self.d = d or {}

    def get(self, key, default=None):
        cur = self.d
        for part in key.split("."):
            if isinstance(cur, dict) and part in cur:
                cur = cur[part]
            else:
                return default
        return cur


def _trunk():
    net = mobilenet_v3_small(weights=None)
    net.classifier = nn.Identity()
    return net


def _frames(video):
    b, c, t, h, w = video.shape
    return video.permute(0, 2, 1, 3, 4).contiguous().view(b * t, c, h, w), b, t


class MidFusionFastOracle(nn.Module):
    def __init__(self, num_classes, config=None):
        super().__init__()
        config = config or DictConfig()
        in_ch = config.get("dataset.audio_channels", 1)
        feat = config.get("model.audio_feature_dim", 128)
        self.audio_cnn = nn.Sequential(nn.Conv2d(in_ch, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2))
        self.audio_fc = nn.Linear(16 * 40 * 58, feat)
        self.video_cnn = _trunk()
        self.video_lstm = nn.LSTM(576, 128, 1, batch_first=True, bidirectional=True)
        self.classifier = nn.Sequential(nn.Linear(128 + 256, 256), nn.ReLU(), nn.Linear(256, num_classes))

    def forward(self, audio, video):
        a = self.audio_cnn(audio.unsqueeze(1))
        a = self.audio_fc(a.flatten(1))
        frames, b, t = _frames(video)
        seq, _ = self.video_lstm(self.video_cnn(frames).view(b, t, -1))
        return self.classifier(torch.cat([a, seq[:, -1]], dim=1))


class EarlyFusionMobileNetOracle(nn.Module):
    """early_fusion.py:14-110.  ``lstm_dropout`` / ``head_dropout`` default to the reference's 0.2 / 0.3;
    parity tests pass 0.0 (bit-matching torch's Philox stream is not a goal)."""

    class _Audio(nn.Module):
        def __init__(self, config):
            super().__init__()
            cin = config.get("dataset.audio_channels", 1)
            dim = config.get("model.audio_feature_dim", 256)
            layers = []
            for i, (ci, co) in enumerate([(cin, 32), (32, 64), (64, 128)]):
                layers += [nn.Conv2d(ci, co, 3, padding=1), nn.BatchNorm2d(co), nn.ReLU(),
                           nn.MaxPool2d((2, 2)) if i < 2 else nn.AdaptiveAvgPool2d((1, 1))]
            self.cnn = nn.Sequential(*layers)
            self.fc = nn.Linear(128, dim)
            self.output_dim = dim

        def forward(self, x):
            return self.fc(self.cnn(x).flatten(1))

    class _Video(nn.Module):
        def __init__(self, config, dropout):
            super().__init__()
            hid = config.get("video.lstm_hidden", 256)
            self.cnn = _trunk()
            self.lstm = nn.LSTM(576, hid, 2, batch_first=True, bidirectional=True, dropout=dropout)
            self.output_dim = 2 * hid

        def forward(self, x):
            frames, b, t = _frames(x)
            seq, _ = self.lstm(self.cnn(frames).view(b, t, -1))
            return seq[:, -1]

    def __init__(self, num_classes, config=None, lstm_dropout=0.2, head_dropout=0.3):
        super().__init__()
        config = config or DictConfig()
        self.audio_encoder = self._Audio(config)
        self.video_encoder = self._Video(config, lstm_dropout)
        dim = self.audio_encoder.output_dim + self.video_encoder.output_dim
        self.classifier = nn.Sequential(nn.Linear(dim, 512), nn.ReLU(), nn.Dropout(head_dropout),
                                        nn.Linear(512, num_classes))

    def forward(self, audio, video):
        a = self.audio_encoder(audio.unsqueeze(1))
        v = self.video_encoder(video)
        return self.classifier(torch.cat([a, v], dim=1))


def train_step(model, optimizer, audio, video, labels):
    """One iteration of audio_video/train.py:61-72 (zero_grad, forward, CE mean, backward, step)."""
    optimizer.zero_grad()
    logits = model(audio, video)
    loss = nn.functional.cross_entropy(logits, labels)
    loss.backward()
    optimizer.step()
    return logits.detach(), float(loss.item())
