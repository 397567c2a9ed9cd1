"""torch / torchvision restatement of the AV fusion models on the hot path.  TEST INFRASTRUCTURE ONLY.

The reference's models are thin compositions of third-party modules that are not vendored
under /root/reference (torchvision==0.21.0 ``mobilenet_v3_small``, torch==2.6.0 ``nn.LSTM`` /
``nn.Linear`` / ``nn.Conv2d``; requirements.txt:88,90).  This file restates the composition --
same sub-module names (so ``state_dict`` keys match), same construction order (so a seeded
init draws the same random numbers), same forward arithmetic -- with ``weights=None`` because
there is no network for the ImageNet checkpoint:

  MidFusionFastOracle      audio_video/models/middle_fusion_fast.py:5-39
  EarlyFusionMobileNetOracle  audio_video/models/early_fusion.py:14-110  (dropout p passed in)
  EarlyFusionResNetOracle     audio_video/models/ef_cnn_lstm_resnet.py:14-127
  ResNet2DBiLSTMOracle        video/models/resnet_lstm.py:56-156
  AudioResNetOracle           audio/models/resnet_model.py:5-39
  LateFusionMobileOracle      audio_cues_video/models/late_fusion_mobile.py:6-107
  LateFusionResNetOracle      audio_cues_video/models/late_fusion_resnet.py:6-99
  MobileNetLSTMOracle         video/models/mobilenet_lstm.py:18-68
  VGGLSTMOracle               video/models/vgg_lstm.py:14-88
  CNNOnlyOracle               video/models/cnn.py:5-69
  AudioResNetLSTMOracle       audio/models/resnet_lstm_model.py:5-59
  VGGAudioOracle              audio/models/vgg_model.py:5-58
  VGGLstmAudioOracle          audio/models/vgg_lstm_model.py:5-75
  LSTMResNetOracle            audio/models/lstm_resnet_model.py:5-71
  LSTMResNetAttnOracle        audio/models/lstm_resnet_attn_model.py:17-88
  ResNet2DAttentionOracle     video/models/resnet_attn.py:38-111
  ResNet2DTransformerOracle   video/models/resnet_trans.py:45-129
  ShuffleNet2DBiLSTMOracle    video/models/shufflenet_lstm.py:27-109
  LSTMResNetTransOracle       audio/models/lstm_resnet_trans_model.py:22-104
  AttentionFusionACVOracle    audio_cues_video/models/{middle_fusion_mobile,middle_fusion_resnet,early_fusion_mobile,early_fusion_resnet}.py
  LateFusionAVMobileNetOracle audio_video/models/late_fusion.py:10-93
  MidFusionAVMobileNetOracle  audio_video/models/middle_fusion.py:11-85
  EarlyFusionFastOracle       audio_video/models/early_fusion_fast.py:6-76
  LateFusionFastOracle        audio_video/models/late_fusion_fast.py:5-59

``tests/golden/make_golden.py`` imports the *real* reference modules in the build container and
records their outputs; ``tests/test_oracle_golden.py`` pins these restatements to them.
"""
import torch
import torch.nn as nn
from torchvision.models import mobilenet_v2, mobilenet_v3_small, resnet18, vgg11_bn, vgg13_bn, vgg16_bn, vgg19_bn


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


class EarlyFusionResNetOracle(EarlyFusionMobileNetOracle):
    """ef_cnn_lstm_resnet.py:14-127: the same composition with a ResNet-18 video trunk (fc = Identity)."""

    class _Video(nn.Module):
        def __init__(self, config, dropout):
            super().__init__()
            hid = config.get("video.lstm_hidden", 256)
            base = resnet18(weights=None)
            base.fc = nn.Identity()
            self.cnn = base
            self.lstm = nn.LSTM(512, hid, 2, batch_first=True, bidirectional=True, dropout=dropout)
            self.output_dim = 2 * hid

        def forward(self, x):
            frames, b, t = _frames(x)
            seq, _ = self.lstm(self.cnn(frames).view(b, t, -1))
            return seq[:, -1]


class _TimeDistributed(nn.Module):
    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, x):
        frames, b, t = _frames(x)
        return self.module(frames).view(b, t, -1)


class ResNet2DBiLSTMOracle(nn.Module):
    """video/models/resnet_lstm.py:56-156 (model.resnet_version 18 / 34 / 50, :79-86): conv1 re-initialised, CNN registered
    twice (cnn_features / time_distributed_cnn.module.0), 2-layer BiLSTM(512 [2048 for resnet50], feature_dim/2),
    x[:, -1] -> ReLU -> Dropout -> fc."""

    def __init__(self, num_classes, config=None):
        super().__init__()
        config = config or DictConfig()
        feature_dim = config.get("model.feature_dim", 1024)
        dropout = config.get("model.dropout", 0.5)
        version = config.get("model.resnet_version", 18)
        if version not in (18, 34, 50):
            raise ValueError(f"Unsupported ResNet version: {version}")
        import torchvision.models as tvm
        base = {18: tvm.resnet18, 34: tvm.resnet34, 50: tvm.resnet50}[version](weights=None)
        base.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.cnn_features = nn.Sequential(*list(base.children())[:-2])
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        # resnet_lstm.py:98-103: the LSTM input size is measured with a dummy pass through the freshly built CNN, which
        # is in train mode -- every BatchNorm's running statistics move once and num_batches_tracked becomes 1
        with torch.no_grad():
            cnn_dim = self.global_pool(self.cnn_features(torch.zeros(1, 3, 44, 44))).view(-1).size(0)
        self.time_distributed_cnn = _TimeDistributed(nn.Sequential(self.cnn_features, self.global_pool, nn.Flatten()))
        self.bilstm = nn.LSTM(cnn_dim, feature_dim // 2, num_layers=2, bidirectional=True, batch_first=True,
                              dropout=dropout if dropout > 0 else 0)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.fc = nn.Linear(feature_dim, num_classes)

    def forward(self, x):
        x, _ = self.bilstm(self.time_distributed_cnn(x))
        return self.fc(self.dropout(self.relu(x[:, -1, :])))


class ResNet2DAttentionOracle(nn.Module):
    """video/models/resnet_attn.py:38-111: ResNet-18 per frame (conv1 re-initialised, constructor-time dummy pass),
    proj_in, nn.MultiheadAttention over time, mean over time, ReLU, Dropout, fc."""

    class _TemporalAttention(nn.Module):
        def __init__(self, embed_dim, num_heads, dropout):
            super().__init__()
            self.attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout, batch_first=True)

        def forward(self, x):
            return self.attn(x, x, x)[0]

    def __init__(self, num_classes, config=None):
        super().__init__()
        config = config or DictConfig()
        base = resnet18(weights=None)
        base.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.cnn_features = nn.Sequential(*list(base.children())[:-2])
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        with torch.no_grad():                                              # :63-67
            self.global_pool(self.cnn_features(torch.zeros(1, 3, 44, 44)))
        self.time_cnn = _TimeDistributed(nn.Sequential(self.cnn_features, self.global_pool, nn.Flatten()))
        attn_dim = config.get("model.attention_dim", 512)
        dropout = config.get("model.dropout", 0.3)
        self.proj_in = nn.Linear(512, attn_dim)
        self.attention = self._TemporalAttention(attn_dim, config.get("model.num_heads", 4), dropout)
        self.dropout = nn.Dropout(dropout)
        self.relu = nn.ReLU()
        self.fc = nn.Linear(attn_dim, num_classes)

    def forward(self, x):
        x = self.attention(self.proj_in(self.time_cnn(x))).mean(dim=1)
        return self.fc(self.dropout(self.relu(x)))


class ShuffleNet2DBiLSTMOracle(nn.Module):
    """video/models/shufflenet_lstm.py:27-109: ShuffleNetV2 (0.5x) conv1 .. conv5 per frame (constructor-time dummy
    pass), 2-layer BiLSTM, x[:, -1] -> ReLU -> Dropout -> fc."""

    def __init__(self, num_classes, config=None):
        super().__init__()
        from torchvision.models import shufflenet_v2_x0_5, shufflenet_v2_x1_0
        config = config or DictConfig()
        version = config.get("model.shufflenet_version", "0.5x")
        base = shufflenet_v2_x0_5(weights=None) if version == "0.5x" else shufflenet_v2_x1_0(weights=None)
        self.cnn_features = nn.Sequential(base.conv1, base.maxpool, base.stage2, base.stage3, base.stage4, base.conv5)
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        with torch.no_grad():                                              # :60-64
            dim = self.global_pool(self.cnn_features(torch.zeros(1, 3, 44, 44))).view(-1).shape[0]
        self.time_cnn = _TimeDistributed(nn.Sequential(self.cnn_features, self.global_pool, nn.Flatten()))
        feature_dim = config.get("model.feature_dim", 512)
        dropout = config.get("model.dropout", 0.4)
        self.lstm = nn.LSTM(dim, feature_dim // 2, num_layers=2, batch_first=True, bidirectional=True, dropout=dropout)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.fc = nn.Linear(feature_dim, num_classes)

    def forward(self, x):
        x, _ = self.lstm(self.time_cnn(x))
        return self.fc(self.dropout(self.relu(x[:, -1])))


class ResNet2DTransformerOracle(nn.Module):
    """video/models/resnet_trans.py:45-129: ResNet-18 per frame, proj_in, sinusoidal positions, 2-layer post-norm
    TransformerEncoder (dim_feedforward = 4 * dim), mean over time, ReLU, Dropout, fc."""

    def __init__(self, num_classes, config=None):
        super().__init__()
        config = config or DictConfig()
        base = resnet18(weights=None)
        base.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.cnn_features = nn.Sequential(*list(base.children())[:-2])
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        with torch.no_grad():                                              # :69-73
            self.global_pool(self.cnn_features(torch.zeros(1, 3, 44, 44)))
        self.time_cnn = _TimeDistributed(nn.Sequential(self.cnn_features, self.global_pool, nn.Flatten()))
        dim = config.get("model.transformer_dim", 256)
        dropout = config.get("model.dropout", 0.2)
        self.proj_in = nn.Linear(512, dim)
        pe = torch.zeros(200, dim)
        pos = torch.arange(0, 200).unsqueeze(1)
        div = torch.exp(torch.arange(0, dim, 2) * (-torch.log(torch.tensor(10000.0)) / dim))
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.pos_encoding = nn.Module()
        self.pos_encoding.pe = pe.unsqueeze(0)
        layer = nn.TransformerEncoderLayer(d_model=dim, nhead=config.get("model.num_heads", 4), dropout=dropout,
                                           batch_first=True, dim_feedforward=dim * 4)
        self.transformer = nn.TransformerEncoder(layer, num_layers=config.get("model.num_layers", 2))
        self.dropout = nn.Dropout(dropout)
        self.relu = nn.ReLU()
        self.fc = nn.Linear(dim, num_classes)

    def forward(self, x):
        x = self.proj_in(self.time_cnn(x))
        x = self.transformer(x + self.pos_encoding.pe[:, :x.size(1), :]).mean(dim=1)
        return self.fc(self.dropout(self.relu(x)))


class AudioResNetOracle(nn.Module):
    """audio/models/resnet_model.py:5-39: resnet18 with a 1-channel conv1 and fc = Linear-BN1d-ReLU-Dropout-Linear."""

    def __init__(self, num_classes=40, dropout_rate=0.5, use_batchnorm=True):
        super().__init__()
        self.use_bn = use_batchnorm
        self.resnet = resnet18(weights=None)
        self.resnet.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        layers = [nn.Linear(self.resnet.fc.in_features, 512)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(512))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate), nn.Linear(512, num_classes)])
        self.resnet.fc = nn.Sequential(*layers)

    def forward(self, x):
        return self.resnet(x.unsqueeze(1))


class LateFusionMobileOracle(nn.Module):
    """audio_cues_video/models/late_fusion_mobile.py:84-107 (pretrained=False): ResNet-18 audio encoder, cue MLP with
    BatchNorm1d, MobileNetV2 + 2-layer BiLSTM video encoder, per-modality classifiers, attention fusion of the logits."""

    class _Attn(nn.Module):
        def __init__(self, dim):
            super().__init__()
            self.attn = nn.Sequential(nn.Linear(dim, dim // 2), nn.ReLU(), nn.Linear(dim // 2, 1))

        def forward(self, feats):
            stacked = torch.stack(feats, dim=1)
            weights = torch.softmax(self.attn(stacked).squeeze(-1), dim=1)
            return (stacked * weights.unsqueeze(-1)).sum(dim=1), weights

    class _Video(nn.Module):
        def __init__(self, feature_dim, dropout):
            super().__init__()
            base = mobilenet_v2(weights=None)
            base.classifier = nn.Identity()
            self.cnn = nn.Sequential(base.features, nn.AdaptiveAvgPool2d(1), nn.Flatten())
            self.td = _TimeDistributed(self.cnn)
            self.lstm = nn.LSTM(1280, feature_dim // 2, num_layers=2, bidirectional=True, batch_first=True, dropout=dropout)
            self.output_dim = feature_dim

        def forward(self, x):
            x, _ = self.lstm(self.td(x))
            return x[:, -1, :]

    class _Audio(nn.Module):
        def __init__(self):
            super().__init__()
            net = resnet18(weights=None)
            net.conv1 = nn.Conv2d(1, 64, 7, 2, 3, bias=False)
            net.fc = nn.Identity()
            self.enc = net

        def forward(self, x):
            return self.enc(x.unsqueeze(1))

    class _Cue(nn.Module):
        def __init__(self, input_dim):
            super().__init__()
            self.net = nn.Sequential(nn.Linear(input_dim, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Linear(256, 256))

        def forward(self, x):
            return self.net(x)

    def __init__(self, num_classes, cue_dim=768, vdim=256, lstm_dropout=0.3):
        super().__init__()
        self.audio = self._Audio()
        self.cue = self._Cue(cue_dim)
        self.video = self._Video(vdim, lstm_dropout)
        self.afc = nn.Linear(512, num_classes)
        self.cfc = nn.Linear(256, num_classes)
        self.vfc = nn.Linear(vdim, num_classes)
        self.attn = self._Attn(num_classes)

    def forward(self, mel, cue, lip):
        a = self.afc(self.audio(mel))
        c = self.cfc(self.cue(cue))
        v = self.vfc(self.video(lip))
        fused, _ = self.attn([a, c, v])
        return fused


class LateFusionResNetOracle(LateFusionMobileOracle):
    """audio_cues_video/models/late_fusion_resnet.py:76-99: the same composition with a ResNet-18 video trunk
    (nn.Sequential(resnet), fc = Identity) and the heads registered in the order afc, vfc, cfc."""

    class _Video(nn.Module):
        def __init__(self, feature_dim, dropout):
            super().__init__()
            resnet = resnet18(weights=None)
            resnet.fc = nn.Identity()
            self.cnn = nn.Sequential(resnet)
            self.td = _TimeDistributed(self.cnn)
            self.lstm = nn.LSTM(512, feature_dim // 2, num_layers=2, bidirectional=True, batch_first=True, dropout=dropout)
            self.output_dim = feature_dim

        def forward(self, x):
            x, _ = self.lstm(self.td(x))
            return x[:, -1, :]

    def __init__(self, num_classes, cue_dim=768, vdim=256, lstm_dropout=0.3):
        nn.Module.__init__(self)
        self.audio = self._Audio()
        self.cue = self._Cue(cue_dim)
        self.video = self._Video(vdim, lstm_dropout)
        self.afc = nn.Linear(512, num_classes)
        self.vfc = nn.Linear(vdim, num_classes)
        self.cfc = nn.Linear(256, num_classes)
        self.attn = self._Attn(num_classes)


class _SafeCheckpoint(nn.Module):
    """early_fusion_mobile.py:62-72.  Recomputation never changes values, and it does not engage at all for a frozen
    backbone whose input needs no gradient -- the only way the reference uses it."""

    def __init__(self, module, enabled=True):
        super().__init__()
        self.module = module
        self.enabled = bool(enabled)

    def forward(self, x):
        return self.module(x)


class _TimeDistributedChunked(nn.Module):
    """early_fusion_mobile.py:31-56: chunk_size time steps of every clip per CNN call."""

    def __init__(self, module, chunk_size=4):
        super().__init__()
        self.module = module
        self.chunk_size = int(chunk_size)

    def forward(self, x):
        B, C, T, H, W = x.shape
        outs = []
        for i in range(0, T, self.chunk_size):
            frames = x[:, :, i:min(i + self.chunk_size, T)].permute(0, 2, 1, 3, 4).reshape(-1, C, H, W)
            out = self.module(frames)
            outs.append(out.view(B, out.size(0) // B, -1))
        return torch.cat(outs, dim=1)


class AttentionFusionACVOracle(nn.Module):
    """The four early / middle audio+cue+video attention models of audio_cues_video/models/:
      kind "middle_mobile"  middle_fusion_mobile.py:84-110    trainable encoders, 2-layer BiLSTM, cls with BatchNorm1d
      kind "middle_resnet"  middle_fusion_resnet.py:164-191   frozen ResNet-18 encoders, chunked TimeDistributed
      kind "early_mobile"   early_fusion_mobile.py:179-213    frozen audio ResNet-18 + MobileNetV2 features, cp, classifier
      kind "early_resnet"   early_fusion_resnet.py:158-191    the same with a frozen ResNet-18 video trunk"""

    def __init__(self, kind, num_classes, cue_dim=768, vdim=256, lstm_dropout=0.3, cue_dropout=0.3, head_dropout=0.4):
        super().__init__()
        self.kind = kind
        frozen = kind != "middle_mobile"
        early = kind.startswith("early")
        # audio
        net = resnet18(weights=None)
        net.conv1 = nn.Conv2d(1, 64, 7, 2, 3, bias=False)
        net.fc = nn.Identity()
        self.audio = nn.Module()
        if frozen:
            for p in net.parameters():
                p.requires_grad = False
            setattr(self.audio, "encoder" if early else "enc", _SafeCheckpoint(net))
        else:
            self.audio.enc = net
        # cue
        self.cue = nn.Module()
        if early:
            self.cue.net = nn.Sequential(nn.Linear(cue_dim, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(cue_dropout),
                                         nn.Linear(256, 256), nn.ReLU())
        else:
            self.cue.net = nn.Sequential(nn.Linear(cue_dim, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Linear(256, 256))
        # video
        self.video = nn.Module()
        if kind.endswith("mobile"):
            base = mobilenet_v2(weights=None)
            base.classifier = nn.Identity()
            seq = nn.Sequential(base.features, nn.AdaptiveAvgPool2d(1), nn.Flatten())
            if frozen:
                for p in base.features.parameters():
                    p.requires_grad = False
            width = 1280
        else:
            base = resnet18(weights=None)
            base.fc = nn.Identity()
            for p in base.parameters():
                p.requires_grad = False
            seq = nn.Sequential(base)
            width = 512
        if frozen:
            self.video.cnn = _SafeCheckpoint(seq)
            self.video.td = _TimeDistributedChunked(self.video.cnn, 4)
            self.video.lstm = nn.LSTM(width, vdim // 2, num_layers=1, bidirectional=True, batch_first=True, dropout=0.0)
        else:
            self.video.cnn = seq
            self.video.td = _TimeDistributed(seq)
            self.video.lstm = nn.LSTM(width, vdim // 2, num_layers=2, bidirectional=True, batch_first=True, dropout=lstm_dropout)
        self.ap = nn.Linear(512, 256)
        self.vp = nn.Linear(vdim, 256)
        if early:
            self.cp = nn.Linear(256, 256)
        self.attn = LateFusionMobileOracle._Attn(256)
        if early:
            self.classifier = nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.Dropout(head_dropout), nn.Linear(256, num_classes))
        elif kind == "middle_mobile":
            self.cls = nn.Sequential(nn.Linear(256, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Dropout(head_dropout),
                                     nn.Linear(512, num_classes))
        else:
            self.cls = nn.Sequential(nn.Linear(256, 512), nn.ReLU(), nn.Dropout(head_dropout), nn.Linear(512, num_classes))

    def forward(self, mel, cue, lip):
        enc = getattr(self.audio, "enc", None) or self.audio.encoder
        a = self.ap(enc(mel.unsqueeze(1)))
        c = self.cue.net(cue)
        if hasattr(self, "cp"):
            c = self.cp(c)
        seq, _ = self.video.lstm(self.video.td(lip))
        v = self.vp(seq[:, -1, :])
        fused, _ = self.attn([a, c, v])
        return (self.classifier if hasattr(self, "classifier") else self.cls)(fused)


class MobileNetLSTMOracle(nn.Module):
    """video/models/mobilenet_lstm.py:18-68."""

    def __init__(self, num_classes, config=None):
        super().__init__()
        config = config or DictConfig()
        feature_dim = config.get("model.feature_dim", 256)
        dropout = config.get("model.dropout", 0.3)
        base = mobilenet_v2(weights=None)
        base.classifier = nn.Identity()
        self.cnn = nn.Sequential(base.features, nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten())
        self.td = _TimeDistributed(self.cnn)
        self.lstm = nn.LSTM(1280, feature_dim // 2, num_layers=2, bidirectional=True, batch_first=True, dropout=dropout)
        self.relu = nn.ReLU()
        self.drop = nn.Dropout(dropout)
        self.fc = nn.Linear(feature_dim, num_classes)

    def forward(self, x):
        x, _ = self.lstm(self.td(x))
        return self.fc(self.drop(self.relu(x[:, -1, :])))


class VGGLSTMOracle(nn.Module):
    """video/models/vgg_lstm.py:14-88: VGGLite per frame (5 convs + ReLU, 2 max pools, global average), 2-layer BiLSTM."""

    class _VGGLite(nn.Module):
        def __init__(self):
            super().__init__()
            self.features = nn.Sequential(
                nn.Conv2d(3, 32, 3, padding=1), nn.ReLU(True), nn.Conv2d(32, 32, 3, padding=1), nn.ReLU(True), nn.MaxPool2d(2),
                nn.Conv2d(32, 64, 3, padding=1), nn.ReLU(True), nn.Conv2d(64, 64, 3, padding=1), nn.ReLU(True), nn.MaxPool2d(2),
                nn.Conv2d(64, 128, 3, padding=1), nn.ReLU(True), nn.AdaptiveAvgPool2d((1, 1)))

        def forward(self, x):
            return self.features(x).flatten(1)

    def __init__(self, num_classes, config=None):
        super().__init__()
        config = config or DictConfig()
        feature_dim = config.get("model.feature_dim", 256)
        dropout = config.get("model.dropout", 0.5)
        self.td = _TimeDistributed(self._VGGLite())
        self.lstm = nn.LSTM(128, feature_dim // 2, num_layers=2, bidirectional=True, batch_first=True, dropout=dropout)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.fc = nn.Linear(feature_dim, num_classes)

    def forward(self, x):
        x, _ = self.lstm(self.td(x))
        return self.fc(self.dropout(self.relu(x[:, -1, :])))


class CNNOnlyOracle(nn.Module):
    """video/models/cnn.py:5-69: per-frame CNN (3 x conv + BN + ReLU), two Conv1d + BatchNorm1d + ReLU over time, mean."""

    def __init__(self, num_classes, config=None):
        super().__init__()
        config = config or DictConfig()
        self.frame_cnn = nn.Sequential(
            nn.Conv2d(3, 32, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(32), nn.ReLU(inplace=True), nn.MaxPool2d(2),
            nn.Conv2d(32, 64, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.MaxPool2d(2),
            nn.Conv2d(64, 128, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
            nn.AdaptiveAvgPool2d((1, 1)))
        tc = config.get("model.temporal_channels", 128)
        self.temporal_conv = nn.Sequential(
            nn.Conv1d(128, tc, kernel_size=3, padding=1), nn.BatchNorm1d(tc), nn.ReLU(inplace=True),
            nn.Conv1d(tc, tc, kernel_size=3, padding=1), nn.BatchNorm1d(tc), nn.ReLU(inplace=True))
        self.dropout = nn.Dropout(config.get("model.dropout", 0.3))
        self.fc = nn.Linear(tc, num_classes)

    def forward(self, x):
        frames, b, t = _frames(x)
        f = self.frame_cnn(frames).view(b, t, -1).permute(0, 2, 1)
        return self.fc(self.dropout(self.temporal_conv(f).mean(dim=2)))


class AudioResNetLSTMOracle(nn.Module):
    """audio/models/resnet_lstm_model.py:5-59: ResNet-18 features as a length-1 sequence through a 2-layer BiLSTM."""

    def __init__(self, num_classes=40, lstm_hidden=128, lstm_layers=2, dropout_rate=0.3, use_batchnorm=True):
        super().__init__()
        self.use_bn = use_batchnorm
        self.resnet = resnet18(weights=None)
        self.resnet.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.resnet.fc = nn.Identity()
        self.lstm = nn.LSTM(512, lstm_hidden, num_layers=lstm_layers, bidirectional=True, batch_first=True)
        layers = [nn.Linear(2 * lstm_hidden, 256)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(256))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate), nn.Linear(256, num_classes)])
        self.classifier = nn.Sequential(*layers)

    def forward(self, x):
        out, _ = self.lstm(self.resnet(x.unsqueeze(1)).unsqueeze(1))
        return self.classifier(out[:, -1, :])


class VGGAudioOracle(nn.Module):
    """audio/models/vgg_model.py:5-58."""

    def __init__(self, num_classes=40, version=11, dropout_rate=0.5, use_batchnorm=True):
        super().__init__()
        self.use_bn = use_batchnorm
        self.vgg = {11: vgg11_bn, 13: vgg13_bn, 16: vgg16_bn, 19: vgg19_bn}[version](weights=None, init_weights=False)
        self.vgg.features[0] = nn.Conv2d(1, 64, kernel_size=3, padding=1)
        self.adaptive_pool = nn.AdaptiveAvgPool2d((2, 3))
        layers = [nn.Linear(512 * 2 * 3, 256)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(256))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate), nn.Linear(256, num_classes)])
        self.vgg.classifier = nn.Sequential(*layers)

    def forward(self, x):
        x = self.adaptive_pool(self.vgg.features(x.unsqueeze(1)))
        return self.vgg.classifier(torch.flatten(x, 1))


class VGGLstmAudioOracle(nn.Module):
    """audio/models/vgg_lstm_model.py:5-75: vgg features, mean over the width, the height as the LSTM's time axis."""

    def __init__(self, num_classes=40, lstm_hidden_size=128, lstm_layers=2, version=11, dropout_rate=0.3, use_batchnorm=True):
        super().__init__()
        self.use_bn = use_batchnorm
        vgg = {11: vgg11_bn, 13: vgg13_bn, 16: vgg16_bn, 19: vgg19_bn}[version](weights=None, init_weights=False)
        vgg.features[0] = nn.Conv2d(1, 64, kernel_size=3, padding=1)
        self.vgg_features = vgg.features
        self.adaptive_pool = nn.AdaptiveAvgPool2d((None, 1))
        self.cnn_output_dim = 512
        self.lstm = nn.LSTM(512, lstm_hidden_size, num_layers=lstm_layers, bidirectional=True, batch_first=True)
        layers = [nn.Linear(2 * lstm_hidden_size, 128)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(128))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate), nn.Linear(128, num_classes)])
        self.classifier = nn.Sequential(*layers)

    def forward(self, x):
        x = self.adaptive_pool(self.vgg_features(x.unsqueeze(1))).squeeze(-1).permute(0, 2, 1)
        out, _ = self.lstm(x)
        return self.classifier(out[:, -1, :])


class LSTMResNetOracle(nn.Module):
    """audio/models/lstm_resnet_model.py:5-71: every mel row through a BiLSTM as a length-1 sequence, the result as a
    1-channel 80 x 128 image through ResNet-18, fc (+BN1d), a second length-1 BiLSTM, classifier."""

    def __init__(self, num_classes=40, input_size=117, dropout_rate=0.3, use_batchnorm=True):
        super().__init__()
        self.use_bn = use_batchnorm
        self.initial_bilstm = nn.LSTM(input_size, 64, num_layers=2, bidirectional=True, batch_first=True)
        self.resnet = resnet18(weights=None)
        self.resnet.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.resnet.fc = nn.Identity()
        layers = [nn.Linear(512, 256)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(256))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate)])
        self.fc = nn.Sequential(*layers)
        self.final_bilstm = nn.LSTM(256, 128, num_layers=2, bidirectional=True, batch_first=True)
        self._tail(num_classes)

    def _tail(self, num_classes):
        self.classifier = nn.Linear(256, num_classes)

    def forward(self, x):
        b = x.size(0)
        x1, _ = self.initial_bilstm(x.view(b * 80, 117).unsqueeze(1))
        x1 = x1.squeeze(1).view(b, 1, 80, -1)
        out, _ = self.final_bilstm(self.fc(self.resnet(x1)).unsqueeze(1))
        return self.classifier(out[:, -1, :])


class LSTMResNetAttnOracle(LSTMResNetOracle):
    """audio/models/lstm_resnet_attn_model.py:17-88: the fc output repeated over 10 steps, a full 2-layer BiLSTM,
    additive attention pooling over the steps, classifier."""

    class _Attention(nn.Module):
        def __init__(self, input_dim):
            super().__init__()
            self.attn = nn.Linear(input_dim, 1)

        def forward(self, x):
            weights = torch.softmax(self.attn(x).squeeze(-1), dim=1)
            return torch.sum(x * weights.unsqueeze(-1), dim=1), weights

    def _tail(self, num_classes):                            # construction (= RNG) order of the reference: :54-58
        self.attention = self._Attention(256)
        self.classifier = nn.Linear(256, num_classes)

    def forward(self, x):
        b = x.size(0)
        x1, _ = self.initial_bilstm(x.view(b * 80, 117).unsqueeze(1))
        x1 = x1.squeeze(1).view(b, 1, 80, -1)
        out, _ = self.final_bilstm(self.fc(self.resnet(x1)).unsqueeze(1).repeat(1, 10, 1))
        pooled, _ = self.attention(out)
        return self.classifier(pooled)


class LSTMResNetTransOracle(nn.Module):
    """audio/models/lstm_resnet_trans_model.py:22-104: LSTMResNet front with fc -> transformer_dim, the projection repeated
    over seq_len steps + positional encoding (a registered buffer), 2-layer TransformerEncoder (torch defaults:
    dim_feedforward 2048, dropout 0.1), mean over the steps, classifier."""

    def __init__(self, num_classes=40, input_size=117, transformer_dim=256, num_heads=4, num_layers=2, seq_len=10,
                 dropout_rate=0.3, use_batchnorm=True, encoder_dropout=0.1):
        super().__init__()
        import numpy as np
        self.seq_len = seq_len
        self.initial_bilstm = nn.LSTM(input_size, 64, num_layers=2, bidirectional=True, batch_first=True)
        self.resnet = resnet18(weights=None)
        self.resnet.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.resnet.fc = nn.Identity()
        layers = [nn.Linear(512, transformer_dim)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(transformer_dim))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate)])
        self.fc = nn.Sequential(*layers)
        pe = torch.zeros(seq_len, transformer_dim)
        position = torch.arange(0, seq_len).unsqueeze(1).float()
        div_term = torch.exp(torch.arange(0, transformer_dim, 2).float() * (-np.log(10000.0) / transformer_dim))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.pos_encoder = nn.Module()
        self.pos_encoder.register_buffer("pe", pe.unsqueeze(0))
        layer = nn.TransformerEncoderLayer(d_model=transformer_dim, nhead=num_heads, batch_first=True, dropout=encoder_dropout)
        self.transformer = nn.TransformerEncoder(layer, num_layers=num_layers)
        self.classifier = nn.Linear(transformer_dim, num_classes)

    def forward(self, x):
        b = x.size(0)
        if x.dim() == 4:
            x = x.squeeze(1)
        x1, _ = self.initial_bilstm(x.view(b * x.shape[1], x.shape[2]).unsqueeze(1))
        x1 = x1.squeeze(1).view(b, 1, 80, -1)
        seq = self.fc(self.resnet(x1)).unsqueeze(1).repeat(1, self.seq_len, 1)
        seq = seq + self.pos_encoder.pe[:, :seq.size(1)]
        return self.classifier(self.transformer(seq).mean(dim=1))


class _VideoLstm(nn.Module):
    """MobileNetV3-small + single-layer BiLSTM video encoder; head "hn": cat(h_n[0], h_n[1]) (late_fusion.py:54-62,
    early_fusion_fast.py:44-54); head "last": out[:, -1] (middle_fusion.py:50-57)."""

    def __init__(self, hidden, head):
        super().__init__()
        self.cnn = _trunk()
        self.lstm = nn.LSTM(576, hidden, 1, batch_first=True, bidirectional=True)
        self.output_dim, self.head = 2 * hidden, head

    def forward(self, x):
        frames, b, t = _frames(x)
        seq, (h_n, _) = self.lstm(self.cnn(frames).view(b, t, -1))
        return torch.cat([h_n[0], h_n[1]], dim=1) if self.head == "hn" else seq[:, -1]


class _AudioCnnFc(nn.Module):
    def __init__(self, cnn, fc_in, dim):
        super().__init__()
        self.cnn = cnn
        self.fc = nn.Linear(fc_in, dim)
        self.output_dim = dim

    def forward(self, x):
        if x.dim() == 3:
            x = x.unsqueeze(1)
        return self.fc(self.cnn(x).flatten(1))


class LateFusionAVMobileNetOracle(nn.Module):
    """audio_video/models/late_fusion.py:10-93."""

    def __init__(self, num_classes, config=None):
        super().__init__()
        config = config or DictConfig()
        cin = config.get("dataset.audio_channels", 1)
        cnn = nn.Sequential(nn.Conv2d(cin, 32, 3, padding=1), nn.BatchNorm2d(32), nn.ReLU(), nn.MaxPool2d(2),
                            nn.Conv2d(32, 64, 3, padding=1), nn.BatchNorm2d(64), nn.ReLU(), nn.AdaptiveAvgPool2d((1, 1)))
        self.audio_encoder = _AudioCnnFc(cnn, 64, config.get("model.audio_feature_dim", 256))
        self.video_encoder = _VideoLstm(config.get("video.lstm_hidden", 256), "hn")
        self.audio_classifier = nn.Linear(self.audio_encoder.output_dim, num_classes)
        self.video_classifier = nn.Linear(self.video_encoder.output_dim, num_classes)
        self.alpha = nn.Parameter(torch.tensor(0.5))

    def forward(self, audio, video):
        a = self.audio_classifier(self.audio_encoder(audio))
        v = self.video_classifier(self.video_encoder(video))
        return self.alpha * a + (1 - self.alpha) * v


class MidFusionAVMobileNetOracle(nn.Module):
    """audio_video/models/middle_fusion.py:11-85 (head dropout passed in; the audio map is flattened channel-major)."""

    class _Audio(nn.Module):
        def __init__(self, cin):
            super().__init__()
            self.cnn = nn.Sequential(nn.Conv2d(cin, 32, kernel_size=3, padding=1), nn.BatchNorm2d(32), nn.ReLU(), nn.MaxPool2d(2),
                                     nn.Conv2d(32, 64, kernel_size=3, padding=1), nn.BatchNorm2d(64), nn.ReLU(), nn.MaxPool2d(2))
            self.output_dim = 64 * 20 * 29

        def forward(self, x):
            return self.cnn(x).flatten(1)

    def __init__(self, num_classes, config=None, head_dropout=0.3):
        super().__init__()
        config = config or DictConfig()
        self.audio_encoder = self._Audio(config.get("dataset.audio_channels", 1))
        self.video_encoder = _VideoLstm(config.get("video.lstm_hidden", 256), "last")
        dim = self.audio_encoder.output_dim + self.video_encoder.output_dim
        self.classifier = nn.Sequential(nn.Linear(dim, 512), nn.ReLU(), nn.Dropout(head_dropout), nn.Linear(512, num_classes))

    def forward(self, audio, video):
        return self.classifier(torch.cat([self.audio_encoder(audio.unsqueeze(1)), self.video_encoder(video)], dim=1))


class EarlyFusionFastOracle(nn.Module):
    """audio_video/models/early_fusion_fast.py:6-76."""

    def __init__(self, num_classes, config=None):
        super().__init__()
        config = config or DictConfig()
        cin = config.get("dataset.audio_channels", 1)
        cnn = nn.Sequential(nn.Conv2d(cin, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
                            nn.Conv2d(16, 32, 3, padding=1), nn.ReLU(), nn.AdaptiveAvgPool2d((1, 1)))
        self.audio_encoder = _AudioCnnFc(cnn, 32, config.get("model.audio_feature_dim", 128))
        self.video_encoder = _VideoLstm(config.get("video.lstm_hidden", 128), "hn")
        dim = self.audio_encoder.output_dim + self.video_encoder.output_dim
        self.classifier = nn.Sequential(nn.Linear(dim, 256), nn.ReLU(), nn.Linear(256, num_classes))

    def forward(self, audio, video):
        return self.classifier(torch.cat([self.audio_encoder(audio), self.video_encoder(video)], dim=1))


class LateFusionFastOracle(nn.Module):
    """audio_video/models/late_fusion_fast.py:5-59."""

    def __init__(self, num_classes, config=None):
        super().__init__()
        config = config or DictConfig()
        cin = config.get("dataset.audio_channels", 1)
        dim = config.get("model.audio_feature_dim", 128)
        self.audio_cnn = nn.Sequential(nn.Conv2d(cin, 16, 3, padding=1), nn.ReLU(), nn.AdaptiveAvgPool2d((1, 1)))
        self.audio_fc = nn.Linear(16, dim)
        self.audio_classifier = nn.Linear(dim, num_classes)
        self.video_cnn = _trunk()
        self.video_lstm = nn.LSTM(576, 128, 1, batch_first=True, bidirectional=True)
        self.video_classifier = nn.Linear(256, num_classes)
        self.alpha = nn.Parameter(torch.tensor(0.5))

    def forward(self, audio, video):
        if audio.dim() == 3:
            audio = audio.unsqueeze(1)
        a = self.audio_classifier(self.audio_fc(self.audio_cnn(audio).flatten(1)))
        frames, b, t = _frames(video)
        _, (h_n, _) = self.video_lstm(self.video_cnn(frames).view(b, t, -1))
        v = self.video_classifier(torch.cat([h_n[0], h_n[1]], dim=1))
        return self.alpha * a + (1 - self.alpha) * v


def train_step_generic(model, optimizer, inputs, labels):
    """zero_grad, forward, CrossEntropyLoss(mean), backward, step -- the loop body shared by audio/train.py:67-78,
    video/train.py:93-104, audio_video/train.py:61-67, audio_cues_video/train.py:60-72."""
    optimizer.zero_grad()
    logits = model(*inputs)
    loss = nn.functional.cross_entropy(logits, labels)
    loss.backward()
    optimizer.step()
    return logits.detach(), float(loss.item())


def train_step(model, optimizer, audio, video, labels):
    """One iteration of audio_video/train.py:61-72 (zero_grad, forward, CE mean, backward, step)."""
    optimizer.zero_grad()
    logits = model(audio, video)
    loss = nn.functional.cross_entropy(logits, labels)
    loss.backward()
    optimizer.step()
    return logits.detach(), float(loss.item())
