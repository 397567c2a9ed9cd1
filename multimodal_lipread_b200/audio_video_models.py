"""Audio-visual fusion models behind the reference's nn.Module surface (audio_video/models/*.py).

Same class / factory names, constructor signature `(num_classes, config)`, `forward(audio, video)` contract and
`state_dict` keys as the reference; the sub-modules (torch / torchvision classes) are kept ONLY as parameter
containers so that seeded initialisation and checkpoints are interchangeable with the reference -- their torch
forward() is never called.  All arithmetic runs through the launch plans of engine.py (hand-written CUDA).

  MidFusionFast / create_mid_fusion_fast                       audio_video/models/middle_fusion_fast.py:5-42
  EarlyFusionAVMobileNet / create_early_fusion_mobilenet_model audio_video/models/early_fusion.py:14-117
  EarlyFusionAV / create_early_fusion_resnet_model             audio_video/models/ef_cnn_lstm_resnet.py:14-133
"""
import torch
import torch.nn as nn
from torchvision.models import mobilenet_v3_small, resnet18

from . import engine
from ._lib import ACT_NONE, ACT_RELU
from .model_base import Cfg, ModelPlan, PlanModel, N_MELS, N_FRAMES_OUT, video_layout

_Cfg = Cfg
_video_layout = video_layout


class MidFusionPlan(ModelPlan):
    """Launch plan of MidFusionFast at one batch shape."""

    def build(self, m, spec):
        B = self.B
        flat, with_backward = self.flat, self.with_backward
        video, layout, scale = self.video_input()
        T, H, W = layout[2], layout[3], layout[4]
        mel = self.audio_input()

        # ---- audio branch: conv+relu+pool -> [B,37120] -> audio_fc -> fused[:, 0:FA]
        FA = m.audio_fc.out_features
        HL = m.video_lstm.hidden_size
        FD = FA + 2 * HL
        self.fused = self.alloc(B * FD)
        dfused = self.alloc(B * FD) if with_backward else None
        KA = m.audio_fc.in_features
        a_pool = self.alloc(B * KA)
        a_arg = self.alloc(B * KA, torch.uint8)
        conv = m.audio_cnn[0]
        self.fwd.add("lr_audio_conv_fwd", mel, conv.weight, conv.bias, a_pool, KA, a_arg, B, N_MELS, N_FRAMES_OUT)
        ks = engine._ksplit(B, FA, KA, self.sms)
        if ks > 1:
            self.fwd.add("lr_memset", self.fused, B * FD * 4)          # split-K accumulates onto zeros
        self.linear(a_pool, KA, B, m.audio_fc.weight, m.audio_fc.bias, self.fused, FD, ksplit=ks)
        if with_backward:
            d_pool = self.alloc(B * KA)
            g = self.bgroup()
            self.linear_bwd(g, a_pool, KA, B, m.audio_fc.weight, m.audio_fc.bias, dfused, FD, dx=d_pool, ldx=KA)
            g.add("lr_audio_conv_bwd", mel, d_pool, KA, a_arg, flat.g(conv.weight), flat.g(conv.bias), B, N_MELS,
                  N_FRAMES_OUT, leaf=True)

        # ---- video trunk: MobileNetV3-small features + avgpool -> feat [B*T, 576]
        last = self.mbv3_features(m.video_cnn.features, video, layout, scale, B, T, H, W)
        feat, dfeat = self.avgpool(last)
        # ---- BiLSTM with an out[:, -1] head, written straight into the fusion row (no torch.cat)
        self.bilstm_last(feat, dfeat, last.C, B, T, m.video_lstm, self.fused.data_ptr() + 4 * FA, FD,
                         (dfused.data_ptr() + 4 * FA) if with_backward else 0)
        # ---- classifier: Linear(FD,256)+ReLU -> Linear(256,C)
        logits, dlogits = self.mlp(self.fused, dfused, B, m.classifier)
        self.set_logits(logits, dlogits)


class MidFusionFast(PlanModel):
    """audio_video/models/middle_fusion_fast.py:5-39."""
    INPUTS = ("audio", "video")
    PLAN = MidFusionPlan

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        # construction order == the reference's, so a seeded init draws identical values
        self.audio_cnn = nn.Sequential(
            nn.Conv2d(config.get("dataset.audio_channels", 1), 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2))
        if self.audio_cnn[0].in_channels != 1:
            raise ValueError("the fused audio kernel handles dataset.audio_channels == 1 (the reference default)")
        self.audio_fc = nn.Linear(16 * 40 * 58, config.get("model.audio_feature_dim", 128))
        # the reference loads MobileNet_V3_Small_Weights.IMAGENET1K_V1 from the network; offline the same
        # checkpoint can be passed as `pretrained_state_dict` (torchvision key names)
        base = mobilenet_v3_small(weights=None)
        if pretrained_state_dict is not None:
            base.load_state_dict(pretrained_state_dict)
        base.classifier = nn.Identity()
        self.video_cnn = base
        self.video_lstm = nn.LSTM(576, 128, 1, batch_first=True, bidirectional=True)
        self.classifier = nn.Sequential(nn.Linear(128 + 256, 256), nn.ReLU(), nn.Linear(256, num_classes))


def create_mid_fusion_fast(num_classes, config=None):
    """audio_video/models/middle_fusion_fast.py:41-42."""
    return MidFusionFast(num_classes, config)


# ------------------------------------------------------------------------------------------------------------
# Early fusion: AudioEncoder (3 x [conv3x3 + BN + ReLU], pools) ++ VideoEncoder (CNN trunk + 2-layer BiLSTM)
# ------------------------------------------------------------------------------------------------------------
class AudioEncoder(nn.Module):
    """Parameter container of audio_video/models/early_fusion.py:14-45 (== ef_cnn_lstm_resnet.py:14-47)."""

    def __init__(self, config):
        super().__init__()
        in_channels = config.get("dataset.audio_channels", 1)
        feature_dim = config.get("model.audio_feature_dim", 256)
        self.cnn = nn.Sequential(
            nn.Conv2d(in_channels, 32, kernel_size=3, padding=1), nn.BatchNorm2d(32), nn.ReLU(), nn.MaxPool2d((2, 2)),
            nn.Conv2d(32, 64, kernel_size=3, padding=1), nn.BatchNorm2d(64), nn.ReLU(), nn.MaxPool2d((2, 2)),
            nn.Conv2d(64, 128, kernel_size=3, padding=1), nn.BatchNorm2d(128), nn.ReLU(), nn.AdaptiveAvgPool2d((1, 1)))
        self.fc = nn.Linear(128, feature_dim)
        self.output_dim = feature_dim


def audio_encoder_plan(plan, enc, mel, B, out, ldo, dout):
    """AudioEncoder forward/backward on mel (B,80,117) viewed as a 1-channel NHWC image; writes out[b, 0:D]."""
    if enc.cnn[0].in_channels != 1:
        raise ValueError("AudioEncoder plan handles dataset.audio_channels == 1 (the reference default)")
    mods = list(enc.cnn)
    # (B,80,117) contiguous == NHWC with C = 1
    frames = (mel, (0, B, 1, N_MELS, N_FRAMES_OUT, N_MELS * N_FRAMES_OUT, 0, 0, N_FRAMES_OUT, 1), 1.0)
    cur = None
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Conv2d):
            raw = plan.dense_conv(cur, m, frames=frames if cur is None else None)
            if plan.with_backward:
                plan.dense_conv_bwd(raw)
            a = engine.T2(plan, raw.F, raw.H, raw.W, raw.C)
            plan.bn_act(raw, mods[i + 1], ACT_RELU, a)
            cur = a
            i += 3
        elif isinstance(m, nn.MaxPool2d):
            cur = plan.maxpool(cur, 2, 2, 0)
            i += 1
        elif isinstance(m, nn.AdaptiveAvgPool2d):
            pooled, dpooled = plan.avgpool(cur)
            i += 1
        else:
            raise NotImplementedError(type(m).__name__)
    fc = enc.fc
    plan.linear(pooled, fc.in_features, B, fc.weight, fc.bias, out, ldo)
    if plan.with_backward:
        plan.linear_bwd(plan.bgroup(), pooled, fc.in_features, B, fc.weight, fc.bias, dout, ldo, dx=dpooled,
                        ldx=fc.in_features)


class EarlyFusionPlan(ModelPlan):
    """Plan of EarlyFusionAVMobileNet / EarlyFusionAV: fused = [audio_encoder(audio) | video_encoder(video)]."""

    def build(self, m, spec):
        B = self.B
        wb = self.with_backward
        video, layout, scale = self.video_input()
        T, H, W = layout[2], layout[3], layout[4]
        mel = self.audio_input()
        DA, DV = m.audio_encoder.output_dim, m.video_encoder.output_dim
        FD = DA + DV
        self.fused = self.alloc(B * FD)
        dfused = self.alloc(B * FD) if wb else None
        audio_encoder_plan(self, m.audio_encoder, mel, B, self.fused, FD, dfused)
        ve = m.video_encoder
        if m.backbone == "mobilenet_v3_small":
            last = self.mbv3_features(ve.cnn.features, video, layout, scale, B, T, H, W)
        else:
            last = self.resnet_features(ve.cnn, (video, layout, scale))
        feat, dfeat = self.avgpool(last)
        self.bilstm_last(feat, dfeat, last.C, B, T, ve.lstm, self.fused.data_ptr() + 4 * DA, FD,
                         (dfused.data_ptr() + 4 * DA) if wb else 0)
        logits, dlogits = self.mlp(self.fused, dfused, B, m.classifier)
        self.set_logits(logits, dlogits)


class VideoEncoder(nn.Module):
    """Parameter container of early_fusion.py:51-72 (MobileNetV3-small) / ef_cnn_lstm_resnet.py:53-78 (ResNet-18)."""

    def __init__(self, config, backbone, pretrained_state_dict=None):
        super().__init__()
        lstm_hidden = config.get("video.lstm_hidden", 256)
        if backbone == "mobilenet_v3_small":
            base = mobilenet_v3_small(weights=None)
            feat = 576
            if pretrained_state_dict is not None:
                base.load_state_dict(pretrained_state_dict)
            base.classifier = nn.Identity()
        else:
            base = resnet18(weights=None)
            feat = 512
            if pretrained_state_dict is not None:
                base.load_state_dict(pretrained_state_dict)
            base.fc = nn.Identity()
        self.cnn = base
        self.lstm = nn.LSTM(input_size=feat, hidden_size=lstm_hidden, num_layers=2, batch_first=True,
                            bidirectional=True, dropout=config.get("video.lstm_dropout", 0.2))
        self.output_dim = lstm_hidden * 2


class _EarlyFusionBase(PlanModel):
    INPUTS = ("audio", "video")
    PLAN = EarlyFusionPlan
    backbone = None

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        self.audio_encoder = AudioEncoder(config)
        self.video_encoder = VideoEncoder(config, self.backbone, pretrained_state_dict)
        fusion_dim = self.audio_encoder.output_dim + self.video_encoder.output_dim
        self.classifier = nn.Sequential(
            nn.Linear(fusion_dim, 512), nn.ReLU(), nn.Dropout(config.get("model.classifier_dropout", 0.3)),
            nn.Linear(512, num_classes))


class EarlyFusionAVMobileNet(_EarlyFusionBase):
    """audio_video/models/early_fusion.py:88-110."""
    backbone = "mobilenet_v3_small"


class EarlyFusionAV(_EarlyFusionBase):
    """audio_video/models/ef_cnn_lstm_resnet.py:90-127 (ResNet-18 video encoder, pretrained conv1 kept)."""
    backbone = "resnet18"


def create_early_fusion_mobilenet_model(num_classes, config=None):
    """audio_video/models/early_fusion.py:116-117."""
    return EarlyFusionAVMobileNet(num_classes, config)


def create_early_fusion_resnet_model(num_classes, config=None):
    """audio_video/models/ef_cnn_lstm_resnet.py:132-133."""
    return EarlyFusionAV(num_classes, config)
