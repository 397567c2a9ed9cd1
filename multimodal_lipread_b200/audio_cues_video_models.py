"""Audio + cue + video triple-fusion models behind the reference's nn.Module surface (audio_cues_video/models/*.py).

  MultimodalAttentionLate        audio_cues_video/models/late_fusion_mobile.py:6-107   (train.model_name == "late_fusion_mobile")
  MultimodalAttentionLateResNet  audio_cues_video/models/late_fusion_resnet.py:6-99    (train.model_name == "late_fusion_resnet")

forward(mel (B,80,117), cue (B,768), lip (B,3,T,H,W) [or uint8 (B,T,H,W,3)]) -> (B, num_classes).
Sub-modules are parameter containers (reference names / construction order / state_dict keys)."""
import torch.nn as nn
from torchvision.models import mobilenet_v2, resnet18

from ._lib import ACT_RELU
from .model_base import ModelPlan, PlanModel, N_MELS, N_FRAMES_OUT
from .video_models import TimeDistributed


class AttentionFusion(nn.Module):
    """late_fusion_mobile.py:6-19 (parameters only)."""

    def __init__(self, dim):
        super().__init__()
        self.attn = nn.Sequential(nn.Linear(dim, dim // 2), nn.ReLU(), nn.Linear(dim // 2, 1))


class MobileNetLSTM(nn.Module):
    """late_fusion_mobile.py:31-54."""

    def __init__(self, feature_dim=256, pretrained_state_dict=None, dropout=0.3):
        super().__init__()
        base = mobilenet_v2(weights=None)
        if pretrained_state_dict is not None:
            base.load_state_dict(pretrained_state_dict)
        base.classifier = nn.Identity()
        self.cnn = nn.Sequential(base.features, nn.AdaptiveAvgPool2d(1), nn.Flatten())
        self.td = TimeDistributed(self.cnn)
        self.lstm = nn.LSTM(1280, feature_dim // 2, num_layers=2, bidirectional=True, batch_first=True, dropout=dropout)
        self.output_dim = feature_dim


class AudioEncoder(nn.Module):
    """late_fusion_mobile.py:57-66: resnet18 with a 1-channel conv1 and fc = Identity."""

    def __init__(self, pretrained_state_dict=None):
        super().__init__()
        net = resnet18(weights=None)
        if pretrained_state_dict is not None:
            net.load_state_dict(pretrained_state_dict)
        net.conv1 = nn.Conv2d(1, 64, 7, 2, 3, bias=False)
        net.fc = nn.Identity()
        self.enc = net
        self.output_dim = 512


class CueEncoder(nn.Module):
    """late_fusion_mobile.py:69-80."""

    def __init__(self, input_dim=768):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(input_dim, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Linear(256, 256))
        self.output_dim = 256


class LateFusionPlan(ModelPlan):
    def build(self, m, spec):
        B, wb, C = self.B, self.with_backward, self.num_classes
        mel = self.audio_input()
        cue = self.vector_input("cue", m.cue.net[0].in_features)
        video, layout, scale = self.video_input()
        T = layout[2]
        S = 3
        stacked = self.alloc(B * S * C)                       # [B, 3, C]: the three heads write their rows directly
        dstacked = self.alloc(B * S * C) if wb else None
        ldS = S * C

        def head(x, dx, fc, slot):
            """logits of one modality -> stacked[:, slot, :]"""
            self.linear(x, fc.in_features, B, fc.weight, fc.bias, stacked.data_ptr() + 4 * slot * C, ldS)
            if wb:
                self.linear_bwd(self.bgroup(), x, fc.in_features, B, fc.weight, fc.bias,
                                dstacked.data_ptr() + 4 * slot * C, ldS, dx=dx, ldx=fc.in_features)

        # ---- audio: ResNet-18 on the 1-channel log-mel image -> afc
        frames = (mel, (0, B, 1, N_MELS, N_FRAMES_OUT, N_MELS * N_FRAMES_OUT, 0, 0, N_FRAMES_OUT, 1), 1.0)
        a_last = self.resnet_features(m.audio.enc, frames)
        a_feat, a_dfeat = self.avgpool(a_last)
        head(a_feat, a_dfeat, m.afc, 0)
        # ---- cue: Linear -> BatchNorm1d -> ReLU -> Linear -> cfc
        net = m.cue.net
        h, dh = self.linear_bn_act(cue, None, B, net[0], net[1], ACT_RELU)
        c_out = self.alloc(B * net[3].out_features)
        c_dout = self.alloc(B * net[3].out_features) if wb else None
        self.linear(h, net[3].in_features, B, net[3].weight, net[3].bias, c_out, net[3].out_features)
        if wb:
            self.linear_bwd(self.bgroup(), h, net[3].in_features, B, net[3].weight, net[3].bias, c_dout,
                            net[3].out_features, dx=dh, ldx=net[3].in_features)
        head(c_out, c_dout, m.cfc, 1)
        # ---- video: MobileNetV2 + 2-layer BiLSTM, out[:, -1] -> vfc
        trunk = m.video.cnn[0]
        v_last = (self.resnet_features(trunk, (video, layout, scale)) if hasattr(trunk, "conv1")
                  else self.mbv2_features(trunk, (video, layout, scale)))
        v_feat, v_dfeat = self.avgpool(v_last)
        D = m.video.output_dim
        v_out = self.alloc(B * D)
        v_dout = self.alloc(B * D) if wb else None
        self.bilstm_last(v_feat, v_dfeat, v_last.C, B, T, m.video.lstm, v_out, D, v_dout if wb else 0)
        head(v_out, v_dout, m.vfc, 2)
        # ---- attention fusion over the three heads
        att = m.attn.attn
        Hh = att[0].out_features
        ah = self.alloc(B * S * Hh)
        scores, weights = self.alloc(B * S), self.alloc(B * S)
        fused = self.alloc(B * C)
        self.linear(stacked, C, B * S, att[0].weight, att[0].bias, ah, Hh, act=ACT_RELU)
        self.linear(ah, Hh, B * S, att[2].weight, att[2].bias, scores, 1)
        self.fwd.add("lr_attn_fuse_fwd", stacked, scores, weights, fused, B, S, C)
        dfused = None
        if wb:
            dfused = self.alloc(B * C)
            dah, dscores = self.alloc(B * S * Hh), self.alloc(B * S)
            g = self.bgroup()
            g.add("lr_attn_fuse_bwd", stacked, weights, dfused, dstacked, dscores, B, S, C)
            self.linear_bwd(g, ah, Hh, B * S, att[2].weight, att[2].bias, dscores, 1, dx=dah, ldx=Hh)
            g.add("lr_act_bwd", dah, ah, B * S * Hh, ACT_RELU)
            self.linear_bwd(g, stacked, C, B * S, att[0].weight, att[0].bias, dah, Hh, dx=dstacked, ldx=C,
                            dx_residual=dstacked, ldr=C)
        self.set_logits(fused, dfused)


class MultimodalAttentionLate(PlanModel):
    """audio_cues_video/models/late_fusion_mobile.py:84-107."""
    INPUTS = ("audio", "cue", "video")
    PLAN = LateFusionPlan
    DEFAULT_LR = 1e-5            # audio_cues_video/configs/acv_config.yaml:14

    def __init__(self, num_classes, cue_dim=768, video_cfg=None, pretrained=False, precision=None, lstm_dropout=0.3):
        super().__init__()
        if pretrained:
            raise ValueError("no network here: pass ImageNet weights through the sub-modules' pretrained_state_dict")
        self._init_base(num_classes, type("C", (), {"get": staticmethod(lambda k, d=None: d)})(), precision)
        self.audio = AudioEncoder()
        self.cue = CueEncoder(cue_dim)
        vdim = int(video_cfg.get("model", {}).get("feature_dim", 256)) if video_cfg else 256
        self.video = MobileNetLSTM(vdim, dropout=lstm_dropout)
        self.afc = nn.Linear(512, num_classes)
        self.cfc = nn.Linear(256, num_classes)
        self.vfc = nn.Linear(vdim, num_classes)
        self.attn = AttentionFusion(num_classes)


class ResNetLSTM(nn.Module):
    """late_fusion_resnet.py:31-46: resnet18 (fc = Identity, pretrained conv1 kept) + 2-layer BiLSTM."""

    def __init__(self, feature_dim=256, pretrained_state_dict=None, dropout=0.3):
        super().__init__()
        resnet = resnet18(weights=None)
        if pretrained_state_dict is not None:
            resnet.load_state_dict(pretrained_state_dict)
        resnet.fc = nn.Identity()
        self.cnn = nn.Sequential(resnet)
        self.td = TimeDistributed(self.cnn)
        self.lstm = nn.LSTM(512, feature_dim // 2, num_layers=2, bidirectional=True, batch_first=True, dropout=dropout)
        self.output_dim = feature_dim


class MultimodalAttentionLateResNet(PlanModel):
    """audio_cues_video/models/late_fusion_resnet.py:76-99 (parameter order afc, vfc, cfc as in the reference)."""
    INPUTS = ("audio", "cue", "video")
    PLAN = LateFusionPlan
    DEFAULT_LR = 1e-5

    def __init__(self, num_classes, cue_dim=768, video_cfg=None, pretrained=False, precision=None, lstm_dropout=0.3):
        super().__init__()
        if pretrained:
            raise ValueError("no network here: pass ImageNet weights through the sub-modules' pretrained_state_dict")
        self._init_base(num_classes, type("C", (), {"get": staticmethod(lambda k, d=None: d)})(), precision)
        self.audio = AudioEncoder()
        self.cue = CueEncoder(cue_dim)
        vdim = int(video_cfg.get("model", {}).get("feature_dim", 256)) if video_cfg else 256
        self.video = ResNetLSTM(vdim, dropout=lstm_dropout)
        self.afc = nn.Linear(512, num_classes)
        self.vfc = nn.Linear(vdim, num_classes)
        self.cfc = nn.Linear(256, num_classes)
        self.attn = AttentionFusion(num_classes)
