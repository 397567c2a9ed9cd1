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
from .model_base import Cfg, ModelPlan, PlanModel, N_MELS, N_FRAMES_OUT, video_layout, load_torchvision_weights

_Cfg = Cfg
_video_layout = video_layout


class MidFusionPlan(ModelPlan):
    """Launch plan of MidFusionFast at one batch shape."""

    def build(self, m, spec):
        B = self.B
        flat, with_backward = self.flat, self.with_backward
        video, layout, scale = self.video_input()
        T, H, W = layout[2], layout[3], layout[4]
        FA = m.audio_fc.out_features
        HL = m.video_lstm.hidden_size
        FD = FA + 2 * HL
        self.fused = self.alloc(B * FD)
        dfused = self.alloc(B * FD) if with_backward else None
        KA = m.audio_fc.in_features
        ks = engine._ksplit(B, FA, KA, self.sms)

        # ---- audio branch (independent of the video trunk: its own branch of the step graph):
        #      [log-mel ->] conv+relu+pool -> [B,37120] -> audio_fc -> fused[:, 0:FA]
        with self.fwd.side_branch():
            mel = self.audio_input()
            a_pool = self.alloc(B * KA)
            a_arg = self.alloc(B * KA, torch.uint8)
            conv = m.audio_cnn[0]
            self.fwd.add("lr_audio_conv_fwd", mel, conv.weight, conv.bias, a_pool, KA, a_arg, B, N_MELS, N_FRAMES_OUT)
            self.linear(a_pool, KA, B, m.audio_fc.weight, m.audio_fc.bias, self.fused, FD, ksplit=ks)

        # ---- video trunk: MobileNetV3-small features + avgpool -> feat [B*T, 576]
        last = self.mbv3_features(m.video_cnn.features, video, layout, scale, B, T, H, W)
        feat, dfeat = self.avgpool(last)
        # ---- BiLSTM with an out[:, -1] head, written straight into the fusion row (no torch.cat)
        self.bilstm_last(feat, dfeat, last.C, B, T, m.video_lstm, self.fused.data_ptr() + 4 * FA, FD,
                         (dfused.data_ptr() + 4 * FA) if with_backward else 0)
        if with_backward:
            # the audio branch's backward is registered HERE so that it runs right after the classifier's (groups run in
            # reverse registration order), all of it on the side branch, next to the LSTM / trunk backward
            d_pool = self.alloc(B * KA)
            g = self.bgroup()
            with g.side_branch():
                self.linear_bwd(g, a_pool, KA, B, m.audio_fc.weight, m.audio_fc.bias, dfused, FD, dx=d_pool, ldx=KA)
                g.add("lr_audio_conv_bwd", mel, d_pool, KA, a_arg, flat.g(conv.weight), flat.g(conv.bias), B, N_MELS,
                      N_FRAMES_OUT)
        self.fwd.join()
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


def create_mid_fusion_fast(num_classes, config=None, pretrained_state_dict=None):
    """audio_video/models/middle_fusion_fast.py:41-42."""
    return MidFusionFast(num_classes, config, pretrained_state_dict=pretrained_state_dict)


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


def audio_cnn_plan(plan, cnn, mel, B):
    """An audio CNN Sequential on the mel (B,80,117) viewed as a 1-channel NHWC image (see Plan.cnn_sequential)."""
    mods = list(cnn)
    if mods[0].in_channels != 1:
        raise ValueError("audio CNN plans handle dataset.audio_channels == 1 (the reference default)")
    # (B,80,117) contiguous == NHWC with C = 1
    frames = (mel, (0, B, 1, N_MELS, N_FRAMES_OUT, N_MELS * N_FRAMES_OUT, 0, 0, N_FRAMES_OUT, 1), 1.0)
    # the small audio encoders keep fp32 storage in every precision mode: their consumers (flatten to the fusion row,
    # fp32 heads) read fp32, and 32 x 80 x 117 images are a sliver of the step's traffic
    return plan.cnn_sequential(mods, frames, h=False)


def audio_encoder_plan(plan, enc, mel, B, out, ldo, dout):
    """AudioEncoder* (cnn + fc) forward/backward; writes out[b, 0:D] (row stride ldo)."""
    kind, pooled, dpooled, _ = audio_cnn_plan(plan, enc.cnn, mel, B)
    assert kind == "pooled"
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
        DA, DV = m.audio_encoder.output_dim, m.video_encoder.output_dim
        FD = DA + DV
        self.fused = self.alloc(B * FD)
        dfused = self.alloc(B * FD) if wb else None
        with self.fwd.side_branch():                       # the audio encoder is independent of the video trunk
            mel = self.audio_input()
            audio_encoder_plan(self, m.audio_encoder, mel, B, self.fused, FD, dfused)
        ve = m.video_encoder
        if m.backbone == "mobilenet_v3_small":
            last = self.mbv3_features(ve.cnn.features, video, layout, scale, B, T, H, W)
        else:
            last = self.resnet_features(ve.cnn, (video, layout, scale))
        feat, dfeat = self.avgpool(last)
        self.bilstm_last(feat, dfeat, last.C, B, T, ve.lstm, self.fused.data_ptr() + 4 * DA, FD,
                         (dfused.data_ptr() + 4 * DA) if wb else 0)
        self.fwd.join()
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


def create_early_fusion_mobilenet_model(num_classes, config=None, pretrained_state_dict=None):
    """audio_video/models/early_fusion.py:116-117."""
    return EarlyFusionAVMobileNet(num_classes, config, pretrained_state_dict=pretrained_state_dict)


def create_early_fusion_resnet_model(num_classes, config=None, pretrained_state_dict=None):
    """audio_video/models/ef_cnn_lstm_resnet.py:132-133."""
    return EarlyFusionAV(num_classes, config, pretrained_state_dict=pretrained_state_dict)


# ------------------------------------------------------------------------------------------------------------
# The remaining audio_video models (all MobileNetV3-small video trunks, single-layer BiLSTM)
# ------------------------------------------------------------------------------------------------------------
def _load_mbv3(model):
    """ImageNet mobilenet_v3_small checkpoint (torchvision key names) into the model's video trunk, when one was given."""
    sd = getattr(model, "_pretrained_sd", None)
    model._pretrained_sd = None
    if sd is not None:
        with torch.no_grad():
            if load_torchvision_weights(model, sd) == 0:
                raise ValueError("pretrained_state_dict matches no tensor of the video trunk")


def _mbv3_lstm(config, hidden_default, pretrained_state_dict=None):
    base = mobilenet_v3_small(weights=None)
    if pretrained_state_dict is not None:
        base.load_state_dict(pretrained_state_dict)
    base.classifier = nn.Identity()
    hid = config.get("video.lstm_hidden", hidden_default)
    return base, nn.LSTM(input_size=576, hidden_size=hid, num_layers=1, batch_first=True, bidirectional=True), hid


class _VideoEncoderLstm(nn.Module):
    """VideoEncoderLate / VideoEncoderMid / VideoEncoderFast parameter container (cnn, lstm)."""

    def __init__(self, config, hidden_default):
        super().__init__()
        self.cnn, self.lstm, hid = _mbv3_lstm(config, hidden_default)
        self.output_dim = hid * 2


class AudioEncoderLate(nn.Module):
    """late_fusion.py:10-34."""

    def __init__(self, config):
        super().__init__()
        cin = config.get("dataset.audio_channels", 1)
        self.cnn = nn.Sequential(nn.Conv2d(cin, 32, 3, padding=1), nn.BatchNorm2d(32), nn.ReLU(), nn.MaxPool2d(2),
                                 nn.Conv2d(32, 64, 3, padding=1), nn.BatchNorm2d(64), nn.ReLU(), nn.AdaptiveAvgPool2d((1, 1)))
        self.fc = nn.Linear(64, config.get("model.audio_feature_dim", 256))
        self.output_dim = config.get("model.audio_feature_dim", 256)


class AudioEncoderMid(nn.Module):
    """middle_fusion.py:11-30: ends on the (64, 20, 29) feature map, flattened channel-major."""

    def __init__(self, config):
        super().__init__()
        cin = config.get("dataset.audio_channels", 1)
        self.cnn = nn.Sequential(nn.Conv2d(cin, 32, kernel_size=3, padding=1), nn.BatchNorm2d(32), nn.ReLU(), nn.MaxPool2d(2),
                                 nn.Conv2d(32, 64, kernel_size=3, padding=1), nn.BatchNorm2d(64), nn.ReLU(), nn.MaxPool2d(2))
        self.output_dim = 64 * 20 * 29


class AudioEncoderFast(nn.Module):
    """early_fusion_fast.py:6-25."""

    def __init__(self, config):
        super().__init__()
        cin = config.get("dataset.audio_channels", 1)
        self.cnn = nn.Sequential(nn.Conv2d(cin, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
                                 nn.Conv2d(16, 32, 3, padding=1), nn.ReLU(), nn.AdaptiveAvgPool2d((1, 1)))
        self.fc = nn.Linear(32, config.get("model.audio_feature_dim", 128))
        self.output_dim = config.get("model.audio_feature_dim", 128)


class _AlphaLatePlan(ModelPlan):
    """fused = alpha * audio_logits + (1 - alpha) * video_logits (late_fusion.py:88-93, late_fusion_fast.py:36-59)."""

    def build(self, m, spec):
        B, wb, C = self.B, self.with_backward, self.num_classes
        video, layout, scale = self.video_input()
        T, H, W = layout[2], layout[3], layout[4]
        mel = self.audio_input()
        cnn, fc, acls, vcnn, vlstm, vcls = m._parts()
        DA = fc.out_features
        a_feat = self.alloc(B * DA)
        da_feat = self.alloc(B * DA) if wb else None
        kind, pooled, dpooled, _ = audio_cnn_plan(self, cnn, mel, B)
        self.linear(pooled, fc.in_features, B, fc.weight, fc.bias, a_feat, DA)
        if wb:
            self.linear_bwd(self.bgroup(), pooled, fc.in_features, B, fc.weight, fc.bias, da_feat, DA, dx=dpooled,
                            ldx=fc.in_features)
        a_log = self.alloc(B * C)
        da_log = self.alloc(B * C) if wb else None
        self.linear(a_feat, DA, B, acls.weight, acls.bias, a_log, C)
        if wb:
            self.linear_bwd(self.bgroup(), a_feat, DA, B, acls.weight, acls.bias, da_log, C, dx=da_feat, ldx=DA)
        last = self.mbv3_features(vcnn.features, video, layout, scale, B, T, H, W)
        feat, dfeat = self.avgpool(last)
        DV = 2 * vlstm.hidden_size
        v_feat = self.alloc(B * DV)
        dv_feat = self.alloc(B * DV) if wb else None
        self.bilstm_hn(feat, dfeat, last.C, B, T, vlstm, v_feat, DV, dv_feat if wb else 0)
        v_log = self.alloc(B * C)
        dv_log = self.alloc(B * C) if wb else None
        self.linear(v_feat, DV, B, vcls.weight, vcls.bias, v_log, C)
        if wb:
            self.linear_bwd(self.bgroup(), v_feat, DV, B, vcls.weight, vcls.bias, dv_log, C, dx=dv_feat, ldx=DV)
        fused = self.alloc(B * C)
        dfused = self.alloc(B * C) if wb else None
        self.fwd.add("lr_alpha_fuse_fwd", a_log, v_log, m.alpha, fused, B * C)
        if wb:
            self.bgroup().add("lr_alpha_fuse_bwd", a_log, v_log, m.alpha, dfused, da_log, dv_log, self.flat.g(m.alpha), B * C)
        self.set_logits(fused, dfused)


class LateFusionAVMobileNet(PlanModel):
    """audio_video/models/late_fusion.py:70-93."""
    PLAN = _AlphaLatePlan

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        self._pretrained_sd = pretrained_state_dict
        self.audio_encoder = AudioEncoderLate(config)
        self.video_encoder = _VideoEncoderLstm(config, 256)
        self.audio_classifier = nn.Linear(self.audio_encoder.output_dim, num_classes)
        self.video_classifier = nn.Linear(self.video_encoder.output_dim, num_classes)
        self.alpha = nn.Parameter(torch.tensor(0.5))
        _load_mbv3(self)

    def _parts(self):
        return (self.audio_encoder.cnn, self.audio_encoder.fc, self.audio_classifier, self.video_encoder.cnn,
                self.video_encoder.lstm, self.video_classifier)


class LateFusionFast(PlanModel):
    """audio_video/models/late_fusion_fast.py:5-59."""
    PLAN = _AlphaLatePlan

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        self._pretrained_sd = pretrained_state_dict
        cin = config.get("dataset.audio_channels", 1)
        self.audio_cnn = nn.Sequential(nn.Conv2d(cin, 16, 3, padding=1), nn.ReLU(), nn.AdaptiveAvgPool2d((1, 1)))
        self.audio_fc = nn.Linear(16, config.get("model.audio_feature_dim", 128))
        self.audio_classifier = nn.Linear(config.get("model.audio_feature_dim", 128), num_classes)
        base = mobilenet_v3_small(weights=None)
        base.classifier = nn.Identity()
        self.video_cnn = base
        self.video_lstm = nn.LSTM(input_size=576, hidden_size=128, num_layers=1, batch_first=True, bidirectional=True)
        self.video_classifier = nn.Linear(128 * 2, num_classes)
        self.alpha = nn.Parameter(torch.tensor(0.5))
        _load_mbv3(self)

    def _parts(self):
        return self.audio_cnn, self.audio_fc, self.audio_classifier, self.video_cnn, self.video_lstm, self.video_classifier


class _ConcatFusionPlan(ModelPlan):
    """fused = [audio features | video features] -> classifier (middle_fusion.py:72-85, early_fusion_fast.py:63-76)."""

    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        video, layout, scale = self.video_input()
        T, H, W = layout[2], layout[3], layout[4]
        mel = self.audio_input()
        ae, ve = m.audio_encoder, m.video_encoder
        DA, DV = ae.output_dim, ve.output_dim
        FD = DA + DV
        fused = self.alloc(B * FD)
        dfused = self.alloc(B * FD) if wb else None
        if hasattr(ae, "fc"):
            audio_encoder_plan(self, ae, mel, B, fused, FD, dfused)
        else:
            kind, fmap = audio_cnn_plan(self, ae.cnn, mel, B)
            assert kind == "map" and fmap.H * fmap.W * fmap.C == DA
            # x.view(B, -1) of the NCHW map: channel-major flatten straight into the fused row
            self.fwd.add("lr_flatten_nchw", fmap.val, fused, FD, B, fmap.H * fmap.W, fmap.C, 1)
            if wb:
                self.bgroup().add("lr_flatten_nchw", fmap.grad, dfused, FD, B, fmap.H * fmap.W, fmap.C, 0)
        last = self.mbv3_features(ve.cnn.features, video, layout, scale, B, T, H, W)
        feat, dfeat = self.avgpool(last)
        vptr = fused.data_ptr() + 4 * DA
        dvptr = (dfused.data_ptr() + 4 * DA) if wb else 0
        if m.head == "hn":
            self.bilstm_hn(feat, dfeat, last.C, B, T, ve.lstm, vptr, FD, dvptr)
        else:
            self.bilstm_last(feat, dfeat, last.C, B, T, ve.lstm, vptr, FD, dvptr)
        logits, dlogits = self.mlp(fused, dfused, B, m.classifier)
        self.set_logits(logits, dlogits)


class MidFusionAVMobileNet(PlanModel):
    """audio_video/models/middle_fusion.py:66-85 (video head: feats[:, -1])."""
    PLAN = _ConcatFusionPlan
    head = "last"

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        self._pretrained_sd = pretrained_state_dict
        self.audio_encoder = AudioEncoderMid(config)
        self.video_encoder = _VideoEncoderLstm(config, 256)
        fusion_dim = self.audio_encoder.output_dim + self.video_encoder.output_dim
        self.classifier = nn.Sequential(nn.Linear(fusion_dim, 512), nn.ReLU(),
                                        nn.Dropout(config.get("model.classifier_dropout", 0.3)), nn.Linear(512, num_classes))
        _load_mbv3(self)


class EarlyFusionFast(PlanModel):
    """audio_video/models/early_fusion_fast.py:57-76 (video head: cat(h_n[0], h_n[1]))."""
    PLAN = _ConcatFusionPlan
    head = "hn"

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        self._pretrained_sd = pretrained_state_dict
        self.audio_encoder = AudioEncoderFast(config)
        self.video_encoder = _VideoEncoderLstm(config, 128)
        fusion_dim = self.audio_encoder.output_dim + self.video_encoder.output_dim
        self.classifier = nn.Sequential(nn.Linear(fusion_dim, 256), nn.ReLU(), nn.Linear(256, num_classes))
        _load_mbv3(self)


def create_late_fusion_mobilenet_model(num_classes, config=None, pretrained_state_dict=None):
    """audio_video/models/late_fusion.py:98-99."""
    return LateFusionAVMobileNet(num_classes, config, pretrained_state_dict=pretrained_state_dict)


def create_mid_fusion_mobilenet_model(num_classes, config=None, pretrained_state_dict=None):
    """audio_video/models/middle_fusion.py:91-92."""
    return MidFusionAVMobileNet(num_classes, config, pretrained_state_dict=pretrained_state_dict)


def create_early_fusion_fast(num_classes, config=None, pretrained_state_dict=None):
    """audio_video/models/early_fusion_fast.py:79-80."""
    return EarlyFusionFast(num_classes, config, pretrained_state_dict=pretrained_state_dict)


def create_late_fusion_fast(num_classes, config=None, pretrained_state_dict=None):
    """audio_video/models/late_fusion_fast.py:62-63."""
    return LateFusionFast(num_classes, config, pretrained_state_dict=pretrained_state_dict)
