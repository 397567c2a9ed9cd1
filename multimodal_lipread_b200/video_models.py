"""Video-only lip readers behind the reference's nn.Module surface (video/models/*.py).

  ResNet2DBiLSTM / create_model      video/models/resnet_lstm.py:56-177   (model.name == "resnet_lstm")
  MobileNetLSTM                      video/models/mobilenet_lstm.py:18-72  (model.name == "mobilenet_lstm")
  VGGLSTM                            video/models/vgg_lstm.py:14-92        (model.name == "vgg_lstm")
  CNNOnly                            video/models/cnn.py:5-73              (model.name == "cnn")

Sub-modules are parameter containers only (same names, construction order and `state_dict` keys as the reference,
including the CNN appearing twice -- `cnn_features.*` and `time_distributed_cnn.module.0.*` share tensors);
arithmetic runs through the launch plans of engine.py."""
import types

import torch
import torch.nn as nn
from torchvision.models import mobilenet_v2, resnet18, resnet34, resnet50

from . import engine
from ._lib import ACT_RELU
from .model_base import Cfg, ModelPlan, PlanModel


class TimeDistributed(nn.Module):
    """Parameter container mirroring video/models/resnet_lstm.py:15-53 (its reshape is folded into the stem's addressing)."""

    def __init__(self, module):
        super().__init__()
        self.module = module


def _resnet_view(seq):
    """children()[:-2] Sequential -> object with the attribute names Plan.resnet_features expects."""
    return types.SimpleNamespace(conv1=seq[0], bn1=seq[1], maxpool=seq[3], layer1=seq[4], layer2=seq[5],
                                 layer3=seq[6], layer4=seq[7])


class ResNetLstmPlan(ModelPlan):
    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        video, layout, scale = self.video_input()
        T = layout[2]
        if isinstance(m, ShuffleNet2DBiLSTM):
            last = self.shufflenet_features(m.cnn_features, (video, layout, scale))
            lstm, drop = m.lstm, m.dropout
            feat, dfeat = self.avgpool(last)
            feat_dim = last.C
        elif hasattr(m, "cnn_features"):
            last = self.resnet_features(_resnet_view(m.cnn_features), (video, layout, scale))
            lstm, drop = m.bilstm, m.dropout
            feat, dfeat = self.avgpool(last)
            feat_dim = last.C
        elif isinstance(m, VGGLSTM):
            kind, feat, dfeat, feat_dim = self.cnn_sequential(list(m.td.module.features), (video, layout, scale))
            assert kind == "pooled"
            lstm, drop = m.lstm, m.dropout
        else:
            last = self.mbv2_features(m.cnn[0], (video, layout, scale))
            lstm, drop = m.lstm, m.drop
            feat, dfeat = self.avgpool(last)
            feat_dim = last.C
        D = 2 * lstm.hidden_size
        seq_last = self.alloc(B * D)
        h = self.alloc(B * D)
        dh = self.alloc(B * D) if wb else None
        # dh is turned into the gradient of x[:, -1] in place by the ReLU backward below, which runs first
        self.bilstm_last(feat, dfeat, feat_dim, B, T, lstm, seq_last, D, dh if wb else 0)
        # x[:, -1] -> ReLU -> Dropout -> fc   (resnet_lstm.py:151-154)
        self.fwd.add("lr_act_fwd", seq_last, h, B * D, ACT_RELU)
        if wb:
            self.bgroup().add("lr_act_bwd", dh, h, B * D, ACT_RELU)
        p = drop.p if isinstance(drop, nn.Dropout) else 0.0
        hd, dhd = self.dropout(h, dh, B * D, p)
        logits = self.alloc(B * self.num_classes)
        dlogits = self.alloc(B * self.num_classes) if wb else None
        self.linear(hd, D, B, m.fc.weight, m.fc.bias, logits, self.num_classes)
        if wb:
            self.linear_bwd(self.bgroup(), hd, D, B, m.fc.weight, m.fc.bias, dlogits, self.num_classes, dx=dhd, ldx=D)
        self.set_logits(logits, dlogits)


class ShuffleNet2DBiLSTM(PlanModel):
    """video/models/shufflenet_lstm.py:27-109 (model.name == "shufflenet_lstm")."""
    INPUTS = ("video",)
    PLAN = None                 # ResNetLstmPlan, set below
    DEFAULT_LR = 5e-5
    DEFAULT_WD = 1e-5

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        from torchvision.models import shufflenet_v2_x0_5, shufflenet_v2_x1_0
        version = config.get("model.shufflenet_version", "0.5x")
        base = shufflenet_v2_x0_5(weights=None) if version == "0.5x" else shufflenet_v2_x1_0(weights=None)
        if pretrained_state_dict is not None:
            base.load_state_dict(pretrained_state_dict)
        self.cnn_features = nn.Sequential(base.conv1, base.maxpool, base.stage2, base.stage3, base.stage4, base.conv5)
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        with torch.no_grad():                                # :60-64, constructor-time pass through the train-mode CNN
            cnn_output_dim = self.global_pool(self.cnn_features(torch.zeros(1, 3, 44, 44))).view(-1).shape[0]
        self.time_cnn = TimeDistributed(nn.Sequential(self.cnn_features, self.global_pool, nn.Flatten()))
        feature_dim = config.get("model.feature_dim", 512)
        dropout = config.get("model.dropout", 0.4)
        self.lstm = nn.LSTM(input_size=cnn_output_dim, hidden_size=feature_dim // 2, num_layers=2, batch_first=True,
                            bidirectional=True, dropout=dropout)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.fc = nn.Linear(feature_dim, num_classes)


class ResNetAttnPlan(ModelPlan):
    """video/models/resnet_attn.py:95-111: per-frame ResNet features -> proj_in -> multi-head self-attention over time ->
    mean over time -> ReLU -> Dropout -> fc."""

    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        video, layout, scale = self.video_input()
        T = layout[2]
        last = self.resnet_features(_resnet_view(m.cnn_features), (video, layout, scale))
        feat, dfeat = self.avgpool(last)
        E = m.proj_in.out_features
        x = self.alloc(B * T * E)
        dx = self.alloc(B * T * E) if wb else None
        self.linear(feat, last.C, B * T, m.proj_in.weight, m.proj_in.bias, x, E)
        if wb:
            self.linear_bwd(self.bgroup(), feat, last.C, B * T, m.proj_in.weight, m.proj_in.bias, dx, E, dx=dfeat, ldx=last.C)
        att, datt = self.multihead_attention(x, dx, B, T, m.attention.attn)
        pooled, dpooled = self.avgpool(engine.T2.of(B, 1, T, E, att, datt))            # x.mean(dim=1)
        h = self.alloc(B * E)
        self.fwd.add("lr_act_fwd", pooled, h, B * E, ACT_RELU)
        if wb:
            self.bgroup().add("lr_act_bwd", dpooled, h, B * E, ACT_RELU)               # dpooled doubles as dh
        hd, dhd = self.dropout(h, dpooled, B * E, m.dropout.p)
        logits = self.alloc(B * self.num_classes)
        dlogits = self.alloc(B * self.num_classes) if wb else None
        self.linear(hd, E, B, m.fc.weight, m.fc.bias, logits, self.num_classes)
        if wb:
            self.linear_bwd(self.bgroup(), hd, E, B, m.fc.weight, m.fc.bias, dlogits, self.num_classes, dx=dhd, ldx=E)
        self.set_logits(logits, dlogits)


class ResNetTransPlan(ModelPlan):
    """video/models/resnet_trans.py:108-129: per-frame ResNet features -> proj_in -> + sinusoidal positions ->
    TransformerEncoder -> mean over time -> ReLU -> Dropout -> fc."""

    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        video, layout, scale = self.video_input()
        T = layout[2]
        last = self.resnet_features(_resnet_view(m.cnn_features), (video, layout, scale))
        feat, dfeat = self.avgpool(last)
        E = m.proj_in.out_features
        pe = m.pos_encoding.pe[0, :T].to(self.dev, torch.float32).repeat(B, 1).contiguous()      # [B*T, E] constant
        self.bufs.append(pe)
        x = self.alloc(B * T * E)
        dx = self.alloc(B * T * E) if wb else None
        # x = proj_in(feat) + pe: the positions ride the GEMM epilogue's residual
        self.gemm_auto(self.fwd, feat, last.C, 0, m.proj_in.weight, last.C, 0, x, E, B * T, E, last.C, bias=m.proj_in.bias,
                       R=pe, ldr=E)
        if wb:
            self.linear_bwd(self.bgroup(), feat, last.C, B * T, m.proj_in.weight, m.proj_in.bias, dx, E, dx=dfeat, ldx=last.C)
        for layer in m.transformer.layers:
            x, dx = self.transformer_encoder_layer(x, dx, B, T, layer)
        if m.transformer.norm is not None:
            raise NotImplementedError("final norm of nn.TransformerEncoder")
        pooled, dpooled = self.avgpool(engine.T2.of(B, 1, T, E, x, dx))                          # x.mean(dim=1)
        h = self.alloc(B * E)
        self.fwd.add("lr_act_fwd", pooled, h, B * E, ACT_RELU)
        if wb:
            self.bgroup().add("lr_act_bwd", dpooled, h, B * E, ACT_RELU)
        hd, dhd = self.dropout(h, dpooled, B * E, m.dropout.p)
        logits = self.alloc(B * self.num_classes)
        dlogits = self.alloc(B * self.num_classes) if wb else None
        self.linear(hd, E, B, m.fc.weight, m.fc.bias, logits, self.num_classes)
        if wb:
            self.linear_bwd(self.bgroup(), hd, E, B, m.fc.weight, m.fc.bias, dlogits, self.num_classes, dx=dhd, ldx=E)
        self.set_logits(logits, dlogits)


class PositionalEncoding(nn.Module):
    """resnet_trans.py:23-42: sinusoidal table kept as a plain attribute (not in the state_dict), as the reference has it."""

    def __init__(self, d_model, max_len=200):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        pos = torch.arange(0, max_len).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2) * (-torch.log(torch.tensor(10000.0)) / d_model))
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.pe = pe.unsqueeze(0)


class ResNet2DTransformer(PlanModel):
    """video/models/resnet_trans.py:45-129 (model.name == "resnet_trans")."""
    INPUTS = ("video",)
    PLAN = ResNetTransPlan
    DEFAULT_LR = 5e-5
    DEFAULT_WD = 1e-5

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        base = resnet18(weights=None) if config.get("model.resnet_version", 18) == 18 else resnet34(weights=None)
        if pretrained_state_dict is not None:
            base.load_state_dict(pretrained_state_dict)
        base.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.cnn_features = nn.Sequential(*list(base.children())[:-2])
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        with torch.no_grad():                                # :69-73, constructor-time pass through the train-mode CNN
            self.cnn_features(torch.zeros(1, 3, 44, 44))
        cnn_out_dim = 512
        self.time_cnn = TimeDistributed(nn.Sequential(self.cnn_features, self.global_pool, nn.Flatten()))
        transformer_dim = config.get("model.transformer_dim", 256)
        num_layers = config.get("model.num_layers", 2)
        num_heads = config.get("model.num_heads", 4)
        dropout = config.get("model.dropout", 0.2)
        self.proj_in = nn.Linear(cnn_out_dim, transformer_dim)
        self.pos_encoding = PositionalEncoding(transformer_dim)
        encoder_layer = nn.TransformerEncoderLayer(d_model=transformer_dim, nhead=num_heads, dropout=dropout,
                                                   batch_first=True, dim_feedforward=transformer_dim * 4)
        self.transformer = nn.TransformerEncoder(encoder_layer, num_layers=num_layers)
        self.dropout = nn.Dropout(dropout)
        self.relu = nn.ReLU()
        self.fc = nn.Linear(transformer_dim, num_classes)


class TemporalAttention(nn.Module):
    """resnet_attn.py:23-35 (parameter container)."""

    def __init__(self, embed_dim, num_heads=4, dropout=0.1):
        super().__init__()
        self.attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout, batch_first=True)


class ResNet2DAttention(PlanModel):
    """video/models/resnet_attn.py:38-111 (model.name == "resnet_attn")."""
    INPUTS = ("video",)
    PLAN = ResNetAttnPlan
    DEFAULT_LR = 5e-5
    DEFAULT_WD = 1e-5

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        base = resnet18(weights=None) if config.get("model.resnet_version", 18) == 18 else resnet34(weights=None)
        if pretrained_state_dict is not None:
            base.load_state_dict(pretrained_state_dict)
        base.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.cnn_features = nn.Sequential(*list(base.children())[:-2])
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        with torch.no_grad():                                # :63-67, the same constructor-time pass as resnet_lstm.py
            self.cnn_features(torch.zeros(1, 3, 44, 44))
        cnn_output_dim = 512
        self.time_cnn = TimeDistributed(nn.Sequential(self.cnn_features, self.global_pool, nn.Flatten()))
        attn_dim = config.get("model.attention_dim", cnn_output_dim)
        num_heads = config.get("model.num_heads", 4)
        dropout = config.get("model.dropout", 0.3)
        self.proj_in = nn.Linear(cnn_output_dim, attn_dim)
        self.attention = TemporalAttention(attn_dim, num_heads=num_heads, dropout=dropout)
        self.dropout = nn.Dropout(dropout)
        self.relu = nn.ReLU()
        self.fc = nn.Linear(attn_dim, num_classes)


class ResNet2DBiLSTM(PlanModel):
    """video/models/resnet_lstm.py:56-156.  forward(x (B,3,T,H,W) f32 [or uint8 (B,T,H,W,3)]) -> (B, num_classes)."""
    INPUTS = ("video",)
    PLAN = ResNetLstmPlan
    DEFAULT_LR = 5e-5            # video/config/visual_config.yaml:25
    DEFAULT_WD = 1e-5            # video/train.py:210

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        resnet_version = config.get("model.resnet_version", 18)
        feature_dim = config.get("model.feature_dim", 1024)
        dropout = config.get("model.dropout", 0.5)
        if resnet_version == 18:
            base = resnet18(weights=None)
        elif resnet_version == 34:
            base = resnet34(weights=None)
        elif resnet_version == 50:
            base = resnet50(weights=None)
        else:
            raise ValueError(f"Unsupported ResNet version: {resnet_version}")
        if pretrained_state_dict is not None:
            base.load_state_dict(pretrained_state_dict)
        base.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)      # re-initialised (:90)
        self.cnn_features = nn.Sequential(*list(base.children())[:-2])
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        # video/models/resnet_lstm.py:98-103 sizes the LSTM input with a dummy pass through the freshly built (train-mode)
        # CNN: that pass moves every BatchNorm's running statistics once and sets num_batches_tracked to 1.  Reproduced
        # at construction (host, once) so that state_dicts and eval-mode outputs agree with the reference's from step 0.
        with torch.no_grad():
            cnn_output_dim = self.cnn_features(torch.zeros(1, 3, 44, 44)).shape[1]     # 512, or 2048 for resnet50
        self.time_distributed_cnn = TimeDistributed(nn.Sequential(self.cnn_features, self.global_pool, nn.Flatten()))
        self.bilstm = nn.LSTM(input_size=cnn_output_dim, hidden_size=feature_dim // 2, num_layers=2, bidirectional=True,
                              batch_first=True, dropout=dropout if dropout > 0 else 0)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.fc = nn.Linear(feature_dim, num_classes)


class MobileNetLSTM(PlanModel):
    """video/models/mobilenet_lstm.py:18-68: MobileNetV2 features + avgpool, 2-layer BiLSTM, x[:, -1] -> ReLU -> Dropout -> fc."""
    INPUTS = ("video",)
    PLAN = ResNetLstmPlan
    DEFAULT_LR = 5e-5
    DEFAULT_WD = 1e-5

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        feature_dim = config.get("model.feature_dim", 256)
        dropout = config.get("model.dropout", 0.3)
        base = mobilenet_v2(weights=None)
        if pretrained_state_dict is not None:
            base.load_state_dict(pretrained_state_dict)
        base.classifier = nn.Identity()
        self.cnn = nn.Sequential(base.features, nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten())
        self.td = TimeDistributed(self.cnn)
        self.lstm = nn.LSTM(input_size=1280, hidden_size=feature_dim // 2, num_layers=2, bidirectional=True,
                            batch_first=True, dropout=dropout)
        self.relu = nn.ReLU()
        self.drop = nn.Dropout(dropout)
        self.fc = nn.Linear(feature_dim, num_classes)


class VGGLite(nn.Module):
    """video/models/vgg_lstm.py:16-45 (parameter container)."""

    def __init__(self):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(3, 32, 3, padding=1), nn.ReLU(True), nn.Conv2d(32, 32, 3, padding=1), nn.ReLU(True), nn.MaxPool2d(2),
            nn.Conv2d(32, 64, 3, padding=1), nn.ReLU(True), nn.Conv2d(64, 64, 3, padding=1), nn.ReLU(True), nn.MaxPool2d(2),
            nn.Conv2d(64, 128, 3, padding=1), nn.ReLU(True), nn.AdaptiveAvgPool2d((1, 1)))


class VGGLSTM(PlanModel):
    """video/models/vgg_lstm.py:48-88: VGGLite per frame, 2-layer BiLSTM, x[:, -1] -> ReLU -> Dropout -> fc."""
    INPUTS = ("video",)
    PLAN = None                  # set below (the plan class is shared with the ResNet / MobileNet variants)
    DEFAULT_LR = 5e-5
    DEFAULT_WD = 1e-5

    def __init__(self, num_classes, config=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        feature_dim = config.get("model.feature_dim", 256)
        dropout = config.get("model.dropout", 0.5)
        self.td = TimeDistributed(VGGLite())
        self.lstm = nn.LSTM(input_size=128, hidden_size=feature_dim // 2, num_layers=2, bidirectional=True,
                            batch_first=True, dropout=dropout)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.fc = nn.Linear(feature_dim, num_classes)


VGGLSTM.PLAN = ResNetLstmPlan


class CnnOnlyPlan(ModelPlan):
    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        video, layout, scale = self.video_input()
        T = layout[2]
        kind, feat, dfeat, C = self.cnn_sequential(list(m.frame_cnn), (video, layout, scale))
        assert kind == "pooled"
        # x.view(B, T, -1).permute(0, 2, 1) -> Conv1d over time: on the channels-last [B*T, C] features that is a 1 x 3
        # window over a [B, 1, T, C] map (no permute), BatchNorm1d over the B*T rows
        x = engine.T2.of(B, 1, T, C, feat, dfeat)
        tc = list(m.temporal_conv)
        for conv, bn in ((tc[0], tc[1]), (tc[3], tc[4])):
            raw = self.dense_conv(x, conv)
            if wb:
                self.dense_conv_bwd(raw)
            a = engine.T2(self, raw.F, raw.H, raw.W, raw.C, h=raw.h)
            self.bn_act(raw, bn, ACT_RELU, a)
            x = a
        pooled, dpooled = self.avgpool(x)                           # x.mean(dim=2): over time
        hd, dhd = self.dropout(pooled, dpooled, B * x.C, m.dropout.p)
        logits = self.alloc(B * self.num_classes)
        dlogits = self.alloc(B * self.num_classes) if wb else None
        self.linear(hd, x.C, B, m.fc.weight, m.fc.bias, logits, self.num_classes)
        if wb:
            self.linear_bwd(self.bgroup(), hd, x.C, B, m.fc.weight, m.fc.bias, dlogits, self.num_classes, dx=dhd, ldx=x.C)
        self.set_logits(logits, dlogits)


class CNNOnly(PlanModel):
    """video/models/cnn.py:5-69: per-frame CNN, two Conv1d + BatchNorm1d + ReLU over time, mean over time, fc."""
    INPUTS = ("video",)
    PLAN = CnnOnlyPlan
    DEFAULT_LR = 5e-5
    DEFAULT_WD = 1e-5

    def __init__(self, num_classes, config=None, precision=None):
        super().__init__()
        config = config or Cfg()
        self._init_base(num_classes, config, precision)
        self.frame_cnn = nn.Sequential(
            nn.Conv2d(3, 32, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(32), nn.ReLU(inplace=True), nn.MaxPool2d(2),
            nn.Conv2d(32, 64, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.MaxPool2d(2),
            nn.Conv2d(64, 128, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
            nn.AdaptiveAvgPool2d((1, 1)))
        temporal_channels = config.get("model.temporal_channels", 128)
        self.temporal_conv = nn.Sequential(
            nn.Conv1d(128, temporal_channels, kernel_size=3, padding=1), nn.BatchNorm1d(temporal_channels), nn.ReLU(inplace=True),
            nn.Conv1d(temporal_channels, temporal_channels, kernel_size=3, padding=1), nn.BatchNorm1d(temporal_channels),
            nn.ReLU(inplace=True))
        self.dropout = nn.Dropout(config.get("model.dropout", 0.3))
        self.fc = nn.Linear(temporal_channels, num_classes)


def create_model(num_classes, config=None):
    """video/models/resnet_lstm.py:165-177."""
    return ResNet2DBiLSTM(num_classes, config)


ShuffleNet2DBiLSTM.PLAN = ResNetLstmPlan
