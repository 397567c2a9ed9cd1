"""Audio-only classifiers behind the reference's nn.Module surface (audio/models/*.py).

  AudioResNet          audio/models/resnet_model.py:5-39        (model.name == "resnet": BASELINE config 1)
  AudioResNetLSTM      audio/models/resnet_lstm_model.py:5-59   (model.name == "resnet_lstm")
  VGGAudioClassifier   audio/models/vgg_model.py:5-58           (model.name == "vgg")
  VGGWithLSTMClassifier audio/models/vgg_lstm_model.py:5-75     (model.name == "vgg_lstm")
  LSTMResNet           audio/models/lstm_resnet_model.py:5-71   (model.name == "lstm_resnet")
  DeepAudioNetWithAttention audio/models/lstm_resnet_attn_model.py:17-88 (model.name == "lstm_resnet_attn")
  LSTMResNetWithTransformer audio/models/lstm_resnet_trans_model.py:22-104 (model.name == "lstm_resnet_trans")

forward(spec (B,80,117) f32 log-mel) -> (B, num_classes); with a raw (B,20000) waveform the fused log-mel kernel
runs first.  Sub-modules are parameter containers (reference names / construction order / state_dict keys)."""
import types

import torch
import torch.nn as nn
from torchvision.models import resnet18, vgg11_bn, vgg13_bn, vgg16_bn, vgg19_bn

from . import engine
from ._lib import ACT_NONE, ACT_RELU
from .model_base import ModelPlan, PlanModel, N_MELS, N_FRAMES_OUT


def bn_head(plan, feat, dfeat, B, head):
    """Linear -> [BatchNorm1d] -> ReLU -> Dropout -> Linear (resnet_model.py:19-33, resnet_lstm_model.py:32-44,
    vgg_model.py:20-32).  Returns (logits, dlogits) buffers."""
    wb = plan.with_backward
    fc1 = head[0]
    D = fc1.out_features
    i = 1
    if isinstance(head[i], nn.BatchNorm1d):
        cur, dcur = plan.linear_bn_act(feat, dfeat, B, fc1, head[i], ACT_RELU)
        i += 2                                           # BatchNorm1d, ReLU
    else:
        cur = plan.alloc(B * D)
        dcur = plan.alloc(B * D) if wb else None
        plan.linear(feat, fc1.in_features, B, fc1.weight, fc1.bias, cur, D, act=ACT_RELU)
        if wb:
            g = plan.bgroup()
            g.add("lr_act_bwd", dcur, cur, B * D, ACT_RELU)
            plan.linear_bwd(g, feat, fc1.in_features, B, fc1.weight, fc1.bias, dcur, D, dx=dfeat, ldx=fc1.in_features)
        i += 1                                           # ReLU
    drop, fc2 = head[i], head[i + 1]
    cur, dcur = plan.dropout(cur, dcur, B * D, drop.p)
    C = fc2.out_features
    logits = plan.alloc(B * C)
    dlogits = plan.alloc(B * C) if wb else None
    plan.linear(cur, D, B, fc2.weight, fc2.bias, logits, C, act=ACT_NONE)
    if wb:
        plan.linear_bwd(plan.bgroup(), cur, D, B, fc2.weight, fc2.bias, dlogits, C, dx=dcur, ldx=D)
    return logits, dlogits


def _mel_frames(mel, B):
    # x.unsqueeze(1): (B,1,80,117) NCHW with one channel == NHWC with C = 1
    return (mel, (0, B, 1, N_MELS, N_FRAMES_OUT, N_MELS * N_FRAMES_OUT, 0, 0, N_FRAMES_OUT, 1), 1.0)


class AudioResNetPlan(ModelPlan):
    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        mel = self.audio_input()
        net = m.resnet
        # x.unsqueeze(1): (B,1,80,117) NCHW with one channel == NHWC with C = 1
        frames = (mel, (0, B, 1, N_MELS, N_FRAMES_OUT, N_MELS * N_FRAMES_OUT, 0, 0, N_FRAMES_OUT, 1), 1.0)
        last = self.resnet_features(net, frames)
        feat, dfeat = self.avgpool(last)
        logits, dlogits = bn_head(self, feat, dfeat, B, list(net.fc))
        self.set_logits(logits, dlogits)


class AudioResNet(PlanModel):
    """audio/models/resnet_model.py:5-39."""
    INPUTS = ("audio",)
    PLAN = AudioResNetPlan
    DEFAULT_LR = 5e-4            # audio/configs/audio_config.yaml:20-21
    DEFAULT_WD = 1e-4

    def __init__(self, num_classes=40, dropout_rate=0.5, use_batchnorm=True, pretrained_state_dict=None, precision=None):
        super().__init__()
        self._init_base(num_classes, types.SimpleNamespace(get=lambda k, d=None: d), precision)
        self.use_bn = use_batchnorm
        self.resnet = resnet18(weights=None)
        if pretrained_state_dict is not None:
            self.resnet.load_state_dict(pretrained_state_dict)
        self.resnet.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        num_features = self.resnet.fc.in_features
        layers = [nn.Linear(num_features, 512)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(512))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate), nn.Linear(512, num_classes)])
        self.resnet.fc = nn.Sequential(*layers)


class AudioResNetLstmPlan(ModelPlan):
    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        mel = self.audio_input()
        last = self.resnet_features(m.resnet, _mel_frames(mel, B))
        feat, dfeat = self.avgpool(last)
        # features.unsqueeze(1): a sequence of length 1 through the 2-layer BiLSTM, lstm_out[:, -1, :]
        D = 2 * m.lstm.hidden_size
        seq = self.alloc(B * D)
        dseq = self.alloc(B * D) if wb else None
        self.bilstm_last(feat, dfeat, last.C, B, 1, m.lstm, seq, D, dseq if wb else 0)
        logits, dlogits = bn_head(self, seq, dseq, B, list(m.classifier))
        self.set_logits(logits, dlogits)


class AudioResNetLSTM(PlanModel):
    """audio/models/resnet_lstm_model.py:5-59."""
    INPUTS = ("audio",)
    PLAN = AudioResNetLstmPlan
    DEFAULT_LR = 5e-4
    DEFAULT_WD = 1e-4

    def __init__(self, num_classes=40, lstm_hidden=128, lstm_layers=2, dropout_rate=0.3, use_batchnorm=True,
                 pretrained_state_dict=None, precision=None):
        super().__init__()
        self._init_base(num_classes, types.SimpleNamespace(get=lambda k, d=None: d), precision)
        self.use_bn = use_batchnorm
        self.resnet = resnet18(weights=None)
        if pretrained_state_dict is not None:
            self.resnet.load_state_dict(pretrained_state_dict)
        self.resnet.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.resnet.fc = nn.Identity()
        self.lstm = nn.LSTM(input_size=512, hidden_size=lstm_hidden, num_layers=lstm_layers, bidirectional=True,
                            batch_first=True)
        layers = [nn.Linear(2 * lstm_hidden, 256)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(256))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate), nn.Linear(256, num_classes)])
        self.classifier = nn.Sequential(*layers)


class VGGAudioPlan(ModelPlan):
    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        mel = self.audio_input()
        kind, fmap = self.cnn_sequential(list(m.vgg.features), _mel_frames(mel, B), h=False)      # flattened to fp32 below
        assert kind == "map"
        osz = m.adaptive_pool.output_size
        if (fmap.H, fmap.W) != tuple(osz):
            raise NotImplementedError(f"AdaptiveAvgPool2d{tuple(osz)} on a {fmap.H}x{fmap.W} map")
        # torch.flatten(x, 1) of the NCHW map: channel-major
        K = fmap.H * fmap.W * fmap.C
        flat_in = self.alloc(B * K)
        dflat = self.alloc(B * K) if wb else None
        self.fwd.add("lr_flatten_nchw", fmap.val, flat_in, K, B, fmap.H * fmap.W, fmap.C, 1)
        if wb:
            self.bgroup().add("lr_flatten_nchw", fmap.grad, dflat, K, B, fmap.H * fmap.W, fmap.C, 0)
        logits, dlogits = bn_head(self, flat_in, dflat, B, list(m.vgg.classifier))
        self.set_logits(logits, dlogits)


class VGGAudioClassifier(PlanModel):
    """audio/models/vgg_model.py:5-58 (torchvision vgg{11,13,16,19}_bn features with a 1-channel first conv)."""
    INPUTS = ("audio",)
    PLAN = VGGAudioPlan
    DEFAULT_LR = 5e-4
    DEFAULT_WD = 1e-4

    def __init__(self, num_classes=40, version=11, dropout_rate=0.5, use_batchnorm=True, pretrained_state_dict=None,
                 precision=None):
        super().__init__()
        self._init_base(num_classes, types.SimpleNamespace(get=lambda k, d=None: d), precision)
        self.use_bn = use_batchnorm
        ctor = {11: vgg11_bn, 13: vgg13_bn, 16: vgg16_bn, 19: vgg19_bn}.get(version)
        if ctor is None:
            raise ValueError(f"Invalid VGG version: {version}")
        # init_weights=False: what torchvision does when (as in the reference) pretrained weights are requested
        self.vgg = ctor(weights=None, init_weights=False)
        if pretrained_state_dict is not None:
            self.vgg.load_state_dict(pretrained_state_dict)
        self.vgg.features[0] = nn.Conv2d(1, 64, kernel_size=3, padding=1)
        self.adaptive_pool = nn.AdaptiveAvgPool2d((2, 3))
        layers = [nn.Linear(512 * 2 * 3, 256)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(256))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate), nn.Linear(256, num_classes)])
        self.vgg.classifier = nn.Sequential(*layers)


class VGGLstmAudioPlan(ModelPlan):
    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        mel = self.audio_input()
        kind, fmap = self.cnn_sequential(list(m.vgg_features), _mel_frames(mel, B), h=False)      # viewed as fp32 rows below
        assert kind == "map"
        # AdaptiveAvgPool2d((None, 1)) + squeeze + permute: mean over W per (clip, row) -> a sequence of H steps of C
        # features; on the channels-last map that is a per-"frame" pooling with frames = (clip, row)
        T = fmap.H
        feat, dfeat = self.avgpool(engine.T2.of(B * T, 1, fmap.W, fmap.C, fmap.val, fmap.grad))
        D = 2 * m.lstm.hidden_size
        seq = self.alloc(B * D)
        dseq = self.alloc(B * D) if wb else None
        self.bilstm_last(feat, dfeat, fmap.C, B, T, m.lstm, seq, D, dseq if wb else 0)
        logits, dlogits = bn_head(self, seq, dseq, B, list(m.classifier))
        self.set_logits(logits, dlogits)


class VGGWithLSTMClassifier(PlanModel):
    """audio/models/vgg_lstm_model.py:5-75."""
    INPUTS = ("audio",)
    PLAN = VGGLstmAudioPlan
    DEFAULT_LR = 5e-4
    DEFAULT_WD = 1e-4

    def __init__(self, num_classes=40, lstm_hidden_size=128, lstm_layers=2, version=11, dropout_rate=0.3,
                 use_batchnorm=True, pretrained_state_dict=None, precision=None):
        super().__init__()
        self._init_base(num_classes, types.SimpleNamespace(get=lambda k, d=None: d), precision)
        self.use_bn = use_batchnorm
        ctor = {11: vgg11_bn, 13: vgg13_bn, 16: vgg16_bn, 19: vgg19_bn}.get(version)
        if ctor is None:
            raise ValueError(f"Invalid VGG version: {version}")
        vgg = ctor(weights=None, init_weights=False)
        if pretrained_state_dict is not None:
            vgg.load_state_dict(pretrained_state_dict)
        vgg.features[0] = nn.Conv2d(1, 64, kernel_size=3, padding=1)
        self.vgg_features = vgg.features
        self.adaptive_pool = nn.AdaptiveAvgPool2d((None, 1))
        self.cnn_output_dim = 512
        self.lstm = nn.LSTM(input_size=512, hidden_size=lstm_hidden_size, num_layers=lstm_layers, bidirectional=True,
                            batch_first=True)
        layers = [nn.Linear(2 * lstm_hidden_size, 128)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(128))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate), nn.Linear(128, num_classes)])
        self.classifier = nn.Sequential(*layers)


def lstm_resnet_front(plan, m):
    """mel rows -> initial_bilstm (length-1 sequences) -> 1-channel image -> ResNet-18 -> fc (+BN1d) -> ReLU -> Dropout:
    the part audio/models/lstm_resnet_model.py:39-60, lstm_resnet_attn_model.py:61-75 and lstm_resnet_trans_model.py
    share.  Returns (h, dh, width)."""
    self, B, wb = plan, plan.B, plan.with_backward
    mel = self.audio_input()
    I = m.initial_bilstm.input_size
    if I != N_FRAMES_OUT:
        raise ValueError(f"LSTMResNet input_size {I} != {N_FRAMES_OUT} mel frames")
    # x.view(B*80, 117).unsqueeze(1): every mel row is a length-1 sequence through the 2-layer BiLSTM(117 -> 64)
    Bq = B * N_MELS
    D0 = 2 * m.initial_bilstm.hidden_size
    img = self.alloc(Bq * D0)
    dimg = self.alloc(Bq * D0) if wb else None
    self.bilstm_last(mel, None, I, Bq, 1, m.initial_bilstm, img, D0, dimg if wb else 0)
    # .view(B, 1, 80, 128): a computed 1-channel image; the stem needs its input gradient
    x = engine.T2.of(B, N_MELS, D0, 1, img, dimg)
    last = self.resnet_features(m.resnet, None, x=x)
    feat, dfeat = self.avgpool(last)
    head = list(m.fc)
    Dh = head[0].out_features
    if isinstance(head[1], nn.BatchNorm1d):
        h, dh = self.linear_bn_act(feat, dfeat, B, head[0], head[1], ACT_RELU)
        drop = head[3]
    else:
        h = self.alloc(B * Dh)
        dh = self.alloc(B * Dh) if wb else None
        self.linear(feat, head[0].in_features, B, head[0].weight, head[0].bias, h, Dh, act=ACT_RELU)
        if wb:
            g = self.bgroup()
            g.add("lr_act_bwd", dh, h, B * Dh, ACT_RELU)
            self.linear_bwd(g, feat, head[0].in_features, B, head[0].weight, head[0].bias, dh, Dh, dx=dfeat,
                            ldx=head[0].in_features)
        drop = head[2]
    h, dh = self.dropout(h, dh, B * Dh, drop.p)
    return h, dh, Dh


class LstmResNetPlan(ModelPlan):
    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        h, dh, Dh = lstm_resnet_front(self, m)
        D1 = 2 * m.final_bilstm.hidden_size
        seq = self.alloc(B * D1)
        dseq = self.alloc(B * D1) if wb else None
        self.bilstm_last(h, dh, Dh, B, 1, m.final_bilstm, seq, D1, dseq if wb else 0)
        C = self.num_classes
        logits = self.alloc(B * C)
        dlogits = self.alloc(B * C) if wb else None
        self.linear(seq, D1, B, m.classifier.weight, m.classifier.bias, logits, C)
        if wb:
            self.linear_bwd(self.bgroup(), seq, D1, B, m.classifier.weight, m.classifier.bias, dlogits, C, dx=dseq, ldx=D1)
        self.set_logits(logits, dlogits)


class LstmResNetAttnPlan(ModelPlan):
    """audio/models/lstm_resnet_attn_model.py:61-88: the projection repeated over 10 steps, a full 2-layer BiLSTM over
    them, additive attention pooling (Linear(256, 1) -> softmax over time -> weighted sum), classifier."""
    STEPS = 10

    def build(self, m, spec):
        B, wb, T = self.B, self.with_backward, self.STEPS
        h, dh, Dh = lstm_resnet_front(self, m)
        # fc_out.unsqueeze(1).repeat(1, 10, 1)
        xs = self.alloc(B * T * Dh)
        dxs = self.alloc(B * T * Dh) if wb else None
        self.fwd.add("lr_add_bcast", h, 0, xs, B, T, Dh)
        if wb:
            ones = torch.ones(B * T * Dh, dtype=torch.float32, device=self.dev)
            self.bufs.append(ones)
            self.bgroup().add("lr_frame_reduce", dxs, ones, dh, B, T, Dh, 1)          # dh[b] = sum_t dxs[b, t]
        seq, dseq = self.bilstm_seq(xs, dxs, Dh, B, T, m.final_bilstm)
        D1 = 2 * m.final_bilstm.hidden_size
        att = m.attention.attn
        scores, weights = self.alloc(B * T), self.alloc(B * T)
        pooled = self.alloc(B * D1)
        dpooled = self.alloc(B * D1) if wb else None
        self.linear(seq, D1, B * T, att.weight, att.bias, scores, 1)
        self.fwd.add("lr_attn_fuse_fwd", seq, scores, weights, pooled, B, T, D1)
        if wb:
            dscores = self.alloc(B * T)
            g = self.bgroup()
            g.add("lr_attn_fuse_bwd", seq, weights, dpooled, dseq, dscores, B, T, D1)
            self.linear_bwd(g, seq, D1, B * T, att.weight, att.bias, dscores, 1, dx=dseq, ldx=D1, dx_residual=dseq, ldr=D1)
        C = self.num_classes
        logits = self.alloc(B * C)
        dlogits = self.alloc(B * C) if wb else None
        self.linear(pooled, D1, B, m.classifier.weight, m.classifier.bias, logits, C)
        if wb:
            self.linear_bwd(self.bgroup(), pooled, D1, B, m.classifier.weight, m.classifier.bias, dlogits, C, dx=dpooled, ldx=D1)
        self.set_logits(logits, dlogits)


class LSTMResNet(PlanModel):
    """audio/models/lstm_resnet_model.py:5-71."""
    INPUTS = ("audio",)
    PLAN = LstmResNetPlan
    DEFAULT_LR = 5e-4
    DEFAULT_WD = 1e-4

    def __init__(self, num_classes=40, input_size=117, dropout_rate=0.3, use_batchnorm=True, pretrained_state_dict=None,
                 precision=None):
        super().__init__()
        self._init_base(num_classes, types.SimpleNamespace(get=lambda k, d=None: d), precision)
        self.use_bn = use_batchnorm
        self.initial_bilstm = nn.LSTM(input_size=input_size, hidden_size=64, num_layers=2, bidirectional=True, batch_first=True)
        self.resnet = resnet18(weights=None)
        if pretrained_state_dict is not None:
            self.resnet.load_state_dict(pretrained_state_dict)
        self.resnet.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.resnet.fc = nn.Identity()
        layers = [nn.Linear(512, 256)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(256))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate)])
        self.fc = nn.Sequential(*layers)
        self.final_bilstm = nn.LSTM(input_size=256, hidden_size=128, num_layers=2, bidirectional=True, batch_first=True)
        self._tail(num_classes)

    def _tail(self, num_classes):
        self.classifier = nn.Linear(2 * 128, num_classes)


class LstmResNetTransPlan(ModelPlan):
    """audio/models/lstm_resnet_trans_model.py:88-104: the projection repeated over seq_len steps + positional encoding,
    TransformerEncoder, mean over the steps, classifier."""

    def build(self, m, spec):
        B, wb, T = self.B, self.with_backward, m.seq_len
        h, dh, E = lstm_resnet_front(self, m)
        pe = m.pos_encoder.pe[0, :T]
        xs = self.alloc(B * T * E)
        dxs = self.alloc(B * T * E) if wb else None
        self.fwd.add("lr_add_bcast", h, pe, xs, B, T, E)               # fc_out.unsqueeze(1).repeat(1, T, 1) + pe[:, :T]
        if wb:
            ones = torch.ones(B * T * E, dtype=torch.float32, device=self.dev)
            self.bufs.append(ones)
            self.bgroup().add("lr_frame_reduce", dxs, ones, dh, B, T, E, 1)             # dh[b] = sum_t dxs[b, t]
        x, dx = xs, dxs
        for layer in m.transformer.layers:
            x, dx = self.transformer_encoder_layer(x, dx, B, T, layer)
        pooled, dpooled = self.avgpool(engine.T2.of(B, 1, T, E, x, dx))                 # x_encoded.mean(dim=1)
        C = self.num_classes
        logits = self.alloc(B * C)
        dlogits = self.alloc(B * C) if wb else None
        self.linear(pooled, E, B, m.classifier.weight, m.classifier.bias, logits, C)
        if wb:
            self.linear_bwd(self.bgroup(), pooled, E, B, m.classifier.weight, m.classifier.bias, dlogits, C, dx=dpooled, ldx=E)
        self.set_logits(logits, dlogits)


class PositionalEncodingBuffer(nn.Module):
    """lstm_resnet_trans_model.py:6-19: the table is a registered buffer (state_dict key `pos_encoder.pe`)."""

    def __init__(self, dim, max_len=5000):
        super().__init__()
        import numpy as np
        pe = torch.zeros(max_len, dim)
        position = torch.arange(0, max_len).unsqueeze(1).float()
        div_term = torch.exp(torch.arange(0, dim, 2).float() * (-np.log(10000.0) / dim))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class LSTMResNetWithTransformer(PlanModel):
    """audio/models/lstm_resnet_trans_model.py:22-104 (model.name == "lstm_resnet_trans")."""
    INPUTS = ("audio",)
    PLAN = LstmResNetTransPlan
    DEFAULT_LR = 5e-4
    DEFAULT_WD = 1e-4

    def __init__(self, num_classes=40, input_size=117, transformer_dim=256, num_heads=4, num_layers=2, seq_len=10,
                 dropout_rate=0.3, use_batchnorm=True, pretrained_state_dict=None, precision=None, encoder_dropout=0.1):
        super().__init__()
        self._init_base(num_classes, types.SimpleNamespace(get=lambda k, d=None: d), precision)
        self.seq_len = seq_len
        self.use_bn = use_batchnorm
        self.initial_bilstm = nn.LSTM(input_size=input_size, hidden_size=64, num_layers=2, bidirectional=True, batch_first=True)
        self.resnet = resnet18(weights=None)
        if pretrained_state_dict is not None:
            self.resnet.load_state_dict(pretrained_state_dict)
        self.resnet.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.resnet.fc = nn.Identity()
        layers = [nn.Linear(512, transformer_dim)]
        if use_batchnorm:
            layers.append(nn.BatchNorm1d(transformer_dim))
        layers.extend([nn.ReLU(), nn.Dropout(dropout_rate)])
        self.fc = nn.Sequential(*layers)
        self.pos_encoder = PositionalEncodingBuffer(dim=transformer_dim, max_len=seq_len)
        # nn.TransformerEncoderLayer's own default dropout (0.1) applies in the reference; encoder_dropout exposes it
        encoder_layer = nn.TransformerEncoderLayer(d_model=transformer_dim, nhead=num_heads, batch_first=True,
                                                   dropout=encoder_dropout)
        self.transformer = nn.TransformerEncoder(encoder_layer, num_layers=num_layers)
        self.classifier = nn.Linear(transformer_dim, num_classes)


class Attention(nn.Module):
    """lstm_resnet_attn_model.py:5-14 (parameters only)."""

    def __init__(self, input_dim):
        super().__init__()
        self.attn = nn.Linear(input_dim, 1)


class DeepAudioNetWithAttention(LSTMResNet):
    """audio/models/lstm_resnet_attn_model.py:17-88 (model.name == "lstm_resnet_attn")."""
    PLAN = LstmResNetAttnPlan

    def _tail(self, num_classes):                            # construction order of the reference (:54-58)
        self.attention = Attention(input_dim=256)
        self.classifier = nn.Linear(256, num_classes)


def get_model(num_classes, input_size, model_name, version=None):
    """audio/train.py:118-134 (the variants with a lipread_b200 plan)."""
    if model_name == "resnet":
        return AudioResNet(num_classes=num_classes)
    if model_name == "resnet_lstm":
        return AudioResNetLSTM(num_classes=num_classes)
    if model_name == "vgg":
        return VGGAudioClassifier(num_classes=num_classes, version=version or 11)
    if model_name == "vgg_lstm":
        return VGGWithLSTMClassifier(num_classes=num_classes, version=version or 11)
    if model_name == "lstm_resnet":
        return LSTMResNet(num_classes=num_classes, input_size=input_size)
    if model_name == "lstm_resnet_attn":
        return DeepAudioNetWithAttention(num_classes=num_classes, input_size=input_size)
    if model_name == "lstm_resnet_trans":
        return LSTMResNetWithTransformer(num_classes=num_classes, input_size=input_size)
    raise ValueError(f"Invalid model name: {model_name}")
