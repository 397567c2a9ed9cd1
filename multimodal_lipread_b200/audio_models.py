"""Audio-only classifiers behind the reference's nn.Module surface (audio/models/*.py).

  AudioResNet      audio/models/resnet_model.py:5-39   (model.name == "resnet": BASELINE config 1)

forward(spec (B,80,117) f32 log-mel) -> (B, num_classes); with a raw (B,20000) waveform the fused log-mel kernel
runs first.  Sub-modules are parameter containers (reference names / construction order / state_dict keys)."""
import types

import torch.nn as nn
from torchvision.models import resnet18

from . import engine
from ._lib import ACT_NONE, ACT_RELU
from .model_base import ModelPlan, PlanModel, N_MELS, N_FRAMES_OUT


class AudioResNetPlan(ModelPlan):
    def build(self, m, spec):
        B, wb = self.B, self.with_backward
        mel = self.audio_input()
        net = m.resnet
        # x.unsqueeze(1): (B,1,80,117) NCHW with one channel == NHWC with C = 1
        frames = (mel, (0, B, 1, N_MELS, N_FRAMES_OUT, N_MELS * N_FRAMES_OUT, 0, 0, N_FRAMES_OUT, 1), 1.0)
        last = self.resnet_features(net, frames)
        feat, dfeat = self.avgpool(last)
        head = list(net.fc)
        fc1 = head[0]
        D = fc1.out_features
        i = 1
        if isinstance(head[i], nn.BatchNorm1d):
            cur, dcur = self.linear_bn_act(feat, dfeat, B, fc1, head[i], ACT_RELU)
            i += 2                                           # BatchNorm1d, ReLU
        else:
            cur = self.alloc(B * D)
            dcur = self.alloc(B * D) if wb else None
            self.linear(feat, fc1.in_features, B, fc1.weight, fc1.bias, cur, D, act=ACT_RELU)
            if wb:
                g = self.bgroup()
                g.add("lr_act_bwd", dcur, cur, B * D, ACT_RELU)
                self.linear_bwd(g, feat, fc1.in_features, B, fc1.weight, fc1.bias, dcur, D, dx=dfeat, ldx=fc1.in_features)
            i += 1                                           # ReLU
        drop, fc2 = head[i], head[i + 1]
        cur, dcur = self.dropout(cur, dcur, B * D, drop.p)
        logits = self.alloc(B * self.num_classes)
        dlogits = self.alloc(B * self.num_classes) if wb else None
        self.linear(cur, D, B, fc2.weight, fc2.bias, logits, self.num_classes, act=ACT_NONE)
        if wb:
            self.linear_bwd(self.bgroup(), cur, D, B, fc2.weight, fc2.bias, dlogits, self.num_classes, dx=dcur, ldx=D)
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


def get_model(num_classes, input_size, model_name, version=None):
    """audio/train.py:118-134 (the variants with a lipread_b200 plan)."""
    if model_name == "resnet":
        return AudioResNet(num_classes=num_classes)
    raise ValueError(f"Invalid model name: {model_name}")
