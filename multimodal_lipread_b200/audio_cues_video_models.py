"""Audio + cue + video triple-fusion models behind the reference's nn.Module surface (audio_cues_video/models/*.py).

  MultimodalAttentionLate        audio_cues_video/models/late_fusion_mobile.py:6-107   (train.model_name == "late_fusion_mobile")
  MultimodalAttentionLateResNet  audio_cues_video/models/late_fusion_resnet.py:6-99    (train.model_name == "late_fusion_resnet")
  MultimodalAttentionMiddle      audio_cues_video/models/middle_fusion_mobile.py:84-110  ("middle_fusion_mobile")
  MultimodalAttentionMiddleResNet audio_cues_video/models/middle_fusion_resnet.py:164-191 ("middle_fusion_resnet", frozen encoders)
  MultimodalAttentionEarly       audio_cues_video/models/early_fusion_mobile.py:179-213  ("early_fusion_mobile", frozen encoders)
  MultimodalAttentionEarlyResNet audio_cues_video/models/early_fusion_resnet.py:158-191  ("early_fusion_resnet", frozen encoders)

forward(mel (B,80,117), cue (B,768), lip (B,3,T,H,W) [or uint8 (B,T,H,W,3)]) -> (B, num_classes).
Sub-modules are parameter containers (reference names / construction order / state_dict keys)."""
import torch
import torch.nn as nn
from torchvision.models import mobilenet_v2, resnet18

from . import engine
from ._lib import ACT_RELU
from .model_base import ModelPlan, PlanModel, N_MELS, N_FRAMES_OUT, load_torchvision_weights
from .video_models import TimeDistributed


def _load_pretrained_trunks(model):
    """pretrained=True is the reference's default (audio_cues_video/models/*: resnet18 / mobilenet_v2 IMAGENET1K_V1 for
    the audio and video encoders).  Offline the checkpoints come in as pretrained_state_dicts = {"audio": resnet18
    state_dict, "video": mobilenet_v2 / resnet18 state_dict}; asked for and not supplied, the deviation is announced."""
    pretrained, dicts = model._pretrained
    if dicts:
        with torch.no_grad():
            for attr, sd in dicts.items():
                if load_torchvision_weights(getattr(model, attr), sd) == 0:
                    raise ValueError(f"pretrained_state_dicts[{attr!r}] matches no tensor of the {attr} encoder")
    elif pretrained:
        import warnings
        warnings.warn(f"{type(model).__name__}: the reference initialises its audio / video trunks from ImageNet "
                      "(pretrained=True); no pretrained_state_dicts were supplied, so they keep their random init "
                      "(the frozen-backbone variants then train on frozen random features)", stacklevel=3)


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

    def __init__(self, num_classes, cue_dim=768, video_cfg=None, pretrained=False, precision=None, lstm_dropout=0.3,
                 pretrained_state_dicts=None):
        super().__init__()
        self._pretrained = (pretrained, pretrained_state_dicts)
        self._init_base(num_classes, type("C", (), {"get": staticmethod(lambda k, d=None: d)})(), precision)
        self.audio = AudioEncoder()
        self.cue = CueEncoder(cue_dim)
        vdim = int(video_cfg.get("model", {}).get("feature_dim", 256)) if video_cfg else 256
        self.video = MobileNetLSTM(vdim, dropout=lstm_dropout)
        self.afc = nn.Linear(512, num_classes)
        self.cfc = nn.Linear(256, num_classes)
        self.vfc = nn.Linear(vdim, num_classes)
        self.attn = AttentionFusion(num_classes)
        _load_pretrained_trunks(self)


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

    def __init__(self, num_classes, cue_dim=768, video_cfg=None, pretrained=False, precision=None, lstm_dropout=0.3,
                 pretrained_state_dicts=None):
        super().__init__()
        self._pretrained = (pretrained, pretrained_state_dicts)
        self._init_base(num_classes, type("C", (), {"get": staticmethod(lambda k, d=None: d)})(), precision)
        self.audio = AudioEncoder()
        self.cue = CueEncoder(cue_dim)
        vdim = int(video_cfg.get("model", {}).get("feature_dim", 256)) if video_cfg else 256
        self.video = ResNetLSTM(vdim, dropout=lstm_dropout)
        self.afc = nn.Linear(512, num_classes)
        self.vfc = nn.Linear(vdim, num_classes)
        self.cfc = nn.Linear(256, num_classes)
        self.attn = AttentionFusion(num_classes)
        _load_pretrained_trunks(self)


# ----------------------------------------------------------------------------------------------------------------
# Early / middle attention fusion (256-d embeddings through AttentionFusion, then a classifier head) and the
# frozen-backbone variants with the chunked TimeDistributed.
# ----------------------------------------------------------------------------------------------------------------
class SafeCheckpoint(nn.Module):
    """early_fusion_mobile.py:62-72 (parameter container: keys `<name>.module.*`).  Activation checkpointing only
    changes what autograd stores; it never engages for a frozen backbone fed with inputs that need no gradient."""

    def __init__(self, module, enabled=True):
        super().__init__()
        self.module = module
        self.enabled = bool(enabled)


class TimeDistributedChunked(nn.Module):
    """early_fusion_mobile.py:31-56: the CNN sees `chunk_size` time steps of every clip at a time, so its train-mode
    BatchNorm statistics are per chunk (B * chunk frames) and the running statistics move once per chunk."""

    def __init__(self, module, chunk_size=4):
        super().__init__()
        self.module = module
        self.chunk_size = int(chunk_size)


def _freeze(params):
    for p in params:
        p.requires_grad = False


class FrozenAudioEncoder(nn.Module):
    """early_fusion_mobile.py:126-154 (`encoder`), middle_fusion_resnet.py:114-138 (`enc`): resnet18, 1-channel conv1,
    fc = Identity, every parameter frozen, wrapped in SafeCheckpoint."""

    def __init__(self, attr):
        super().__init__()
        net = resnet18(weights=None)
        net.conv1 = nn.Conv2d(1, 64, 7, 2, 3, bias=False)
        net.fc = nn.Identity()
        _freeze(net.parameters())
        setattr(self, attr, SafeCheckpoint(net, True))
        self.output_dim = 512


class FrozenVideoLSTM(nn.Module):
    """early_fusion_mobile.py:78-121 (MobileNetV2, `features` frozen), early_fusion_resnet.py / middle_fusion_resnet.py
    (ResNet-18, all frozen): SafeCheckpoint(cnn) under a chunked TimeDistributed + 1-layer BiLSTM."""

    def __init__(self, trunk, feature_dim=256):
        super().__init__()
        if trunk == "mobilenet":
            base = mobilenet_v2(weights=None)
            base.classifier = nn.Identity()
            seq = nn.Sequential(base.features, nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten())
            _freeze(base.features.parameters())
            width = 1280
        else:
            base = resnet18(weights=None)
            base.fc = nn.Identity()
            _freeze(base.parameters())
            seq = nn.Sequential(base)
            width = 512
        self.cnn = SafeCheckpoint(seq, True)
        self.td = TimeDistributedChunked(self.cnn, chunk_size=4)
        self.lstm = nn.LSTM(width, feature_dim // 2, num_layers=1, bidirectional=True, batch_first=True, dropout=0.0)
        self.output_dim = feature_dim


class CueEncoderEarly(nn.Module):
    """early_fusion_mobile.py:160-173: Linear, BatchNorm1d, ReLU, Dropout(0.3), Linear, ReLU."""

    def __init__(self, input_dim=768, dropout=0.3):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(input_dim, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(dropout),
                                 nn.Linear(256, 256), nn.ReLU())
        self.output_dim = 256


def _unwrap(mod):
    return mod.module if isinstance(mod, SafeCheckpoint) else mod


class AttentionFusionPlan(ModelPlan):
    """a = ap(audio(mel)), c = [cp](cue(cue)), v = vp(video(lip)) -> AttentionFusion([a, c, v]) -> classifier head.
    middle_fusion_mobile.py:103-110, middle_fusion_resnet.py:184-191, early_fusion_mobile.py:205-213,
    early_fusion_resnet.py:183-191."""

    def build(self, m, spec):
        from .audio_models import bn_head
        B, wb = self.B, self.with_backward
        mel = self.audio_input()
        cue = self.vector_input("cue", m.cue.net[0].in_features)
        video, layout, scale = self.video_input()
        is_u8, _, T, H, W, sb, st, sc, sh, sw = layout
        S, D = 3, m.attn.attn[0].in_features
        stacked = self.alloc(B * S * D)                       # [B, 3, D]: every modality writes its row in place
        dstacked = self.alloc(B * S * D) if wb else None
        ldS = S * D

        def into_stack(x, dx, fc, slot, act=engine.ACT_NONE):
            dst = stacked.data_ptr() + 4 * slot * D
            self.linear(x, fc.in_features, B, fc.weight, fc.bias, dst, ldS, act=act)
            if wb:
                g = self.bgroup()
                dptr = dstacked.data_ptr() + 4 * slot * D
                if act != engine.ACT_NONE:
                    raise NotImplementedError("activation on a strided stack row")
                self.linear_bwd(g, x, fc.in_features, B, fc.weight, fc.bias, dptr, ldS, dx=dx, ldx=fc.in_features)

        # ---- audio: ResNet-18 on the 1-channel log-mel image (frozen in the early / middle-resnet variants)
        enc = _unwrap(m.audio.enc if hasattr(m.audio, "enc") else m.audio.encoder)
        frames = (mel, (0, B, 1, N_MELS, N_FRAMES_OUT, N_MELS * N_FRAMES_OUT, 0, 0, N_FRAMES_OUT, 1), 1.0)
        with self.frozen(not any(p.requires_grad for p in enc.parameters())):
            a_feat, a_dfeat = self.avgpool(self.resnet_features(enc, frames))
        into_stack(a_feat, a_dfeat, m.ap, 0)
        # ---- cue: Linear -> BatchNorm1d -> ReLU -> [Dropout] -> Linear [-> ReLU -> cp]
        net = list(m.cue.net)
        h, dh = self.linear_bn_act(cue, None, B, net[0], net[1], ACT_RELU)
        i = 3
        if isinstance(net[i], nn.Dropout):
            h, dh = self.dropout(h, dh, B * net[0].out_features, net[i].p)
            i += 1
        fc = net[i]
        if hasattr(m, "cp"):
            act = ACT_RELU if (i + 1 < len(net) and isinstance(net[i + 1], nn.ReLU)) else engine.ACT_NONE
            c_out = self.alloc(B * fc.out_features)
            c_dout = self.alloc(B * fc.out_features) if wb else None
            self.linear(h, fc.in_features, B, fc.weight, fc.bias, c_out, fc.out_features, act=act)
            if wb:
                g = self.bgroup()
                if act != engine.ACT_NONE:
                    g.add("lr_act_bwd", c_dout, c_out, B * fc.out_features, act)
                self.linear_bwd(g, h, fc.in_features, B, fc.weight, fc.bias, c_dout, fc.out_features, dx=dh, ldx=fc.in_features)
            into_stack(c_out, c_dout, m.cp, 1)
        else:
            if i + 1 != len(net):
                raise NotImplementedError("cue encoder ending in an activation without a cp projection")
            into_stack(h, dh, fc, 1)
        # ---- video: MobileNetV2 / ResNet-18 per frame (whole clip, or `chunk_size` time steps at a time) + BiLSTM
        trunk = _unwrap(m.video.cnn)[0]
        chunk = getattr(m.video.td, "chunk_size", None) or T
        Cf = m.video.lstm.input_size
        tdim = 1 if is_u8 else 2

        def features(frames_):
            last = (self.resnet_features(trunk, frames_) if hasattr(trunk, "conv1") else self.mbv2_features(trunk, frames_))
            return self.avgpool(last)
        with self.frozen(not any(p.requires_grad for p in trunk.parameters())):
            if chunk >= T:
                v_feat, v_dfeat = features((video, layout, scale))
            else:
                v_feat, v_dfeat = self.alloc(B * T * Cf), None
                if self.with_backward:
                    raise NotImplementedError("chunked TimeDistributed over a trainable backbone")
                for t0 in range(0, T, chunk):
                    tc = min(chunk, T - t0)
                    part, _ = features((video.narrow(tdim, t0, tc), (is_u8, B, tc, H, W, sb, st, sc, sh, sw), scale))
                    self.fwd.add("lr_copy2d", v_feat.data_ptr() + 4 * t0 * Cf, T * Cf, part, tc * Cf, B, tc * Cf)
        Dv = m.video.output_dim
        v_out = self.alloc(B * Dv)
        v_dout = self.alloc(B * Dv) if wb else None
        self.bilstm_last(v_feat, v_dfeat, Cf, B, T, m.video.lstm, v_out, Dv, v_dout if wb else 0)
        into_stack(v_out, v_dout, m.vp, 2)
        # ---- attention fusion over the three embeddings
        att = m.attn.attn
        Hh = att[0].out_features
        ah = self.alloc(B * S * Hh)
        scores, weights = self.alloc(B * S), self.alloc(B * S)
        fused = self.alloc(B * D)
        dfused = self.alloc(B * D) if wb else None
        self.linear(stacked, D, B * S, att[0].weight, att[0].bias, ah, Hh, act=ACT_RELU)
        self.linear(ah, Hh, B * S, att[2].weight, att[2].bias, scores, 1)
        self.fwd.add("lr_attn_fuse_fwd", stacked, scores, weights, fused, B, S, D)
        if wb:
            dah, dscores = self.alloc(B * S * Hh), self.alloc(B * S)
            g = self.bgroup()
            g.add("lr_attn_fuse_bwd", stacked, weights, dfused, dstacked, dscores, B, S, D)
            self.linear_bwd(g, ah, Hh, B * S, att[2].weight, att[2].bias, dscores, 1, dx=dah, ldx=Hh)
            g.add("lr_act_bwd", dah, ah, B * S * Hh, ACT_RELU)
            self.linear_bwd(g, stacked, D, B * S, att[0].weight, att[0].bias, dah, Hh, dx=dstacked, ldx=D,
                            dx_residual=dstacked, ldr=D)
        head = m.cls if hasattr(m, "cls") else m.classifier
        logits, dlogits = bn_head(self, fused, dfused, B, list(head))
        self.set_logits(logits, dlogits)


class _AttentionFusionModel(PlanModel):
    INPUTS = ("audio", "cue", "video")
    PLAN = AttentionFusionPlan
    DEFAULT_LR = 1e-4            # audio_cues_video/train.py:162  cfg.get("train.lr", 1e-4)

    def _start(self, num_classes, pretrained, precision, pretrained_state_dicts=None):
        super().__init__()
        self._pretrained = (pretrained, pretrained_state_dicts)
        self._init_base(num_classes, type("C", (), {"get": staticmethod(lambda k, d=None: d)})(), precision)

    @staticmethod
    def _vdim(video_cfg):
        return int(video_cfg.get("model", {}).get("feature_dim", 256)) if video_cfg else 256


class MultimodalAttentionMiddle(_AttentionFusionModel):
    """audio_cues_video/models/middle_fusion_mobile.py:84-110 (train.model_name == "middle_fusion_mobile")."""

    def __init__(self, num_classes, cue_dim=768, video_cfg=None, pretrained=False, precision=None, lstm_dropout=0.3,
                 head_dropout=0.4, pretrained_state_dicts=None):
        self._start(num_classes, pretrained, precision, pretrained_state_dicts)
        self.audio = AudioEncoder()
        self.cue = CueEncoder(cue_dim)
        vdim = self._vdim(video_cfg)
        self.video = MobileNetLSTM(vdim, dropout=lstm_dropout)
        self.ap = nn.Linear(512, 256)
        self.vp = nn.Linear(vdim, 256)
        self.attn = AttentionFusion(256)
        self.cls = nn.Sequential(nn.Linear(256, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Dropout(head_dropout),
                                 nn.Linear(512, num_classes))
        _load_pretrained_trunks(self)


class MultimodalAttentionMiddleResNet(_AttentionFusionModel):
    """audio_cues_video/models/middle_fusion_resnet.py:164-191 ("middle_fusion_resnet"): frozen ResNet-18 encoders."""

    def __init__(self, num_classes, cue_dim=768, video_cfg=None, pretrained=False, precision=None, head_dropout=0.4,
                 pretrained_state_dicts=None):
        self._start(num_classes, pretrained, precision, pretrained_state_dicts)
        self.audio = FrozenAudioEncoder("enc")
        self.cue = CueEncoder(cue_dim)
        vdim = self._vdim(video_cfg)
        self.video = FrozenVideoLSTM("resnet", vdim)
        self.ap = nn.Linear(512, 256)
        self.vp = nn.Linear(vdim, 256)
        self.attn = AttentionFusion(256)
        self.cls = nn.Sequential(nn.Linear(256, 512), nn.ReLU(), nn.Dropout(head_dropout), nn.Linear(512, num_classes))
        _load_pretrained_trunks(self)


class MultimodalAttentionEarly(_AttentionFusionModel):
    """audio_cues_video/models/early_fusion_mobile.py:179-213 ("early_fusion_mobile"): frozen ResNet-18 audio encoder,
    frozen MobileNetV2 features under the chunked TimeDistributed, cp projection."""

    def __init__(self, num_classes, cue_dim=768, video_cfg=None, pretrained=False, precision=None, cue_dropout=0.3,
                 head_dropout=0.4, trunk="mobilenet", pretrained_state_dicts=None):
        self._start(num_classes, pretrained, precision, pretrained_state_dicts)
        self.audio = FrozenAudioEncoder("encoder")
        self.cue = CueEncoderEarly(cue_dim, cue_dropout)
        vdim = self._vdim(video_cfg)
        self.video = FrozenVideoLSTM(trunk, vdim)
        self.ap = nn.Linear(512, 256)
        self.vp = nn.Linear(vdim, 256)
        self.cp = nn.Linear(256, 256)
        self.attn = AttentionFusion(256)
        self.classifier = nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.Dropout(head_dropout), nn.Linear(256, num_classes))
        _load_pretrained_trunks(self)


class MultimodalAttentionEarlyResNet(MultimodalAttentionEarly):
    """audio_cues_video/models/early_fusion_resnet.py:158-191 ("early_fusion_resnet")."""

    def __init__(self, num_classes, cue_dim=768, video_cfg=None, pretrained=False, precision=None, cue_dropout=0.3,
                 head_dropout=0.4, pretrained_state_dicts=None):
        super().__init__(num_classes, cue_dim, video_cfg, pretrained, precision, cue_dropout, head_dropout, trunk="resnet",
                         pretrained_state_dicts=pretrained_state_dicts)
