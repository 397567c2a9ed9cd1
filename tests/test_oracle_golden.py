"""Pin the oracle (oracle/) against vectors produced by the reference's own modules
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel as olm
from oracle.frontend import AudioProcessorPort, lips_u8_to_model_input
from oracle.av_models import MidFusionFastOracle
from multimodal_lipread_b200 import synthetic


@pytest.fixture(scope="module")
def lg(golden_dir):
    return np.load(os.path.join(golden_dir, "logmel_golden.npz"))


@pytest.fixture(scope="module")
def mg(golden_dir):
    return np.load(os.path.join(golden_dir, "midfusion_golden.npz"))


def test_window_and_filterbank_match_torchaudio(lg):
    assert np.abs(olm.hann_window() - lg["window"]).max() < 5e-7   # torch builds it in fp32
    fb = olm.melscale_fbanks()
    assert fb.shape == (201, 80)
    assert np.abs(fb - lg["fb"]).max() < 1e-5          # torchaudio builds fb in fp32
    assert (lg["fb"] > 0).sum() == 392 and ((lg["fb"] > 0).sum(0) > 0).all()
    assert abs(float((lg["window"].astype(np.float64) ** 2).sum()) - 150.0) < 1e-4


def test_port_reproduces_reference_bitwise(lg):
    ap = AudioProcessorPort()
    wave = torch.from_numpy(lg["wave"])
    out = ap.batch_frontend_loop(wave).numpy()
    ref = lg["out"]
    assert np.abs(out - ref).max() == 0.0              # same torch, same call sequence


def test_float64_restatement_matches_reference(lg):
    """fp32 reference vs float64 restatement: fp32 round-off only (<= ~2e-4 abs)."""
    ref = lg["out"].astype(np.float64)
    out = olm.logmel_frontend(lg["wave"], window=lg["window"], fb=lg["fb"])
    n = ref.shape[0] - 1                                # last clip is silence, checked below
    err = np.abs(out[:n] - ref[:n]).max(axis=(1, 2))
    scale = np.abs(ref[:n]).max(axis=(1, 2))
    assert (err <= 1e-4 * scale).all(), (err, scale)
    # Silent clip: every log-mel value is ln(1e-9), so (x - mean) is pure fp32 round-off in the
    # reference (a constant -0.9994...), amplified by the 1e-9-regularised division.  Exact
    # arithmetic gives 0; the restatement (and the CUDA kernel) return that.
    assert np.unique(ref[n]).size == 1 and abs(ref[n]).max() <= 1.0
    assert (out[n] == 0.0).all()


def test_raw_logmel_and_padding_value(lg):
    raw = olm.log_mel(lg["wave"], window=lg["window"], fb=lg["fb"])
    ref = lg["logmel_raw"].astype(np.float64)
    assert raw.shape == ref.shape == (8, 80, 126)
    assert np.abs(raw - ref).max() < 2e-3               # un-normalised log power, values ~ 5..20
    assert np.allclose(ref[-1], np.log(1e-9), atol=1e-5)


def test_frames_reflect_padding():
    x = np.arange(20000, dtype=np.float64)
    fr = olm.frames(x)
    assert fr.shape == (126, 400)
    assert fr[0, 0] == 200 and fr[0, 199] == 1 and fr[0, 200] == 0 and fr[0, 399] == 199
    assert fr[125, 399] == 19998 - 199 and fr[125, 200] == 19998
    assert fr[125, 199] == 19999


@pytest.mark.parametrize("size", [44, 88])
def test_midfusion_oracle_matches_reference(mg, size):
    torch.manual_seed(0)
    model = MidFusionFastOracle(40)
    model.train()
    assert [n for n, _ in model.named_parameters()] == list(mg["param_names"])
    assert list(model.state_dict().keys()) == list(mg["state_keys"])
    wsum0 = np.array([p.detach().double().sum().item() for p in model.parameters()])
    np.testing.assert_allclose(wsum0, mg[f"wsum_before_{size}"], rtol=0, atol=0)   # same seeded init
    B = 2
    wav = synthetic.make_waveforms(B, pad_fraction=0.5)
    mel = AudioProcessorPort().batch_frontend_loop(wav)
    video = lips_u8_to_model_input(synthetic.make_lips_u8(B, size=size))
    labels = synthetic.make_labels(B, 40)
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    opt.zero_grad()
    logits = model(mel, video)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), mg[f"logits_{size}"], rtol=1e-5, atol=1e-6)
    assert abs(loss.item() - float(mg[f"loss_{size}"])) < 1e-6
    gnorm = np.array([p.grad.double().norm().item() for p in model.parameters()])
    np.testing.assert_allclose(gnorm, mg[f"grad_norm_{size}"], rtol=1e-4, atol=1e-7)
    opt.step()
    wsum1 = np.array([p.detach().double().sum().item() for p in model.parameters()])
    np.testing.assert_allclose(wsum1, mg[f"wsum_after_{size}"], rtol=1e-5, atol=1e-4)
    sd = model.state_dict()
    np.testing.assert_allclose(sd["video_cnn.features.0.1.running_mean"].numpy(), mg[f"rm_stem_{size}"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(sd["video_cnn.features.12.1.running_var"].numpy(), mg[f"rv_last_{size}"], rtol=1e-5, atol=1e-7)
    model.eval()
    with torch.no_grad():
        np.testing.assert_allclose(model(mel, video).numpy(), mg[f"logits_eval_{size}"], rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------
# configs 1, 2, 4: oracle ports pinned to the reference's own AudioResNet / ResNet2DBiLSTM / EarlyFusionAVMobileNet
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def models_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "models_golden.npz"))


@pytest.mark.parametrize("name", ["early_fusion_mobilenet", "video_resnet_lstm", "video_resnet34_lstm", "video_resnet50_lstm", "audio_resnet", "acv_late_fusion_mobile", "video_mobilenet_lstm",
                                  "acv_late_fusion_resnet", "video_vgg_lstm", "video_cnn", "video_resnet_attn", "video_resnet_trans", "video_shufflenet_lstm", "audio_resnet_lstm", "audio_vgg", "audio_vgg_lstm", "audio_lstm_resnet", "audio_lstm_resnet_attn", "audio_lstm_resnet_trans",
                                  "late_fusion_mobilenet", "middle_fusion_mobilenet", "early_fusion_fast", "late_fusion_fast",
                                  "acv_middle_fusion_mobile", "acv_middle_fusion_resnet", "acv_early_fusion_mobile", "acv_early_fusion_resnet"])
def test_model_oracles_match_reference(models_golden, name):
    from oracle import av_models as O
    g = models_golden
    B, T, size = int(g[f"{name}_B"]), int(g[f"{name}_T"]), int(g[f"{name}_size"])
    C = 8 if name.startswith("audio_") else 40
    torch.manual_seed(0)
    if name == "early_fusion_mobilenet":
        model, lr, wd = O.EarlyFusionMobileNetOracle(C, lstm_dropout=0.0, head_dropout=0.0), 3e-4, 0.0
    elif name == "video_resnet_lstm":
        model, lr, wd = O.ResNet2DBiLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0}})), 5e-5, 1e-5
    elif name in ("video_resnet34_lstm", "video_resnet50_lstm"):
        model, lr, wd = O.ResNet2DBiLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0, "resnet_version": int(name[12:14])}})), 5e-5, 1e-5
    elif name == "acv_late_fusion_mobile":
        model, lr, wd = O.LateFusionMobileOracle(C, lstm_dropout=0.0), 1e-5, 0.0
    elif name == "video_vgg_lstm":
        model, lr, wd = O.VGGLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0}})), 5e-5, 1e-5
    elif name == "video_cnn":
        model, lr, wd = O.CNNOnlyOracle(C, O.DictConfig({"model": {"dropout": 0.0}})), 5e-5, 1e-5
    elif name == "video_shufflenet_lstm":
        model, lr, wd = O.ShuffleNet2DBiLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0}})), 5e-5, 1e-5
    elif name == "video_resnet_trans":
        model, lr, wd = O.ResNet2DTransformerOracle(C, O.DictConfig({"model": {"dropout": 0.0}})), 5e-5, 1e-5
    elif name == "audio_lstm_resnet_trans":
        model, lr, wd = O.LSTMResNetTransOracle(C, dropout_rate=0.0, encoder_dropout=0.0), 5e-4, 1e-4
    elif name == "video_resnet_attn":
        model, lr, wd = O.ResNet2DAttentionOracle(C, O.DictConfig({"model": {"dropout": 0.0}})), 5e-5, 1e-5
    elif name == "audio_resnet_lstm":
        model, lr, wd = O.AudioResNetLSTMOracle(C, dropout_rate=0.0), 5e-4, 1e-4
    elif name == "audio_vgg":
        model, lr, wd = O.VGGAudioOracle(C, version=11, dropout_rate=0.0), 5e-4, 1e-4
    elif name == "audio_vgg_lstm":
        model, lr, wd = O.VGGLstmAudioOracle(C, version=11, dropout_rate=0.0), 5e-4, 1e-4
    elif name == "audio_lstm_resnet":
        model, lr, wd = O.LSTMResNetOracle(C, dropout_rate=0.0), 5e-4, 1e-4
    elif name == "audio_lstm_resnet_attn":
        model, lr, wd = O.LSTMResNetAttnOracle(C, dropout_rate=0.0), 5e-4, 1e-4
    elif name == "video_mobilenet_lstm":
        model, lr, wd = O.MobileNetLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0}})), 5e-5, 1e-5
    elif name == "acv_late_fusion_resnet":
        model, lr, wd = O.LateFusionResNetOracle(C, lstm_dropout=0.0), 1e-5, 0.0
    elif name == "late_fusion_mobilenet":
        model, lr, wd = O.LateFusionAVMobileNetOracle(C), 3e-4, 0.0
    elif name == "middle_fusion_mobilenet":
        model, lr, wd = O.MidFusionAVMobileNetOracle(C, head_dropout=0.0), 3e-4, 0.0
    elif name == "early_fusion_fast":
        model, lr, wd = O.EarlyFusionFastOracle(C), 3e-4, 0.0
    elif name == "late_fusion_fast":
        model, lr, wd = O.LateFusionFastOracle(C), 3e-4, 0.0
    elif name in ("acv_middle_fusion_mobile", "acv_middle_fusion_resnet", "acv_early_fusion_mobile", "acv_early_fusion_resnet"):
        kind = "_".join(name.split("_")[1::2])                       # acv_middle_fusion_mobile -> middle_mobile
        model, lr, wd = O.AttentionFusionACVOracle(kind, C, lstm_dropout=0.0, cue_dropout=0.0, head_dropout=0.0), 1e-4, 0.0
    else:
        model, lr, wd = O.AudioResNetOracle(C, dropout_rate=0.0), 5e-4, 1e-4
    model.train()
    assert [n for n, _ in model.named_parameters()] == list(g[f"{name}_param_names"])
    assert list(model.state_dict().keys()) == list(g[f"{name}_state_keys"])
    wsum0 = np.array([p.detach().double().sum().item() for p in model.parameters()])
    np.testing.assert_allclose(wsum0, g[f"{name}_wsum_before"], rtol=0, atol=0)     # same seeded init
    wav = synthetic.make_waveforms(B, pad_fraction=0.5)
    mel = AudioProcessorPort().batch_frontend_loop(wav)
    video = lips_u8_to_model_input(synthetic.make_lips_u8(B, size=size)[:, :T].contiguous())
    labels = synthetic.make_labels(B, C)
    inputs = {"video_resnet_lstm": (video,), "video_mobilenet_lstm": (video,), "video_vgg_lstm": (video,), "video_cnn": (video,), "video_resnet_attn": (video,), "video_resnet_trans": (video,), "video_shufflenet_lstm": (video,), "audio_lstm_resnet_trans": (mel,),
              "audio_resnet": (mel,), "audio_resnet_lstm": (mel,), "audio_vgg": (mel,), "audio_vgg_lstm": (mel,), "audio_lstm_resnet": (mel,), "audio_lstm_resnet_attn": (mel,),
              "acv_late_fusion_mobile": (mel, synthetic.make_cues(B), video),
              "acv_late_fusion_resnet": (mel, synthetic.make_cues(B), video)}.get(name, (mel, video))
    if name.startswith("acv_"):
        inputs = (mel, synthetic.make_cues(B), video)
    if name.startswith("video_"):
        inputs = (video,)
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    logits, loss = O.train_step_generic(model, opt, inputs, labels)
    np.testing.assert_allclose(logits.numpy(), g[f"{name}_logits"], rtol=1e-5, atol=1e-6)
    assert abs(loss - float(g[f"{name}_loss"])) < 1e-6
    gnorm = np.array([0.0 if p.grad is None else p.grad.double().norm().item() for p in model.parameters()])
    np.testing.assert_allclose(gnorm, g[f"{name}_grad_norm"], rtol=1e-4, atol=1e-7)
    assert [p.grad is None for p in model.parameters()] == g[f"{name}_frozen"].tolist()      # frozen backbones stay frozen
    sd = model.state_dict()
    assert [int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")] == g[f"{name}_nbt"].tolist()
    np.testing.assert_allclose([v.double().sum().item() for k, v in sd.items() if k.endswith("running_mean")],
                               g[f"{name}_running_mean_sum"], rtol=1e-4, atol=1e-5)
    wsum1 = np.array([p.detach().double().sum().item() for p in model.parameters()])
    np.testing.assert_allclose(wsum1, g[f"{name}_wsum_after"], rtol=1e-5, atol=1e-4)
