"""Host-side train flow (no GPU): config loading, model-name dispatch and its error behaviour mirror the reference's
train scripts (audio_video/train.py:112-127, video/train.py:189-204, audio/train.py:118-134,
audio_cues_video/train.py:144-155, config/config.py:33-34)."""
import os

import pytest
import torch


def test_config_dotted_get_and_missing_file(tmp_path):
    from multimodal_lipread_b200.train import Config
    p = tmp_path / "av_config.yaml"
    p.write_text("dataset:\n  num_classes: 40\n  audio_input_size: 117\nmodel:\n  name: middle_fusion_fast\ntraining:\n  learning_rate: 0.0003\n")
    c = Config(str(p))
    assert c.get("dataset.num_classes") == 40 and c.get("model.name") == "middle_fusion_fast"
    assert c.get("model.audio_feature_dim", 128) == 128 and c.get("nope.deeper") is None
    with pytest.raises(FileNotFoundError):
        Config(str(tmp_path / "missing.yaml"))


def test_model_name_dispatch_and_errors():
    from multimodal_lipread_b200 import train as T
    from multimodal_lipread_b200.model_base import Cfg
    cfg = Cfg()
    m = T.create_av_model("middle_fusion_fast", 40, cfg)
    assert type(m).__name__ == "MidFusionFast" and len(m.state_dict()) == 256
    assert type(T.create_av_model("early_fusion_mobilenet", 40, cfg)).__name__ == "EarlyFusionAVMobileNet"
    assert type(T.create_av_model("early_fusion_resnet", 40, cfg)).__name__ == "EarlyFusionAV"
    assert type(T.create_video_model("resnet_lstm", 40, cfg)).__name__ == "ResNet2DBiLSTM"
    assert type(T.create_audio_model("resnet", 8)).__name__ == "AudioResNet"
    assert type(T.create_acv_model("late_fusion_mobile", 40)).__name__ == "MultimodalAttentionLate"
    with pytest.raises(ValueError, match="Unknown model name"):
        T.create_av_model("not_a_model", 40, cfg)
    with pytest.raises(ValueError, match="Invalid model name"):
        T.create_audio_model("not_a_model", 8)
    for name in T.AV_MODELS:                             # every audio_video model name of av_config.yaml:10 has a plan
        assert T.create_av_model(name, 40, cfg).num_classes == 40
    assert type(T.create_video_model("vgg_lstm", 40, cfg)).__name__ == "VGGLSTM"
    assert type(T.create_video_model("cnn", 40, cfg)).__name__ == "CNNOnly"
    assert type(T.create_audio_model("resnet_lstm", 8)).__name__ == "AudioResNetLSTM"
    assert type(T.create_audio_model("vgg", 8, version=11)).__name__ == "VGGAudioClassifier"
    assert type(T.create_audio_model("vgg_lstm", 8, version=11)).__name__ == "VGGWithLSTMClassifier"
    assert type(T.create_audio_model("lstm_resnet", 8, input_size=117)).__name__ == "LSTMResNet"
    assert type(T.create_audio_model("lstm_resnet_attn", 8, input_size=117)).__name__ == "DeepAudioNetWithAttention"
    for name in T.ACV_MODELS:                            # every audio_cues_video name (train.py:144-155) has a plan
        m = T.create_acv_model(name, 40)
        assert m.num_classes == 40 and m.INPUTS == ("audio", "cue", "video")
    frozen = T.create_acv_model("early_fusion_mobile", 40)
    assert not any(p.requires_grad for p in frozen.audio.parameters())
    assert not any(p.requires_grad for p in frozen.video.cnn.parameters()) and all(p.requires_grad for p in frozen.video.lstm.parameters())
    with pytest.raises(NotImplementedError, match="frozen"):
        frozen.configure_optimizer(lr=1e-4, weight_decay=1e-5)
    with pytest.raises(ValueError, match="Unknown model name"):
        T.create_acv_model("not_a_model", 40)
    assert type(T.create_video_model("resnet_attn", 40, cfg)).__name__ == "ResNet2DAttention"
    assert type(T.create_video_model("resnet_trans", 40, cfg)).__name__ == "ResNet2DTransformer"
    for name in T.AUDIO_MODELS:                          # every audio model name (audio/train.py:118-134) has a plan
        assert T.create_audio_model(name, 8, input_size=117, version=11).num_classes == 8
    for name in T.VIDEO_MODELS:                          # every video model name (video/train.py:189-204) has a plan
        assert T.create_video_model(name, 40, cfg).num_classes == 40
    with pytest.raises(ValueError, match="Unknown model"):
        T.create_video_model("not_a_model", 40, cfg)


def test_models_refuse_cpu_tensors():
    from multimodal_lipread_b200 import train as T
    from multimodal_lipread_b200.model_base import Cfg
    m = T.create_av_model("middle_fusion_fast", 8, Cfg())
    with pytest.raises(Exception):
        m(torch.zeros(1, 80, 117), torch.zeros(1, 3, 4, 44, 44))


def test_default_batch_adapters():
    from multimodal_lipread_b200.train import default_batch_to_inputs
    a, b, c = torch.zeros(2, 80, 117), torch.zeros(2, 3, 29, 44, 44), torch.zeros(2, dtype=torch.int64)
    assert default_batch_to_inputs((a, b, c))[0] == (a, b)
    ins, lab = default_batch_to_inputs({"lip_regions": b, "label": c})
    assert ins == (b,) and lab is c


def test_plateau_schedule_follows_torch():
    """ReduceLROnPlateau drives model.set_lr exactly as torch's scheduler drives the reference's optimizer
    (video/train.py:213-215 mode max / patience 5; audio_cues_video/train.py:163 mode min / patience 3)."""
    import torch
    from multimodal_lipread_b200 import train as T

    class Dummy:
        def __init__(self, lr):
            self._opt = {"lr": lr}

        def set_lr(self, lr):
            self._opt["lr"] = lr

    g = torch.Generator().manual_seed(0)
    for mode, patience in (("max", 5), ("min", 3), ("min", 0)):
        metrics = (torch.rand(60, generator=g) * 0.2 + torch.linspace(0.5, 0.6, 60)).tolist()
        metrics[20:40] = [metrics[19]] * 20                                   # a real plateau
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=1e-3)
        ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode=mode, factor=0.5, patience=patience)
        ours_model = Dummy(1e-3)
        ours = T.ReduceLROnPlateau(ours_model, mode=mode, factor=0.5, patience=patience)
        for m in metrics:
            ref.step(m)
            ours.step(m)
            assert ours_model._opt["lr"] == opt.param_groups[0]["lr"]
            assert ours.num_bad_epochs == ref.num_bad_epochs and ours.best == ref.best
        assert ours_model._opt["lr"] < 1e-3


def test_log_files_have_the_reference_schema(tmp_path):
    import csv
    from multimodal_lipread_b200 import train as T
    out = str(tmp_path / "metrics")
    T.init_log_files("m", out)
    T.log_to_files("m", 1, 2.5, 10.0, 2.25, 12.5, 2.0, 15.0, out)
    T.init_log_files("m", out)                                               # a second run appends, never truncates
    T.log_to_files("m", 2, 1.5, 30.0, 1.25, 32.5, 1.0, 35.0, out)
    T.log_final_results("m", 1.0, 35.0, out)
    rows = list(csv.reader(open(os.path.join(out, "m_training_log.csv"))))
    assert rows[0] == ["epoch", "train_loss", "train_acc", "val_loss", "val_acc", "test_loss", "test_acc"]
    assert rows[1] == ["1", "2.5", "10.0", "2.25", "12.5", "2.0", "15.0"] and len(rows) == 3
    txt = open(os.path.join(out, "m_training_log.txt")).read()
    assert txt.startswith("Training Log\n\nEpoch 1\n  Train Loss: 2.5000, Train Acc: 10.00%\n  Val Loss:   2.2500, Val Acc:   12.50%\n")
    assert txt.endswith("Final Test Loss: 1.0000, Final Test Acc: 35.00%\n")


def test_epoch_bookkeeping_follows_each_train_script():
    """Loss / accuracy aggregation of the epoch loops with a ragged last batch: mean of batch means
    (audio_video/train.py:57-75,78-90) vs size-weighted mean (audio_cues_video/train.py:52-81)."""
    import torch
    from multimodal_lipread_b200 import train as T

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.eye(4))

        def forward(self, x):
            return x @ self.w

        def train_step(self, x, labels):
            out = self(x)
            return torch.nn.functional.cross_entropy(out, labels).detach().reshape(1), out.detach()

        def eval_step(self, x, labels):                      # PlanModel.eval_step: (loss, correct, logits) tensors
            out = self(x)
            return (torch.nn.functional.cross_entropy(out, labels).detach().reshape(1),
                    (out.argmax(1) == labels).sum().reshape(1), out.detach())

    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(n, 4, generator=g), torch.randint(0, 4, (n,), generator=g)) for n in (5, 5, 2)]
    losses = [torch.nn.functional.cross_entropy(x, y).item() for x, y in batches]
    correct = sum((x.argmax(1) == y).sum().item() for x, y in batches)
    m = Stub()
    for fn in (T.train_epoch, T.validate):
        loss, acc = fn(m, batches, "cpu")
        assert loss == pytest.approx(sum(losses) / 3, rel=1e-6) and acc == pytest.approx(100.0 * correct / 12)
        loss, acc = fn(m, batches, "cpu", per_sample_loss=True)
        assert loss == pytest.approx((5 * losses[0] + 5 * losses[1] + 2 * losses[2]) / 12, rel=1e-6)
    dict_batches = [{"lip_regions": x, "label": y} for x, y in batches]                  # video/data_utils items
    assert T.validate(m, dict_batches, "cpu")[0] == pytest.approx(sum(losses) / 3, rel=1e-6)


def test_factories_take_the_imagenet_checkpoint_or_warn(tmp_path):
    """The reference initialises every torchvision trunk from ImageNet (weights=...IMAGENET1K_V1 / pretrained=True);
    offline the checkpoint is an argument or the YAML key model.pretrained_weights, and its absence is announced."""
    import warnings
    from torchvision.models import mobilenet_v2, mobilenet_v3_small, resnet18
    from multimodal_lipread_b200 import train as T
    from multimodal_lipread_b200.model_base import Cfg
    torch.manual_seed(11)
    v3, r18, v2 = mobilenet_v3_small(weights=None).state_dict(), resnet18(weights=None).state_dict(), mobilenet_v2(weights=None).state_dict()
    with pytest.warns(UserWarning, match="ImageNet"):
        T.create_av_model("middle_fusion_fast", 8, Cfg())
    with pytest.warns(UserWarning, match="ImageNet"):
        T.create_acv_model("early_fusion_mobile", 8)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        m = T.create_av_model("middle_fusion_fast", 8, Cfg(), pretrained_state_dict=v3)
        assert torch.equal(m.state_dict()["video_cnn.features.3.block.0.0.weight"], v3["features.3.block.0.0.weight"])
        path = str(tmp_path / "r18.pt")
        torch.save(r18, path)
        m = T.create_video_model("resnet_lstm", 8, Cfg({"model": {"pretrained_weights": path, "feature_dim": 256}}))
        # video/models/resnet_lstm.py:90-93 keeps children()[:-2] as a Sequential: index 6 is layer3
        assert torch.equal(m.state_dict()["cnn_features.6.1.conv2.weight"], r18["layer3.1.conv2.weight"])
        m = T.create_acv_model("early_fusion_mobile", 8, pretrained_state_dicts={"audio": r18, "video": v2})
        sd = m.state_dict()
        ka = [k for k in sd if k.startswith("audio") and k.endswith("layer2.0.conv1.weight")][0]
        kv = [k for k in sd if k.startswith("video") and k.endswith("3.conv.0.0.weight")][0]
        assert torch.equal(sd[ka], r18["layer2.0.conv1.weight"]) and torch.equal(sd[kv], v2["features.3.conv.0.0.weight"])
        # the 1-channel audio conv1 the reference re-creates after loading keeps its own init
        k1 = [k for k in sd if k.startswith("audio") and k.endswith("conv1.weight") and "layer" not in k][0]
        assert sd[k1].shape[1] == 1
    with pytest.raises(FileNotFoundError):
        T.create_video_model("resnet_lstm", 8, Cfg({"model": {"pretrained_weights": str(tmp_path / "missing.pt")}}))
