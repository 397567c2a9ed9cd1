"""Generate the committed golden vectors by running the REAL reference modules.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden.py
Writes tests/golden/dataset_golden.npz, tests/golden/logmel_golden.npz, tests/golden/midfusion_golden.npz and tests/golden/models_golden.npz.

Recipe: stub the unused top-level imports `librosa` / `pydub`, make the
ImageNet weight download a no-op (seeded random init instead), and load one reference
package at a time because they import siblings by bare names (`from models.x import ...`).
Nothing from the reference is copied: its modules are imported, called and their outputs saved.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from multimodal_lipread_b200 import synthetic  # noqa: E402


def _stub_unused_imports():
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    pd = types.ModuleType("pydub")
    pd.AudioSegment = object
    sys.modules.setdefault("pydub", pd)


def _offline_weights():
    from torchvision.models import _api
    _api.WeightsEnum.get_state_dict = lambda self, *a, **k: None
    orig = torch.nn.Module.load_state_dict

    def load_state_dict(self, sd, *a, **k):
        if sd is None:
            return None
        return orig(self, sd, *a, **k)
    torch.nn.Module.load_state_dict = load_state_dict


def load_ref(pkg, module):
    for name in list(sys.modules):
        if name.split(".")[0] in ("models", "config", "configs", "utils", "data_utils"):
            del sys.modules[name]
    sys.path[:] = [p for p in sys.path if not p.startswith(REF)]
    sys.path.insert(0, os.path.join(REF, pkg))
    return importlib.import_module(module)


class Cfg:
    def get(self, key, default=None):
        return default


def golden_logmel():
    ap_mod = load_ref("audio_video", "utils.audio_processor")
    ap = ap_mod.AudioProcessor()
    waves = torch.cat([
        synthetic.make_waveforms(3, kind="pcm", pad_fraction=0.0),
        synthetic.make_waveforms(2, seed=77, kind="pcm", pad_fraction=1.0),
        synthetic.make_waveforms(1, kind="unit", pad_fraction=0.0),
        synthetic.make_waveforms(1, kind="tone", pad_fraction=0.0),
        torch.zeros(1, synthetic.N_SAMPLES),                      # silent clip: std == 0 edge case
    ])
    outs, raws = [], []
    for w in waves:
        # exactly audio_video/data_utils/dataset_av.py:58-66
        mel = ap.compute_melspectrogram(w)
        raws.append(mel.clone())
        mel = ap.normalize_spectrogram(mel)
        mel = mel[:80, :117].float()
        outs.append(mel)
    np.savez_compressed(
        os.path.join(HERE, "logmel_golden.npz"),
        wave=waves.numpy(), logmel_raw=torch.stack(raws).numpy(), out=torch.stack(outs).numpy(),
        window=ap.mel_transform.spectrogram.window.numpy(), fb=ap.mel_transform.mel_scale.fb.numpy())
    print("logmel golden:", torch.stack(outs).shape)


def golden_midfusion():
    mod = load_ref("audio_video", "models.middle_fusion_fast")
    out = {}
    for size in (44, 88):
        torch.manual_seed(0)
        model = mod.create_mid_fusion_fast(40, Cfg())
        model.train()
        B = 2
        wav = synthetic.make_waveforms(B, pad_fraction=0.5)
        apm = load_ref("audio_video", "utils.audio_processor").AudioProcessor()
        mel = torch.stack([apm.normalize_spectrogram(apm.compute_melspectrogram(w))[:80, :117].float() for w in wav])
        lips = synthetic.make_lips_u8(B, size=size)
        video = (lips.float() / 255.0).permute(0, 4, 1, 2, 3).contiguous()
        labels = synthetic.make_labels(B, 40)
        opt = torch.optim.Adam(model.parameters(), lr=3e-4)
        opt.zero_grad()
        logits = model(mel, video)
        loss = torch.nn.CrossEntropyLoss()(logits, labels)
        loss.backward()
        names = [n for n, _ in model.named_parameters()]
        gnorm = np.array([p.grad.double().norm().item() for _, p in model.named_parameters()])
        gsum = np.array([p.grad.double().sum().item() for _, p in model.named_parameters()])
        wsum0 = np.array([p.detach().double().sum().item() for _, p in model.named_parameters()])
        opt.step()
        wsum1 = np.array([p.detach().double().sum().item() for _, p in model.named_parameters()])
        sd = model.state_dict()
        out[f"logits_{size}"] = logits.detach().numpy()
        out[f"loss_{size}"] = np.array(loss.item())
        out[f"grad_norm_{size}"] = gnorm
        out[f"grad_sum_{size}"] = gsum
        out[f"wsum_before_{size}"] = wsum0
        out[f"wsum_after_{size}"] = wsum1
        out[f"rm_stem_{size}"] = sd["video_cnn.features.0.1.running_mean"].numpy()
        out[f"rv_stem_{size}"] = sd["video_cnn.features.0.1.running_var"].numpy()
        out[f"rm_last_{size}"] = sd["video_cnn.features.12.1.running_mean"].numpy()
        out[f"rv_last_{size}"] = sd["video_cnn.features.12.1.running_var"].numpy()
        out["param_names"] = np.array(names)
        out["state_keys"] = np.array(list(sd.keys()))
        model.eval()
        with torch.no_grad():
            out[f"logits_eval_{size}"] = model(mel, video).numpy()
        print(size, "loss", loss.item(), "n_params", sum(p.numel() for p in model.parameters()))
    np.savez_compressed(os.path.join(HERE, "midfusion_golden.npz"), **out)


class DCfg:
    def __init__(self, d):
        self.d = d

    def get(self, key, default=None):
        return self.d.get(key, default)


def golden_models(only_resnet_variants=False):
    """Configs 1, 2, 4: the reference's own AudioResNet / ResNet2DBiLSTM / EarlyFusionAVMobileNet with dropout set to 0
    (the only change: p is a constructor argument / module attribute), one train step each."""
    out = {}
    apm = load_ref("audio_video", "utils.audio_processor").AudioProcessor()

    def data(B, size, T, C):
        wav = synthetic.make_waveforms(B, pad_fraction=0.5)
        mel = torch.stack([apm.normalize_spectrogram(apm.compute_melspectrogram(w))[:80, :117].float() for w in wav])
        lips = synthetic.make_lips_u8(B, size=size)[:, :T].contiguous()
        video = (lips.float() / 255.0).permute(0, 4, 1, 2, 3).contiguous()
        return mel, video, synthetic.make_labels(B, C)

    def record(name, model, inputs, labels, lr, wd, B, T, size):
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
        opt.zero_grad()
        logits = model(*inputs)
        loss = torch.nn.CrossEntropyLoss()(logits, labels)
        loss.backward()
        out[f"{name}_logits"] = logits.detach().numpy()
        out[f"{name}_loss"] = np.array(loss.item())
        out[f"{name}_grad_norm"] = np.array([0.0 if p.grad is None else p.grad.double().norm().item()
                                             for _, p in model.named_parameters()])
        out[f"{name}_frozen"] = np.array([p.grad is None for _, p in model.named_parameters()])
        out[f"{name}_wsum_before"] = np.array([p.detach().double().sum().item() for _, p in model.named_parameters()])
        opt.step()
        out[f"{name}_wsum_after"] = np.array([p.detach().double().sum().item() for _, p in model.named_parameters()])
        out[f"{name}_param_names"] = np.array([n for n, _ in model.named_parameters()])
        out[f"{name}_state_keys"] = np.array(list(model.state_dict().keys()))
        sd = model.state_dict()
        out[f"{name}_nbt"] = np.array([int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")])
        out[f"{name}_running_mean_sum"] = np.array([v.double().sum().item() for k, v in sd.items() if k.endswith("running_mean")])
        out[f"{name}_B"], out[f"{name}_T"], out[f"{name}_size"] = np.array(B), np.array(T), np.array(size)
        print(name, "loss", loss.item(), "n_params", sum(p.numel() for p in model.parameters()))

    # video resnet_lstm with model.resnet_version 34 / 50 (video/models/resnet_lstm.py:79-86): BasicBlock [3,4,6,3] and
    # Bottleneck trunks.  `--resnet-variants` adds just these records to the committed file.
    # (resnet34 at 88 px / 18 frames: with 8 frames of 44 px its 36 BatchNorms see <= 32 values per channel in layer4 and
    # the fp32 gradients of the REFERENCE itself move by ~1 % under round-off -- not a parity case)
    mod = load_ref("video", "models.resnet_lstm")
    for version, (B, T, size, C) in ((34, (3, 6, 88, 40)), (50, (2, 4, 44, 40))):
        mel, video, labels = data(B, size, T, C)
        torch.manual_seed(0)
        model = mod.ResNet2DBiLSTM(C, DCfg({"model.dropout": 0.0, "model.resnet_version": version}))
        record(f"video_resnet{version}_lstm", model, (video,), labels, 5e-5, 1e-5, B, T, size)
    if only_resnet_variants:
        old = dict(np.load(os.path.join(HERE, "models_golden.npz")))
        old.update(out)
        np.savez_compressed(os.path.join(HERE, "models_golden.npz"), **old)
        return

    # config 4: audio_video early_fusion_mobilenet (lr 3e-4, av_config.yaml:23)
    B, T, size, C = 3, 8, 44, 40
    mod = load_ref("audio_video", "models.early_fusion")
    torch.manual_seed(0)
    model = mod.create_early_fusion_mobilenet_model(C, Cfg())
    model.video_encoder.lstm.dropout = 0.0
    model.classifier[2].p = 0.0
    mel, video, labels = data(B, size, T, C)
    record("early_fusion_mobilenet", model, (mel, video), labels, 3e-4, 0.0, B, T, size)

    # config 2: video resnet_lstm (lr 5e-5, weight_decay 1e-5: visual_config.yaml:25, video/train.py:210)
    B, T, size, C = 2, 5, 44, 40
    mod = load_ref("video", "models.resnet_lstm")
    torch.manual_seed(0)
    model = mod.ResNet2DBiLSTM(C, DCfg({"model.dropout": 0.0}))
    mel, video, labels = data(B, size, T, C)
    record("video_resnet_lstm", model, (video,), labels, 5e-5, 1e-5, B, T, size)

    # config 1: audio-only resnet (lr 5e-4, weight_decay 1e-4: audio_config.yaml:20-21)
    B, C = 4, 8
    mod = load_ref("audio", "models.resnet_model")
    torch.manual_seed(0)
    model = mod.AudioResNet(num_classes=C, dropout_rate=0.0)
    mel, video, labels = data(B, 44, 1, C)
    record("audio_resnet", model, (mel,), labels, 5e-4, 1e-4, B, 1, 44)
    # config 5: audio_cues_video late_fusion_mobile (lr 1e-5, acv_config.yaml:14), pretrained=False as its own
    # constructor allows (late_fusion_mobile.py:86)
    B, T, size, C = 3, 6, 44, 40
    mod = load_ref("audio_cues_video", "models.late_fusion_mobile")
    torch.manual_seed(0)
    model = mod.MultimodalAttentionLate(C, cue_dim=768, video_cfg=None, pretrained=False)
    model.video.lstm.dropout = 0.0
    mel, video, labels = data(B, size, T, C)
    cue = synthetic.make_cues(B)
    record("acv_late_fusion_mobile", model, (mel, cue, video), labels, 1e-5, 0.0, B, T, size)
    # video mobilenet_lstm (dropout 0 through its own config key) and audio_cues_video late_fusion_resnet
    B, T, size, C = 3, 6, 44, 40
    mod = load_ref("video", "models.mobilenet_lstm")
    torch.manual_seed(0)
    model = mod.MobileNetLSTM(C, DCfg({"model.dropout": 0.0}))
    mel, video, labels = data(B, size, T, C)
    record("video_mobilenet_lstm", model, (video,), labels, 5e-5, 1e-5, B, T, size)
    mod = load_ref("audio_cues_video", "models.late_fusion_resnet")
    torch.manual_seed(0)
    model = mod.MultimodalAttentionLateResNet(C, cue_dim=768, video_cfg=None, pretrained=False)
    model.video.lstm.dropout = 0.0
    record("acv_late_fusion_resnet", model, (mel, synthetic.make_cues(B), video), labels, 1e-5, 0.0, B, T, size)
    mod = load_ref("video", "models.resnet_attn")
    torch.manual_seed(0)
    model = mod.ResNet2DAttention(C, DCfg({"model.dropout": 0.0}))
    record("video_resnet_attn", model, (video,), labels, 5e-5, 1e-5, B, T, size)
    mod = load_ref("video", "models.shufflenet_lstm")
    torch.manual_seed(0)
    model = mod.ShuffleNet2DBiLSTM(C, DCfg({"model.dropout": 0.0}))
    record("video_shufflenet_lstm", model, (video,), labels, 5e-5, 1e-5, B, T, size)
    mod = load_ref("video", "models.resnet_trans")
    torch.manual_seed(0)
    model = mod.ResNet2DTransformer(C, DCfg({"model.dropout": 0.0}))
    record("video_resnet_trans", model, (video,), labels, 5e-5, 1e-5, B, T, size)
    # audio_cues_video early / middle attention fusion; the early and middle-resnet variants freeze their backbones and
    # feed the CNN 4 time steps at a time (T = 6: one chunk of 4 and one of 2)
    cue = synthetic.make_cues(B)
    for key, module, cls_name in (("middle_fusion_mobile", "models.middle_fusion_mobile", "MultimodalAttentionMiddle"),
                                  ("middle_fusion_resnet", "models.middle_fusion_resnet", "MultimodalAttentionMiddleResNet"),
                                  ("early_fusion_mobile", "models.early_fusion_mobile", "MultimodalAttentionEarly"),
                                  ("early_fusion_resnet", "models.early_fusion_resnet", "MultimodalAttentionEarlyResNet")):
        mod = load_ref("audio_cues_video", module)
        torch.manual_seed(0)
        model = getattr(mod, cls_name)(C, cue_dim=768, video_cfg=None, pretrained=False)
        for m_ in model.modules():
            if isinstance(m_, torch.nn.Dropout):
                m_.p = 0.0
        model.video.lstm.dropout = 0.0
        record("acv_" + key, model, (mel, cue, video), labels, 1e-4, 0.0, B, T, size)
    # video vgg_lstm, audio resnet_lstm (dropout is a constructor argument) and audio vgg (version 11)
    B, T, size, C = 3, 6, 44, 40
    mod = load_ref("video", "models.vgg_lstm")
    torch.manual_seed(0)
    model = mod.VGGLSTM(C, DCfg({"model.dropout": 0.0}))
    mel, video, labels = data(B, size, T, C)
    record("video_vgg_lstm", model, (video,), labels, 5e-5, 1e-5, B, T, size)
    mod = load_ref("video", "models.cnn")
    torch.manual_seed(0)
    model = mod.CNNOnly(C, DCfg({"model.dropout": 0.0}))
    record("video_cnn", model, (video,), labels, 5e-5, 1e-5, B, T, size)
    B, C = 4, 8
    mel, video, labels = data(B, 44, 1, C)
    mod = load_ref("audio", "models.resnet_lstm_model")
    torch.manual_seed(0)
    model = mod.AudioResNetLSTM(num_classes=C, dropout_rate=0.0)
    record("audio_resnet_lstm", model, (mel,), labels, 5e-4, 1e-4, B, 1, 44)
    mod = load_ref("audio", "models.vgg_model")
    torch.manual_seed(0)
    model = mod.VGGAudioClassifier(num_classes=C, version=11, dropout_rate=0.0)
    record("audio_vgg", model, (mel,), labels, 5e-4, 1e-4, B, 1, 44)
    mod = load_ref("audio", "models.vgg_lstm_model")
    torch.manual_seed(0)
    model = mod.VGGWithLSTMClassifier(num_classes=C, version=11, dropout_rate=0.0)
    record("audio_vgg_lstm", model, (mel,), labels, 5e-4, 1e-4, B, 1, 44)
    mod = load_ref("audio", "models.lstm_resnet_model")
    torch.manual_seed(0)
    model = mod.LSTMResNet(num_classes=C, input_size=117, dropout_rate=0.0)
    record("audio_lstm_resnet", model, (mel,), labels, 5e-4, 1e-4, B, 1, 44)
    mod = load_ref("audio", "models.lstm_resnet_attn_model")
    torch.manual_seed(0)
    model = mod.DeepAudioNetWithAttention(num_classes=C, input_size=117, dropout_rate=0.0)
    record("audio_lstm_resnet_attn", model, (mel,), labels, 5e-4, 1e-4, B, 1, 44)
    mod = load_ref("audio", "models.lstm_resnet_trans_model")
    torch.manual_seed(0)
    model = mod.LSTMResNetWithTransformer(num_classes=C, input_size=117, dropout_rate=0.0)
    for m_ in model.modules():                               # the encoder layers' own dropout (torch default 0.1)
        if isinstance(m_, torch.nn.Dropout):
            m_.p = 0.0
        if isinstance(m_, torch.nn.MultiheadAttention):
            m_.dropout = 0.0
    record("audio_lstm_resnet_trans", model, (mel,), labels, 5e-4, 1e-4, B, 1, 44)
    # the remaining audio_video models (av_config.yaml:10), lr 3e-4
    B, T, size, C = 3, 8, 44, 40
    for name, module, factory, drop in (("late_fusion_mobilenet", "models.late_fusion", "create_late_fusion_mobilenet_model", False),
                                        ("middle_fusion_mobilenet", "models.middle_fusion", "create_mid_fusion_mobilenet_model", True),
                                        ("early_fusion_fast", "models.early_fusion_fast", "create_early_fusion_fast", False),
                                        ("late_fusion_fast", "models.late_fusion_fast", "create_late_fusion_fast", False)):
        mod = load_ref("audio_video", module)
        torch.manual_seed(0)
        model = getattr(mod, factory)(C, Cfg())
        if drop:
            model.classifier[2].p = 0.0
        mel, video, labels = data(B, size, T, C)
        record(name, model, (mel, video), labels, 3e-4, 0.0, B, T, size)
    np.savez_compressed(os.path.join(HERE, "models_golden.npz"), **out)


def golden_dataset():
    """The reference's own VisualDataset / GLipsMultimodalDataset over synthetic.write_dataset_tree: sample lists and
    items.  The m4a decode (pydub + ffmpeg, absent here) is the one stubbed call: AudioSegment.from_file returns the
    int16 samples the tree's audio files hold, so AudioProcessor.load_audio's own float conversion and pad / truncate
    (audio_processor.py:29,37-44) run unmodified."""
    import tempfile

    class FakeSegment:
        def __init__(self, a):
            self.a = a

        @classmethod
        def from_file(cls, path, format=None):
            return cls(np.load(path))

        def set_frame_rate(self, r):
            return self

        def set_channels(self, c):
            return self

        def get_array_of_samples(self):
            return self.a

    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        root = os.path.join(tmp, "GLips_4")
        synthetic.write_dataset_tree(root)
        sys.modules["pydub"].AudioSegment = FakeSegment
        mod = load_ref("audio_video", "data_utils.dataset_av")
        sys.modules["utils.audio_processor"].AudioSegment = FakeSegment
        for split in ("train", "val"):
            ds = mod.GLipsMultimodalDataset(root, 117, split=split)
            keys = []
            for i in range(len(ds)):
                mel, lips, label = ds[i]
                rel = os.path.relpath(ds.samples[i]["video_path"], root + "_lip_regions")
                key = os.path.splitext(rel)[0].split(os.sep, 1)[1].replace(os.sep, "/")
                keys.append(key)
                out[f"mel|{key}"] = mel.numpy()
                out[f"lips|{key}"] = lips.numpy()
                out[f"label|{key}"] = np.array(int(label))
            out[f"keys_av|{split}"] = np.array(sorted(keys))
            vds = mod.VisualDataset(root, root + "_lip_regions", split=split)
            vkeys = []
            for i in range(len(vds)):
                item = vds[i]
                rel = os.path.relpath(vds.samples[i][0], root + "_lip_regions")
                key = os.path.splitext(rel)[0].split(os.sep, 1)[1].replace(os.sep, "/")
                vkeys.append(key)
                out[f"vlabel|{key}"] = np.array(int(item["label"]))
                out[f"vsum|{key}"] = np.array(item["lip_regions"].double().sum().item())
            out[f"keys_video|{split}"] = np.array(sorted(vkeys))
            out[f"classes|{split}"] = np.array(vds.classes)
        # audio + cue + video: the reference's MultimodalTripleDataset with SentenceTransformer stubbed by the seeded
        # fake embedder (the text encoder is outside the path) and the same fake m4a decode
        st = types.ModuleType("sentence_transformers")

        class FakeST:
            def __init__(self, *a, **k):
                pass

            def encode(self, descs, show_progress_bar=False):
                return synthetic.fake_sentence_embedding(list(descs))
        st.SentenceTransformer = FakeST
        sys.modules["sentence_transformers"] = st
        glips, cue_root, lip_root = synthetic.write_triple_tree(os.path.join(tmp, "triple"))
        load_ref("audio_cues_video", "utils.audio_processor")
        sys.path.insert(0, os.path.join(REF, "audio_cues_video", "data_utils"))     # dataset.py does `from audio_data import`
        tmod = importlib.import_module("data_utils.dataset")
        sys.modules["utils.audio_processor"].AudioSegment = FakeSegment
        for split in ("train", "val"):
            ds = tmod.MultimodalTripleDataset(glips, cue_root, lip_root, 117, split=split,
                                              cache_dir=os.path.join(tmp, "cache"))
            keys = []
            for i in range(len(ds)):
                mel, cue, lip, label = ds[i]
                key = f"{ds.samples[i]['word']}/{split}/{ds.samples[i]['sid']}"
                keys.append(key)
                out[f"tmel|{key}"], out[f"tcue|{key}"] = mel.numpy(), cue.numpy()
                out[f"tlipsum|{key}"] = np.array(lip.double().sum().item())
                out[f"tlabel|{key}"] = np.array(int(label))
            out[f"keys_triple|{split}"] = np.array(sorted(keys))
            batch = tmod.collate_fn_triple([ds[i] for i in range(min(2, len(ds)))])
            out[f"collate_shapes|{split}"] = np.array([list(t.shape) + [0] * (5 - t.dim()) for t in batch])
    np.savez_compressed(os.path.join(HERE, "dataset_golden.npz"), **out)
    print("dataset golden:", {k: v.tolist() for k, v in out.items() if k.startswith("keys_av")})


if __name__ == "__main__":
    _stub_unused_imports()
    _offline_weights()
    torch.set_num_threads(8)
    if "--dataset-only" in sys.argv:
        golden_dataset()
        sys.exit(0)
    if "--resnet-variants" in sys.argv:
        golden_models(only_resnet_variants=True)
        sys.exit(0)
    if "--models-only" not in sys.argv:
        golden_dataset()
        golden_logmel()
        golden_midfusion()
    golden_models()
