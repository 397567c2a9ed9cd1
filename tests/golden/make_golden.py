"""Generate the committed golden vectors by running the REAL reference modules.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden.py
Writes tests/golden/logmel_golden.npz and tests/golden/midfusion_golden.npz.

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


if __name__ == "__main__":
    _stub_unused_imports()
    _offline_weights()
    torch.set_num_threads(8)
    golden_logmel()
    golden_midfusion()
