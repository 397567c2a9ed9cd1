"""CPU restatement of the reference's dataset classes for the input path.  TEST INFRASTRUCTURE ONLY.

  scan_visual          video/data_utils/dataset_loader.py:36-83     VisualDataset.__init__ / _build_samples
  scan_multimodal      audio_video/data_utils/dataset_av.py:28-51   GLipsMultimodalDataset.__init__
  load_audio           audio/utils/audio_processor.py:29,37-46      decoded samples -> (20000,) float
  getitem_multimodal   audio_video/data_utils/dataset_av.py:54-77   (mel, lips, label)
  getitem_visual       video/data_utils/dataset_loader.py:87-101    {"lip_regions", "label"}
Container decoding (pydub m4a, :25-28) is a parameter: `decode(path) -> int16 samples, (n,) mono or (channels, n)`
with `scale` 1 for the pydub branch and 1/32768 for the torchaudio.load branch.
Pinned by tests/golden/dataset_golden.npz, produced by the reference's own classes (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from .frontend import AudioProcessorPort


def scan_visual(root_dir, lip_regions_dir, split="train"):
    class_dir = os.path.join(root_dir, "lipread_files")
    lip_dir = os.path.join(lip_regions_dir, "lipread_files")
    classes = sorted(d for d in os.listdir(class_dir) if os.path.isdir(os.path.join(class_dir, d)))
    samples = []
    for idx, name in enumerate(classes):
        split_dir = os.path.join(class_dir, name, split)
        if not os.path.exists(split_dir):
            continue
        for f in [f for f in os.listdir(split_dir) if f.endswith(".mp4")]:
            p = os.path.join(lip_dir, name, split, os.path.splitext(f)[0] + ".npy")
            if os.path.exists(p):
                samples.append((p, idx))
    return classes, samples


def scan_multimodal(root_dir, split="train", audio_ext=".m4a"):
    _, vis = scan_visual(root_dir, root_dir + "_lip_regions", split)
    out = []
    for video_path, label in vis:
        base = os.path.splitext(os.path.basename(video_path))[0]
        audio_path = os.path.join(root_dir, "lipread_files", video_path.split(os.sep)[-3], split, base + audio_ext)
        if os.path.exists(audio_path):
            out.append({"audio_path": audio_path, "video_path": video_path, "label": label})
    return out


def load_audio(samples_i16, scale=1.0, target=20000):
    samples = torch.as_tensor(np.asarray(samples_i16)).float() * scale
    audio = samples.mean(dim=0) if samples.dim() > 1 else samples
    if audio.size(0) > target:
        audio = audio[:target]
    elif audio.size(0) < target:
        audio = torch.nn.functional.pad(audio, (0, target - audio.size(0)))
    return audio


_AP = None


def getitem_multimodal(sample, decode, n_out=117, scale=1.0):
    global _AP
    _AP = _AP or AudioProcessorPort()
    mel = _AP.clip_frontend(load_audio(decode(sample["audio_path"]), scale), n_out)
    lips = torch.tensor(np.load(sample["video_path"]).astype(np.float32) / 255.0).permute(3, 0, 1, 2)
    return mel, lips, torch.tensor(sample["label"], dtype=torch.long)


def getitem_visual(sample):
    path, label = sample
    lips = torch.tensor(np.load(path).astype(np.float32) / 255.0).permute(3, 0, 1, 2)
    return {"lip_regions": lips, "label": torch.tensor(label, dtype=torch.long)}
