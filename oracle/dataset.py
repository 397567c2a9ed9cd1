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


def scan_audio(root_dir, split="train", audio_ext=".m4a"):
    """audio_cues_video/data_utils/audio_data.py:20-37 (= audio/data_utils/dataset.py): (classes, samples)."""
    class_dir = os.path.join(root_dir, "lipread_files")
    classes = sorted(e.name for e in os.scandir(class_dir) if e.is_dir())
    samples = []
    for idx, word in enumerate(classes):
        word_dir = os.path.join(class_dir, word, split)
        if os.path.exists(word_dir):
            for f in [f for f in os.listdir(word_dir) if f.endswith(audio_ext)]:
                samples.append({"audio_path": os.path.join(word_dir, f), "label": idx})
    return classes, samples


def scan_triple(root_dir, cue_root, lip_root, split="train", cue_mode="emotion", audio_ext=".m4a"):
    """audio_cues_video/data_utils/dataset.py:72-205: cues by split, lip-region index, strict alignment."""
    import json
    import pathlib
    import re
    sid_regex = re.compile(r"\d{4}-\d{4}")
    classes, audio = scan_audio(root_dir, split, audio_ext)
    cues = {}
    folder = os.path.join(cue_root, f"Descriptions_{cue_mode.capitalize()}")
    for file in os.listdir(folder):
        if split not in file.lower():
            continue
        for entry in json.load(open(os.path.join(folder, file))):
            cues[(entry["word"], entry["sequence_id"], split)] = entry["description"]
    index = {}
    for f in pathlib.Path(lip_root).rglob("*.npy"):
        m = sid_regex.search(f.name)
        parts = [p.lower() for p in f.parts]
        if not m or split not in parts:
            continue
        word = next((c for c in classes if c.lower() in parts), None)
        if word is None:
            continue
        key = (word, m.group(), split)
        if key in index:
            raise RuntimeError(f"Duplicate video entries for {key}")
        index[key] = str(f)
    aligned = []
    for s in audio:
        m = sid_regex.search(s["audio_path"])
        if not m:
            continue
        key = (classes[s["label"]], m.group(), split)
        if key in cues and key in index:
            aligned.append({"audio_path": s["audio_path"], "label": s["label"], "word": key[0], "sid": key[1],
                            "desc": cues[key], "lip_path": index[key]})
    if not aligned:
        raise RuntimeError("No aligned samples were built. Check folder structure and naming!")
    return classes, aligned


def getitem_triple(sample, desc2vec, decode, n_out=117):
    """audio_cues_video/data_utils/dataset.py:229-273 -> (mel, cue, lip, label)."""
    global _AP
    _AP = _AP or AudioProcessorPort()
    mel = _AP.clip_frontend(load_audio(decode(sample["audio_path"])), n_out)
    cue = torch.tensor(desc2vec[sample["desc"]], dtype=torch.float32)
    arr = np.load(sample["lip_path"]).astype(np.float32)
    if arr.max() > 1.0:
        arr = arr / 255.0
    return mel, cue, torch.tensor(arr).permute(3, 0, 1, 2).float(), torch.tensor(sample["label"], dtype=torch.long)


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
