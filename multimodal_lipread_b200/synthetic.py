"""Deterministic synthetic GLips-shaped inputs: one seed per tensor family.

Shapes follow the reference's data contracts:
  waveform  (N, 20000) f32, int16-valued PCM floats   audio/utils/audio_processor.py:29,40-44
  lips      (N, 29, H, W, 3) uint8                    video/data_utils/visual_preprocessing.py:211
  labels    (N,) int64                                audio_video/data_utils/dataset_av.py:75
All generators run on the CPU with torch.Generator so they give identical bytes everywhere.
"""
import math
import torch

SEED_WAVE, SEED_LIPS, SEED_CUE, SEED_LABEL = 1234, 2345, 3456, 4567
N_SAMPLES = 20000
T_FRAMES = 29


def make_waveforms(n, seed=SEED_WAVE, kind="pcm", pad_fraction=0.25):
    """kind: 'pcm' round(3000*randn); 'unit' 0.1*randn; 'tone' 3000*sin(2*pi*440 t)+50*randn.
    A `pad_fraction` of the clips get a right zero-pad of 0..8000 samples (short recordings)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, N_SAMPLES, generator=g)
    if kind == "pcm":
        x = torch.round(3000.0 * x)
    elif kind == "unit":
        x = 0.1 * x
    elif kind == "tone":
        t = torch.arange(N_SAMPLES, dtype=torch.float64) / 16000.0
        x = (3000.0 * torch.sin(2 * math.pi * 440.0 * t)).float()[None, :] + 50.0 * x
    else:
        raise ValueError(f"unknown waveform kind {kind!r}")
    pick = torch.rand(n, generator=g) < pad_fraction
    lens = torch.randint(0, 8001, (n,), generator=g)
    for i in range(n):
        if pick[i] and lens[i] > 0:
            x[i, N_SAMPLES - int(lens[i]):] = 0.0
    return x.contiguous()


def make_lips_u8(n, size=44, seed=SEED_LIPS, grayscale=False):
    """(n, 29, size, size, 3) uint8; grayscale=True replicates one plane to the 3 channels."""
    g = torch.Generator().manual_seed(seed)
    if grayscale:
        plane = torch.randint(0, 256, (n, T_FRAMES, size, size, 1), generator=g, dtype=torch.uint8)
        return plane.expand(-1, -1, -1, -1, 3).contiguous()
    return torch.randint(0, 256, (n, T_FRAMES, size, size, 3), generator=g, dtype=torch.uint8)


def make_labels(n, num_classes, seed=SEED_LABEL):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, num_classes, (n,), generator=g, dtype=torch.int64)


def make_cues(n, dim=768, seed=SEED_CUE):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(n, dim, generator=g)
    return c / c.norm(dim=1, keepdim=True)


def write_dataset_tree(root, classes=("aufgaben", "besser", "danke"), per_split=None, T=5, size=12, seed=97,
                       audio_ext=".m4a", missing_every=5):
    """A tiny GLips-shaped tree in the reference's on-disk layout (README "Data"; dataset_loader.py:38-83,
    dataset_av.py:28-51):  <root>/lipread_files/<class>/<split>/<base>.mp4 (empty marker) + <base><audio_ext>,
    <root>_lip_regions/lipread_files/<class>/<split>/<base>.npy  uint8 (T, size, size, 3).
    The audio files hold int16 mono PCM in .npy format (what pydub's get_array_of_samples() would return after the
    m4a decode that stays outside the path) with ragged lengths on both sides of 20 000 samples.  Every
    `missing_every`-th clip lacks its lip regions and the one after it lacks its audio, as incomplete preprocessing
    leaves them.  Returns {relative sample key: (pcm int16, lips uint8, class index)} for the complete clips."""
    import os
    import numpy as np
    per_split = per_split or {"train": 4, "val": 2}
    g = torch.Generator().manual_seed(seed)
    truth, k = {}, 0
    for ci, cname in enumerate(sorted(classes)):
        for split, n in per_split.items():
            vdir = os.path.join(root, "lipread_files", cname, split)
            ldir = os.path.join(root + "_lip_regions", "lipread_files", cname, split)
            os.makedirs(vdir, exist_ok=True)
            os.makedirs(ldir, exist_ok=True)
            for i in range(n):
                base = f"{cname}_{i:03d}"
                k += 1
                n_samp = int(torch.randint(9000, 26000, (1,), generator=g))
                pcm = torch.round(3000.0 * torch.randn(n_samp, generator=g)).clamp(-32768, 32767).to(torch.int16).numpy()
                lips = torch.randint(0, 256, (T, size, size, 3), generator=g, dtype=torch.uint8).numpy()
                open(os.path.join(vdir, base + ".mp4"), "wb").close()
                has_lips = k % missing_every != 0
                has_audio = k % missing_every != 1 or k == 1
                if has_lips:
                    np.save(os.path.join(ldir, base + ".npy"), lips)
                if has_audio:
                    with open(os.path.join(vdir, base + audio_ext), "wb") as f:
                        np.save(f, pcm)
                if has_lips and has_audio:
                    truth[f"{cname}/{split}/{base}"] = (pcm, lips, ci)
    return truth


def write_triple_tree(root, classes=("aufgaben", "besser", "danke"), per_split=None, T=5, size=12, seed=131, audio_ext=".m4a"):
    """A tiny tree for the audio + cue + video dataset (audio_cues_video/data_utils/dataset.py:19-205):
      <root>/GLips/lipread_files/<word>/<split>/<word>_<sid>.m4a            int16 PCM (.npy payload, see write_dataset_tree)
      <root>/lips/<word>/<split>/<word>_<sid>.npy                           uint8 (T, size, size, 3)
      <root>/cues/Descriptions_Emotion/cues_<split>.json                    [{"word", "sequence_id", "description"}]
    with sid = "dddd-dddd".  Some clips lack a cue, some a lip file, one audio file has no sequence id in its name.
    Returns (glips_root, cue_root, lip_root)."""
    import json
    import os
    import numpy as np
    per_split = per_split or {"train": 5, "val": 2}
    g = torch.Generator().manual_seed(seed)
    glips, cue_root, lip_root = os.path.join(root, "GLips"), os.path.join(root, "cues"), os.path.join(root, "lips")
    os.makedirs(os.path.join(cue_root, "Descriptions_Emotion"), exist_ok=True)
    moods = ("calm", "tense", "happy", "tired", "angry")
    k = 0
    for split, n in per_split.items():
        entries = []
        for cname in sorted(classes):
            adir = os.path.join(glips, "lipread_files", cname, split)
            ldir = os.path.join(lip_root, cname, split)
            os.makedirs(adir, exist_ok=True)
            os.makedirs(ldir, exist_ok=True)
            for i in range(n):
                k += 1
                sid = f"{1000 + k:04d}-{2000 + 7 * k:04d}"
                base = f"{cname}_{sid}" if k % 7 else f"{cname}_nosid{k}"
                pcm = torch.round(3000.0 * torch.randn(int(torch.randint(9000, 26000, (1,), generator=g)), generator=g))
                with open(os.path.join(adir, base + audio_ext), "wb") as f:
                    np.save(f, pcm.clamp(-32768, 32767).to(torch.int16).numpy())
                if k % 5 != 0:
                    np.save(os.path.join(ldir, base + ".npy"),
                            torch.randint(0, 256, (T, size, size, 3), generator=g, dtype=torch.uint8).numpy())
                if k % 4 != 0:
                    entries.append({"word": cname, "sequence_id": sid,
                                    "description": f"The speaker sounds {moods[k % len(moods)]} saying {cname}."})
        with open(os.path.join(cue_root, "Descriptions_Emotion", f"cues_{split}.json"), "w") as f:
            json.dump(entries, f)
    return glips, cue_root, lip_root


def fake_sentence_embedding(descs, dim=768):
    """Deterministic stand-in for SentenceTransformer.encode (the embedder itself is out of scope): unit vectors
    seeded by the text."""
    import hashlib
    import numpy as np
    out = np.empty((len(descs), dim), dtype=np.float32)
    for i, d in enumerate(descs):
        rng = np.random.default_rng(int(hashlib.md5(d.encode()).hexdigest()[:8], 16))
        v = rng.standard_normal(dim).astype(np.float32)
        out[i] = v / np.linalg.norm(v)
    return out
