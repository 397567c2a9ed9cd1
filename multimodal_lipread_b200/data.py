"""On-GPU input path from the reference's on-disk formats.

The reference decodes, scales and transposes every clip on the CPU inside DataLoader worker processes
(`np.load(...).astype(np.float32) / 255.0`, `.permute(3, 0, 1, 2)`, video/data_utils/dataset_loader.py:87-101;
log-mel per clip, audio_video/data_utils/dataset_av.py:55-66) and ships fp32 tensors to the device.  Here the host
only moves BYTES: worker threads read the `.npy` payloads (uint8, (T, H, W, 3)) and the 16-bit PCM straight into
pinned ring slots, one async copy per slot puts them in HBM, and the arithmetic happens there --
lr_pcm_ingest (mono mean, truncate / right zero-pad) -> lr_logmel_fwd for the audio, and the `/255` + layout
change folded into the first convolution's gather for the frames (the models take the uint8 (B, T, H, W, 3) batch
as it is).  4x fewer bytes over PCIe for the frames, no per-clip CPU FFT.

Same directory layout, sample order, class indexing and error conventions as the reference:
  VisualDataset             video/data_utils/dataset_loader.py:18-83
  GLipsMultimodalDataset    audio_video/data_utils/dataset_av.py:16-51
Decoding compressed containers (m4a: pydub / ffmpeg, audio/utils/audio_processor.py:25-29) stays on the host and is
the caller's `audio_decoder`; the built-in decoder reads 16-bit PCM `.wav` (stdlib `wave`) and int16 `.npy`."""
import collections
import os
import struct
import wave
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

TARGET_SAMPLES = 20000            # audio/utils/audio_processor.py:13  int(1.25 * 16000)
SAMPLE_RATE = 16000


# ---------------------------------------------------------------------------------------------- sample lists
class VisualDataset:
    """Sample list of video/data_utils/dataset_loader.py:18-83: classes = sorted sub-directories of
    root_dir/lipread_files; one sample per `<class>/<split>/*.mp4` whose
    `<lip_regions_dir>/lipread_files/<class>/<split>/<base>.npy` exists, in os.listdir order."""

    def __init__(self, root_dir, lip_regions_dir, split="train", transform=None):
        if transform is not None:
            raise NotImplementedError("host-side transforms are not part of the device input path")
        self.root_dir, self.lip_regions_dir, self.split = root_dir, lip_regions_dir, split
        self.class_dir = os.path.join(root_dir, "lipread_files")
        self.lip_regions_class_dir = os.path.join(lip_regions_dir, "lipread_files")
        self.classes = sorted(d for d in os.listdir(self.class_dir) if os.path.isdir(os.path.join(self.class_dir, d)))
        self.class_to_idx = {c: i for i, c in enumerate(self.classes)}
        self.samples = self._build_samples()

    def _build_samples(self):
        samples = []
        for name in self.classes:
            split_dir = os.path.join(self.class_dir, name, self.split)
            if not os.path.exists(split_dir):
                continue
            for f in os.listdir(split_dir):
                if not f.endswith(".mp4"):
                    continue
                p = os.path.join(self.lip_regions_class_dir, name, self.split, os.path.splitext(f)[0] + ".npy")
                if os.path.exists(p):
                    samples.append((p, self.class_to_idx[name]))
        return samples

    def __len__(self):
        return len(self.samples)

    def paths(self, idx):
        p, label = self.samples[idx]
        return None, p, label


class GLipsMultimodalDataset:
    """Sample list of audio_video/data_utils/dataset_av.py:16-51: the VisualDataset samples that also have
    `<root>/lipread_files/<class>/<split>/<base><audio_ext>`; lip regions under `<root>_lip_regions`."""

    def __init__(self, root_dir, input_size_audio, split="train", transform_audio=None, transform_video=None,
                 audio_ext=".m4a"):
        if transform_audio is not None or transform_video is not None:
            raise NotImplementedError("host-side transforms are not part of the device input path")
        self.root_dir, self.input_size_audio, self.split = root_dir, input_size_audio, split
        self.video_dataset = VisualDataset(root_dir, root_dir + "_lip_regions", split)
        self.samples = []
        for video_path, label in self.video_dataset.samples:
            base = os.path.splitext(os.path.basename(video_path))[0]
            audio_path = os.path.join(root_dir, "lipread_files", video_path.split(os.sep)[-3], split, base + audio_ext)
            if os.path.exists(audio_path):
                self.samples.append({"audio_path": audio_path, "video_path": video_path, "label": label})

    def __len__(self):
        return len(self.samples)

    def paths(self, idx):
        s = self.samples[idx]
        return s["audio_path"], s["video_path"], s["label"]


class GLipsDataset:
    """Sample list of the audio-only dataset (audio/data_utils/dataset.py:11-40, audio_cues_video/data_utils/
    audio_data.py:11-37): classes = sorted sub-directories, one sample per `<class>/<split>/*<audio_ext>` in
    os.listdir order.  The loader yields (mel, label)."""

    def __init__(self, root_dir, input_size, split="train", transform=None, audio_ext=".m4a"):
        if transform is not None:
            raise NotImplementedError("host-side transforms are not part of the device input path")
        self.root_dir, self.input_size, self.split = root_dir, input_size, split
        self.input_size_audio = input_size
        self.class_dir = os.path.join(root_dir, "lipread_files")
        self.classes = sorted(e.name for e in os.scandir(self.class_dir) if e.is_dir())
        self.class_to_idx = {c: i for i, c in enumerate(self.classes)}
        self.samples = []
        for word in self.classes:
            word_dir = os.path.join(self.class_dir, word, split)
            if os.path.exists(word_dir):
                for f in os.listdir(word_dir):
                    if f.endswith(audio_ext):
                        self.samples.append({"audio_path": os.path.join(word_dir, f), "label": self.class_to_idx[word]})

    def __len__(self):
        return len(self.samples)

    def paths(self, idx):
        s = self.samples[idx]
        return s["audio_path"], None, s["label"]


class MultimodalTripleDataset:
    """Sample list of audio_cues_video/data_utils/dataset.py:19-205: audio samples (the source of labels) strictly
    aligned with cue descriptions and lip-region files by (word, sequence id "dddd-dddd", split); RuntimeError for
    duplicate lip files and for an empty alignment, as the reference raises them.  Cue embeddings come from the
    reference's own cache file `<cache_dir>/<cue_mode>_<md5 of the descriptions>.npz` when it exists, otherwise from
    `embedder(list_of_descriptions) -> (n, dim) array` (SentenceTransformer.encode in the reference, :207-221; the
    text encoder itself is outside the path).  The loader yields (mel, cue, lips, label) like collate_fn_triple (:279-284).
    The reference divides a clip by 255 only `if arr.max() > 1.0` (:256-258): `unit_clip_rule` makes the loader run
    lr_u8_unit_clips on every batch so that a clip of 0 / 1 pixels reaches the model as 0.0 / 1.0 here too."""
    unit_clip_rule = True

    def __init__(self, root_dir, cue_root, lip_regions_root, input_size=117, split="train", cue_mode="emotion",
                 cache_dir=".cache_cues", embedder=None, audio_ext=".m4a"):
        import hashlib
        import json
        import pathlib
        import re
        sid_regex = re.compile(r"\d{4}-\d{4}")
        self.split = split.lower()
        self.cue_root, self.cue_mode, self.cache_dir = cue_root, cue_mode, cache_dir
        self.audio_ds = GLipsDataset(root_dir, input_size, split=self.split, audio_ext=audio_ext)
        self.input_size_audio = input_size
        self.classes, self.class_to_idx = self.audio_ds.classes, self.audio_ds.class_to_idx
        # cues (:72-97)
        folder = os.path.join(cue_root, f"Descriptions_{cue_mode.capitalize()}")
        self.cues = {}
        for file in os.listdir(folder):
            if self.split not in file.lower():
                continue
            with open(os.path.join(folder, file), "r") as f:
                for entry in json.load(f):
                    self.cues[(entry["word"], entry["sequence_id"], self.split)] = entry["description"]
        # lip-region index (:102-144)
        self.video_index = {}
        for npy_file in pathlib.Path(lip_regions_root).rglob("*.npy"):
            m = sid_regex.search(npy_file.name)
            if not m:
                continue
            parts = [p.lower() for p in npy_file.parts]
            if self.split not in parts:
                continue
            word = next((c for c in self.classes if c.lower() in parts), None)
            if word is None:
                continue
            key = (word, m.group(), self.split)
            if key in self.video_index:
                raise RuntimeError(f"Duplicate video entries for {key}:\n  Existing: {self.video_index[key]}\n  New:      {npy_file}")
            self.video_index[key] = str(npy_file)
        # strict alignment (:150-205)
        self.samples = []
        for s in self.audio_ds.samples:
            m = sid_regex.search(s["audio_path"])
            if not m:
                continue
            word = self.classes[s["label"]]
            key = (word, m.group(), self.split)
            if key not in self.cues or key not in self.video_index:
                continue
            self.samples.append({"audio_path": s["audio_path"], "label": s["label"], "word": word, "sid": m.group(),
                                 "desc": self.cues[key], "lip_path": self.video_index[key]})
        if len(self.samples) == 0:
            raise RuntimeError("No aligned samples were built. Check folder structure and naming!")
        # embeddings (:207-221)
        descs = sorted(set(s["desc"] for s in self.samples))
        sig = hashlib.md5("".join(descs).encode()).hexdigest()
        cache = os.path.join(cache_dir, f"{cue_mode}_{sig}.npz")
        if os.path.exists(cache):
            d = np.load(cache, allow_pickle=True)
            self.desc2vec = dict(zip(d["desc"], d["emb"]))
        elif embedder is not None:
            emb = np.asarray(embedder(descs))
            os.makedirs(cache_dir, exist_ok=True)
            np.savez(cache, desc=descs, emb=emb)
            self.desc2vec = dict(zip(descs, emb))
        else:
            raise FileNotFoundError(f"no cached cue embeddings at {cache} and no embedder given "
                                    "(the reference computes them with sentence-transformers)")
        self.cue_dim = int(np.asarray(next(iter(self.desc2vec.values()))).shape[0])

    def __len__(self):
        return len(self.samples)

    def paths(self, idx):
        s = self.samples[idx]
        return s["audio_path"], s["lip_path"], s["label"]

    def cue(self, idx):
        return np.asarray(self.desc2vec[self.samples[idx]["desc"]], dtype=np.float32)


# ---------------------------------------------------------------------------------------------- file readers
def npy_header(f):
    """Parse a .npy header from an open binary file -> (dtype, fortran_order, shape); leaves f at the payload."""
    magic = f.read(8)
    if len(magic) != 8 or magic[:6] != b"\x93NUMPY":
        raise ValueError("not a .npy file")
    major = magic[6]
    if major == 1:
        (hlen,) = struct.unpack("<H", f.read(2))
    elif major in (2, 3):
        (hlen,) = struct.unpack("<I", f.read(4))
    else:
        raise ValueError(f"unsupported .npy version {major}")
    import ast
    d = ast.literal_eval(f.read(hlen).decode("latin1" if major < 3 else "utf8"))
    return np.dtype(d["descr"]), bool(d["fortran_order"]), tuple(d["shape"])


def read_npy_u8_into(path, dst):
    """Read the payload of a uint8 C-order .npy directly into `dst` (a writable uint8 ndarray of the same shape,
    typically a slice of a pinned ring slot).  video/data_utils/dataset_loader.py:90 minus the float conversion."""
    with open(path, "rb") as f:
        dtype, fortran, shape = npy_header(f)
        if dtype != np.uint8 or fortran:
            raise ValueError(f"{path}: expected C-order uint8 lip regions, got {dtype} fortran_order={fortran}")
        if shape != tuple(dst.shape):
            raise ValueError(f"{path}: lip regions {shape} differ from the batch shape {tuple(dst.shape)}")
        n = f.readinto(memoryview(dst).cast("B"))
        if n != dst.size:
            raise ValueError(f"{path}: truncated payload ({n} of {dst.size} bytes)")
    return shape


def decode_pcm16(path):
    """Built-in decoder -> (interleaved int16 ndarray, channels, scale).  `.wav`: 16-bit PCM at 16 kHz via the stdlib
    (the torchaudio.load branch of audio/utils/audio_processor.py:30-35 normalises int16 by 1/32768 and, because of
    the overwrite at :35, never resamples -- so another rate is refused rather than silently mis-timed);
    `.npy`: int16 samples as pydub's get_array_of_samples() returns them, (n,) or (n, channels), scale 1 (:26-29)."""
    if path.endswith(".npy"):
        a = np.load(path)
        if a.dtype != np.int16 or a.ndim not in (1, 2):
            raise ValueError(f"{path}: expected int16 PCM of shape (n,) or (n, channels), got {a.dtype} {a.shape}")
        ch = 1 if a.ndim == 1 else a.shape[1]
        return np.ascontiguousarray(a).reshape(-1), ch, 1.0
    if path.endswith(".wav"):
        with wave.open(path, "rb") as w:
            if w.getsampwidth() != 2:
                raise ValueError(f"{path}: only 16-bit PCM wav is decoded here")
            if w.getframerate() != SAMPLE_RATE:
                raise ValueError(f"{path}: {w.getframerate()} Hz; resample to {SAMPLE_RATE} Hz in the audio_decoder")
            ch = w.getnchannels()
            a = np.frombuffer(w.readframes(min(w.getnframes(), TARGET_SAMPLES)), dtype="<i2")
        return a, ch, 1.0 / 32768.0
    raise ValueError(f"{path}: no built-in decoder for this container; pass audio_decoder= (m4a decoding stays on the "
                     "host, audio/utils/audio_processor.py:25-29)")


class HostSlot:
    """One ring slot of host staging memory: frames uint8 [B,T,H,W,3], packed PCM int16 [B, cap] and per-clip
    metadata.  Pinned when CUDA is there (async H2D), plain memory otherwise (host-logic tests)."""

    def __init__(self, batch, frame_shape, with_audio, max_channels, pin, cue_dim=0):
        def buf(shape, dtype):
            t = torch.empty(shape, dtype=dtype)
            return t.pin_memory() if pin else t
        self.frames = buf((batch,) + tuple(frame_shape), torch.uint8) if frame_shape is not None else None
        self.cues = buf((batch, cue_dim), torch.float32) if cue_dim else None
        self.labels = buf((batch,), torch.int64)
        self.cap = (TARGET_SAMPLES * max_channels + 7) // 8 * 8          # 16-byte aligned clip starts
        self.with_audio = with_audio
        if with_audio:
            self.pcm = buf((batch, self.cap), torch.int16)
            self.meta = buf((3, batch), torch.int64)                      # offset, n_frames, channels
        self.scale = 1.0
        self.n = 0

    def stage_batch(self, items, audio_decoder, n_threads):
        """Fill the slot with `items` = [(audio_path, video_path, label)]: the .npy payloads are read by the library's
        native thread pool (lr_host_read_npy_u8 / lr_host_read_npy_pcm16, no interpreter in the loop); audio in
        another container goes through `audio_decoder` clip by clip.  Returns the PCM scale of the batch."""
        import ctypes
        from ._lib import lib, check
        n = self.n = len(items)

        def c_paths(paths):
            return (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
        if self.frames is not None:
            shape = (ctypes.c_longlong * (self.frames.dim() - 1))(*self.frames.shape[1:])
            check(lib.lr_host_read_npy_u8(c_paths([it[1] for it in items]), n, self.frames.data_ptr(), shape,
                                          self.frames.dim() - 1, n_threads))
        if self.cues is not None:
            self.cues[:n] = torch.from_numpy(np.stack([it[3] for it in items]))
        self.labels[:n] = torch.tensor([int(it[2]) for it in items], dtype=torch.int64)
        if not self.with_audio:
            return None
        if audio_decoder is decode_pcm16 and all(it[0].endswith(".npy") for it in items):
            meta = torch.zeros(3, n, dtype=torch.int64)
            check(lib.lr_host_read_npy_pcm16(c_paths([it[0] for it in items]), n, self.pcm.data_ptr(), self.cap,
                                             TARGET_SAMPLES, meta.data_ptr(), n_threads))
            self.meta[:, :n] = meta
            return 1.0
        scales = {self._stage_audio(j, it[0], audio_decoder) for j, it in enumerate(items)}
        if len(scales) != 1:
            raise ValueError("clips of one batch must share the PCM scale (one decoder branch)")
        return scales.pop()

    def _stage_audio(self, j, audio_path, audio_decoder):
        pcm, ch, scale = audio_decoder(audio_path)
        if ch < 1 or ch * TARGET_SAMPLES > self.cap:
            raise ValueError(f"{audio_path}: {ch} channels exceed the slot capacity")
        n_frames = min(pcm.size // ch, TARGET_SAMPLES)                    # later samples are truncated anyway (:40-41)
        self.pcm[j].numpy()[:n_frames * ch] = pcm[:n_frames * ch]
        self.meta[0, j], self.meta[1, j], self.meta[2, j] = j * self.cap, n_frames, ch
        return float(scale)

    def stage_clip(self, j, audio_path, video_path, label, audio_decoder):
        """Pure-Python staging of one clip (the readable restatement of stage_batch; used by the host tests)."""
        read_npy_u8_into(video_path, self.frames[j].numpy())
        self.labels[j] = int(label)
        if not self.with_audio:
            return None
        return self._stage_audio(j, audio_path, audio_decoder)


def batch_indices(n, batch_size, shuffle, drop_last, generator, rank=0, world=1):
    """Per-rank batches of sample indices.  world == 1: DataLoader order (randperm when shuffled, ragged last batch
    unless drop_last).  world > 1: the same global order on every rank (same seed), cut into GLOBAL
    batches of batch_size * world clips of which rank r takes clips [r * batch_size, (r + 1) * batch_size); a ragged
    global tail is dropped so that every rank takes the same number of equal steps (the in-graph allreduce needs it)."""
    order = torch.randperm(n, generator=generator).tolist() if shuffle else list(range(n))
    if world > 1:
        G = batch_size * world
        return [order[i + rank * batch_size:i + (rank + 1) * batch_size] for i in range(0, n - G + 1, G)]
    out = [order[i:i + batch_size] for i in range(0, n, batch_size)]
    if drop_last and out and len(out[-1]) < batch_size:
        out.pop()
    return out


# ---------------------------------------------------------------------------------------------- the loader
class DeviceBatchLoader:
    """DataLoader replacement for the reference's datasets whose batches arrive in HBM ready for the models:
      GLipsMultimodalDataset -> (mel (B,80,n_out) f32, lips (B,T,H,W,3) uint8, labels (B,) i64)   [dataset_av.py:77]
      VisualDataset          -> {"lip_regions": lips uint8, "label": labels}                      [dataset_loader.py:98-101]
      GLipsDataset           -> (mel, labels)                                                      [audio/data_utils/dataset.py:52]
      MultimodalTripleDataset -> (mel, cue (B,dim) f32, lips uint8, labels)     [audio_cues_video/data_utils/dataset.py:273-284]
    `depth` batches are in flight: worker threads fill pinned slots while the copy stream uploads the previous one
    and runs the audio kernels.  Contract: enqueue the work that reads a batch (e.g. model.train_step, which copies its
    inputs first) BEFORE asking for the next batch; its ring buffers are reused `depth` batches later, ordered after
    that work on the device."""

    def __init__(self, dataset, batch_size, shuffle=False, drop_last=False, device="cuda", depth=3, workers=8, seed=0,
                 audio_decoder=None, max_channels=2, rank=0, world=1):
        if len(dataset) == 0:
            raise RuntimeError("empty dataset")
        self.ds, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.rank, self.world = int(rank), int(world)        # data parallel: batch_size is per rank, same seed everywhere
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NotImplementedError("DeviceBatchLoader feeds the GPU path; there is no CPU frontend here")
        self.depth = max(2, int(depth))
        a0, v0, _ = dataset.paths(0)
        self.with_audio, self.with_video = a0 is not None, v0 is not None
        self.cue_dim = int(getattr(dataset, "cue_dim", 0))
        self.n_out = int(dataset.input_size_audio) if self.with_audio else 0
        self.decoder = audio_decoder or decode_pcm16
        self.gen = torch.Generator().manual_seed(seed)
        self.frame_shape = None
        if self.with_video:
            with open(v0, "rb") as f:
                dtype, fortran, self.frame_shape = npy_header(f)
            if len(self.frame_shape) != 4 or self.frame_shape[3] != 3:
                raise ValueError(f"lip regions must be (T, H, W, 3), got {self.frame_shape}")
        self.workers = int(workers)
        self.pool = ThreadPoolExecutor(max_workers=self.depth)        # one staging job per ring slot in flight
        self.slots = [HostSlot(self.batch_size, self.frame_shape, self.with_audio, max_channels, pin=True, cue_dim=self.cue_dim)
                      for _ in range(self.depth)]
        B = self.batch_size
        self.dev = [dict(frames=(torch.empty((B,) + self.frame_shape, dtype=torch.uint8, device=self.device)
                                 if self.with_video else None),
                         cues=torch.empty(B, self.cue_dim, dtype=torch.float32, device=self.device) if self.cue_dim else None,
                         labels=torch.empty(B, dtype=torch.int64, device=self.device),
                         pcm=torch.empty(B, self.slots[0].cap, dtype=torch.int16, device=self.device) if self.with_audio else None,
                         meta=torch.empty(3, B, dtype=torch.int64, device=self.device) if self.with_audio else None,
                         ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(self.depth)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        if self.with_audio:
            from .audio_processor import AudioProcessor
            self.ap = AudioProcessor(device=self.device)

    def __len__(self):
        n = len(self.ds)
        if self.world > 1:
            return n // (self.batch_size * self.world)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _stage(self, slot, idxs):
        items = [self.ds.paths(i) + ((self.ds.cue(i),) if self.cue_dim else ()) for i in idxs]
        return self.pool.submit(slot.stage_batch, items, self.decoder, self.workers)

    def _upload(self, k, futures):
        """Slot k: wait for its files, then (copy stream) H2D + pcm_ingest + log-mel; returns the device batch."""
        from . import ops
        slot, dev = self.slots[k], self.dev[k]
        scale = futures.result()
        n = slot.n
        cur = torch.cuda.current_stream(self.device)
        # Whatever the consumer enqueued for the PREVIOUS batch is on its stream by now: mark that batch's slot as
        # released there.  This slot's own release mark dates from depth - 1 requests ago, so the upload below overlaps
        # the steps still running on the batches in between instead of queueing behind them.
        self.dev[(k - 1) % self.depth]["free"].record(cur)
        self.copy_stream.wait_event(dev["free"])
        with torch.cuda.stream(self.copy_stream):
            if self.with_video:
                dev["frames"][:n].copy_(slot.frames[:n], non_blocking=True)
                if getattr(self.ds, "unit_clip_rule", False):
                    from ._lib import lib, check
                    clip_bytes = dev["frames"][0].numel()
                    if clip_bytes % 16:
                        raise ValueError("unit_clip_rule needs clips of a multiple of 16 bytes")
                    check(lib.lr_u8_unit_clips(dev["frames"].data_ptr(), n, clip_bytes, self.copy_stream.cuda_stream))
            if self.cue_dim:
                dev["cues"][:n].copy_(slot.cues[:n], non_blocking=True)
            dev["labels"][:n].copy_(slot.labels[:n], non_blocking=True)
            mel = None
            if self.with_audio:
                dev["pcm"][:n].copy_(slot.pcm[:n], non_blocking=True)
                dev["meta"].copy_(slot.meta, non_blocking=True)
                meta = dev["meta"]
                wav = ops.pcm_ingest(dev["pcm"].view(-1), meta[0, :n].contiguous(), meta[1, :n].to(torch.int32),
                                     meta[2, :n].to(torch.int32), scale, TARGET_SAMPLES)
                mel = ops.logmel(wav, self.ap.plan, self.n_out, 0)
            dev["ready"].record(self.copy_stream)
        cur.wait_event(dev["ready"])
        frames = dev["frames"][:n] if self.with_video else None
        labels = dev["labels"][:n]
        if mel is not None:
            mel.record_stream(cur)                   # allocated on the copy stream, consumed on the caller's
        if self.cue_dim:
            return mel, dev["cues"][:n], frames, labels
        if self.with_audio and self.with_video:
            return mel, frames, labels
        if self.with_audio:
            return mel, labels
        return {"lip_regions": frames, "label": labels}

    def __iter__(self):
        batches = batch_indices(len(self.ds), self.batch_size, self.shuffle, self.drop_last, self.gen, self.rank, self.world)
        inflight = collections.deque()
        nxt = 0
        while nxt < len(batches) and len(inflight) < self.depth - 1:
            inflight.append((nxt % self.depth, self._stage(self.slots[nxt % self.depth], batches[nxt])))
            nxt += 1
        while inflight:
            k, futures = inflight.popleft()
            out = self._upload(k, futures)
            if nxt < len(batches):
                kk = nxt % self.depth
                self.dev[kk]["ready"].synchronize()             # its last upload has left the pinned slot
                inflight.append((kk, self._stage(self.slots[kk], batches[nxt])))
                nxt += 1
            yield out

    def close(self):
        self.pool.shutdown(wait=True)


def get_data_loaders(config_path, device="cuda", **loader_kw):
    """video/data_utils/dataset_loader.py:129-186: (train, val, test) loaders over the lip regions next to
    `dataset.root_dir` (`<root>_lip_regions`), batch size `training.batch_size` (default 4), only the train split
    shuffled; FileNotFoundError when the preprocessed lip regions are missing (:144-148)."""
    from .train import Config
    config = Config(config_path)
    dataset_path = config.get("dataset.root_dir")
    lip_regions_dir = os.path.join(os.path.dirname(dataset_path), os.path.basename(dataset_path) + "_lip_regions")
    if not os.path.exists(lip_regions_dir):
        raise FileNotFoundError(f"Preprocessed lip regions not found at {lip_regions_dir}. "
                                f"Run visual_preprocessing.py first.")
    batch_size = config.get("training.batch_size", 4)
    return tuple(DeviceBatchLoader(VisualDataset(dataset_path, lip_regions_dir, split=s), batch_size, shuffle=(s == "train"),
                                   device=device, **loader_kw) for s in ("train", "val", "test"))
