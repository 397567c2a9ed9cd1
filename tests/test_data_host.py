"""Host side of the on-GPU input path (multimodal_lipread_b200/data.py) and its oracle (oracle/dataset.py) against the
golden items the reference's own GLipsMultimodalDataset / VisualDataset produced over the same synthetic tree
(tests/golden/make_golden.py golden_dataset).  No GPU: sample lists, file readers, ring-slot staging."""
import os

import numpy as np
import pytest
import torch

from multimodal_lipread_b200 import data, synthetic
from oracle import dataset as ods


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("glips") / "GLips_4")
    truth = synthetic.write_dataset_tree(root)
    return root, truth


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "dataset_golden.npz"))


def _key(video_path, root):
    rel = os.path.relpath(video_path, root + "_lip_regions")
    return os.path.splitext(rel)[0].split(os.sep, 1)[1].replace(os.sep, "/")


def _decode(path):
    return np.load(path)


@pytest.mark.parametrize("split", ["train", "val"])
def test_oracle_items_match_the_reference(tree, golden, split):
    root, _ = tree
    samples = ods.scan_multimodal(root, split)
    assert sorted(_key(s["video_path"], root) for s in samples) == golden[f"keys_av|{split}"].tolist()
    for s in samples:
        k = _key(s["video_path"], root)
        mel, lips, label = ods.getitem_multimodal(s, _decode)
        assert int(label) == int(golden[f"label|{k}"])
        assert np.array_equal(lips.numpy(), golden[f"lips|{k}"])
        assert mel.shape == (80, 117) and np.abs(mel.numpy() - golden[f"mel|{k}"]).max() < 1e-5
    classes, vis = ods.scan_visual(root, root + "_lip_regions", split)
    assert classes == golden[f"classes|{split}"].tolist()
    assert sorted(_key(p, root) for p, _ in vis) == golden[f"keys_video|{split}"].tolist()
    for smp in vis:
        item = ods.getitem_visual(smp)
        k = _key(smp[0], root)
        assert int(item["label"]) == int(golden[f"vlabel|{k}"])
        assert item["lip_regions"].double().sum().item() == float(golden[f"vsum|{k}"])


@pytest.mark.parametrize("split", ["train", "val", "test"])
def test_sample_lists_equal_the_oracle_scan(tree, split):
    root, _ = tree
    vds = data.VisualDataset(root, root + "_lip_regions", split)
    classes, vis = ods.scan_visual(root, root + "_lip_regions", split)
    assert vds.classes == classes and vds.samples == vis and len(vds) == len(vis)
    ds = data.GLipsMultimodalDataset(root, 117, split)
    assert ds.samples == ods.scan_multimodal(root, split)
    if split == "test":
        assert len(ds) == 0


def test_ring_slot_staging_is_byte_exact(tree):
    root, truth = tree
    ds = data.GLipsMultimodalDataset(root, 117, "train")
    slot = data.HostSlot(len(ds), (5, 12, 12, 3), True, 2, pin=False)
    for j in range(len(ds)):
        scale = slot.stage_clip(j, *ds.paths(j), lambda p: (np.load(p), 1, 1.0))
        assert scale == 1.0
        pcm, lips, label = truth[_key(ds.samples[j]["video_path"], root)]
        assert np.array_equal(slot.frames[j].numpy(), lips) and int(slot.labels[j]) == label
        off, n, ch = (int(slot.meta[r, j]) for r in range(3))
        assert off == j * slot.cap and off % 8 == 0 and ch == 1 and n == min(len(pcm), 20000)
        assert np.array_equal(slot.pcm.view(-1).numpy()[off:off + n], pcm[:n])


def test_native_batch_staging_equals_the_python_restatement(tree, tmp_path):
    root, truth = tree
    ds = data.GLipsMultimodalDataset(root, 117, "train")
    items = [ds.paths(j) for j in range(len(ds))]
    # the tree's audio files are .npy payloads under an .m4a name: give the native reader .npy names
    for a, _, _ in items:
        os.link(a, a[:-4] + ".npy") if not os.path.exists(a[:-4] + ".npy") else None
    items = [(a[:-4] + ".npy", v, l) for a, v, l in items]
    py = data.HostSlot(len(ds), (5, 12, 12, 3), True, 2, pin=False)
    for j, it in enumerate(items):
        py.stage_clip(j, *it, data.decode_pcm16)
    for threads in (1, 4):
        nat = data.HostSlot(len(ds), (5, 12, 12, 3), True, 2, pin=False)
        nat.pcm.zero_(), py.pcm.mul_(1)
        assert nat.stage_batch(items, data.decode_pcm16, threads) == 1.0 and nat.n == len(ds)
        assert torch.equal(nat.frames, py.frames) and torch.equal(nat.labels, py.labels)
        assert torch.equal(nat.meta, py.meta)
        for j in range(len(ds)):
            n = int(py.meta[1, j])
            assert torch.equal(nat.pcm[j, :n], py.pcm[j, :n])
    # a custom decoder keeps the per-clip path for the audio and the native path for the frames
    nat = data.HostSlot(len(ds), (5, 12, 12, 3), True, 2, pin=False)
    assert nat.stage_batch([ds.paths(j) for j in range(len(ds))], lambda p: (np.load(p), 1, 1.0), 2) == 1.0
    assert torch.equal(nat.frames, py.frames) and torch.equal(nat.meta, py.meta)
    # errors name the file
    from multimodal_lipread_b200._lib import LipreadError
    bad = list(items)
    bad[2] = (bad[2][0], str(tmp_path / "missing.npy"), 0)
    with pytest.raises(LipreadError, match="missing.npy"):
        nat.stage_batch(bad, data.decode_pcm16, 3)
    np.save(tmp_path / "float.npy", np.zeros((5, 12, 12, 3), np.float32))
    bad[2] = (bad[2][0], str(tmp_path / "float.npy"), 0)
    with pytest.raises(LipreadError, match="uint8"):
        nat.stage_batch(bad, data.decode_pcm16, 3)
    np.save(tmp_path / "shape.npy", np.zeros((5, 12, 13, 3), np.uint8))
    bad[2] = (bad[2][0], str(tmp_path / "shape.npy"), 0)
    with pytest.raises(LipreadError, match="batch shape"):
        nat.stage_batch(bad, data.decode_pcm16, 3)
    raw = open(items[0][1], "rb").read()
    (tmp_path / "cut.npy").write_bytes(raw[:-7])
    bad[2] = (bad[2][0], str(tmp_path / "cut.npy"), 0)
    with pytest.raises(LipreadError, match="truncated"):
        nat.stage_batch(bad, data.decode_pcm16, 3)
    np.save(tmp_path / "stereo.npy", np.arange(60, dtype=np.int16).reshape(30, 2))
    ok = [(str(tmp_path / "stereo.npy"), items[0][1], 1)]
    assert nat.stage_batch(ok, data.decode_pcm16, 1) == 1.0
    assert nat.meta[:, 0].tolist() == [0, 30, 2] and nat.pcm[0, :60].tolist() == list(range(60))


def test_npy_reader_refuses_what_the_kernels_cannot_take(tmp_path):
    good = np.arange(2 * 3 * 4 * 3, dtype=np.uint8).reshape(2, 3, 4, 3)
    np.save(tmp_path / "a.npy", good)
    dst = np.zeros_like(good)
    assert data.read_npy_u8_into(str(tmp_path / "a.npy"), dst) == good.shape and np.array_equal(dst, good)
    with pytest.raises(ValueError, match="differ from the batch shape"):
        data.read_npy_u8_into(str(tmp_path / "a.npy"), np.zeros((2, 3, 4, 4), np.uint8))
    np.save(tmp_path / "f.npy", good.astype(np.float32))
    with pytest.raises(ValueError, match="uint8"):
        data.read_npy_u8_into(str(tmp_path / "f.npy"), dst)
    np.save(tmp_path / "t.npy", good)
    raw = (tmp_path / "t.npy").read_bytes()
    (tmp_path / "t.npy").write_bytes(raw[:-5])
    with pytest.raises(ValueError, match="truncated"):
        data.read_npy_u8_into(str(tmp_path / "t.npy"), dst)
    (tmp_path / "x.npy").write_bytes(b"not numpy at all")
    with pytest.raises(ValueError, match="not a .npy"):
        data.read_npy_u8_into(str(tmp_path / "x.npy"), dst)


def test_builtin_decoder(tmp_path):
    import wave
    pcm = (np.arange(3000) % 200 - 100).astype(np.int16)
    st = np.stack([pcm, -pcm], axis=1)
    with wave.open(str(tmp_path / "s.wav"), "wb") as w:
        w.setnchannels(2), w.setsampwidth(2), w.setframerate(16000)
        w.writeframes(st.tobytes())
    a, ch, scale = data.decode_pcm16(str(tmp_path / "s.wav"))
    assert ch == 2 and scale == 1.0 / 32768.0 and np.array_equal(a.reshape(-1, 2), st)
    with wave.open(str(tmp_path / "r.wav"), "wb") as w:
        w.setnchannels(1), w.setsampwidth(2), w.setframerate(8000)
        w.writeframes(pcm.tobytes())
    with pytest.raises(ValueError, match="resample"):
        data.decode_pcm16(str(tmp_path / "r.wav"))
    with pytest.raises(ValueError, match="audio_decoder"):
        data.decode_pcm16(str(tmp_path / "clip.m4a"))
    np.save(tmp_path / "p.npy", st)
    a, ch, scale = data.decode_pcm16(str(tmp_path / "p.npy"))
    assert ch == 2 and scale == 1.0 and np.array_equal(a, st.reshape(-1))


def test_batch_order_and_loader_needs_the_gpu(tree):
    g = torch.Generator().manual_seed(3)
    b = data.batch_indices(10, 4, False, False, g)
    assert b == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    assert data.batch_indices(10, 4, False, True, g) == [[0, 1, 2, 3], [4, 5, 6, 7]]
    s = data.batch_indices(10, 4, True, False, torch.Generator().manual_seed(3))
    assert sorted(sum(s, [])) == list(range(10)) and s == data.batch_indices(10, 4, True, False, torch.Generator().manual_seed(3))
    root, _ = tree
    ds = data.GLipsMultimodalDataset(root, 117, "train")
    with pytest.raises(NotImplementedError):
        data.DeviceBatchLoader(ds, 4, device="cpu")
    with pytest.raises(RuntimeError, match="empty"):
        data.DeviceBatchLoader(data.GLipsMultimodalDataset(root, 117, "test"), 4)


@pytest.fixture(scope="module")
def triple_tree(tmp_path_factory):
    return synthetic.write_triple_tree(str(tmp_path_factory.mktemp("triple")))


@pytest.mark.parametrize("split", ["train", "val"])
def test_triple_dataset_alignment_and_items_match_the_reference(triple_tree, golden, tmp_path, split):
    """audio + cue + video: sample alignment, cue vectors and items equal what the reference's
    MultimodalTripleDataset produced (golden), for the oracle port and for data.MultimodalTripleDataset."""
    glips, cue_root, lip_root = triple_tree
    classes, aligned = ods.scan_triple(glips, cue_root, lip_root, split)
    want = golden[f"keys_triple|{split}"].tolist()
    assert sorted(f"{s['word']}/{split}/{s['sid']}" for s in aligned) == want
    ds = data.MultimodalTripleDataset(glips, cue_root, lip_root, 117, split, cache_dir=str(tmp_path / "cache"),
                                      embedder=synthetic.fake_sentence_embedding)
    assert ds.samples == aligned and ds.classes == classes and ds.cue_dim == 768
    for i, s in enumerate(aligned):
        k = f"{s['word']}/{split}/{s['sid']}"
        mel, cue, lip, label = ods.getitem_triple(s, ds.desc2vec, _decode)
        assert np.abs(mel.numpy() - golden[f"tmel|{k}"]).max() < 1e-5
        assert np.array_equal(cue.numpy(), golden[f"tcue|{k}"]) and np.array_equal(ds.cue(i), golden[f"tcue|{k}"])
        assert lip.double().sum().item() == float(golden[f"tlipsum|{k}"]) and int(label) == int(golden[f"tlabel|{k}"])
        assert ds.paths(i) == (s["audio_path"], s["lip_path"], s["label"])
    # the second construction finds the cache the first one wrote (the reference's file name and layout)
    again = data.MultimodalTripleDataset(glips, cue_root, lip_root, 117, split, cache_dir=str(tmp_path / "cache"))
    assert all(np.array_equal(again.desc2vec[d], ds.desc2vec[d]) for d in ds.desc2vec)
    with pytest.raises(FileNotFoundError, match="embedder"):
        data.MultimodalTripleDataset(glips, cue_root, lip_root, 117, split, cache_dir=str(tmp_path / "empty"))


def test_triple_dataset_error_conventions(triple_tree, tmp_path):
    import shutil
    glips, cue_root, lip_root = triple_tree
    with pytest.raises(RuntimeError, match="No aligned samples"):
        data.MultimodalTripleDataset(glips, cue_root, lip_root, 117, "test", embedder=synthetic.fake_sentence_embedding,
                                     cache_dir=str(tmp_path / "c"))
    dup = tmp_path / "lips"
    shutil.copytree(lip_root, dup)
    first = sorted((dup / "aufgaben" / "train").glob("*.npy"))[0]
    other = dup / "extra" / "aufgaben" / "train"
    other.mkdir(parents=True)
    shutil.copy(first, other / first.name)
    with pytest.raises(RuntimeError, match="Duplicate video entries"):
        data.MultimodalTripleDataset(glips, cue_root, str(dup), 117, "train", embedder=synthetic.fake_sentence_embedding,
                                     cache_dir=str(tmp_path / "c"))
    classes, samples = ods.scan_audio(glips, "train")
    ads = data.GLipsDataset(glips, 117, "train")
    assert ads.classes == classes and ads.samples == samples and ads.paths(0)[1] is None


def test_get_data_loaders_error_convention(tmp_path):
    """video/data_utils/dataset_loader.py:144-148: missing preprocessed lip regions -> FileNotFoundError."""
    (tmp_path / "GLips_4" / "lipread_files").mkdir(parents=True)
    cfg = tmp_path / "visual_config.yaml"
    cfg.write_text(f"dataset:\n  root_dir: {tmp_path / 'GLips_4'}\ntraining:\n  batch_size: 8\n")
    with pytest.raises(FileNotFoundError, match="Preprocessed lip regions not found"):
        data.get_data_loaders(str(cfg))
    with pytest.raises(FileNotFoundError, match="Config file not found"):
        data.get_data_loaders(str(tmp_path / "nope.yaml"))


@pytest.mark.parametrize("n,B,world,shuffle", [(37, 4, 2, True), (64, 8, 4, False), (10, 4, 4, True), (33, 2, 8, True)])
def test_rank_sharding_of_the_batch_order(n, B, world, shuffle):
    """rank r takes clips [r*B, (r+1)*B) of every global batch; equal step counts on all ranks; no clip
    twice; the global order is the single-process order of the same seed."""
    per_rank = [data.batch_indices(n, B, shuffle, False, torch.Generator().manual_seed(5), r, world) for r in range(world)]
    single = data.batch_indices(n, B * world, shuffle, True, torch.Generator().manual_seed(5))
    steps = n // (B * world)
    assert all(len(b) == steps for b in per_rank) and len(single) == steps
    for step in range(steps):
        glob = sum((per_rank[r][step] for r in range(world)), [])
        assert glob == single[step] and all(len(per_rank[r][step]) == B for r in range(world))
    flat = [i for r in per_rank for b in r for i in b]
    assert len(flat) == len(set(flat)) == steps * B * world


def test_native_npy_reader_property(tmp_path):
    """lr_host_read_npy_u8 / lr_host_read_npy_pcm16 against numpy for random shapes, .npy format versions 1.0 / 2.0 /
    3.0 and every thread count (hypothesis)."""
    import ctypes
    from hypothesis import given, settings, strategies as st, HealthCheck
    from numpy.lib import format as npf
    from multimodal_lipread_b200._lib import lib, check

    def write(path, arr, version):
        with open(path, "wb") as f:
            npf.write_array(f, arr, version=version)

    counter = [0]

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(shape=st.lists(st.integers(1, 7), min_size=1, max_size=5), n=st.integers(1, 6), threads=st.integers(1, 9),
           version=st.sampled_from([(1, 0), (2, 0), (3, 0)]), seed=st.integers(0, 2 ** 16))
    def frames(shape, n, threads, version, seed):
        rng = np.random.default_rng(seed)
        counter[0] += 1
        arrs = [rng.integers(0, 256, size=shape, dtype=np.uint8) for _ in range(n)]
        paths = []
        for i, a in enumerate(arrs):
            p = str(tmp_path / f"f{counter[0]}_{i}.npy")
            write(p, a, version)
            paths.append(p)
        dst = np.zeros((n,) + tuple(shape), np.uint8)
        cp = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
        cs = (ctypes.c_longlong * len(shape))(*shape)
        check(lib.lr_host_read_npy_u8(cp, n, dst.ctypes.data, cs, len(shape), threads))
        assert np.array_equal(dst, np.stack(arrs))

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(lens=st.lists(st.integers(0, 90), min_size=1, max_size=5), ch=st.integers(1, 2), threads=st.integers(1, 4),
           version=st.sampled_from([(1, 0), (2, 0)]), seed=st.integers(0, 2 ** 16))
    def pcm(lens, ch, threads, version, seed):
        rng = np.random.default_rng(seed)
        counter[0] += 1
        target, cap = 64, 64 * 2
        n = len(lens)
        arrs = [rng.integers(-32768, 32768, size=(L,) if ch == 1 else (L, ch), dtype=np.int16) for L in lens]
        paths = []
        for i, a in enumerate(arrs):
            p = str(tmp_path / f"p{counter[0]}_{i}.npy")
            write(p, a, version)
            paths.append(p)
        dst = np.full((n, cap), 7, np.int16)
        meta = np.zeros((3, n), np.int64)
        cp = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
        check(lib.lr_host_read_npy_pcm16(cp, n, dst.ctypes.data, cap, target, meta.ctypes.data, threads))
        for i, a in enumerate(arrs):
            keep = min(lens[i], target)
            assert meta[:, i].tolist() == [i * cap, keep, ch]
            assert np.array_equal(dst[i, :keep * ch], a.reshape(-1)[:keep * ch])
            assert (dst[i, keep * ch:] == 7).all()                      # nothing written past the kept frames

    frames()
    pcm()
