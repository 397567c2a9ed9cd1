"""On-GPU input path, through the C ABI (lr_pcm_ingest -> lr_logmel_fwd, uint8 frames into the stem gather), against
the golden items of the reference's own GLipsMultimodalDataset and against oracle/dataset.py."""
import os

import numpy as np
import pytest
import torch

from multimodal_lipread_b200 import data, synthetic
from oracle import dataset as ods
from oracle import logmel as olm

pytestmark = pytest.mark.gpu


def _key(video_path, root):
    rel = os.path.relpath(video_path, root + "_lip_regions")
    return os.path.splitext(rel)[0].split(os.sep, 1)[1].replace(os.sep, "/")


def _decode(path):
    a = np.load(path)
    return a, 1, 1.0


@pytest.mark.parametrize("case", ["mono_aligned", "mono_odd_offsets", "stereo", "mixed_channels", "scaled"])
def test_pcm_ingest_is_bit_exact(cuda_device, case):
    """Ragged lengths on both sides of 20 000, empty clips, offsets of any parity, 1-3 channels: equals
    audio_processor.py:29,37-44 restated in oracle/dataset.py load_audio, bit for bit."""
    from multimodal_lipread_b200 import ops
    g = np.random.default_rng(5)
    lens = [0, 1, 3, 19999, 20000, 20001, 26000, 12345, 7, 20003]
    chans = {"mono_aligned": [1] * 10, "mono_odd_offsets": [1] * 10, "stereo": [2] * 10,
             "mixed_channels": [1, 2, 3, 1, 2, 3, 1, 2, 3, 2], "scaled": [2, 1] * 5}[case]
    scale = 1.0 / 32768.0 if case == "scaled" else 1.0
    clips = [g.integers(-32768, 32768, size=(n, c), dtype=np.int16) for n, c in zip(lens, chans)]
    parts, offs, pos = [], [], 0
    for a in clips:
        pad = (-pos) % 8 if case == "mono_aligned" else (1 if case == "mono_odd_offsets" else 0)
        parts.append(np.zeros(pad, np.int16))
        pos += pad
        offs.append(pos)
        parts.append(a.reshape(-1))
        pos += a.size
    packed = torch.from_numpy(np.concatenate(parts + [np.zeros(8, np.int16)])).to(cuda_device)
    out = ops.pcm_ingest(packed, torch.tensor(offs, dtype=torch.int64, device=cuda_device),
                         torch.tensor(lens, dtype=torch.int32, device=cuda_device),
                         torch.tensor(chans, dtype=torch.int32, device=cuda_device), scale, 20000).cpu()
    for b, a in enumerate(clips):
        ref = ods.load_audio(a[:, 0] if a.shape[1] == 1 else a.T, scale)
        assert torch.equal(out[b], ref), (case, b, (out[b] - ref).abs().max())
    # all-mono shortcut (channels == NULL) and the empty batch
    if case.startswith("mono"):
        out2 = ops.pcm_ingest(packed, torch.tensor(offs, dtype=torch.int64, device=cuda_device),
                              torch.tensor(lens, dtype=torch.int32, device=cuda_device),
                              torch.empty(0, dtype=torch.int32, device=cuda_device), scale, 20000).cpu()
        assert torch.equal(out2, out)
    e = ops.pcm_ingest(packed, torch.empty(0, dtype=torch.int64, device=cuda_device),
                       torch.empty(0, dtype=torch.int32, device=cuda_device),
                       torch.empty(0, dtype=torch.int32, device=cuda_device), scale, 20000)
    assert e.shape == (0, 20000)


@pytest.mark.parametrize("batch_size,shuffle", [(4, False), (2, True), (16, False)])
def test_loader_batches_match_the_reference_items(cuda_device, golden_dir, tmp_path, batch_size, shuffle):
    golden = np.load(os.path.join(golden_dir, "dataset_golden.npz"))
    root = str(tmp_path / "GLips_4")
    synthetic.write_dataset_tree(root)
    for split in ("train", "val"):
        ds = data.GLipsMultimodalDataset(root, 117, split)
        loader = data.DeviceBatchLoader(ds, batch_size, shuffle=shuffle, device=cuda_device, depth=2, workers=4, seed=11,
                                        audio_decoder=_decode)
        order = data.batch_indices(len(ds), batch_size, shuffle, False, torch.Generator().manual_seed(11))
        seen = []
        assert len(loader) == len(order)
        for idxs, (mel, lips, labels) in zip(order, loader):
            assert mel.is_cuda and lips.dtype == torch.uint8 and tuple(lips.shape) == (len(idxs), 5, 12, 12, 3)
            mel_h, lips_h, lab_h = mel.cpu(), lips.cpu(), labels.cpu()       # copy out: the ring reuses the buffers
            for j, i in enumerate(idxs):
                k = _key(ds.samples[i]["video_path"], root)
                seen.append(k)
                ref = golden[f"mel|{k}"]                                    # the reference's own fp32 item
                ref64 = olm.logmel_frontend(ods.load_audio(np.load(ds.samples[i]["audio_path"])).numpy()[None])[0]
                noise = np.abs(ref - ref64).max()
                assert np.abs(mel_h[j].numpy() - ref64).max() <= 1e-4 * np.abs(ref64).max()
                assert np.abs(mel_h[j].numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + noise
                got = (lips_h[j].float() / 255.0).permute(3, 0, 1, 2)
                assert np.array_equal(got.numpy(), golden[f"lips|{k}"])
                assert int(lab_h[j]) == int(golden[f"label|{k}"])
        assert sorted(seen) == golden[f"keys_av|{split}"].tolist()
        loader.close()
    vds = data.VisualDataset(root, root + "_lip_regions", "train")
    vl = data.DeviceBatchLoader(vds, batch_size, device=cuda_device, workers=2)
    n = 0
    for batch in vl:
        assert set(batch) == {"lip_regions", "label"}
        for j in range(batch["label"].numel()):
            k = _key(vds.samples[n][0], root)
            assert int(batch["label"][j]) == int(golden[f"vlabel|{k}"])
            assert (batch["lip_regions"][j].double() / 255.0).sum().item() == pytest.approx(float(golden[f"vsum|{k}"]), rel=1e-6)
            n += 1
    assert n == len(vds)
    vl.close()


def test_loader_feeds_the_train_step(cuda_device, tmp_path):
    """Files -> pinned ring -> HBM -> MidFusionFast.train_step: the loss equals the oracle model's on the items the
    oracle dataset (pinned to the reference's) yields for the same files."""
    from multimodal_lipread_b200 import audio_video_models as M
    from oracle import av_models as O
    root = str(tmp_path / "GLips_4")
    synthetic.write_dataset_tree(root, per_split={"train": 2}, T=6, size=44, missing_every=1000)
    ds = data.GLipsMultimodalDataset(root, 117, "train")
    assert len(ds) == 6
    torch.manual_seed(0)
    ref = O.MidFusionFastOracle(3).train()
    ours = M.create_mid_fusion_fast(3, O.DictConfig()).to(cuda_device)
    ours.load_state_dict(ref.state_dict())
    ours.train()
    ours.configure_optimizer(lr=3e-4)
    loader = data.DeviceBatchLoader(ds, 6, device=cuda_device, audio_decoder=_decode)
    n = 0
    for mel, lips, labels in loader:
        loss, logits = ours.train_step(mel, lips, labels)
        items = [ods.getitem_multimodal(s, lambda p: np.load(p)) for s in ds.samples]
        rlogits = ref(torch.stack([i[0] for i in items]), torch.stack([i[1] for i in items]))
        rl = torch.nn.functional.cross_entropy(rlogits, torch.stack([i[2] for i in items]))
        assert abs(float(loss) - float(rl.detach())) < 5e-4 * max(1.0, abs(float(rl.detach())))
        assert (logits.cpu() - rlogits.detach()).abs().max() < 5e-3 * rlogits.abs().max()
        n += 1
    assert n == 1
    loader.close()


def test_fit_run_checkpoints_logs_and_resume_interop(cuda_device, tmp_path):
    """The reference's main() loop on files: per-epoch CSV rows, `<model>_checkpoint.pth` / `model_best.pth` in the
    reference's dict format whose "optimizer" entry loads into torch.optim.Adam (and back), resume continues the
    Adam state exactly."""
    import csv
    from multimodal_lipread_b200 import audio_video_models as M, train as T
    from oracle import av_models as O
    root = str(tmp_path / "GLips_4")
    synthetic.write_dataset_tree(root, per_split={"train": 4, "val": 2, "test": 2}, T=6, size=44, missing_every=1000)
    loaders = tuple(data.DeviceBatchLoader(data.GLipsMultimodalDataset(root, 117, s), 4, shuffle=(s == "train"),
                                           device=cuda_device, audio_decoder=_decode) for s in ("train", "val", "test"))
    torch.manual_seed(0)
    model = M.create_mid_fusion_fast(3, O.DictConfig()).to(cuda_device)
    model.configure_optimizer(lr=3e-4)
    save_dir, out_dir = str(tmp_path / "ckpt"), str(tmp_path / "metrics")
    res = T.fit(model, "middle_fusion_fast", loaders, cuda_device, 2, save_dir, out_dir, schedule=("max", 5), log=lambda s: None)
    rows = list(csv.reader(open(os.path.join(out_dir, "middle_fusion_fast_training_log.csv"))))
    assert rows[0] == T.LOG_HEADER and [r[0] for r in rows[1:]] == ["1", "2"]
    assert all(np.isfinite(float(v)) for r in rows[1:] for v in r[1:])
    assert 0.0 <= res["test_acc"] <= 100.0 and os.path.exists(os.path.join(save_dir, "test_results.txt"))
    ck = torch.load(os.path.join(save_dir, "middle_fusion_fast_checkpoint.pth"), map_location="cpu")
    # the reference's four keys (video/train.py:246-251) + the dropout mask counter, which its resume code never reads
    assert set(ck) == {"epoch", "state_dict", "optimizer", "best_val_acc", "rng_step"} and ck["epoch"] == 3
    # the reference side can resume from it: same keys, and torch's Adam accepts the optimizer entry
    ref = O.MidFusionFastOracle(3)
    ref.load_state_dict(ck["state_dict"])
    opt = torch.optim.Adam(ref.parameters(), lr=1.0)
    opt.load_state_dict(ck["optimizer"])
    assert opt.param_groups[0]["lr"] == pytest.approx(3e-4) and len(opt.state) == len(list(ref.parameters()))
    steps_done = 2 * len(loaders[0])
    assert all(float(s["step"]) == steps_done for s in opt.state.values())
    # ... take one step on the reference side, hand the state back, and take the same step here from the checkpoint
    items = [ods.getitem_multimodal(s, lambda p: np.load(p)) for s in loaders[0].ds.samples[:4]]
    mel, lips, lab = (torch.stack([i[k] for i in items]) for k in range(3))
    ref.train()
    torch.nn.functional.cross_entropy(ref(mel, lips), lab).backward()
    opt.step()
    torch.manual_seed(1)
    again = M.create_mid_fusion_fast(3, O.DictConfig({"precision": {"compute": "fp32"}})).to(cuda_device)
    start, best = T.resume(again, os.path.join(save_dir, "middle_fusion_fast_checkpoint.pth"))
    assert start == 3 and best == ck["best_val_acc"]
    again.train()
    again.train_step(mel.to(cuda_device), lips.to(cuda_device), lab.to(cuda_device))
    # one Adam step moves a weight by at most ~lr; both sides start from the same moments, so the step agrees to
    # a small fraction of lr on average (single elements may sit on a ReLU kink: bounded by lr itself)
    for (n, p), q in zip(ref.named_parameters(), again.parameters()):
        d = (p.detach() - q.detach().cpu()).abs()
        assert d.mean() <= 0.02 * 3e-4 and d.max() <= 1.01 * 3e-4, (n, d.mean(), d.max())
    back = again.optimizer_state_dict()
    assert float(back["state"][0]["step"]) == steps_done + 1
    m_ref = opt.state_dict()["state"][0]["exp_avg"]
    assert (back["state"][0]["exp_avg"].cpu() - m_ref).abs().max() <= 1e-3 * m_ref.abs().max() + 1e-7
    for ld in loaders:
        ld.close()


def test_triple_and_audio_loaders_match_the_reference_items(cuda_device, golden_dir, tmp_path):
    """(mel, cue, lips, label) batches of the audio + cue + video dataset and (mel, label) batches of the audio-only
    one, against the items of the reference's own MultimodalTripleDataset (collate_fn_triple layout)."""
    golden = np.load(os.path.join(golden_dir, "dataset_golden.npz"))
    glips, cue_root, lip_root = synthetic.write_triple_tree(str(tmp_path / "triple"))
    for split in ("train", "val"):
        ds = data.MultimodalTripleDataset(glips, cue_root, lip_root, 117, split, cache_dir=str(tmp_path / "cache"),
                                          embedder=synthetic.fake_sentence_embedding)
        loader = data.DeviceBatchLoader(ds, 4, device=cuda_device, audio_decoder=_decode, workers=2)
        i = 0
        for mel, cue, lips, labels in loader:
            assert cue.shape[1] == 768 and lips.dtype == torch.uint8 and mel.shape[1:] == (80, 117)
            mel_h, cue_h, lips_h, lab_h = mel.cpu(), cue.cpu(), lips.cpu(), labels.cpu()
            for j in range(lab_h.numel()):
                s = ds.samples[i]
                k = f"{s['word']}/{split}/{s['sid']}"
                ref = golden[f"tmel|{k}"]
                assert np.abs(mel_h[j].numpy() - ref).max() <= 2e-4 * np.abs(ref).max()
                assert np.array_equal(cue_h[j].numpy(), golden[f"tcue|{k}"])
                assert (lips_h[j].double() / 255.0).sum().item() == pytest.approx(float(golden[f"tlipsum|{k}"]), rel=1e-6)
                assert int(lab_h[j]) == int(golden[f"tlabel|{k}"])
                i += 1
        assert i == len(ds) == len(golden[f"keys_triple|{split}"])
        loader.close()
    ads = data.GLipsDataset(glips, 117, "train")
    n = 0
    al = data.DeviceBatchLoader(ads, 8, device=cuda_device, audio_decoder=_decode)
    for mel, labels in al:
        assert mel.shape[1:] == (80, 117) and torch.isfinite(mel).all()
        assert labels.tolist() == [s["label"] for s in ads.samples[n:n + labels.numel()]]
        n += labels.numel()
    assert n == len(ads)
    al.close()


def test_unit_clip_rule_matches_the_reference_division(cuda_device):
    """dataset.py:256-258: `if arr.max() > 1.0: arr = arr / 255.0` -- clips of 0 / 1 pixels are NOT divided."""
    from multimodal_lipread_b200._lib import lib, check
    g = torch.Generator().manual_seed(1)
    clips = torch.randint(0, 256, (5, 3, 8, 8, 3), generator=g, dtype=torch.uint8)
    clips[1] = torch.randint(0, 2, clips[1].shape, generator=g, dtype=torch.uint8)       # max == 1
    clips[2] = 0                                                                          # max == 0
    clips[3] = torch.randint(0, 2, clips[3].shape, generator=g, dtype=torch.uint8)
    clips[3, 2, 7, 7, 2] = 2                                                              # a single 2 in the tail
    ref = []
    for c in clips:
        arr = c.numpy().astype(np.float32)
        ref.append(torch.from_numpy(arr / 255.0 if arr.max() > 1.0 else arr))
    d = clips.to(cuda_device)
    check(lib.lr_u8_unit_clips(d.data_ptr(), 5, clips[0].numel(), torch.cuda.current_stream().cuda_stream))
    ours = d.cpu().float() / 255.0
    for j in range(5):
        assert torch.equal(ours[j], ref[j]), j
    assert lib.lr_u8_unit_clips(d.data_ptr(), 5, 100, None) == -1
