"""K1 parity on the GPU, through the C ABI (ops.logmel -> lr_logmel_fwd), against
  (a) the committed golden vectors produced by the reference's own AudioProcessor, and
  (b) the oracle (float64 restatement + the fp32 torchaudio port) on seeded synthetic clips.

Tolerance (BASELINE.json north_star: "log-mel within 1e-4 relative in fp32"; SURVEY.md 7.3 explains
why element-wise relative error is meaningless for a standardised signal):
    max|out - ref| <= 1e-4 * max|ref|   per clip (norm-wise relative), against the float64 oracle,
    and no worse than 2x the reference's own fp32 round-off against the fp32 reference.
"""
import os

import numpy as np
import pytest
import torch

from oracle import logmel as olm
from oracle.frontend import AudioProcessorPort

pytestmark = pytest.mark.gpu

RTOL_NORMWISE = 1e-4


@pytest.fixture(scope="module")
def ap(cuda_device):
    from multimodal_lipread_b200.audio_processor import AudioProcessor
    return AudioProcessor(device=cuda_device)


def _normwise(out, ref):
    err = np.abs(out - ref).max(axis=(1, 2))
    return err / np.abs(ref).max(axis=(1, 2))


def test_golden_vectors(ap, golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_golden.npz"))
    wave = torch.from_numpy(g["wave"])
    out = ap.frontend(wave).cpu().numpy()
    ref32 = g["out"]
    ref64 = olm.logmel_frontend(g["wave"], window=g["window"], fb=g["fb"])
    assert out.shape == ref32.shape == (8, 80, 117) and out.dtype == np.float32
    n = 7                                    # clip 7 is digital silence (std == 0), checked below
    assert (_normwise(out[:n], ref64[:n]) <= RTOL_NORMWISE).all(), _normwise(out[:n], ref64[:n])
    ref_noise = _normwise(ref32[:n].astype(np.float64), ref64[:n])
    assert (_normwise(out[:n], ref32[:n]) <= RTOL_NORMWISE + ref_noise).all()
    # silent clip: a DOCUMENTED DEVIATION.  Every log-mel value of digital silence equals
    # ln(1e-9); exact arithmetic gives (x - mean) = 0 -> output 0, which is what the kernel returns.  The reference
    # divides its fp32 round-off of (x - mean) (~1e-6) by (std + 1e-9) with std == 0 and returns -0.9994 everywhere;
    # that golden value is kept as an expected failure below (test_silent_clip_golden_value_of_the_reference).
    assert (out[7] == 0.0).all()
    raw = ap.compute_melspectrogram(wave).cpu().numpy()
    assert raw.shape == (8, 80, 126)
    assert np.abs(raw - g["logmel_raw"]).max() < 2e-3
    assert np.allclose(raw[7], np.log(1e-9), atol=1e-5)


@pytest.mark.xfail(strict=True, reason="documented deviation: the reference's silent-clip output is fp32 round-off / 1e-9 "
                   "(-0.9994 everywhere); the kernel returns the exact-arithmetic value 0")
def test_silent_clip_golden_value_of_the_reference(ap, golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_golden.npz"))
    out = ap.frontend(torch.from_numpy(g["wave"][7:8])).cpu().numpy()
    assert np.abs(out[0] - g["out"][7]).max() <= 1e-4 * np.abs(g["out"][7]).max()


@pytest.mark.parametrize("kind", ["pcm", "unit", "tone"])
def test_seeded_clips_match_oracle(ap, kind):
    from multimodal_lipread_b200 import synthetic
    wav = synthetic.make_waveforms(64, seed=99, kind=kind, pad_fraction=0.25)
    out = ap.frontend(wav).cpu().numpy()
    ref64 = olm.logmel_frontend(wav.numpy())
    # window / fb of the oracle are float64 here; the kernel uses the fp32 torchaudio-formula buffers
    port = AudioProcessorPort()
    ref64b = olm.logmel_frontend(wav.numpy(), window=port.window.numpy(), fb=port.fb.numpy())
    assert (_normwise(out, ref64b) <= RTOL_NORMWISE).all(), _normwise(out, ref64b).max()
    # float64 window/fb instead of the fp32 buffers: only the oracle-side buffer rounding is added
    assert (_normwise(out, ref64) <= RTOL_NORMWISE + _normwise(ref64b, ref64)).all()
    ref32 = port.batch_frontend_loop(wav).numpy()
    noise = _normwise(ref32.astype(np.float64), ref64b)
    assert (_normwise(out, ref32) <= RTOL_NORMWISE + noise).all()


def test_window_and_fb_buffers_equal_torchaudio(ap):
    port = AudioProcessorPort()
    assert torch.equal(ap.window.cpu(), port.window)
    assert torch.equal(ap.fb.cpu(), port.fb)


def test_ragged_batches_and_single_clip(ap):
    """Batch sizes around the persistent grid (1, SMs-1, SMs, SMs+1, 3*SMs+5) and a 1-D clip."""
    from multimodal_lipread_b200 import synthetic
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    wav = synthetic.make_waveforms(3 * sms + 5, seed=5, kind="pcm")
    full = ap.frontend(wav)
    for b in (1, sms - 1, sms, sms + 1):
        part = ap.frontend(wav[:b])
        assert torch.equal(part, full[:b])           # a clip's result does not depend on the batch
    one = ap.frontend(wav[3])
    assert one.shape == (80, 117) and torch.equal(one, full[3])
    assert ap.frontend(wav[:0]).shape == (0, 80, 117)


def test_short_and_long_clips_pad_truncate(ap):
    """audio_processor.py:40-44: right zero-pad short clips, truncate long ones."""
    from multimodal_lipread_b200 import synthetic
    wav = synthetic.make_waveforms(2, seed=11, kind="pcm", pad_fraction=0.0)
    short = wav[0, :12345]
    padded = torch.cat([short, torch.zeros(20000 - 12345)])
    assert torch.equal(ap.frontend(short), ap.frontend(padded))
    long = torch.cat([wav[1], wav[0]])
    assert torch.equal(ap.frontend(long), ap.frontend(wav[1]))


def test_normalize_alone_and_crop_sizes(ap):
    from multimodal_lipread_b200 import synthetic
    wav = synthetic.make_waveforms(4, seed=3, kind="pcm")
    raw = ap.compute_melspectrogram(wav)
    norm = ap.normalize_spectrogram(raw).cpu().numpy()
    ref = olm.normalize(raw.cpu().numpy().astype(np.float64))
    assert np.abs(norm - ref).max() <= 1e-5 * np.abs(ref).max()
    for n_out in (1, 100, 126):
        o = ap.frontend(wav, n_out=n_out).cpu().numpy()
        assert o.shape == (4, 80, n_out)
        assert np.abs(o - ref[:, :, :n_out]).max() <= 1e-5 * np.abs(ref).max()


def test_size_independent_properties_full_batch(ap):
    """At a bench-sized batch: per-clip mean 0 / unbiased std 1 over the 126 frames (checked via
    n_out = 126), and scale invariance: frontend(c * x) == frontend(x) up to fp32 round-off because
    log turns the gain into an additive constant that the normalisation removes."""
    from multimodal_lipread_b200 import synthetic
    wav = synthetic.make_waveforms(4096, seed=21, kind="pcm", pad_fraction=0.0)
    o = ap.frontend(wav, n_out=126)
    flat = o.reshape(o.shape[0], -1).double()
    assert flat.mean(dim=1).abs().max().item() < 1e-5
    assert (flat.std(dim=1) - 1.0).abs().max().item() < 1e-5
    o2 = ap.frontend(wav * 4.0, n_out=126)
    assert (o2 - o).abs().max().item() < 2e-4
    assert ap.frontend(wav).shape == (4096, 80, 117)


def test_launch_counter_counts_our_kernels(ap):
    from multimodal_lipread_b200 import _lib, synthetic
    wav = synthetic.make_waveforms(8, seed=1).cuda()
    before = _lib.launch_count()
    ap.frontend(wav)
    torch.cuda.synchronize()
    assert _lib.launch_count() == before + 1
