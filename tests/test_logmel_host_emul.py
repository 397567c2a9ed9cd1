"""CPU check of the K1 kernel's math: the kernel's own __host__ __device__ stage functions
(csrc/logmel_core.cuh) compiled for the host and driven task by task, against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import logmel as olm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(ROOT, "build", "liblogmel_host_emul.so")
    src = os.path.join(ROOT, "tests", "csrc", "logmel_host_emul.cu")
    core = os.path.join(ROOT, "multimodal_lipread_b200", "csrc", "logmel_core.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared",
                        src, "-o", out], check=True)
    return ctypes.CDLL(out)


def test_dft_codelets(emul):
    rng = np.random.default_rng(0)
    for n, fn in ((25, None), (8, None)):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        if n == 25:
            o = np.zeros(25, np.complex64)
            emul.dft25_host(x.ctypes.data_as(FP), o.ctypes.data_as(FP))
        else:
            o = x.copy()
            emul.dft8_host(o.ctypes.data_as(FP))
        assert np.abs(o - np.fft.fft(x.astype(np.complex128))).max() < 5e-6


def test_stages_match_oracle(emul, golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel_golden.npz"))
    window, fb = g["window"], np.ascontiguousarray(g["fb"])
    for i in range(g["wave"].shape[0]):
        wav = np.ascontiguousarray(g["wave"][i])
        out = np.zeros((80, 126), np.float32)
        st = emul.logmel_host_emul(wav.ctypes.data_as(FP), window.ctypes.data_as(FP), fb.ctypes.data_as(FP),
                                   out.ctypes.data_as(FP))
        assert st == 0
        ref64 = olm.log_mel(wav, window=window, fb=fb)
        ref32 = g["logmel_raw"][i]
        # raw log-power values are ~5..25; fp32 FFT round-off shows as ~1e-5 abs, tonal clip ~3e-4
        assert np.abs(out - ref64).max() <= 2.0 * max(np.abs(ref32 - ref64).max(), 1e-5)
        n64 = olm.normalize(out.astype(np.float64))[:, :117]
        r64 = olm.normalize(ref64)[:, :117]
        if i < g["wave"].shape[0] - 1:
            assert np.abs(n64 - r64).max() <= 1e-4 * np.abs(r64).max()
