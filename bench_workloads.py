"""Workloads timed by bench.py (B200 arm only; the CPU arm lives in bench.py:cpu_reference)."""
import torch

from multimodal_lipread_b200 import synthetic
from multimodal_lipread_b200.audio_processor import AudioProcessor

LOGMEL_BYTES_PER_CLIP = 20000 * 4 + 80 * 117 * 4


class _KernelTimer:
    """CUDA-event pairs around one kernel launch, on the launching (current torch) stream."""

    def __init__(self):
        self.pairs = []

    def reset(self):
        self.pairs = []

    def wrap(self, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        self.pairs.append((a, b))
        return out

    def mean_ms(self):
        if not self.pairs:
            return None
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in self.pairs) / len(self.pairs)


class LogmelWorkload:
    dtype = "f32"

    def __init__(self, dev, batch, cfg, rank, world):
        self.dev, self.batch = dev, batch
        self.ap = AudioProcessor(device=dev)
        base = synthetic.make_waveforms(min(batch, 2048), seed=1234 + rank)
        reps = (batch + base.shape[0] - 1) // base.shape[0]
        host = base.repeat(reps, 1)[:batch].contiguous()
        self.host = host.pin_memory()
        self.wav = self.host.to(dev)                    # 80 kB per clip: 16384 clips = 1.3 GB >> L2
        self.stage = torch.empty_like(self.wav)
        self.out_host = torch.empty(batch, 80, 117, dtype=torch.float32).pin_memory()
        self.timer = _KernelTimer()
        self.h2d_bytes = host.numel() * 4
        self.d2h_bytes = self.out_host.numel() * 4

    def units_per_step(self):
        return self.batch * LOGMEL_BYTES_PER_CLIP / 1e9

    def launches_per_step(self):
        return 1

    def reset_kernel_timer(self):
        self.timer.reset()

    def kernel_ms(self):
        return self.timer.mean_ms()

    def step_device(self):
        return self.timer.wrap(lambda: self.ap.frontend(self.wav))

    def step_e2e(self):
        self.stage.copy_(self.host, non_blocking=True)
        out = self.ap.frontend(self.stage)
        self.out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def roofline(self, kernel_ms, ms_step, peaks):
        ms = kernel_ms or ms_step
        achieved = self.batch * LOGMEL_BYTES_PER_CLIP / 1e9 / (ms / 1e3)
        return {"kernel": "lm::logmel_kernel", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm"],
                "unit": "GB/s", "frac": achieved / peaks["hbm"], "traffic": None, "peak_source": peaks["src"],
                "kernel_ms": ms}

    def extra(self):
        return {"clips_per_sec": None}
