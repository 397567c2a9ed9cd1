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
        traffic = _committed_traffic().get("lm::logmel_kernel")
        return {"kernel": "lm::logmel_kernel", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm"],
                "unit": "GB/s", "frac": achieved / peaks["hbm"],
                "traffic": None if traffic is None else traffic["dram_bytes_per_clip"] * self.batch,
                "peak_source": peaks["src"], "kernel_ms": ms,
                "binds": "fp32 issue slots and the shared-memory pipe (a 400-point real FFT + mel + log per frame: ~490 warp-instructions and "
                         "~134 shared-memory wavefronts per frame; ncu: issue 61-65 %, LSU shared wavefronts 62-68 % of peak, DRAM 21 %), "
                         "not HBM: see DESIGN.md section 4 and profiles/README.md"}

    def extra(self):
        return {"clips_per_sec": None}

    def release(self):
        self.wav = self.stage = self.host = self.out_host = None
        torch.cuda.empty_cache()


def _build_model(kind, num_classes, precision="tf32"):
    """(model, input names) of a train workload; dropout keeps the reference's rates (it is part of the step)."""
    if kind == "mid_fusion_fast":
        from multimodal_lipread_b200.audio_video_models import MidFusionFast
        return MidFusionFast(num_classes, precision=precision), ("audio", "video")
    if kind == "early_fusion_mobilenet":
        from multimodal_lipread_b200.audio_video_models import EarlyFusionAVMobileNet
        return EarlyFusionAVMobileNet(num_classes, precision=precision), ("audio", "video")
    if kind == "early_fusion_resnet":
        from multimodal_lipread_b200.audio_video_models import EarlyFusionAV
        return EarlyFusionAV(num_classes, precision=precision), ("audio", "video")
    if kind == "video_resnet_lstm":
        from multimodal_lipread_b200.video_models import ResNet2DBiLSTM
        return ResNet2DBiLSTM(num_classes, precision=precision), ("video",)
    if kind == "audio_resnet":
        from multimodal_lipread_b200.audio_models import AudioResNet
        return AudioResNet(num_classes, precision=precision), ("audio",)
    if kind == "acv_late_fusion_mobile":
        from multimodal_lipread_b200.audio_cues_video_models import MultimodalAttentionLate
        return MultimodalAttentionLate(num_classes, precision=precision), ("audio", "cue", "video")
    raise ValueError(f"unknown train workload {kind!r}")


def _committed_traffic():
    """ncu-measured DRAM traffic committed under profiles/ (dram__bytes_read.sum + dram__bytes_write.sum), keyed by
    kernel / op family / workload; absent entries report `traffic: null`."""
    import json
    import os
    tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_traffic.json")
    return json.load(open(tp)) if os.path.exists(tp) else {}


# forward GFLOP per clip at 44 / 88 px (analytic 2*MAC counts of the reference modules)
FWD_GFLOP = {"mid_fusion_fast": (0.256, 0.630), "early_fusion_mobilenet": (0.569, 0.944),
             "early_fusion_resnet": (5.79, 17.99), "video_resnet_lstm": (6.04, 18.24), "audio_resnet": (0.72, 0.72),
             "acv_late_fusion_mobile": (1.75, 3.81)}


def flush_l2(dev, _buf={}):
    """Overwrite a 512 MB buffer (4x the 126 MB L2) so that the next kernel starts cold."""
    if dev not in _buf:
        _buf[dev] = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    _buf[dev].zero_()


class AvTrainWorkload:
    """One train step of a lipread_b200 model on synthetic GLips-shaped clips (audio_video/train.py:61-67 and its
    video / audio / audio_cues_video siblings): [log-mel ->] forward -> CE -> backward -> (NCCL allreduce) -> Adam,
    all in lipread_b200 kernels.  The headline workload is audio_video middle_fusion_fast."""
    def __init__(self, dev, batch, cfg, rank, world):
        import torch.distributed as dist
        self.dev, self.batch, self.world, self.cfg = dev, batch, world, cfg
        self.kind = cfg.get("model", "mid_fusion_fast")
        self.precision = cfg.get("precision", "tf32")
        # the arithmetic type of the GEMM-shaped work: "tf32" = tcgen05 kind::tf32 products on fp32 storage (fp32
        # accumulate); "bf16" = bf16 activations / kind::f16; "f32" = strict fp32 SIMT everywhere
        self.dtype = {"tf32": "tf32", "fp32": "f32", "bf16": "bf16"}[self.precision]
        torch.manual_seed(0)                                   # identical replicas on every rank
        model, self.names = _build_model(self.kind, cfg["num_classes"], self.precision)
        self.model = model.to(dev).train()
        self.model.configure_optimizer()
        # a ring of distinct input batches larger than L2 (126 MB), resident in HBM
        per_batch = batch * (20000 * 4 + 29 * cfg["size"] * cfg["size"] * 3)
        self.ring = max(2, min(64, (160 * 1024 * 1024) // per_batch + 1))
        g = 1000 + rank
        self.host = []
        for i in range(self.ring):
            t = []
            for n in self.names:
                if n == "audio":
                    t.append(synthetic.make_waveforms(batch, seed=g * 100 + i).pin_memory())
                elif n == "video":
                    t.append(synthetic.make_lips_u8(batch, size=cfg["size"], seed=g * 100 + 50 + i,
                                                    grayscale=cfg["grayscale"]).pin_memory())
                else:
                    t.append(synthetic.make_cues(batch, seed=g * 100 + 30 + i).pin_memory())
            t.append(synthetic.make_labels(batch, cfg["num_classes"], seed=g * 100 + 77 + i).pin_memory())
            self.host.append(tuple(t))
        self.devb = [tuple(x.to(dev) for x in t) for t in self.host]
        # e2e: two staging sets; the next batch's H2D runs on a copy stream while the current step computes
        self.stages = [tuple(torch.empty_like(t) for t in self.devb[0]) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.staged = [torch.cuda.Event() for _ in range(2)]
        self.prefetched = None
        self.i = 0
        self.loss_host = torch.zeros(1).pin_memory()
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.host[0])
        self.d2h_bytes = 4
        self.allreduce = None
        if world > 1:
            def allreduce(grad):
                dist.all_reduce(grad, op=dist.ReduceOp.SUM)
            self.allreduce = allreduce
        self.timer = _KernelTimer()
        self.last_loss = None

    def units_per_step(self):
        return self.batch

    def launches_per_step(self):
        return self.model.launches_per_step()

    def reset_kernel_timer(self):
        self.timer.reset()

    def kernel_ms(self):
        return None

    def _step(self, *tensors):
        loss, _ = self.model.train_step(*tensors, grad_allreduce=self.allreduce, world=self.world)
        self.last_loss = loss
        return loss

    def step_device(self):
        t = self.devb[self.i % self.ring]
        self.i += 1
        return self._step(*t)

    def _prefetch(self, i):
        """H2D of batch i from pinned host memory into staging set i % 2, on the copy stream."""
        k = i % 2
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(self.stages[k], self.host[i % self.ring]):
                dst.copy_(src, non_blocking=True)
            self.staged[k].record(self.copy_stream)
        self.prefetched = i

    def step_e2e(self):
        """One step from HOST buffers: every step's inputs cross PCIe from pinned memory and its loss is read back
        (host sync) inside the timed region; the copy of step i+1 is issued before step i's loss is awaited, so it
        overlaps the compute the way a prefetching loader (data.DeviceBatchLoader) does."""
        i = self.i
        self.i += 1
        cur = torch.cuda.current_stream()
        if self.prefetched != i:
            self.copy_stream.wait_stream(cur)
            self._prefetch(i)
        cur.wait_event(self.staged[i % 2])
        loss = self._step(*self.stages[i % 2])          # train_step copies the inputs into the plan's own buffers
        self._prefetch(i + 1)                           # set (i+1) % 2 was last read by step i-1, host-synced below

        self.loss_host.copy_(loss, non_blocking=True)
        cur.synchronize()
        return self.loss_host

    def files_e2e(self, n_batches=16, workers=4):
        """the same train step fed from the reference's on-disk formats (uint8 .npy lip regions,
        16-bit PCM) through data.DeviceBatchLoader -- pinned ring, async H2D, lr_pcm_ingest + log-mel on the copy
        stream -- instead of from tensors already in memory.  Files are written to a temp dir first (page cache
        warm: this measures the loader and the PCIe path, not the disk).  One epoch warm-up, one epoch timed."""
        import os
        import shutil
        import tempfile
        import time
        import numpy as np
        from multimodal_lipread_b200 import data
        if tuple(self.names) != ("audio", "video"):
            return None
        tmp = tempfile.mkdtemp(prefix="lipread_files_")
        try:
            root = os.path.join(tmp, "GLips")
            k = 0
            for i in range(n_batches):
                wav, lips, labels = self.host[i % self.ring]
                pcm = wav.clamp(-32768, 32767).to(torch.int16).numpy()
                for j in range(self.batch):
                    cname = f"c{int(labels[j]):02d}"
                    vdir = os.path.join(root, "lipread_files", cname, "train")
                    ldir = os.path.join(root + "_lip_regions", "lipread_files", cname, "train")
                    os.makedirs(vdir, exist_ok=True)
                    os.makedirs(ldir, exist_ok=True)
                    base = f"{cname}_{k:05d}"
                    k += 1
                    open(os.path.join(vdir, base + ".mp4"), "wb").close()
                    np.save(os.path.join(vdir, base + ".npy"), pcm[j])
                    np.save(os.path.join(ldir, base + ".npy"), lips[j].numpy())
            ds = data.GLipsMultimodalDataset(root, 117, "train", audio_ext=".npy")
            loader = data.DeviceBatchLoader(ds, self.batch, shuffle=True, drop_last=True, device=self.dev, depth=3,
                                            workers=workers)
            best = None
            for epoch in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n = 0
                for mel, frames, labels in loader:
                    loss = self._step(mel, frames, labels)
                    n += labels.numel()
                self.loss_host.copy_(loss, non_blocking=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                if epoch > 0:
                    best = dt if best is None else min(best, dt)
            loader.close()
            return {"value": n / best, "unit": "clips/s", "clips_per_epoch": n,
                    "source": "uint8 .npy lip regions + int16 PCM .npy on local disk, page cache warm",
                    "loader": f"DeviceBatchLoader depth 3, {workers} reader threads, shuffle",
                    "h2d_bytes_per_step": int(self.batch * (lips[0].numel() + 2 * 20000)), "timing": "host wall clock, best of 2 epochs"}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    def profile_ops(self):
        """Device time of every kernel of one train step, each replayed alone from its own CUDA graph after an L2
        flush (CUDA events on the launching stream, best of 3): no host launch overhead, cold cache like in the step."""
        from multimodal_lipread_b200 import engine
        plan = next(p for p in self.model._plans.values() if p.with_backward)
        rows = []
        for phase, ops in (("fwd", plan.fwd), ("bwd", plan.bwd)):
            for name, args, ms in engine.profile_ops_graph(ops, reps=1, flush=lambda: flush_l2(self.dev)):
                rows.append({"phase": phase, "op": name, "ms": ms, "bytes": engine.op_algorithmic_bytes(name, args),
                             "shape": [a for a in args if isinstance(a, int) and 0 < a < (1 << 31)][-6:]})
        return rows

    def roofline(self, kernel_ms, ms_step, peaks):
        """Dominant kernel = the op FAMILY (all launches of one C-ABI kernel) with the largest share of the step's
        device time; achieved = sum of algorithmic bytes / sum of CUDA-event durations over ALL its launches (each
        timed alone from its own graph after an L2 flush).  `traffic` (per launch, averaged) and `step_frac` use the
        ncu DRAM byte counts committed in profiles/r2_traffic.json for this workload."""
        rows = self.profile_ops()
        self.op_rows = rows
        total = sum(r["ms"] for r in rows)
        fam = {}
        for r in rows:
            f = fam.setdefault(r["op"], {"ms": 0.0, "bytes": 0.0, "n": 0, "n_bytes": 0, "ms_bytes": 0.0})
            f["ms"] += r["ms"]; f["n"] += 1
            if r["bytes"]:
                f["bytes"] += r["bytes"]; f["n_bytes"] += 1; f["ms_bytes"] += r["ms"]
        top_name, top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        achieved = top["bytes"] / 1e9 / (top["ms_bytes"] / 1e3) if top["ms_bytes"] else None
        committed = _committed_traffic().get(f"{self.kind}:{self.precision}", {})
        fam_traffic = committed.get("families", {}).get(top_name)
        step_dram = committed.get("step_dram_bytes")
        f44, f88 = FWD_GFLOP[self.kind]
        fwd_gflop = f88 if self.cfg["size"] == 88 else f44
        slowest = max(rows, key=lambda r: r["ms"])
        return {"kernel": f"{top_name} (op family: {top['n']} launches per step)", "bound": "hbm", "achieved": achieved,
                "peak": peaks["hbm"], "unit": "GB/s", "frac": None if achieved is None else achieved / peaks["hbm"],
                "traffic": None if not fam_traffic else fam_traffic["dram_bytes"] / fam_traffic["launches"],
                "algorithmic_bytes_per_launch": top["bytes"] / max(top["n_bytes"], 1),
                "peak_source": peaks["src"], "kernel_ms": top["ms"] / top["n"], "family_ms_per_step": top["ms"],
                "kernel_share_of_step": top["ms"] / total, "step_ms_sum_of_kernels_cold": total,
                # whole step: DRAM bytes of one step (ncu, committed) / the device-timed step / peak
                "step_dram_bytes": step_dram,
                "step_frac": None if not step_dram else step_dram / 1e9 / (ms_step / 1e3) / peaks["hbm"],
                "slowest_launch": f"{slowest['op']} {slowest['shape']} ({slowest['phase']}): {slowest['ms'] * 1e3:.0f} us",
                "time_by_op_ms": {k: round(v["ms"], 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
                # every op family with a byte model: sum of algorithmic bytes / sum of cold device times, vs the HBM peak
                "hbm_frac_by_op": {k: {"launches": v["n_bytes"], "GBps": round(v["bytes"] / 1e9 / (v["ms_bytes"] / 1e3), 1),
                                       "frac": round(v["bytes"] / 1e9 / (v["ms_bytes"] / 1e3) / peaks["hbm"], 3)}
                                   for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]) if v["ms_bytes"]},
                "step_model_tflops": self.model_tflops(ms_step)}

    def model_tflops(self, ms_step):
        """3 x analytic forward FLOPs of the reference model per device-timed step."""
        f44, f88 = FWD_GFLOP[self.kind]
        return 3 * (f88 if self.cfg["size"] == 88 else f44) * self.batch / (ms_step / 1e3) / 1e3

    def time_without_allreduce(self, steps, barrier):
        """ms per step of the same captured step WITHOUT the collective (grad_allreduce=None keeps 1/world in Adam):
        the difference to the timed step is the exposed collective time."""
        saved, self.allreduce = self.allreduce, None
        for _ in range(3):
            self.step_device()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            self.step_device()
        b.record()
        barrier()
        self.allreduce = saved
        return a.elapsed_time(b) / steps

    def release(self):
        """Drop graphs (they may hold captured NCCL kernels) and every device buffer of this workload."""
        import gc
        self.model._graphs.clear()
        self.model._plans.clear()
        self.devb = self.stages = self.host = None
        self.model = None
        gc.collect()
        torch.cuda.empty_cache()

    def extra(self):
        return {"final_loss": float(self.last_loss.item()) if self.last_loss is not None else None,
                "input_ring_batches": self.ring}

