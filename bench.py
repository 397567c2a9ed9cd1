#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 audio-visual hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload av_train|logmel|<model>] [--impl reference]

One "step" = one pass of the hot path over one per-GPU batch of synthetic GLips-shaped clips:
  av_train : log-mel frontend -> MidFusionFast forward -> CE -> backward -> (allreduce) -> Adam
             (audio_video/train.py:61-67 + audio_video/data_utils/dataset_av.py:58-71)
  logmel   : the log-mel frontend alone (audio/utils/audio_processor.py:48-64 + crop)
  early_fusion_mobilenet | early_fusion_resnet | video_resnet_lstm | audio_resnet | acv_late_fusion_mobile :
             the same train step for the other configs of BASELINE.json (not the headline line)
`value` is device-timed with inputs resident in HBM; `e2e` goes through the public API from pinned HOST
buffers with the H2D copies and a D2H read of the result inside the timed region.
`--impl reference` times the reference's CPU implementation (oracle port: the reference's own
torch/torchaudio call sequence) on the host cores, same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LOGMEL_BYTES_PER_CLIP = 20000 * 4 + 80 * 117 * 4      # 117 440 algorithmic bytes


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained"),
                "src": "measured"}
    return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own call sequence on the host cores (oracle "port").
# ------------------------------------------------------------------------------------------------
def cpu_reference(workload, cfg, steps, warmup, budget_s=25.0):
    """Returns (value, unit, ms_per_step, sample description, cores)."""
    import torch
    from multimodal_lipread_b200 import synthetic
    from oracle.frontend import AudioProcessorPort, lips_u8_to_model_input
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use all the host threads it can
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    torch.set_num_threads(max(torch.get_num_threads(), avail))
    cores = torch.get_num_threads()
    ap = AudioProcessorPort()
    if workload == "logmel":
        n = 64
        wav = synthetic.make_waveforms(n, seed=1)
        def step():
            ap.batch_frontend_loop(wav)
        unit_per_step, scale = n, LOGMEL_BYTES_PER_CLIP / 1e9
        sample = f"{n} clips per step, per-clip AudioProcessor loop (dataset_av.py:58-66 semantics)"
    else:
        from oracle import av_models as O
        kind = cfg.get("model", "mid_fusion_fast")
        n = cfg["cpu_batch"]
        C = cfg["num_classes"]
        torch.manual_seed(0)
        lr, wd = 3e-4, 0.0
        if kind == "mid_fusion_fast":
            model, names = O.MidFusionFastOracle(C), ("audio", "video")
        elif kind == "early_fusion_mobilenet":
            model, names = O.EarlyFusionMobileNetOracle(C), ("audio", "video")
        elif kind == "early_fusion_resnet":
            model, names = O.EarlyFusionResNetOracle(C), ("audio", "video")
        elif kind == "video_resnet_lstm":
            model, names, lr, wd = O.ResNet2DBiLSTMOracle(C), ("video",), 5e-5, 1e-5
        elif kind == "audio_resnet":
            model, names, lr, wd = O.AudioResNetOracle(C), ("audio",), 5e-4, 1e-4
        else:
            model, names, lr = O.LateFusionMobileOracle(C), ("audio", "cue", "video"), 1e-5
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
        wav = synthetic.make_waveforms(n, seed=1)
        lips = synthetic.make_lips_u8(n, size=cfg["size"], grayscale=cfg["grayscale"])
        cue = synthetic.make_cues(n)
        labels = synthetic.make_labels(n, C)
        def step():
            feed = {"cue": cue}
            if "audio" in names:
                feed["audio"] = ap.batch_frontend_loop(wav)
            if "video" in names:
                feed["video"] = lips_u8_to_model_input(lips)
            O.train_step_generic(model, opt, tuple(feed[k] for k in names), labels)
        unit_per_step, scale = n, 1.0
        sample = (f"{n} clips per step: per-clip log-mel + {type(model).__name__} fwd/CE/bwd/Adam in fp32 torch CPU "
                  f"(the reference's train loop body), lips {cfg['size']}x{cfg['size']}")
    t_budget = time.perf_counter()
    for _ in range(warmup):
        step()
        if time.perf_counter() - t_budget > budget_s:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_budget > 2 * budget_s:
            break
    ms = 1e3 * sum(times) / len(times)
    return unit_per_step / (ms / 1e3) * scale, ms, f"{sample}; {len(times)} timed steps", cores


# ------------------------------------------------------------------------------------------------
def main():
    # NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION/INFO; stdout must carry ONE JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "WARN"
    # ... and anything else a library writes to fd 1 goes to stderr: the JSON line is written to the saved fd
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--gpus", type=int, default=1)
    ap_.add_argument("--steps", type=int, default=20)
    ap_.add_argument("--warmup", type=int, default=5)
    ap_.add_argument("--impl", default="b200", choices=["b200", "reference"])
    TRAIN_MODELS = ["early_fusion_mobilenet", "early_fusion_resnet", "video_resnet_lstm", "audio_resnet",
                    "acv_late_fusion_mobile"]
    ap_.add_argument("--workload", default=None, choices=["av_train", "logmel"] + TRAIN_MODELS)
    ap_.add_argument("--batch", type=int, default=None, help="clips per GPU per step")
    ap_.add_argument("--size", type=int, default=88, help="lip frame height = width (88 benchmark, 44 reference)")
    ap_.add_argument("--classes", type=int, default=40)
    ap_.add_argument("--no-cpu-baseline", action="store_true")
    ap_.add_argument("--dump-ops", default=None, help="write the per-op device times of one step (JSON) to this file")
    args = ap_.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    have_train = os.path.exists(os.path.join(ROOT, "multimodal_lipread_b200", "engine.py"))
    workload = args.workload or ("av_train" if have_train else "logmel")
    if workload == "logmel":
        metric, unit = "logmel_frontend_throughput", "GB/s"
        batch = args.batch or 16384
        config = {"workload": "log-mel frontend, 20000-sample 16 kHz clips -> (80,117) normalised log-mel",
                  "clips_per_gpu_per_step": batch, "algorithmic_bytes_per_clip": LOGMEL_BYTES_PER_CLIP,
                  "l2": "inputs larger than L2 (batch * 80 kB >> 126 MB)"}
    else:
        kind = "mid_fusion_fast" if workload == "av_train" else workload
        names = {"mid_fusion_fast": "audio_video middle_fusion_fast", "early_fusion_mobilenet": "audio_video early_fusion_mobilenet",
                 "early_fusion_resnet": "audio_video early_fusion_resnet", "video_resnet_lstm": "video resnet_lstm",
                 "audio_resnet": "audio resnet", "acv_late_fusion_mobile": "audio_cues_video late_fusion_mobile"}
        metric = "av_midfusion_train_clips_per_sec" if workload == "av_train" else f"{workload}_train_clips_per_sec"
        unit = "clips/s"
        batch = args.batch or 32
        config = {"workload": f"{names[kind]} train step, GLips_{args.classes} shape "
                              f"(29x{args.size}x{args.size} lips, 1.25 s 16 kHz audio)",
                  "batch_per_gpu": batch, "global_batch": batch * world, "num_classes": args.classes,
                  "lip_size": args.size, "grayscale_replicated": True, "parallelism": f"dp{world}",
                  "l2": "ring of input batches larger than L2",
                  "e2e_pipeline": "H2D of step i+1 (copy stream) overlaps step i; loss read back and host-synced every step"}
    cfg = {"num_classes": args.classes, "size": args.size, "grayscale": True, "cpu_batch": min(batch, 32),
           "model": "mid_fusion_fast" if workload in ("av_train", "logmel") else workload}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        v, ms, sample, cores = cpu_reference(workload, cfg, args.steps, args.warmup)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from multimodal_lipread_b200 import _lib, synthetic
    from multimodal_lipread_b200.audio_processor import AudioProcessor
    peaks = _peaks()

    if workload == "logmel":
        from bench_workloads import LogmelWorkload as W
    else:
        from bench_workloads import AvTrainWorkload as W
    wl = W(dev, batch, cfg, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # device-resident timing
    for _ in range(max(args.warmup, 3)):
        wl.step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    wl.reset_kernel_timer()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        wl.step_device()
    e1.record()
    barrier()
    launches = (wl.launches_per_step() or 0) * args.steps or (_lib.launch_count() - l0)
    ms_total = e0.elapsed_time(e1)
    kernel_ms = wl.kernel_ms()                      # dominant-kernel time per launch (CUDA events), or None
    # end to end from pinned host memory
    for _ in range(3):
        wl.step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        wl.step_e2e()
    f1.record()
    barrier()
    ms_e2e_total = f0.elapsed_time(f1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total, ms_e2e_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e_total = t.tolist()

    if rank == 0:
        ms_step = ms_total / args.steps
        units = wl.units_per_step() * world            # clips (av_train) or GB (logmel) per step, all ranks
        value = units / (ms_step / 1e3)
        e2e_v = units / (ms_e2e_total / args.steps / 1e3)
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": config,
                "e2e": {"value": e2e_v, "unit": unit, "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": wl.d2h_bytes},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": wl.roofline(kernel_ms, ms_step, peaks)}
        line.update(wl.extra())
        if world == 1 and hasattr(wl, "files_e2e"):
            fe = wl.files_e2e()
            if fe:
                line["files_e2e"] = fe
        if args.dump_ops and getattr(wl, "op_rows", None):
            with open(args.dump_ops, "w") as f:
                json.dump(wl.op_rows, f, indent=0)
        if world == 1 and not args.no_cpu_baseline:
            v, ms, sample, cores = cpu_reference(workload, cfg, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample}
        emit(line)
    if world > 1:
        # Orderly multi-rank exit.  The step graphs hold captured NCCL kernels: they are released BEFORE the
        # communicator goes away (destroying the process group under live graphs, or leaving both to interpreter
        # teardown, hung the job after the JSON line had been printed), then every rank leaves through os._exit once
        # all of them are past the last collective.
        import gc
        wl.model._graphs.clear() if hasattr(wl, "model") else None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
