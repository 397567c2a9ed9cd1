#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 audio-visual hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload av_train|logmel|<model>] [--impl reference]

One "step" = one pass of the hot path over one per-GPU batch of synthetic GLips-shaped clips:
  av_train : log-mel frontend -> MidFusionFast forward -> CE -> backward -> (allreduce) -> Adam
             (audio_video/train.py:61-67 + audio_video/data_utils/dataset_av.py:58-71)          [the headline line]
  logmel   : the log-mel frontend alone (audio/utils/audio_processor.py:48-64 + crop)
  early_fusion_mobilenet | early_fusion_resnet | video_resnet_lstm | audio_resnet | acv_late_fusion_mobile :
             the same train step for the other configs of BASELINE.json
`value` is device-timed with inputs resident in HBM; `e2e` goes through the public API from pinned HOST
buffers with the H2D copies and a D2H read of the result inside the timed region.

The DEFAULT invocation (no --workload) prints the headline line and carries, in the same JSON object, the rest of
BASELINE.json's metric so that the driver sees it: `logmel` (GB/s, fraction of the HBM peak, its own cpu_baseline),
`configs` (BASELINE.json configs 1, 2, 4, 5: clips/s, ms/step [, cpu_baseline at N = 1]), `value_tf32` / `value_fp32` (the
same step on fp32 storage with TF32 tensor-core products / with every GEMM in strict fp32), and at N > 1 `dp_parity` (allreduced gradient against the mean of per-shard oracle
gradients) and `comm` (exposed collective time per step).

`--impl reference` times the reference's CPU implementation on the host cores, same metric/config: the UNMODIFIED
reference modules staged in oracle/_ref (its own train_epoch and AudioProcessor; `kind: "reference"`), or the oracle
port of the same call sequence when the staged copy is absent (`kind: "port"`).
"""
import argparse
import importlib.util
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

TRAIN_MODELS = ["early_fusion_mobilenet", "early_fusion_resnet", "video_resnet_lstm", "audio_resnet",
                "acv_late_fusion_mobile"]
# BASELINE.json `configs` (0-based list there; numbered 1..5 in SURVEY.md): 3 is the headline workload itself
CONFIG_WORKLOADS = {"1": "audio_resnet", "2": "video_resnet_lstm", "4": "early_fusion_mobilenet",
                    "5": "acv_late_fusion_mobile"}
NAMES = {"mid_fusion_fast": "audio_video middle_fusion_fast", "early_fusion_mobilenet": "audio_video early_fusion_mobilenet",
         "early_fusion_resnet": "audio_video early_fusion_resnet", "video_resnet_lstm": "video resnet_lstm",
         "audio_resnet": "audio resnet", "acv_late_fusion_mobile": "audio_cues_video late_fusion_mobile"}
# clips per CPU-baseline step: a BOUNDED sample of the workload (10-30 s of host work per config)
CPU_BATCH = {"mid_fusion_fast": 32, "early_fusion_mobilenet": 32, "audio_resnet": 32, "video_resnet_lstm": 8,
             "early_fusion_resnet": 8, "acv_late_fusion_mobile": 8}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained"),
                "src": "measured"}
    return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


def _load_synthetic():
    """multimodal_lipread_b200/synthetic.py loaded BY PATH: importing the package dlopens liblipread_b200.so, which
    the CPU arm must not map."""
    spec = importlib.util.spec_from_file_location("_lr_synthetic", os.path.join(ROOT, "multimodal_lipread_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


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
# CPU arm: the reference's own implementation of the path on the host cores.
# ------------------------------------------------------------------------------------------------
def cpu_reference(workload, cfg, steps, warmup, budget_s=25.0):
    """Returns (value, ms_per_step, sample description, cores, kind).  kind "reference": the unmodified reference
    modules staged in oracle/_ref (AudioProcessor per clip as dataset_av.py:58-66 does, then ITS train_epoch on the
    batch); kind "port": the oracle's restatement of the same call sequence, when nothing is staged."""
    import torch
    synthetic = _load_synthetic()
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use all the host threads it can
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    torch.set_num_threads(max(torch.get_num_threads(), avail))
    cores = torch.get_num_threads()
    from oracle import ref_loader
    staged = ref_loader.ref_root() is not None
    kind = "reference" if staged else "port"
    if staged:
        rap = ref_loader.audio_processor()

        def frontend(wav):
            # exactly audio_video/data_utils/dataset_av.py:58-66, one clip at a time, then the default collate
            return torch.stack([rap.normalize_spectrogram(rap.compute_melspectrogram(w))[:80, :117].float() for w in wav])
    else:
        from oracle.frontend import AudioProcessorPort
        frontend = AudioProcessorPort().batch_frontend_loop

    def lips_to_input(lips):                        # dataset_av.py:70-71 + default collate
        return (lips.to(torch.float32) / 255.0).permute(0, 4, 1, 2, 3).contiguous()

    if workload == "logmel":
        n = 64
        wav = synthetic.make_waveforms(n, seed=1)

        def step():
            frontend(wav)
        unit_per_step, scale = n, LOGMEL_BYTES_PER_CLIP / 1e9
        sample = (f"{n} clips per step, per-clip AudioProcessor loop (dataset_av.py:58-66 semantics), "
                  f"{'reference module' if staged else 'oracle port'}")
    else:
        model_kind = cfg.get("model", "mid_fusion_fast")
        n = cfg["cpu_batch"]
        C = cfg["num_classes"]
        lr, wd = {"video_resnet_lstm": (5e-5, 1e-5), "audio_resnet": (5e-4, 1e-4),
                  "acv_late_fusion_mobile": (1e-5, 0.0)}.get(model_kind, (3e-4, 0.0))
        names = {"video_resnet_lstm": ("video",), "audio_resnet": ("audio",),
                 "acv_late_fusion_mobile": ("audio", "cue", "video")}.get(model_kind, ("audio", "video"))
        torch.manual_seed(0)
        if staged:
            _, make_model, run_batch, how = ref_loader.train_loop(model_kind)
            model = make_model(C)
        else:
            from oracle import av_models as O
            model = {"mid_fusion_fast": O.MidFusionFastOracle, "early_fusion_mobilenet": O.EarlyFusionMobileNetOracle,
                     "early_fusion_resnet": O.EarlyFusionResNetOracle, "video_resnet_lstm": O.ResNet2DBiLSTMOracle,
                     "audio_resnet": O.AudioResNetOracle, "acv_late_fusion_mobile": O.LateFusionMobileOracle}[model_kind](C)
            how = "oracle port of the train loop body"

            def run_batch(model, opt, batch):
                *inputs, labels = (batch["lip_regions"], batch["label"]) if isinstance(batch, dict) else batch
                O.train_step_generic(model, opt, tuple(inputs), labels)
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
        wav = synthetic.make_waveforms(n, seed=1)
        lips = synthetic.make_lips_u8(n, size=cfg["size"], grayscale=cfg["grayscale"])
        cue = synthetic.make_cues(n)
        labels = synthetic.make_labels(n, C)

        def step():
            feed = {"cue": cue}
            if "audio" in names:
                feed["audio"] = frontend(wav)
            if "video" in names:
                feed["video"] = lips_to_input(lips)
            if names == ("video",):
                batch = {"lip_regions": feed["video"], "label": labels}      # video/data_utils/dataset_loader.py:98-101
            else:
                batch = tuple(feed[k] for k in names) + (labels,)
            run_batch(model, opt, batch)
        unit_per_step, scale = n, 1.0
        sample = (f"{n} clips per step: per-clip log-mel + {type(model).__name__} fwd/CE/bwd/Adam in fp32 torch CPU "
                  f"({how}), lips {cfg['size']}x{cfg['size']}")
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
    return unit_per_step / (ms / 1e3) * scale, ms, f"{sample}; {len(times)} timed steps", cores, kind


def _cpu_baseline_record(workload, cfg, unit, steps=3, warmup=1, budget_s=25.0):
    v, ms, sample, cores, kind = cpu_reference(workload, cfg, steps=steps, warmup=warmup, budget_s=budget_s)
    return {"value": v, "unit": unit, "cores": cores, "kind": kind, "sample": sample, "ms_per_step": ms}


def _workload_config(workload, args, batch, world):
    if workload == "logmel":
        return {"workload": "log-mel frontend, 20000-sample 16 kHz clips -> (80,117) normalised log-mel",
                "clips_per_gpu_per_step": batch, "algorithmic_bytes_per_clip": LOGMEL_BYTES_PER_CLIP,
                "l2": "inputs larger than L2 (batch * 80 kB >> 126 MB)"}
    kind = "mid_fusion_fast" if workload == "av_train" else workload
    return {"workload": f"{NAMES[kind]} train step, GLips_{args.classes} shape "
                        f"(29x{args.size}x{args.size} lips, 1.25 s 16 kHz audio)",
            "batch_per_gpu": batch, "global_batch": batch * world, "num_classes": args.classes,
            "lip_size": args.size, "grayscale_replicated": True, "parallelism": f"dp{world}",
            "l2": "ring of input batches larger than L2",
            "e2e_pipeline": "H2D of step i+1 (copy stream) overlaps step i; loss read back and host-synced every step"}


# ------------------------------------------------------------------------------------------------
def main():
    # anything a library writes to fd 1 (e.g. NCCL_DEBUG=INFO banners) goes to stderr: the JSON line is written to the
    # saved fd.  NCCL_DEBUG itself is left as the driver set it.
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
    ap_.add_argument("--workload", default=None, choices=["av_train", "logmel"] + TRAIN_MODELS)
    ap_.add_argument("--batch", type=int, default=None, help="clips per GPU per step")
    ap_.add_argument("--size", type=int, default=88, help="lip frame height = width (88 benchmark, 44 reference)")
    ap_.add_argument("--classes", type=int, default=40)
    ap_.add_argument("--precision", default="bf16", choices=["tf32", "fp32", "bf16"],
                     help="bf16 (default; the north star's precision): bf16 activation storage + tcgen05 kind::f16; "
                          "tf32: fp32 storage + tcgen05 kind::tf32; fp32: strict fp32 SIMT everywhere")
    ap_.add_argument("--no-cpu-baseline", action="store_true")
    ap_.add_argument("--no-sub-records", action="store_true", help="headline line only (no logmel / configs / value_fp32)")
    ap_.add_argument("--dump-ops", default=None, help="write the per-op device times of one step (JSON) to this file")
    ap_.add_argument("--torch-gpu-baseline", action="store_true",
                     help="informational: also time the oracle port on the GPU (eager torch + cuDNN, fp32 and bf16 autocast)")
    args = ap_.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    default_run = args.workload is None
    workload = args.workload or "av_train"
    if workload == "logmel":
        metric, unit = "logmel_frontend_throughput", "GB/s"
        batch = args.batch or 16384
    else:
        metric = "av_midfusion_train_clips_per_sec" if workload == "av_train" else f"{workload}_train_clips_per_sec"
        unit = "clips/s"
        batch = args.batch or 32
    config = _workload_config(workload, args, batch, world)

    def cfg_for(wl_name, b):
        model = "mid_fusion_fast" if wl_name in ("av_train", "logmel") else wl_name
        return {"num_classes": 8 if model == "audio_resnet" and default_run else args.classes, "size": args.size,
                "grayscale": True, "cpu_batch": min(b, CPU_BATCH.get(model, 32)), "model": model, "precision": args.precision}
    cfg = cfg_for(workload, batch)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        v, ms, sample, cores, kind = cpu_reference(workload, cfg, args.steps, args.warmup)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
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
    from multimodal_lipread_b200 import _lib
    import bench_workloads as BW
    import bench_checks
    peaks = _peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_workload(wl, steps, warmup, with_e2e=True, sampler=None):
        """W untimed + K timed steps bracketed by barrier + synchronize, CUDA events, max over ranks."""
        for _ in range(max(warmup, 3)):
            wl.step_device()
        barrier()
        if sampler is not None:
            sampler.start()
        l0 = _lib.launch_count()
        wl.reset_kernel_timer()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            wl.step_device()
        e1.record()
        barrier()
        launches = (wl.launches_per_step() or 0) * steps or (_lib.launch_count() - l0)
        ms_total = e0.elapsed_time(e1)
        kernel_ms = wl.kernel_ms()
        ms_e2e_total = 0.0
        if with_e2e:
            for _ in range(3):
                wl.step_e2e()
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(steps):
                wl.step_e2e()
            f1.record()
            barrier()
            ms_e2e_total = f0.elapsed_time(f1)
        t = torch.tensor([ms_total, ms_e2e_total], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e_total = t.tolist()
        return ms_total / steps, (ms_e2e_total / steps if with_e2e else None), launches, kernel_ms

    W = BW.LogmelWorkload if workload == "logmel" else BW.AvTrainWorkload
    wl = W(dev, batch, cfg, rank, world)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step, ms_e2e, launches, kernel_ms = time_workload(wl, args.steps, args.warmup, sampler=sampler)
    clocks = sampler.stop() if rank == 0 else None

    line = None
    if rank == 0:
        units = wl.units_per_step() * world            # clips (train) or GB (logmel) per step, all ranks
        line = {"metric": metric, "value": units / (ms_step / 1e3), "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": config,
                "e2e": {"value": units / (ms_e2e / 1e3), "unit": unit, "h2d_bytes_per_step": wl.h2d_bytes,
                        "d2h_bytes_per_step": wl.d2h_bytes},
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

    # ---- exposed collective time: the same step with the allreduce left out of the graph (replicas drift apart from
    # here on, so this is measured after the headline numbers and on throw-away state)
    if world > 1 and workload != "logmel":
        ms_nocomm = wl.time_without_allreduce(args.steps, barrier)
        t = torch.tensor([ms_nocomm], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            line["comm"] = {"collective": "ncclAllReduce(sum) of the flat fp32 gradient in buckets on a communication branch of "
                                          "the step graph, each issued when the last backward op that writes it has been issued",
                            "payload_bytes": wl.model._flat.grad.numel() * 4, "buckets": wl.model.allreduce_buckets(),
                            "ms_per_step_without_collective": t.item(),
                            "exposed_ms_per_step": max(0.0, ms_step - t.item())}
    wl.release()
    del wl

    # ---- the rest of BASELINE.json's metric, in the default invocation
    if default_run and not args.no_sub_records:
        sub_steps = min(args.steps, 10)
        # (1) the same step in the wider arithmetic types, beside the headline number
        notes = {"fp32": "every GEMM on the fp32 SIMT kernel (precision='fp32'): the reference's arithmetic type",
                 "tf32": "fp32 storage, tcgen05 kind::tf32 products (precision='tf32': round 1's benchmarked mode)",
                 "bf16": "bf16 activation storage, tcgen05 kind::f16 (precision='bf16')"}
        for other in ("tf32", "fp32", "bf16"):
            if other == args.precision:
                continue
            wo = BW.AvTrainWorkload(dev, batch, dict(cfg, precision=other), rank, world)
            ms_o, _, _, _ = time_workload(wo, sub_steps, 3, with_e2e=False)
            if rank == 0:
                line["value_" + other] = {"value": batch * world / (ms_o / 1e3), "unit": unit, "ms_per_step": ms_o, "note": notes[other]}
            wo.release()
            del wo
        # (2) log-mel frontend
        lb = 16384
        lcfg = cfg_for("logmel", lb)
        wlm = BW.LogmelWorkload(dev, lb, lcfg, rank, world)
        ms_l, ms_le, _, k_ms = time_workload(wlm, sub_steps, 3)
        if rank == 0:
            gb = wlm.units_per_step() * world
            rec = {"metric": "logmel_frontend_throughput", "value": gb / (ms_l / 1e3), "unit": "GB/s",
                   "clips_per_sec": lb * world / (ms_l / 1e3), "ms_per_step": ms_l, "dtype": wlm.dtype,
                   "config": _workload_config("logmel", args, lb, world),
                   "e2e": {"value": gb / (ms_le / 1e3), "unit": "GB/s", "h2d_bytes_per_step": wlm.h2d_bytes,
                           "d2h_bytes_per_step": wlm.d2h_bytes},
                   "roofline": wlm.roofline(k_ms, ms_l, peaks)}
            if world == 1 and not args.no_cpu_baseline:
                rec["cpu_baseline"] = _cpu_baseline_record("logmel", lcfg, "GB/s", budget_s=8.0)
            line["logmel"] = rec
        wlm.release()
        del wlm
        # (3) BASELINE.json configs 1, 2, 4, 5
        recs = {}
        for key, name in CONFIG_WORKLOADS.items():
            ccfg = cfg_for(name, batch)
            wc = BW.AvTrainWorkload(dev, batch, ccfg, rank, world)
            ms_c, ms_ce, n_l, _ = time_workload(wc, sub_steps, 3)
            if rank == 0:
                rec = {"workload": _workload_config(name, args, batch, world)["workload"].replace(
                           f"GLips_{args.classes}", f"GLips_{ccfg['num_classes']}"),
                       "metric": f"{name}_train_clips_per_sec", "value": batch * world / (ms_c / 1e3), "unit": "clips/s",
                       "ms_per_step": ms_c, "e2e": batch * world / (ms_ce / 1e3), "dtype": wc.dtype,
                       "batch_per_gpu": batch, "n_gpus": world, "gpu_launches_per_step": wc.launches_per_step(),
                       "model_tflops": wc.model_tflops(ms_c)}
                if world == 1 and not args.no_cpu_baseline:
                    rec["cpu_baseline"] = _cpu_baseline_record(name, ccfg, "clips/s", steps=2, warmup=1, budget_s=15.0)
                recs[key] = rec
            wc.release()
            del wc
        if rank == 0:
            line["configs"] = recs
    if world > 1 and workload != "logmel":
        dp = bench_checks.dp_parity(dev, rank, world)
        if rank == 0:
            line["dp_parity"] = dp
    if args.torch_gpu_baseline and rank == 0 and workload != "logmel":
        line["torch_gpu_baseline"] = bench_checks.torch_gpu_baseline(dev, cfg, batch)

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = _cpu_baseline_record(workload, cfg, unit)
            line["cpu_baseline"].pop("ms_per_step", None)
        emit(line)
    if world > 1:
        # Orderly multi-rank exit.  The step graphs hold captured NCCL kernels: they are released BEFORE the
        # communicator goes away (destroying the process group under live graphs, or leaving both to interpreter
        # teardown, hung the job after the JSON line had been printed), then every rank leaves through os._exit once
        # all of them are past the last collective.
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
