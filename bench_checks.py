"""Checker legs of bench.py -- NOT timed workloads, NOT product code.

  dp_parity          (N > 1)  the data-parallel step on the hardware against the DP oracle of SURVEY.md 8(e)
  torch_gpu_baseline (opt-in) the oracle port stepped by eager torch on the GPU, for information

These are the only functions outside tests/, __graft_entry__.smoke() and bench.py's CPU arm that import oracle/, and
they use it as the checker / as a foreign baseline: nothing here feeds `value`, `e2e` or `roofline`.
"""
import torch

from multimodal_lipread_b200 import synthetic


def dp_parity(dev, rank, world, clips_per_rank=2, T=8, size=44, C=40):
    """SURVEY.md 8(e) DP oracle on the hardware path: every rank runs ONE data-parallel train step of MidFusionFast
    (strict fp32 kernels, so that the collective -- not tensor-core rounding -- is what is compared) on its own shard,
    first eagerly and then as the captured graph with the NCCL allreduce inside it; rank 0 then runs the ORACLE model
    (CPU, same seeded weights) on every shard separately and averages the per-shard gradients.  Reported: the worst
    per-tensor norm-wise deviation of allreduce(sum)/world from that mean, and of the post-Adam weights from torch's
    Adam applied to the averaged gradient.  The oracle is the checker here, never the thing measured."""
    import torch.distributed as dist
    from multimodal_lipread_b200.audio_video_models import MidFusionFast
    B = clips_per_rank
    wav = synthetic.make_waveforms(B, seed=9100 + rank, pad_fraction=0.5).to(dev)
    lips = synthetic.make_lips_u8(B, size=size, seed=9200 + rank)[:, :T].contiguous().to(dev)
    labels = synthetic.make_labels(B, C, seed=9300 + rank).to(dev)

    def allreduce(grad):
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)
    out = {}
    res = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(0)
        model = MidFusionFast(C, precision="fp32").to(dev).train()
        model.configure_optimizer(lr=3e-4)
        w0 = model._ensure_flat(dev).flat.clone()
        if mode == "graph":
            # step 0 of a plan is always eager (warm-up): take it with lr = 0, restore the optimizer state, then the
            # compared step is a graph replay with the collective captured inside
            model.set_lr(0.0)
            model.train_step(wav, lips, labels, grad_allreduce=allreduce, world=world, use_graph=True)
            model.configure_optimizer(lr=3e-4)
            model._flat.flat.copy_(w0)
        model.train_step(wav, lips, labels, grad_allreduce=allreduce, world=world, use_graph=(mode == "graph"))
        torch.cuda.synchronize()
        res[mode] = (model._flat.grad.clone() / world, model._flat.flat.clone(), w0, model)
    gathered = {}
    for name, t in (("wav", wav), ("lips", lips), ("labels", labels)):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        gathered[name] = [p.cpu() for p in parts]
    if rank != 0:
        for v in res.values():
            v[3]._graphs.clear()
        return None
    from oracle.av_models import MidFusionFastOracle
    from oracle.frontend import AudioProcessorPort, lips_u8_to_model_input
    torch.manual_seed(0)
    ref = MidFusionFastOracle(C).train()
    ap = AudioProcessorPort()
    grads = None
    for r in range(world):
        ref.zero_grad()
        mel = ap.batch_frontend_loop(gathered["wav"][r])
        loss = torch.nn.functional.cross_entropy(ref(mel, lips_u8_to_model_input(gathered["lips"][r])), gathered["labels"][r])
        loss.backward()
        g = [p.grad.detach().clone() for p in ref.parameters()]
        grads = g if grads is None else [a + b for a, b in zip(grads, g)]
    grads = [g / world for g in grads]
    floor = 1e-7 / 3e-3
    for mode, (gavg, w1, w0, model) in res.items():
        flat = model._flat
        worst, worst_name = 0.0, None
        for (n, p), off, gr in zip(model.named_parameters(), flat.offsets, grads):
            mine = gavg[off:off + p.numel()].view(p.shape).cpu().double()
            e = (mine - gr.double()).abs().max().item() / (gr.double().abs().max().item() + floor)
            if e > worst:
                worst, worst_name = e, n
        # torch's Adam applied to OUR averaged gradient must land on our new weights (1/world folded into lr_adam_step)
        ps = [w0[off:off + p.numel()].view(p.shape).cpu().clone().requires_grad_(True) for p, off in zip(flat.params, flat.offsets)]
        chk = torch.optim.Adam(ps, lr=3e-4)
        for t, p, off in zip(ps, flat.params, flat.offsets):
            t.grad = gavg[off:off + p.numel()].view(p.shape).cpu().clone()
        chk.step()
        adam = max((t.detach() - w1[off:off + p.numel()].view(p.shape).cpu()).abs().max().item()
                   for t, p, off in zip(ps, flat.params, flat.offsets))
        out[mode] = {"grad_max_rel": worst, "worst_tensor": worst_name, "adam_weights_max_abs": adam}
        model._graphs.clear()
    out["max_rel"] = max(v["grad_max_rel"] for v in out.values() if isinstance(v, dict))
    out["what"] = (f"MidFusionFast fp32 kernels, {B} clips/rank x {world} ranks, T={T}, {size}px: allreduce(sum)/world vs the "
                   "mean of per-shard oracle gradients (per-tensor max|d| / max|ref|); eager collective and in-graph collective")
    return out


def torch_gpu_baseline(dev, cfg, batch, steps=10):
    """INFORMATIONAL ("the kernel to beat is whatever cuDNN/cuBLAS PyTorch dispatches for the same
    nn.Modules on the same box"): the oracle port of the model moved to the GPU and stepped with eager torch (cuDNN /
    cuBLAS), in fp32 and under bf16 autocast, on device-resident inputs.  Not part of the product path."""
    from oracle import av_models as O
    from oracle.frontend import AudioProcessorPort
    kind = cfg.get("model", "mid_fusion_fast")
    C = cfg["num_classes"]
    ctor = {"mid_fusion_fast": O.MidFusionFastOracle, "early_fusion_mobilenet": O.EarlyFusionMobileNetOracle,
            "early_fusion_resnet": O.EarlyFusionResNetOracle, "video_resnet_lstm": O.ResNet2DBiLSTMOracle,
            "audio_resnet": O.AudioResNetOracle, "acv_late_fusion_mobile": O.LateFusionMobileOracle}[kind]
    names = {"video_resnet_lstm": ("video",), "audio_resnet": ("audio",),
             "acv_late_fusion_mobile": ("audio", "cue", "video")}.get(kind, ("audio", "video"))
    out = {}
    for mode in ("fp32", "bf16_autocast"):
        torch.manual_seed(0)
        model = ctor(C).to(dev).train()
        opt = torch.optim.Adam(model.parameters(), lr=3e-4)
        ap = AudioProcessorPort()
        ap.mel_transform = ap.mel_transform.to(dev)
        wav = synthetic.make_waveforms(batch, seed=1).to(dev)
        lips = synthetic.make_lips_u8(batch, size=cfg["size"], grayscale=cfg["grayscale"]).to(dev)
        cue = synthetic.make_cues(batch).to(dev)
        labels = synthetic.make_labels(batch, C).to(dev)

        def step():
            feed = {"cue": cue}
            if "audio" in names:
                feed["audio"] = ap.batch_frontend_batched(wav)
            if "video" in names:
                feed["video"] = (lips.float() / 255.0).permute(0, 4, 1, 2, 3).contiguous()
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
                logits = model(*[feed[k] for k in names])
                loss = torch.nn.functional.cross_entropy(logits.float(), labels)
            loss.backward()
            opt.step()
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        out[mode] = {"ms_per_step": ms, "clips_per_sec": batch / (ms / 1e3)}
        del model, opt
    out["what"] = "oracle port .cuda(), eager torch (cuDNN/cuBLAS/cuFFT), batched log-mel, device-resident inputs, no host sync per step"
    return out
