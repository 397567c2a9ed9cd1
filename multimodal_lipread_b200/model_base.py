"""Shared plumbing of every lipread_b200 model: flat parameters, launch plans keyed by input shape, the autograd
bridge behind the reference's `model(*inputs) -> logits` call, and the fused train step
(zero_grad -> forward -> CrossEntropyLoss -> backward -> [allreduce] -> Adam: audio_video/train.py:61-67,
video/train.py:93-104, audio/train.py:67-78, audio_cues_video/train.py:60-72) replayed as a CUDA graph.

A concrete model keeps the reference's sub-modules as parameter containers (so seeded initialisation and
`state_dict` keys are interchangeable with the reference) and provides a plan class whose constructor writes the
launch lists.  The torch forward() of the sub-modules is never called; there is no CPU path."""
import os

import torch
import torch.nn as nn

from . import _lib, engine
from ._lib import lib

N_MELS, N_FRAMES_OUT, N_SAMPLES = 80, 117, 20000


class Cfg:
    """Stand-in for the reference's Config objects: anything with .get("dotted.key", default)."""

    def __init__(self, values=None):
        self.values = values or {}

    def get(self, key, default=None):
        if key in self.values:
            return self.values[key]
        cur = self.values
        for part in key.split("."):
            if not isinstance(cur, dict) or part not in cur:
                return default
            cur = cur[part]
        return cur


def load_torchvision_weights(module, state_dict):
    """Copy a torchvision checkpoint (e.g. IMAGENET1K_V1 of resnet18 / mobilenet_v2 / mobilenet_v3_small, torchvision key
    names) into `module`, whatever prefix the trunk sits under: every entry of module.state_dict() takes the
    checkpoint tensor whose key equals one of the entry's dotted suffixes (optionally under "features.") and whose
    shape matches.  Layers the reference re-creates after loading (a 1-channel conv1, fc = Identity) simply find
    no match.  Returns the number of tensors copied."""
    own = module.state_dict()
    n = 0
    for key, dst in own.items():
        parts = key.split(".")
        for i in range(len(parts)):
            tail = ".".join(parts[i:])
            for cand in (tail, "features." + tail):
                src = state_dict.get(cand)
                if src is not None and tuple(src.shape) == tuple(dst.shape):
                    dst.copy_(src)
                    n += 1
                    break
            else:
                continue
            break
    return n


def pretrained_weights(config, explicit, what):
    """The ImageNet checkpoint the reference downloads (weights=...IMAGENET1K_V1 / pretrained=True:
    audio_video/models/middle_fusion_fast.py:15, video/models/resnet_lstm.py:80, audio_cues_video/models/*).  There is
    no network here, so it comes from the caller: `explicit` (a state_dict) or the YAML key `model.pretrained_weights`
    (a torch.save'd state_dict).  With neither, the trunk keeps its seeded random init -- a DEVIATION from the
    reference's flow that is announced, not silent (for the frozen-backbone models it means frozen random features)."""
    if explicit is not None:
        return explicit
    path = config.get("model.pretrained_weights", None) if config is not None else None
    if path:
        if not os.path.exists(path):
            raise FileNotFoundError(f"model.pretrained_weights not found: {path}")
        return torch.load(path, map_location="cpu")
    import warnings
    warnings.warn(f"{what}: the reference initialises this trunk from ImageNet (IMAGENET1K_V1); no weights were supplied "
                  "(pretrained_state_dict= or config model.pretrained_weights), so it starts from random init",
                  stacklevel=3)
    return None


def video_layout(video):
    """(kind, B, T, H, W, sb, st, sc, sh, sw), scale of lip frames in the caller's own layout:
    uint8 (B,T,H,W,3) as the .npy files hold them (video/data_utils/dataset_loader.py:87-96) or float32
    (B,3,T,H,W) as the reference's forward() receives them."""
    if video.dtype == torch.uint8:
        if video.dim() != 5 or video.shape[-1] != 3:
            raise ValueError(f"uint8 lip frames must be (B, T, H, W, 3), got {tuple(video.shape)}")
        B, T, H, W, _ = video.shape
        sb, st, sh, sw, sc = video.stride()
        return (1, B, T, H, W, sb, st, sc, sh, sw), 1.0 / 255.0
    if video.dtype != torch.float32:
        raise ValueError(f"lip frames must be uint8 (B,T,H,W,3) or float32 (B,3,T,H,W), got {video.dtype}")
    if video.dim() != 5 or video.shape[1] != 3:
        raise ValueError(f"float lip frames must be (B, 3, T, H, W), got {tuple(video.shape)}")
    B, _, T, H, W = video.shape
    sb, sc, st, sh, sw = video.stride()
    return (0, B, T, H, W, sb, st, sc, sh, sw), 1.0


class ModelPlan(engine.Plan):
    """Launch plan of one model at one input shape.  Sub-classes write the launch lists in build()."""

    def __init__(self, model, flat, spec, device, training, with_backward):
        super().__init__(flat, device, training, with_backward, precision=model.precision)
        self.model, self.spec = model, spec
        self.B = spec["B"]
        self.num_classes = model.num_classes
        self.inputs = {}                               # name -> static device tensor the caller copies into
        self.build(model, spec)
        self._finish()

    # -- static inputs ------------------------------------------------------------------------------------
    def video_input(self):
        kind, B, T, H, W = self.spec["video"]
        if kind == 1:
            v = torch.empty(B, T, H, W, 3, dtype=torch.uint8, device=self.dev)
        else:
            v = torch.empty(B, 3, T, H, W, dtype=torch.float32, device=self.dev)
        self.inputs["video"] = v
        self.bufs.append(v)
        layout, scale = video_layout(v)
        return v, layout, scale

    def audio_input(self):
        """(B,80,117) log-mel buffer; when the plan takes raw waveforms the log-mel kernel fills it first."""
        B = self.B
        mel = torch.empty(B, N_MELS, N_FRAMES_OUT, dtype=torch.float32, device=self.dev)
        self.bufs.append(mel)
        self.mel = mel
        if self.spec.get("from_wav"):
            wav = torch.empty(B, N_SAMPLES, dtype=torch.float32, device=self.dev)
            self.bufs.append(wav)
            self.inputs["audio"] = wav
            self.fwd.add("lr_logmel_fwd", wav, self.model.logmel_plan(self.dev), mel, B, N_FRAMES_OUT, 0)
        else:
            self.inputs["audio"] = mel
        return mel

    def vector_input(self, name, dim):
        v = torch.empty(self.B, dim, dtype=torch.float32, device=self.dev)
        self.inputs[name] = v
        self.bufs.append(v)
        return v

    # -- head --------------------------------------------------------------------------------------------
    def mlp(self, x, dx, B, layers):
        """nn.Sequential of Linear / ReLU / Dropout on a [B, D] buffer -> (out, dout) of the last Linear."""
        cur, dcur, dim = x, dx, None
        mods = list(layers)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                act = engine.ACT_NONE
                if i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU):
                    act = engine.ACT_RELU
                    i += 1
                out = self.alloc(B * m.out_features)
                dout = self.alloc(B * m.out_features) if self.with_backward else None
                self.linear(cur, m.in_features, B, m.weight, m.bias, out, m.out_features, act=act)
                if self.with_backward:
                    g = self.bgroup()
                    if act != engine.ACT_NONE:
                        g.add("lr_act_bwd", dout, out, B * m.out_features, act)
                    self.linear_bwd(g, cur, m.in_features, B, m.weight, m.bias, dout, m.out_features,
                                    dx=dcur, ldx=m.in_features)
                cur, dcur, dim = out, dout, m.out_features
            elif isinstance(m, nn.Dropout):
                if dim is None:
                    raise NotImplementedError("Dropout before the first Linear of a head")
                cur, dcur = self.dropout(cur, dcur, B * dim, m.p)
            elif isinstance(m, (nn.Identity,)):
                pass
            else:
                raise NotImplementedError(f"{type(m).__name__} in a classifier head")
            i += 1
        return cur, dcur

    def linear_bn_act(self, x, dx, B, fc, bn, act):
        """nn.Linear -> nn.BatchNorm1d -> activation on a [B, K] buffer: the GEMM epilogue emits the batch statistics,
        one lr_bn_act pass applies them.  Returns (value, gradient) buffers of the activation output."""
        D, K = fc.out_features, fc.in_features
        raw = engine.T2(self, B, 1, 1, D, h=False)            # per-clip head vectors stay fp32 in every mode
        raw.stat_slot = self.stat_slot(D)
        st = (lambda s=raw.stat_slot: s["fwd"]) if self.training else 0
        self.gemm_auto(self.fwd, x, K, 0, fc.weight, K, 0, raw.val, D, B, D, K, bias=(fc.bias if fc.bias is not None else 0),
                       stats=st)
        if self.with_backward:
            self.linear_bwd(self.bgroup(), x, K, B, fc.weight, fc.bias, raw.grad, D, dx=dx, ldx=K)
        h = engine.T2(self, B, 1, 1, D, h=False)
        self.bn_act(raw, bn, act, h)
        return h.val, h.grad

    def set_logits(self, logits, dlogits):
        self.logits = logits.view(self.B, self.num_classes)
        self.dlogits = dlogits.view(self.B, self.num_classes) if dlogits is not None else None

    def _finish(self):
        B, C = self.B, self.num_classes
        self.labels = torch.zeros(B, dtype=torch.int64, device=self.dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.correct = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.bufs += [self.labels, self.loss, self.correct]
        self.finalize()
        # per-step zeroing: BN statistic arena and (for the backward) the flat gradient
        self.pre = engine.OpList()
        self.pre.add("lr_memset", self.stats, self.stats.numel() * 8)
        if self._uses_shadow:                          # precision "bf16": refresh the bf16 shadow of the weights
            self.pre.add("lr_cast_bf16", self.flat.flat, self.flat.shadow(), self.flat.numel)
        if self.weight_taps:                           # every tap-major bf16 conv operand of the plan, one launch
            rows = [[w.data_ptr(), d.data_ptr(), co | (ci << 32), kk | (mode << 32)] for w, d, co, ci, kk, mode in self.weight_taps]
            self.wtap_table = torch.tensor(rows, dtype=torch.int64, device=self.dev)
            self.bufs.append(self.wtap_table)
            self.pre.add("lr_weight_tap_batch_h", self.wtap_table, len(rows), max(co * ci * kk for _, _, co, ci, kk, _ in self.weight_taps))
        if self.rng_step is not None:
            self.pre.add("lr_rng_tick", self.rng_step)
        self.pre_bwd = engine.OpList()
        if self.with_backward:
            self.pre_bwd.add("lr_memset", self.flat.grad, self.flat.grad.numel() * 4)
        self.ce = engine.OpList()
        self.ce.add("lr_memset", self.loss, 4)
        self.ce.add("lr_memset", self.correct, 4)
        self.ce.add("lr_ce_loss", self.logits, self.labels, self.loss, self.dlogits if self.with_backward else 0,
                    self.correct, B, C, 1.0 / B)

    # -- execution ---------------------------------------------------------------------------------------
    def run_forward(self, stream, forked=None):
        """forked = (main, side) torch streams: ops of an independent branch (OpList.side_branch) run on `side`."""
        self.pre.run(stream)
        if forked is None:
            self.fwd.run(stream)
        else:
            self.fwd.run_forked(*forked)

    def run_backward(self, stream, forked=None, comm_hooks=None):
        """forked = (main, side) torch streams: weight-gradient kernels run on `side` concurrently with the
        dgrad chain (used under CUDA-graph capture, where it becomes a parallel branch of the graph).
        comm_hooks: see OpList.run_forked (bucketed gradient allreduce overlapped with the backward)."""
        self.pre_bwd.run(stream)
        if forked is None:
            self.bwd.run(stream)
        else:
            self.bwd.run_forked(*forked, comm_hooks=comm_hooks)

    def n_launches(self):
        return len(self.fwd) + len(self.bwd) + 1


class _PlanFn(torch.autograd.Function):
    """autograd bridge for the drop-in `model(*inputs)` -> logits call: the backward runs the plan's
    hand-written backward schedule and hands the parameter gradients to autograd."""

    @staticmethod
    def forward(ctx, model, need_backward, n_in, *args):
        inputs, _params = args[:n_in], args[n_in:]
        named = dict(zip(model.INPUTS, inputs))
        plan = model._plan_for(named, training=model.training, with_backward=need_backward)
        for name, t in named.items():
            buf = plan.inputs[name]
            buf.copy_(t.reshape(buf.shape))
        plan.run_forward(torch.cuda.current_stream().cuda_stream)
        ctx.plan, ctx.model, ctx.n_in = plan, model, n_in
        return plan.logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        plan, model = ctx.plan, ctx.model
        if not plan.with_backward:
            raise RuntimeError("forward ran without gradient buffers (torch.no_grad)")
        plan.dlogits.copy_(dlogits)
        plan.run_backward(torch.cuda.current_stream().cuda_stream)
        flat = model._flat
        return (None, None, None) + (None,) * ctx.n_in + tuple(flat.g(p).clone() for p in flat.params)


class PlanModel(nn.Module):
    """Base class: sub-classes set INPUTS (forward argument names, in order), PLAN (a ModelPlan sub-class) and
    build the reference's sub-modules in __init__ (after calling _init_base)."""

    INPUTS = ("audio", "video")
    PLAN = None
    DEFAULT_LR = 3e-4
    DEFAULT_WD = 0.0

    def _init_base(self, num_classes, config, precision):
        self.num_classes = num_classes
        # "tf32": GEMM-shaped work on the tensor cores (tcgen05, TF32 products, fp32 accumulate; >= the bf16 the
        # north star allows); "fp32": every kernel in fp32 SIMT arithmetic (strict parity with the reference);
        # "bf16": the trunk's activations / gradients stored as bfloat16, its GEMMs tcgen05 kind::f16 (fp32 accumulate)
        self.precision = precision or config.get("precision.compute", "tf32")
        self._flat = None
        self._plans = {}
        self._logmel = {}
        self._graphs = {}

    # ------------------------------------------------------------------ plumbing
    def _ensure_flat(self, device):
        if self._flat is None or not self._flat.intact() or self._flat.device != torch.device(device):
            old = self._flat
            self._flat = engine.FlatParams(self, device)
            for b in self.buffers():
                if b.device != self._flat.device:
                    b.data = b.data.to(self._flat.device)
            self._plans.clear()
            self._graphs.clear()
            if old is not None and old.m is not None and old.numel == self._flat.numel and old.device == self._flat.device:
                self._flat.m, self._flat.v, self._flat.adam_state = old.m, old.v, old.adam_state
            if old is not None and old.rng_step is not None and old.device == self._flat.device:
                self._flat.rng_step = old.rng_step
        return self._flat

    def logmel_plan(self, device):
        from .audio_processor import AudioProcessor
        key = str(device)
        if key not in self._logmel:
            self._logmel[key] = AudioProcessor(device=device)
        return self._logmel[key].plan

    def _spec_of(self, named):
        """Hashable description of the input shapes (the plan key)."""
        spec = {}
        dev = None
        for name, t in named.items():
            if t.device.type != "cuda":
                raise _lib.LipreadError("multimodal_lipread_b200 models run on CUDA tensors only (no CPU path)")
            dev = t.device if dev is None else dev
            if t.device != dev:
                raise _lib.LipreadError("all inputs must live on the same CUDA device")
            if name == "video":
                layout, _ = video_layout(t)
                spec["video"] = layout[:5]
                spec["B"] = layout[1]
            elif name == "audio":
                spec["from_wav"] = t.dim() == 2 and t.shape[1] == N_SAMPLES
                if not spec["from_wav"] and t.numel() != t.shape[0] * N_MELS * N_FRAMES_OUT:
                    raise ValueError(f"audio must be (B,80,117) log-mel or (B,20000) waveform, got {tuple(t.shape)}")
                spec.setdefault("B", t.shape[0])
            else:
                spec[name] = tuple(t.shape[1:])
                spec.setdefault("B", t.shape[0])
        return spec, dev

    def _plan_for(self, named, training, with_backward):
        spec, dev = self._spec_of(named)
        flat = self._ensure_flat(dev)
        key = (tuple(sorted((k, v) for k, v in spec.items())), bool(training), bool(with_backward), self.precision)
        plan = self._plans.get(key)
        if plan is None:
            plan = self.PLAN(self, flat, spec, dev, training, with_backward)
            self._plans[key] = plan
        return plan

    # ------------------------------------------------------------------ reference surface
    def forward(self, *inputs):
        if len(inputs) != len(self.INPUTS):
            raise TypeError(f"{type(self).__name__}.forward takes {self.INPUTS}, got {len(inputs)} tensors")
        for t in inputs:
            if t.device.type != "cuda":
                raise _lib.LipreadError("multimodal_lipread_b200 models run on CUDA tensors only (no CPU path)")
        flat = self._ensure_flat(inputs[0].device)
        return _PlanFn.apply(self, torch.is_grad_enabled(), len(inputs), *inputs, *flat.params)

    @torch.no_grad()
    def eval_step(self, *inputs_and_labels):
        """Forward in the module's current mode + CrossEntropyLoss (mean) + number of correct arg-max predictions, all in
        lipread_b200 kernels and without a host sync: the body of the reference's validate()
        (audio_video/train.py:82-88).  Returns (loss, correct, logits) device tensors owned by the plan."""
        *inputs, labels = inputs_and_labels
        named = dict(zip(self.INPUTS, inputs))
        plan = self._plan_for(named, training=self.training, with_backward=False)
        for name, t in named.items():
            buf = plan.inputs[name]
            buf.copy_(t.reshape(buf.shape), non_blocking=True)
        plan.labels.copy_(labels, non_blocking=True)
        s = torch.cuda.current_stream().cuda_stream
        plan.run_forward(s)
        plan.ce.run(s)
        return plan.loss, plan.correct, plan.logits

    # ------------------------------------------------------------------ fused training step
    def configure_optimizer(self, lr=None, betas=(0.9, 0.999), eps=1e-8, weight_decay=None):
        """torch.optim.Adam as the reference's train scripts build it; state lives next to the flat parameters."""
        lr = self.DEFAULT_LR if lr is None else lr
        weight_decay = self.DEFAULT_WD if weight_decay is None else weight_decay
        if weight_decay and any(not p.requires_grad for p in self.parameters()):
            raise NotImplementedError("weight decay with frozen parameters: torch skips them, the flat Adam kernel would "
                                      "decay them (the reference's frozen models train with weight_decay 0)")
        self._opt = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._graphs.clear()                        # betas / eps / weight decay are launch arguments baked into graphs
        if self._flat is not None:
            self._flat.init_adam(lr)

    def set_lr(self, lr):
        """Change the learning rate between steps (what optim.lr_scheduler.ReduceLROnPlateau does to the reference's
        optimizer, video/train.py:213-215, audio/train.py:156): the rate lives in the device-side Adam state, so
        captured step graphs pick it up on their next replay."""
        if not hasattr(self, "_opt"):
            self.configure_optimizer()
        self._opt["lr"] = float(lr)
        if self._flat is not None and self._flat.adam_state is not None:
            self._flat.adam_state[3] = float(lr)

    def optimizer_state_dict(self):
        """The Adam state in the layout of torch.optim.Adam(model.parameters()).state_dict() -- what the reference
        stores under "optimizer" in its checkpoints (video/train.py:246-251, audio_cues_video/train.py:178-183), so a
        run can be resumed by either side.  Parameter indices follow model.parameters() order."""
        if not hasattr(self, "_opt"):
            self.configure_optimizer()
        o, flat = self._opt, self._flat
        n_params = len(list(self.parameters()))
        state = {}
        if flat is not None and flat.m is not None and float(flat.adam_state[0]) > 0:
            step = float(flat.adam_state[0])
            for i, (p, off) in enumerate(zip(flat.params, flat.offsets)):
                if not p.requires_grad:
                    continue                                   # torch keeps no state for parameters without gradients
                state[i] = {"step": torch.tensor(step), "exp_avg": flat.m[off:off + p.numel()].view(p.shape).clone(),
                            "exp_avg_sq": flat.v[off:off + p.numel()].view(p.shape).clone()}
        group = {"lr": o["lr"], "betas": tuple(o["betas"]), "eps": o["eps"], "weight_decay": o["weight_decay"],
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": False, "params": list(range(n_params))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        """Inverse of optimizer_state_dict; accepts a state_dict written by the reference's torch.optim.Adam."""
        (group,) = sd["param_groups"]
        if group.get("amsgrad") or group.get("maximize"):
            raise ValueError("the Adam kernel implements the reference's plain Adam (no amsgrad / maximize)")
        self.configure_optimizer(lr=group["lr"], betas=tuple(group["betas"]), eps=group["eps"],
                                 weight_decay=group["weight_decay"])
        if not sd["state"]:
            return
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("move the model to its CUDA device (model.to(device)) before loading optimizer state")
        flat = self._ensure_flat(dev)
        n_live = sum(p.requires_grad for p in flat.params)
        if len(sd["state"]) != n_live:
            raise ValueError(f"optimizer state has {len(sd['state'])} entries, the model {n_live} trainable parameters")
        flat.init_adam(group["lr"])
        steps = set()
        for i, (p, off) in enumerate(zip(flat.params, flat.offsets)):
            if not p.requires_grad:
                continue
            st = sd["state"][i]
            flat.m[off:off + p.numel()].view(p.shape).copy_(st["exp_avg"])
            flat.v[off:off + p.numel()].view(p.shape).copy_(st["exp_avg_sq"])
            steps.add(float(st["step"]))
        if len(steps) != 1:
            raise ValueError("per-parameter step counts differ; the flat Adam state keeps one")
        flat.adam_state[0] = steps.pop()

    def rng_step(self):
        """Number of dropout mask draws so far (the device counter behind lr_dropout_fwd); saved in checkpoints."""
        f = self._flat
        return 0 if f is None or f.rng_step is None else int(f.rng_step.item())

    def set_rng_step(self, n):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("move the model to its CUDA device before restoring the dropout counter")
        f = self._ensure_flat(dev)
        if f.rng_step is None:
            f.rng_step = torch.zeros(1, dtype=torch.int64, device=dev)
        f.rng_step.fill_(int(n))

    def train_step(self, *inputs_and_labels, grad_allreduce=None, world=1, use_graph=True):
        """One training iteration entirely in lipread_b200 kernels:
        [log-mel if the audio input is a raw (B,20000) waveform] -> forward -> CE -> backward -> [allreduce] -> Adam.
        Call as train_step(*inputs, labels).  Returns (loss, logits) device tensors owned by the plan (no host sync)."""
        if not hasattr(self, "_opt"):
            self.configure_optimizer()
        *inputs, labels = inputs_and_labels
        named = dict(zip(self.INPUTS, inputs))
        plan = self._plan_for(named, training=True, with_backward=True)
        flat = self._flat
        if flat.m is None:
            flat.init_adam(self._opt["lr"])
        for name, t in named.items():
            buf = plan.inputs[name]
            buf.copy_(t.reshape(buf.shape), non_blocking=True)
        plan.labels.copy_(labels, non_blocking=True)
        o = self._opt

        def compute(stream, forked=None, comm_hooks=None):
            plan.run_forward(stream, forked)
            plan.ce.run(stream)
            plan.run_backward(stream, forked, comm_hooks)

        def update(stream):
            _lib.check(lib.lr_adam_step(flat.flat.data_ptr(), flat.grad.data_ptr(), flat.m.data_ptr(), flat.v.data_ptr(),
                                        flat.adam_state.data_ptr(), flat.numel, o["betas"][0], o["betas"][1], o["eps"],
                                        o["weight_decay"], 1.0 / world, stream))

        if not use_graph or not getattr(plan, "warm", False):
            # eager launches; the first step of every plan runs this way, which also serves as the warm-up
            # (function attributes, lazy module loading) that must happen outside graph capture
            s = torch.cuda.current_stream().cuda_stream
            n0 = _lib.launch_count()
            compute(s)
            if grad_allreduce is not None:
                grad_allreduce(flat.grad)
            update(s)
            plan.warm = True
            plan.kernel_launches = _lib.launch_count() - n0
            return plan.loss, plan.logits
        gkey = (id(plan), grad_allreduce is not None)
        graphs = self._graphs.get(gkey)
        if graphs is None:
            graphs = self._capture(compute, update, grad_allreduce, flat, plan)
            self._graphs[gkey] = graphs
        graphs[0].replay()
        if len(graphs) > 1:                                  # split capture: the collective runs between two graphs
            grad_allreduce(flat.grad)
            graphs[1].replay()
        return plan.loss, plan.logits

    def _capture(self, compute, update, grad_allreduce, flat, plan=None):
        """CUDA-graph capture of the step.  Single GPU: one graph.  Data parallel: the NCCL allreduce of the flat
        gradient is captured INSIDE the same graph, in BUCKETS on a communication branch: each contiguous range of
        the flat gradient goes out as soon as the last backward op that writes it has been issued (audio_fc's 19 MB
        at the very start of MidFusionFast's backward, the trunk last), so only the last bucket's collective is
        exposed before Adam.  LIPREAD_ALLREDUCE_BUCKETS=1 restores the single collective after the backward; if the
        collective cannot be captured the step falls back to two graphs with the collective launched between them."""
        main_serial = os.environ.get("LIPREAD_SERIAL_GRAPH", "0") == "1"
        if not hasattr(self, "_side"):
            self._side = torch.cuda.Stream()
            self._comm = torch.cuda.Stream()
        n_buckets = int(os.environ.get("LIPREAD_ALLREDUCE_BUCKETS", "4"))

        def body(with_update, with_allreduce):
            main = torch.cuda.current_stream()
            # LIPREAD_SERIAL_GRAPH=1: one linear chain of kernel nodes (what the ncu launch lists are taken from);
            # default: weight-gradient kernels as parallel branches of the graph
            hooks = None
            if with_allreduce and not main_serial and plan is not None and n_buckets > 1:
                table = {}
                buckets = plan.grad_buckets(max_buckets=n_buckets)
                for idx, lo, hi in buckets:
                    table.setdefault(max(idx, 0), []).append(lambda lo=lo, hi=hi: grad_allreduce(flat.grad[lo:hi]))
                hooks = (self._comm, table)
                self._n_buckets = len(buckets)
            compute(main.cuda_stream, forked=None if main_serial else (main, self._side), comm_hooks=hooks)
            if with_allreduce and hooks is None:
                grad_allreduce(flat.grad)
                self._n_buckets = 1
            if with_update:
                update(main.cuda_stream)

        # The step is captured on a HIGH-priority stream, the side / communication branches on default-priority ones:
        # kernel nodes inherit the priority of the stream they were captured on, so when a main-chain kernel and a
        # weight-gradient kernel are both ready the block scheduler serves the main chain first (the weight gradients
        # fill the gaps instead of delaying the critical path): 3.46 -> 3.36 ms per step at batch 32
        # (LIPREAD_MAIN_PRIORITY=0 restores equal priorities)
        prio = int(os.environ.get("LIPREAD_MAIN_PRIORITY", "-1"))
        cap = torch.cuda.Stream(priority=prio) if prio else None
        torch.cuda.synchronize()
        if grad_allreduce is None:
            g0 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g0, stream=cap):
                body(True, False)
            return (g0,)
        if os.environ.get("LIPREAD_ALLREDUCE_IN_GRAPH", "1") == "1":
            try:
                g0 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g0, stream=cap):
                    body(True, True)
                return (g0,)
            except Exception as e:                           # collective not capturable on this stack: split
                import warnings
                warnings.warn(f"allreduce could not be captured into the step graph ({e}); using two graphs")
                torch.cuda.synchronize()
        g0 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g0, stream=cap):
            body(False, False)
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, stream=cap):
            update(torch.cuda.current_stream().cuda_stream)
        return (g0, g1)

    def allreduce_buckets(self):
        """Number of collectives a data-parallel step issues (the flat gradient goes out in this many pieces)."""
        return getattr(self, "_n_buckets", 1)

    def launches_per_step(self):
        """Kernels of this library launched by one train_step (counted by the library during the eager step)."""
        return max((getattr(p, "kernel_launches", 0) for p in self._plans.values()), default=0)
