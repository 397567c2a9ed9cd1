"""Audio-visual fusion models behind the reference's nn.Module surface (audio_video/models/*.py).

Same class / factory names, constructor signature `(num_classes, config)`, `forward(audio, video)` contract and
`state_dict` keys as the reference; the sub-modules (torch / torchvision classes) are kept ONLY as parameter
containers so that seeded initialisation and checkpoints are interchangeable with the reference -- their torch
forward() is never called.  All arithmetic runs through the launch plans of engine.py (hand-written CUDA).

  MidFusionFast / create_mid_fusion_fast      audio_video/models/middle_fusion_fast.py:5-42
"""
import torch
import torch.nn as nn
from torchvision.models import mobilenet_v3_small

from . import _lib, engine
from ._lib import lib, ACT_NONE, ACT_RELU

N_MELS, N_FRAMES_OUT = 80, 117


class _Cfg:
    def get(self, key, default=None):
        return default


def _video_layout(video):
    """(kind, B, T, H, W, sb, st, sc, sh, sw, scale) of the lip frames in the caller's own layout."""
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


class MidFusionPlan(engine.Plan):
    """Launch plan of MidFusionFast at one batch shape."""

    def __init__(self, model, flat, key, device, training, with_backward, from_wav):
        super().__init__(flat, device, training, with_backward, precision=model.precision)
        kind, B, T, H, W = key[:5]
        self.B, self.T, self.num_classes = B, T, model.num_classes
        self.from_wav = from_wav
        m = model
        C = m.num_classes
        # ---- static input buffers (callers copy into them; pointers are baked into the plan)
        if kind == 1:
            self.video = torch.empty(B, T, H, W, 3, dtype=torch.uint8, device=self.dev)
        else:
            self.video = torch.empty(B, 3, T, H, W, dtype=torch.float32, device=self.dev)
        layout, scale = _video_layout(self.video)
        self.mel = torch.empty(B, N_MELS, N_FRAMES_OUT, dtype=torch.float32, device=self.dev)
        self.wav = torch.empty(B, 20000, dtype=torch.float32, device=self.dev) if from_wav else None
        self.labels = torch.zeros(B, dtype=torch.int64, device=self.dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.correct = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.bufs += [self.video, self.mel, self.labels, self.loss, self.correct]
        if from_wav:
            self.fwd.add("lr_logmel_fwd", self.wav, m.logmel_plan(self.dev), self.mel, B, N_FRAMES_OUT, 0)

        # ---- audio branch: conv+relu+pool -> [B,37120] -> audio_fc -> fused[:, 0:FA]
        FA = m.audio_fc.out_features
        HL = m.video_lstm.hidden_size
        FD = FA + 2 * HL
        self.fused = self.alloc(B * FD)
        dfused = self.alloc(B * FD) if with_backward else None
        KA = m.audio_fc.in_features
        a_pool = self.alloc(B * KA)
        a_arg = self.alloc(B * KA, torch.uint8)
        conv = m.audio_cnn[0]
        self.fwd.add("lr_audio_conv_fwd", self.mel, conv.weight, conv.bias, a_pool, KA, a_arg, B, N_MELS, N_FRAMES_OUT)
        ks = engine._ksplit(B, FA, KA, self.sms)
        if ks > 1:
            self.fwd.add("lr_memset", self.fused, B * FD * 4)          # split-K accumulates onto zeros
        self.linear(a_pool, KA, B, m.audio_fc.weight, m.audio_fc.bias, self.fused, FD, ksplit=ks)
        if with_backward:
            d_pool = self.alloc(B * KA)
            g = self.bgroup()
            self.linear_bwd(g, a_pool, KA, B, m.audio_fc.weight, m.audio_fc.bias, dfused, FD, dx=d_pool, ldx=KA)
            g.add("lr_audio_conv_bwd", self.mel, d_pool, KA, a_arg, flat.g(conv.weight), flat.g(conv.bias), B, N_MELS,
                  N_FRAMES_OUT, leaf=True)

        # ---- video trunk: MobileNetV3-small features + avgpool -> feat [B*T, 576]
        last = self.mbv3_features(m.video_cnn.features, self.video, layout, scale, B, T, H, W)
        feat, dfeat = self.avgpool(last)
        I = last.C
        F = B * T
        L = m.video_lstm
        G4 = 4 * HL
        # ---- BiLSTM with an out[:, -1] head: forward direction over all T steps, reverse direction one step
        xp_f, hs_f = self.alloc(F * G4), self.alloc(F * HL)
        gates_f, c_f, hp_f = self.alloc(F * G4), self.alloc(F * HL), self.alloc(F * HL)
        self.linear(feat, I, F, L.weight_ih_l0, L.bias_ih_l0, xp_f, G4)
        self.fwd.add("lr_lstm_fwd", xp_f, G4, L.bias_hh_l0, L.weight_hh_l0, hs_f, HL, gates_f, c_f, hp_f, B, T, HL, T, 0)
        fused_f = self.fused.data_ptr() + 4 * FA
        fused_r = self.fused.data_ptr() + 4 * (FA + HL)
        self.fwd.add("lr_copy2d", fused_f, FD, hs_f.data_ptr() + 4 * (T - 1) * HL, T * HL, B, HL)
        xp_r, gates_r, c_r = self.alloc(B * G4), self.alloc(B * G4), self.alloc(B * HL)
        feat_last = feat.data_ptr() + 4 * (T - 1) * I
        self.linear(feat_last, T * I, B, L.weight_ih_l0_reverse, L.bias_ih_l0_reverse, xp_r, G4)
        self.fwd.add("lr_lstm_fwd", xp_r, G4, L.bias_hh_l0_reverse, L.weight_hh_l0_reverse, fused_r, FD, gates_r, c_r, 0,
                     B, 1, HL, 1, 1)
        if with_backward:
            dg_f, dg_r = self.alloc(F * G4), self.alloc(B * G4)
            dfused_f = dfused.data_ptr() + 4 * FA
            dfused_r = dfused.data_ptr() + 4 * (FA + HL)
            g = self.bgroup()
            # forward direction
            g.add("lr_lstm_bwd", dfused_f, FD, T - 1, gates_f, c_f, L.weight_hh_l0, dg_f, B, T, HL, T, 0)
            self.linear_bwd(g, hp_f, HL, F, L.weight_hh_l0, L.bias_hh_l0, dg_f, G4)
            self.linear_bwd(g, feat, I, F, L.weight_ih_l0, L.bias_ih_l0, dg_f, G4, dx=dfeat, ldx=I)
            # reverse direction (one step from the zero state: W_hh_reverse gets no gradient)
            g.add("lr_lstm_bwd", dfused_r, FD, 0, gates_r, c_r, L.weight_hh_l0_reverse, dg_r, B, 1, HL, 1, 1)
            g.add("lr_colsum", dg_r, G4, B, G4, flat.g(L.bias_hh_l0_reverse), leaf=True)
            dfeat_last = dfeat.data_ptr() + 4 * (T - 1) * I
            self.linear_bwd(g, feat_last, T * I, B, L.weight_ih_l0_reverse, L.bias_ih_l0_reverse, dg_r, G4,
                            dx=dfeat_last, ldx=T * I, dx_residual=dfeat_last, ldr=T * I)

        # ---- classifier: Linear(FD,256)+ReLU -> Linear(256,C)
        fc0, fc2 = m.classifier[0], m.classifier[2]
        HD = fc0.out_features
        hid = self.alloc(B * HD)
        self.logits = self.alloc(B * C).view(B, C)
        self.linear(self.fused, FD, B, fc0.weight, fc0.bias, hid, HD, act=ACT_RELU)
        self.linear(hid, HD, B, fc2.weight, fc2.bias, self.logits, C)
        if with_backward:
            self.dlogits = self.alloc(B * C).view(B, C)
            dhid = self.alloc(B * HD)
            g = self.bgroup()
            self.linear_bwd(g, hid, HD, B, fc2.weight, fc2.bias, self.dlogits, C, dx=dhid, ldx=HD)
            g.add("lr_act_bwd", dhid, hid, B * HD, ACT_RELU)
            self.linear_bwd(g, self.fused, FD, B, fc0.weight, fc0.bias, dhid, HD, dx=dfused, ldx=FD)
        self.finalize()
        # per-step zeroing: BN statistic arena and (for the backward) the flat gradient
        self.pre = engine.OpList()
        self.pre.add("lr_memset", self.stats, self.stats.numel() * 8)
        self.pre_bwd = engine.OpList()
        if with_backward:
            self.pre_bwd.add("lr_memset", flat.grad, flat.grad.numel() * 4)
        self.ce = engine.OpList()
        self.ce.add("lr_memset", self.loss, 4)
        self.ce.add("lr_memset", self.correct, 4)
        self.ce.add("lr_ce_loss", self.logits, self.labels, self.loss, self.dlogits if with_backward else 0, self.correct,
                    B, C, 1.0 / B)

    # ---- execution
    def run_forward(self, stream):
        self.pre.run(stream)
        self.fwd.run(stream)

    def run_backward(self, stream, forked=None):
        """forked = (main, side) torch streams: weight-gradient kernels run on `side` concurrently with the
        dgrad chain (used under CUDA-graph capture, where it becomes a parallel branch of the graph)."""
        self.pre_bwd.run(stream)
        if forked is None:
            self.bwd.run(stream)
        else:
            self.bwd.run_forked(*forked)

    def n_launches(self):
        return len(self.fwd) + len(self.bwd) + 1


class _PlanFn(torch.autograd.Function):
    """autograd bridge for the drop-in `model(audio, video)` -> logits call: the backward runs the plan's
    hand-written backward schedule and hands the parameter gradients to autograd."""

    @staticmethod
    def forward(ctx, model, need_backward, audio, video, *params):
        plan = model._plan_for(video, training=model.training, with_backward=need_backward, from_wav=False)
        plan.mel.copy_(audio.reshape(plan.mel.shape))
        plan.video.copy_(video)
        plan.run_forward(torch.cuda.current_stream().cuda_stream)
        ctx.plan, ctx.model = plan, model
        return plan.logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        plan, model = ctx.plan, ctx.model
        if not plan.with_backward:
            raise RuntimeError("forward ran without gradient buffers (torch.no_grad)")
        plan.dlogits.copy_(dlogits)
        plan.run_backward(torch.cuda.current_stream().cuda_stream)
        flat = model._flat
        return (None, None, None, None) + tuple(flat.g(p).clone() for p in flat.params)


class MidFusionFast(nn.Module):
    """audio_video/models/middle_fusion_fast.py:5-39."""

    def __init__(self, num_classes, config=None, pretrained_state_dict=None, precision=None):
        super().__init__()
        config = config or _Cfg()
        self.num_classes = num_classes
        # "tf32": trunk GEMMs on the tensor cores (TF32 products, fp32 accumulate; >= the bf16 the north star
        # allows); "fp32": every kernel in fp32 SIMT arithmetic (strict parity with the reference's fp32 path)
        self.precision = precision or config.get("precision.compute", "tf32")
        # construction order == the reference's, so a seeded init draws identical values
        self.audio_cnn = nn.Sequential(
            nn.Conv2d(config.get("dataset.audio_channels", 1), 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2))
        if self.audio_cnn[0].in_channels != 1:
            raise ValueError("the fused audio kernel handles dataset.audio_channels == 1 (the reference default)")
        self.audio_fc = nn.Linear(16 * 40 * 58, config.get("model.audio_feature_dim", 128))
        # the reference loads MobileNet_V3_Small_Weights.IMAGENET1K_V1 from the network; offline the same
        # checkpoint can be passed as `pretrained_state_dict` (torchvision key names)
        base = mobilenet_v3_small(weights=None)
        if pretrained_state_dict is not None:
            base.load_state_dict(pretrained_state_dict)
        base.classifier = nn.Identity()
        self.video_cnn = base
        self.video_lstm = nn.LSTM(576, 128, 1, batch_first=True, bidirectional=True)
        self.classifier = nn.Sequential(nn.Linear(128 + 256, 256), nn.ReLU(), nn.Linear(256, num_classes))
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
        return self._flat

    def logmel_plan(self, device):
        from .audio_processor import AudioProcessor
        key = str(device)
        if key not in self._logmel:
            self._logmel[key] = AudioProcessor(device=device)
        return self._logmel[key].plan

    def _plan_for(self, video, training, with_backward, from_wav):
        dev = video.device
        if dev.type != "cuda":
            raise _lib.LipreadError("multimodal_lipread_b200 models run on CUDA only (no CPU path)")
        flat = self._ensure_flat(dev)
        layout, _ = _video_layout(video)
        key = layout[:5] + (bool(training), bool(with_backward), bool(from_wav), self.precision)
        plan = self._plans.get(key)
        if plan is None:
            plan = MidFusionPlan(self, flat, key, dev, training, with_backward, from_wav)
            self._plans[key] = plan
        return plan

    # ------------------------------------------------------------------ reference surface
    def forward(self, audio, video):
        """audio (B,80,117) f32 log-mel, video (B,3,T,H,W) f32 in [0,1] (or uint8 (B,T,H,W,3)) -> logits (B,C)."""
        if video.device.type != "cuda" or audio.device != video.device:
            raise _lib.LipreadError("multimodal_lipread_b200 models run on CUDA tensors only (no CPU path)")
        flat = self._ensure_flat(video.device)
        return _PlanFn.apply(self, torch.is_grad_enabled(), audio, video, *flat.params)

    # ------------------------------------------------------------------ fused training step
    def configure_optimizer(self, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        """Adam as audio_video/train.py:130 builds it; state lives next to the flat parameter buffer."""
        self._opt = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        if self._flat is not None:
            self._flat.init_adam(lr)

    def train_step(self, audio, video, labels, grad_allreduce=None, world=1, use_graph=True):
        """One iteration of audio_video/train.py:61-67 entirely in lipread_b200 kernels:
        [log-mel if `audio` is a raw (B,20000) waveform] -> forward -> CE -> backward -> [allreduce] -> Adam.
        Returns (loss, logits) as device tensors owned by the plan (no host sync)."""
        if not hasattr(self, "_opt"):
            self.configure_optimizer()
        from_wav = audio.dim() == 2 and audio.shape[1] == 20000
        plan = self._plan_for(video, training=True, with_backward=True, from_wav=from_wav)
        flat = self._flat
        if flat.m is None:
            flat.init_adam(self._opt["lr"])
        (plan.wav if from_wav else plan.mel).copy_(audio.reshape((plan.wav if from_wav else plan.mel).shape), non_blocking=True)
        plan.video.copy_(video, non_blocking=True)
        plan.labels.copy_(labels, non_blocking=True)
        o = self._opt

        def compute(stream, forked=None):
            plan.run_forward(stream)
            plan.ce.run(stream)
            plan.run_backward(stream, forked)

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
            graphs = self._capture(compute, update, split=grad_allreduce is not None)
            self._graphs[gkey] = graphs
        graphs[0].replay()
        if grad_allreduce is not None:
            grad_allreduce(flat.grad)
            graphs[1].replay()
        return plan.loss, plan.logits

    def _capture(self, compute, update, split):
        """CUDA-graph capture of the step (one graph, or compute / update split around the NCCL allreduce)."""
        torch.cuda.synchronize()
        g0 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g0):
            main = torch.cuda.current_stream()
            if not hasattr(self, "_side"):
                self._side = torch.cuda.Stream()
            compute(main.cuda_stream, forked=(main, self._side))
            if not split:
                update(main.cuda_stream)
        if not split:
            return (g0,)
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            update(torch.cuda.current_stream().cuda_stream)
        return (g0, g1)

    def launches_per_step(self):
        """Kernels of this library launched by one train_step (counted by the library during the eager step)."""
        return max((getattr(p, "kernel_launches", 0) for p in self._plans.values()), default=0)


def create_mid_fusion_fast(num_classes, config=None):
    """audio_video/models/middle_fusion_fast.py:41-42."""
    return MidFusionFast(num_classes, config)
