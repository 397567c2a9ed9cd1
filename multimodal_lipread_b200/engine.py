"""Execution engine of the B200 audio-visual path: static launch plans over preallocated HBM buffers.

A *plan* is built once per (model, batch shape, train/eval) and is nothing but two flat lists of C-ABI
kernel launches (forward, backward) with every pointer already resolved, so a step is a tight loop of
ctypes calls -- or, after CUDA-graph capture, a single graph launch.  There is no autograd tape and no
tracing compiler: the backward schedule is written out by hand next to each forward op.

Data layout in HBM (all fp32, channels-last):
  frames      the caller's own layout, uint8 (B,T,H,W,3) or float (B,3,T,H,W); read in place by the stem
  activations [F*H*W, C] row-major per layer (F = B*T frames), one value buffer + one gradient buffer
  BN stats    one float64 arena: per BatchNorm [2C] forward sums + [2C] backward sums, zeroed once per step
  parameters  ONE flat fp32 buffer (module parameters are views into it) + flat gradient + Adam m, v

Reference structure reproduced: torchvision MobileNetV3-small `features` + avgpool
(audio_video/models/middle_fusion_fast.py:15-17,34), nn.LSTM with an out[:, -1] head (:18,35-36), the audio
conv/fc branch (:8-13,28-30), the classifier (:20-25,38-39), CrossEntropyLoss + Adam (audio_video/train.py:129-130).
"""
import contextlib
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, ACT_NONE, ACT_RELU, ACT_HSWISH, ACT_HSIGMOID, ACT_RELU6

_ACT_OF = {nn.ReLU: ACT_RELU, nn.Hardswish: ACT_HSWISH, nn.Hardsigmoid: ACT_HSIGMOID, nn.ReLU6: ACT_RELU6,
           nn.Identity: ACT_NONE}


def _act_code(mod):
    for k, v in _ACT_OF.items():
        if isinstance(mod, k):
            return v
    raise NotImplementedError(f"activation {type(mod).__name__} has no lipread_b200 kernel")


class OpList:
    """A flat list of (C function, argument tuple); run() appends the stream and checks the status."""

    def __init__(self):
        self.ops = []
        self._branch = False

    def add(self, name, *args, leaf=False):
        """leaf=True marks an op whose output nobody reads before the next join (weight / bias gradients: the end of
        the backward): such ops may run on a side stream concurrently with the main chain (run_forked)."""
        self.ops.append((getattr(lib, name), tuple(a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args), name,
                         leaf or self._branch))

    def side_branch(self):
        """Context manager: every op added inside belongs to an independent branch (e.g. the audio encoder next to the
        video trunk) and is marked leaf; `join()` marks where the main chain needs the branch's results."""
        ops = self

        class _Ctx:
            def __enter__(self):
                ops._branch = True

            def __exit__(self, *exc):
                ops._branch = False
        return _Ctx()

    def join(self):
        self.ops.append((None, (), "__join__", False))

    def kernels(self):
        """The launches proper (join markers excluded)."""
        return [op for op in self.ops if op[0] is not None]

    def run(self, stream):
        for fn, args, name, _ in self.ops:
            if fn is None:
                continue
            rc = fn(*args, stream)
            if rc != 0:
                raise _lib.LipreadError(f"{name} failed ({rc}): {lib.lr_last_error().decode()}")

    def run_forked(self, main, side, comm_hooks=None):
        """Same launches, but leaf ops go to `side` (a torch.cuda.Stream) behind an event recorded on `main`
        (their inputs are ready there), and `main` joins `side` at the end.  Under CUDA-graph capture this turns
        the weight-gradient kernels into parallel branches of the graph.
        comm_hooks = (comm_stream, {op index: [callable, ...]}): after the op at that index has been issued, each
        callable runs on `comm_stream` behind everything issued so far on main and side (gradient-bucket allreduces
        that overlap the rest of the backward); main joins the comm stream at the end."""
        ms, ss = main.cuda_stream, side.cuda_stream
        comm, hooks = comm_hooks if comm_hooks else (None, {})
        forked = False
        for i, (fn, args, name, leaf) in enumerate(self.ops):
            if fn is None:                                 # join: the main chain needs everything the side branch produced
                if forked:
                    main.wait_stream(side)
                    forked = False
            elif leaf:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                rc = fn(*args, ss)
                forked = True
                if rc != 0:
                    raise _lib.LipreadError(f"{name} failed ({rc}): {lib.lr_last_error().decode()}")
            else:
                rc = fn(*args, ms)
                if rc != 0:
                    raise _lib.LipreadError(f"{name} failed ({rc}): {lib.lr_last_error().decode()}")
            for hook in hooks.get(i, ()):
                comm.wait_stream(main)
                comm.wait_stream(side)
                with torch.cuda.stream(comm):
                    hook()
        if forked:
            main.wait_stream(side)
        if comm is not None:
            main.wait_stream(comm)

    def __len__(self):
        return len(self.kernels())

    def profile(self, stream_obj, reps=5):
        """Eager launches with a CUDA-event pair around every op on the launching stream; returns
        [(name, args, mean ms)] in launch order (used by bench.py for the per-kernel roofline)."""
        s = stream_obj.cuda_stream
        acc = [0.0] * len(self.kernels())
        for _ in range(reps):
            evs = []
            for fn, args, name, _ in self.kernels():
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream_obj)
                rc = fn(*args, s)
                b.record(stream_obj)
                if rc != 0:
                    raise _lib.LipreadError(f"{name} failed ({rc}): {lib.lr_last_error().decode()}")
                evs.append((a, b))
            stream_obj.synchronize()
            for i, (a, b) in enumerate(evs):
                acc[i] += a.elapsed_time(b)
        return [(name, args, t / reps) for (fn, args, name, _), t in zip(self.kernels(), acc)]


def profile_ops_graph(ops, reps=20, flush=None):
    """Device time of every op of an OpList measured without host launch overhead: each op is captured `reps`
    times back to back into its own CUDA graph and the replay is timed with CUDA events (mean per launch, the
    ~2 us launch-to-launch gap of graph kernel nodes included).  `flush`: optional callable run (untimed) before each
    replay, e.g. an L2 flush.  Returns [(name, args, ms per launch)]."""
    out = []
    s = torch.cuda.Stream()
    for fn, args, name, _ in ops.kernels():
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            rc = fn(*args, s.cuda_stream)                      # warm (lazy module load / attributes) outside capture
            if rc != 0:
                raise _lib.LipreadError(f"{name} failed ({rc}): {lib.lr_last_error().decode()}")
            s.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(reps):
                    fn(*args, s.cuda_stream)
            best = None
            for _ in range(3):
                if flush is not None:
                    flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(s)
                g.replay()
                b.record(s)
                s.synchronize()
                t = a.elapsed_time(b) / reps
                best = t if best is None else min(best, t)
        out.append((name, args, best))
        del g
    return out


def op_algorithmic_bytes(name, args):
    """Algorithmic HBM bytes of one launch (every operand read once, every result written once; 4 bytes per fp32
    element, 2 per bf16 element of the `_h` / lr_gemm_bf16 variants)."""
    if name in ("lr_gemm", "lr_gemm_tf32"):
        M, N, K = args[8], args[9], args[10]
        extra = M * N if args[13] else 0                 # residual / accumulate operand
        return 4 * (M * K + N * K + M * N + extra)
    if name == "lr_gemm_bf16":
        c_h, M, N, K = args[8], args[9], args[10], args[11]
        ce = 2 if c_h else 4
        extra = M * N if args[14] else 0
        return 2 * (M * K + N * K) + ce * (M * N + extra)
    es = 2 if name.endswith("_h") else 4
    base = name[:-2] if name.endswith("_h") else name
    if base == "lr_bn_act_fwd":
        rows, C = args[-2], args[-1]
        return es * rows * C * (3 if args[-4] else 2)
    if base == "lr_bn_act_bwd":
        rows, C = args[-2], args[-1]
        return es * rows * C * 3                         # x, dz in; dx out (the two passes re-read x and dz)
    if base in ("lr_dwconv_fwd", "lr_dwconv_dgrad", "lr_dwconv_wgrad"):
        F, H, W, C, k, st = args[-6:]
        Ho, Wo = (H + 2 * (k // 2) - k) // st + 1, (W + 2 * (k // 2) - k) // st + 1
        return es * F * C * (H * W + Ho * Wo)
    if base == "lr_frame_reduce":
        F, HW, C, mode = args[-4:]
        return es * F * HW * C * (2 if mode else 1) + 4 * F * C
    if base == "lr_frame_scale":
        F, HW, C = args[-3:]
        return es * F * HW * C * (2 if args[0] else 1) + 4 * F * C
    return None


_SMALL_LINEAR = os.environ.get("LIPREAD_SMALL_LINEAR", "1") == "1"      # 0: heads through the tile GEMMs (A/B switch)


def _conv_geometry(conv):
    """(kh, kw, stride, pad_h, pad_w) of an nn.Conv2d -- or of an nn.Conv1d seen as a 1 x k window over [B, 1, T, C]."""
    assert conv.groups == 1
    if isinstance(conv, nn.Conv1d):
        assert conv.dilation == (1,)
        return 1, conv.kernel_size[0], conv.stride[0], 0, conv.padding[0]
    assert conv.stride[0] == conv.stride[1] and conv.dilation == (1, 1)
    return conv.kernel_size[0], conv.kernel_size[1], conv.stride[0], conv.padding[0], conv.padding[1]


def _ksplit(M, N, K, sms, min_k=256):
    tiles = ((M + 63) // 64) * ((N + 63) // 64)
    want = max(1, (2 * sms) // tiles)
    return max(1, min(want, K // min_k if K >= min_k else 1))


def dropout_seed(layer_index):
    """Seed of one dropout layer's counter-based generator: torch's global seed (so torch.manual_seed selects the mask
    sequence, as it does for the reference's nn.Dropout), the data-parallel rank (replicas draw different masks, as
    independent torch processes would) and the layer's index in the plan."""
    base = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    rank = int(os.environ.get("RANK", "0"))
    x = (base * 0x9E3779B97F4A7C15 + (rank + 1) * 0xBF58476D1CE4E5B9 + layer_index * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 31
    return x & 0x7FFFFFFFFFFFFFFF


class FlatParams:
    """All parameters of a module in ONE flat fp32 buffer (+ gradient, Adam moments); the module's Parameters
    become views into it, so state_dict()/load_state_dict() keep working with the reference's key names."""

    def __init__(self, module, device):
        self.module = module
        self.params = [p for p in module.parameters()]
        self.device = torch.device(device)
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4                    # keep every tensor 16-byte aligned
        self.offsets, self.numel = offs, n
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.m = None
        self.v = None
        self.adam_state = None
        self.rng_step = None                           # device counter behind the dropout masks (created on first use)
        for p, o in zip(self.params, offs):
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.detach().to(self.device, torch.float32))
            p.data = view
        self._ptrs = [p.data_ptr() for p in self.params]

    def intact(self):
        return all(p.data_ptr() == q for p, q in zip(self.params, self._ptrs))

    def shadow(self):
        """bf16 copy of the flat parameter buffer (precision "bf16": the B operands of the tcgen05 kind::f16 GEMMs)."""
        if getattr(self, "h16", None) is None:
            self.h16 = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.device)
        return self.h16

    def g(self, p):
        """Gradient view of parameter p inside the flat gradient buffer."""
        for q, o in zip(self.params, self.offsets):
            if q is p:
                return self.grad[o:o + p.numel()].view(p.shape)
        raise KeyError("parameter not owned by this FlatParams")

    def init_adam(self, lr):
        """Zero moments and step counter.  Buffers are reused when they exist: captured step graphs hold their
        addresses."""
        if self.m is None:
            self.m = torch.zeros_like(self.flat)
            self.v = torch.zeros_like(self.flat)
            self.adam_state = torch.zeros(4, dtype=torch.float32, device=self.device)
        else:
            self.m.zero_()
            self.v.zero_()
        self.adam_state.copy_(torch.tensor([0.0, 0.0, 0.0, float(lr)]))


class T2:
    """A [rows, C] channels-last activation (value + gradient buffer) with its frame geometry."""

    def __init__(self, plan, F, H, W, C, need_grad=True, h=None):
        """h: bf16 storage (default: the plan's precision); arithmetic on it is fp32 either way."""
        self.F, self.H, self.W, self.C = F, H, W, C
        self.rows = F * H * W
        self.h = plan.h if h is None else bool(h)
        dt = torch.bfloat16 if self.h else torch.float32
        self.val = plan.alloc(self.rows * C, dt)
        self.grad = plan.alloc(self.rows * C, dt) if (need_grad and plan.with_backward) else None

    @classmethod
    def of(cls, F, H, W, C, val, grad):
        """A view of existing fp32 buffers under another frame geometry (e.g. [B*T, C] features as a [B, 1, T, C] map)."""
        t = cls.__new__(cls)
        t.F, t.H, t.W, t.C, t.rows, t.val, t.grad = F, H, W, C, F * H * W, val, grad
        t.h = isinstance(val, torch.Tensor) and val.dtype == torch.bfloat16
        return t


class Plan:
    """Launch plan for one model at one batch shape."""

    def __init__(self, flat, device, training, with_backward, precision="tf32"):
        """precision: "tf32" -- the GEMMs with enough rows run on the tensor cores (tcgen05, TF32 products, fp32
        accumulate); "fp32" -- every GEMM on the fp32 SIMT kernel (strict-parity mode); "bf16" -- as "tf32" plus
        bf16 storage of the trunk activations and bf16 tensor-core products there."""
        if precision not in ("tf32", "fp32", "bf16"):
            raise ValueError(f"precision must be 'tf32', 'fp32' or 'bf16', got {precision!r}")
        self.flat, self.dev, self.training, self.with_backward = flat, torch.device(device), training, with_backward
        self.tc = precision in ("tf32", "bf16")
        # "bf16": the trunk's per-pixel activations and their gradients are stored as bfloat16 and its GEMMs run
        # tcgen05 kind::f16 on a bf16 shadow of the weights; per-frame vectors, heads, LSTM stay fp32 / tf32
        self.h = precision == "bf16"
        self._uses_shadow = False
        self.weight_taps = []          # (weight, bf16 tap-major operand, Cout, Cin, taps, mode): re-laid out once per step
        self.fwd, self.bwd_rev = OpList(), []          # bwd_rev: groups appended in forward order, run reversed
        self.bufs = []
        self.sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        self._stat_reqs = []
        self.stats = None
        self._ws_floats = 0                            # shared scratch (dgrad patch matrices), sized at finalize()
        self._ws = None
        self.rng_step = None                           # device counter behind the dropout masks
        self._n_dropout = 0

    # ---- memory
    def alloc(self, n, dtype=torch.float32):
        t = torch.empty(max(int(n), 1), dtype=dtype, device=self.dev)
        self.bufs.append(t)
        return t

    def stat_slot(self, C):
        """Reserve [2C] forward + [2C] backward float64 sums; resolved to pointers by finalize()."""
        slot = {"C": C}
        self._stat_reqs.append(slot)
        return slot

    def finalize(self):
        total = sum(4 * s["C"] for s in self._stat_reqs)
        self.stats = torch.zeros(max(total, 1), dtype=torch.float64, device=self.dev)
        self.bufs.append(self.stats)
        off = 0
        base = self.stats.data_ptr()
        for s in self._stat_reqs:
            s["fwd"] = base + 8 * off
            s["bwd"] = base + 8 * (off + 2 * s["C"])
            off += 4 * s["C"]
        if self._ws_floats:
            self._ws = self.alloc(self._ws_floats)
        for ops in [self.fwd] + [g for g in self.bwd_rev]:
            ops.ops = [(fn, tuple(a() if callable(a) else a for a in args), name, leaf) for fn, args, name, leaf in ops.ops]
        self.bwd = OpList()
        for g in reversed(self.bwd_rev):
            self.bwd.ops.extend(g.ops)

    def bgroup(self):
        g = OpList()
        self.bwd_rev.append(g)
        return g

    def grad_buckets(self, max_buckets=4, min_bytes=1 << 20):
        """[(index of the LAST backward op that writes into the bucket, lo, hi)] over the flat gradient: contiguous
        ranges of whole parameters, cut where the readiness (position in the backward schedule of the last op that
        touches the parameter's gradient) changes most, so that each range can be allreduced as soon as it is final
        while the rest of the backward still runs.  Derived from the launch list itself: an op writes a parameter's
        gradient iff one of its pointer arguments lies inside that parameter's slice of flat.grad."""
        flat = self.flat
        base, nbytes = flat.grad.data_ptr(), flat.grad.numel() * 4
        starts = flat.offsets + [flat.numel]
        ready = [-1] * len(flat.params)
        import bisect
        for i, (fn, args, name, leaf) in enumerate(self.bwd.ops):
            if fn is None:
                continue
            for a in args:
                if isinstance(a, int) and base <= a < base + nbytes:
                    k = bisect.bisect_right(flat.offsets, (a - base) // 4) - 1
                    ready[k] = max(ready[k], i)
        # greedy merge of adjacent parameters into at most max_buckets ranges: repeatedly merge the adjacent pair whose
        # merge delays the earlier-ready side least (bytes * delay), never leaving a bucket below min_bytes
        buckets = [[ready[k], starts[k], starts[k + 1]] for k in range(len(flat.params))]
        def cost(a, b):
            r = max(a[0], b[0])
            return (r - a[0]) * (a[2] - a[1]) + (r - b[0]) * (b[2] - b[1])
        while len(buckets) > 1:
            small = [j for j in range(len(buckets)) if (buckets[j][2] - buckets[j][1]) * 4 < min_bytes]
            if len(buckets) <= max_buckets and not small:
                break
            cands = range(len(buckets) - 1)
            if len(buckets) <= max_buckets:                       # only the undersized ones still need a partner
                cands = sorted({j for t in small for j in (t - 1, t) if 0 <= j < len(buckets) - 1})
            j = min(cands, key=lambda j: cost(buckets[j], buckets[j + 1]))
            a, b = buckets[j], buckets[j + 1]
            buckets[j:j + 2] = [[max(a[0], b[0]), a[1], b[2]]]
        return [tuple(b) for b in buckets]

    def workspace(self, n_elems, h=False):
        """A scratch buffer shared by every op that asks (used and consumed within one backward group on the
        main stream); returns a thunk resolved by finalize().  h: the elements are bf16."""
        n_floats = (int(n_elems) + 1) // 2 if h else int(n_elems)
        self._ws_floats = max(self._ws_floats, n_floats)
        return lambda: self._ws.data_ptr()

    # ---- bf16 operands
    def wh(self, w):
        """Pointer to the bf16 shadow of parameter tensor `w` (a view into the flat buffer).  The shadow of the whole
        flat parameter buffer is refreshed by ONE cast launch at the start of every step (ModelPlan._finish)."""
        off = (w.data_ptr() - self.flat.flat.data_ptr()) // 4
        assert 0 <= off and off + w.numel() <= self.flat.numel, "not a parameter of this model"
        self._uses_shadow = True
        return self.flat.shadow().data_ptr() + 2 * off

    def cast_h(self, ops, src, leaf=False):
        """bf16 copy of an fp32 scratch matrix (a re-laid-out weight), emitted into `ops`."""
        dst = self.alloc(src.numel(), torch.bfloat16)
        ops.add("lr_cast_bf16", src, dst, src.numel(), leaf=leaf)
        return dst

    def alloc_like(self, t, n):
        """n elements with T2 t's storage type."""
        return self.alloc(n, torch.bfloat16 if t.h else torch.float32)

    # ---- op emitters ----------------------------------------------------------------------------------
    def lstm_op(self, which, H, nsteps):
        """Name of the recurrence kernel: the tensor-core walk (csrc/lstm_tc.cu: bf16 W_hh / h operands, tcgen05) under
        precision "bf16" for H = 128 / 256 / 512 and walks of more than one step, else the fp32 kernels.  Measured at
        B = 32, T = 29 (tools/microbench.py lstm, profiles/r2_lstm_microbench.txt), forward / backward: H = 512
        483 / 580 -> 137 / 196 us, H = 256 237 / 270 -> 127 / 145 us, H = 128 132 / 144 -> 73 / 74 us (4-CTA cluster)."""
        tc = self.h and H in (128, 256, 512) and nsteps > 1 and os.environ.get("LIPREAD_LSTM_TC", "1") != "0"
        return f"lr_lstm_{which}_tc" if tc else f"lr_lstm_{which}"

    def gemm(self, ops, A, lda, at, B, ldb, bt, C, ldc, M, N, K, bias=0, act=ACT_NONE, R=0, ldr=0, stats=0, ksplit=1,
             leaf=False):
        ops.add("lr_gemm", A, lda, at, B, ldb, bt, C, ldc, M, N, K, bias, act, R, ldr, stats, ksplit, leaf=leaf)

    def use_tc(self, M, N, K, lda, ldb):
        """Tensor-core path for GEMMs that stream many rows (1x1 convs, SE / LSTM projections over all frames);
        tiny GEMMs (head, per-clip vectors) stay on the SIMT kernel.  TMA needs 16-byte row pitches."""
        return self.tc and lda % 4 == 0 and ldb % 4 == 0 and max(M, K) >= 512

    def gemm_auto(self, ops, A, lda, at, B, ldb, bt, C, ldc, M, N, K, bias=0, act=ACT_NONE, R=0, ldr=0, stats=0,
                  split_ok=False, h=False):
        """Emit one GEMM, choosing the tcgen05 kernel when it pays; split_ok: the caller accumulates into a
        pre-initialised C, so the reduction may be split over CTAs with atomic adds.
        h: A and B are bf16 (activations / bf16 weight shadow): tcgen05 kind::f16; the result is bf16 too, except for
        split_ok calls (weight gradients accumulate in fp32)."""
        if h:
            if lda % 8 or ldb % 8 or (not split_ok and ldc % 8):
                raise NotImplementedError(f"bf16 GEMM operands need row pitches that are multiples of 8 (got {lda}, {ldb}, {ldc})")
            ks = 1
            if split_ok:
                bn_tiles = ((M + 127) // 128) * ((N + 255) // 256)
                ks = max(1, min((K + 255) // 256, (2 * self.sms) // bn_tiles))
            if ks > 1:
                ops.add("lr_gemm_bf16", A, lda, at, B, ldb, bt, C, ldc, 0, M, N, K, 0, ACT_NONE, 0, 0, 0, ks, leaf=True)
            else:
                ops.add("lr_gemm_bf16", A, lda, at, B, ldb, bt, C, ldc, int(not split_ok), M, N, K, bias, act,
                        (C if split_ok else R), (ldc if split_ok else ldr), stats, 1, leaf=split_ok)
            return
        if N >= 4 and self.use_tc(M, N, K, lda, ldb):
            ks = 1
            if split_ok:
                bn_tiles = ((M + 127) // 128) * ((N + 255) // 256)
                ks = max(1, min((K + 255) // 256, (2 * self.sms) // bn_tiles))
            if ks > 1:
                ops.add("lr_gemm_tf32", A, lda, at, B, ldb, bt, C, ldc, M, N, K, 0, ACT_NONE, 0, 0, 0, ks, leaf=split_ok)
            else:
                ops.add("lr_gemm_tf32", A, lda, at, B, ldb, bt, C, ldc, M, N, K, bias, act, (C if split_ok else R),
                        (ldc if split_ok else ldr), stats, 1, leaf=split_ok)
            return
        ks = _ksplit(M, N, K, self.sms) if split_ok else 1
        if ks > 1:
            self.gemm(ops, A, lda, at, B, ldb, bt, C, ldc, M, N, K, ksplit=ks, leaf=split_ok)
        else:
            self.gemm(ops, A, lda, at, B, ldb, bt, C, ldc, M, N, K, bias=bias, act=act, R=(C if split_ok else R),
                      ldr=(ldc if split_ok else ldr), stats=stats, leaf=split_ok)

    def linear(self, x, lda, M, w, b, out, ldc, act=ACT_NONE, stats=0, ksplit=1):
        N, K = w.shape[0], w[0].numel()
        if (ksplit == 1 and M <= 32 and not stats and K % 4 == 0 and lda % 4 == 0 and 32 <= K <= 1600 and _SMALL_LINEAR
                and act in (ACT_NONE, ACT_RELU, ACT_RELU6, ACT_HSIGMOID, ACT_HSWISH)):
            # heads (one row per clip): output columns over CTAs instead of a one-CTA tile chain (csrc/rowgemm.cu)
            self.fwd.add("lr_linear_small_fwd", x, lda, w, (b if b is not None else 0), out, ldc, M, N, K, act)
            return
        if ksplit > 1:
            # forward split-K is the ORDERED kind (partials in a workspace, reduced in slice order): bit-reproducible
            ws = self.alloc(ksplit * M * N)
            self.fwd.add("lr_gemm_splitk", x, lda, 0, w, K, 0, out, ldc, M, N, K, (b if b is not None else 0), act, ksplit,
                         ws, ws.numel() * 4)
        else:
            self.gemm_auto(self.fwd, x, lda, 0, w, K, 0, out, ldc, M, N, K, bias=(b if b is not None else 0), act=act,
                           stats=stats)

    def linear_bwd(self, g, x, lda, M, w, b, dy, ldy, dx=None, ldx=0, dx_residual=0, ldr=0, h=False):
        """dw += dy^T x, db += colsum(dy), dx = dy w (+ residual).  Ops are appended to backward group g.
        h: x, dy, dx (and the residual) are bf16 activations; dw / db stay fp32."""
        N, K = w.shape[0], w[0].numel()
        dw = self.flat.g(w)
        self.gemm_auto(g, dy, ldy, 1, x, lda, 1, dw, K, N, K, M, split_ok=True, h=h)       # dw += dy^T x
        if b is not None:
            g.add("lr_colsum_h" if h else "lr_colsum", dy, ldy, M, N, self.flat.g(b), leaf=True)
        if dx is not None:
            if (not h and M <= 32 and N % 4 == 0 and N <= 1088 and K % 32 == 0 and ldy % 4 == 0 and ldx % 4 == 0
                    and ((isinstance(dx_residual, int) and dx_residual == 0) or ldr % 4 == 0) and _SMALL_LINEAR):
                g.add("lr_linear_small_dgrad", dy, ldy, w, dx, ldx, dx_residual, ldr, M, N, K)
                return
            self.gemm_auto(g, dy, ldy, 0, (self.wh(w) if h else w), K, 1, dx, ldx, M, K, N, R=dx_residual, ldr=ldr, h=h)

    def bn_act(self, x, bn, act, out, residual=None, res_pre=False, dres=None):
        """out.val = act(bn(x.val)) (+ residual.val); backward: x.grad from out.grad (residual.grad is out.grad).
        res_pre (ResNet BasicBlock): out.val = act(bn(x.val) + residual.val); the backward takes the derivative
        through out.val and writes the masked gradient (= the identity branch's gradient) to `dres`."""
        slot = x.stat_slot
        st = (lambda s=slot: s["fwd"]) if self.training else 0
        momentum = 0.1 if bn.momentum is None else float(bn.momentum)
        sfx = "_h" if x.h else ""
        self.fwd.add("lr_bn_act_fwd" + sfx, x.val, st, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                     bn.num_batches_tracked, float(bn.eps), momentum, act, int(self.training),
                     residual.val if residual is not None else 0, int(res_pre), out.val, x.rows, x.C)
        if self.with_backward:
            g = self.bgroup()
            g.add("lr_bn_act_bwd" + sfx, x.val, st, bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.eps), act,
                  int(self.training), out.grad, (out.val if res_pre else 0), (dres if dres is not None else 0),
                  (lambda s=slot: s["bwd"]), x.grad, self.flat.g(bn.weight), self.flat.g(bn.bias), x.rows, x.C)

    def pw_conv(self, x, conv, F, H, W):
        """1x1 convolution as a GEMM on [rows, Cin]; returns the raw output tensor (with BN statistics slot)."""
        Cout, Cin = conv.out_channels, conv.in_channels
        y = T2(self, F, H, W, Cout, h=x.h)
        y.stat_slot = self.stat_slot(Cout)
        st = (lambda s=y.stat_slot: s["fwd"]) if self.training else 0
        self.gemm_auto(self.fwd, x.val, Cin, 0, (self.wh(conv.weight) if x.h else conv.weight), Cin, 0, y.val, Cout, x.rows,
                       Cout, Cin, stats=st, h=x.h)
        return y

    def dw_conv(self, x, conv):
        C, k, s = conv.in_channels, conv.kernel_size[0], conv.stride[0]
        Ho = (x.H + 2 * (k // 2) - k) // s + 1
        Wo = (x.W + 2 * (k // 2) - k) // s + 1
        y = T2(self, x.F, Ho, Wo, C, h=x.h)
        y.stat_slot = self.stat_slot(C)
        st = (lambda sl=y.stat_slot: sl["fwd"]) if self.training else self.dummy_stats()
        sfx = "_h" if x.h else ""
        self.fwd.add("lr_dwconv_fwd" + sfx, x.val, conv.weight, y.val, st, x.F, x.H, x.W, C, k, s)
        if self.with_backward:
            g = self.bgroup()
            g.add("lr_dwconv_wgrad" + sfx, y.grad, x.val, self.flat.g(conv.weight), x.F, x.H, x.W, C, k, s, leaf=True)
            g.add("lr_dwconv_dgrad" + sfx, y.grad, conv.weight, x.grad, x.F, x.H, x.W, C, k, s)
        return y

    @contextlib.contextmanager
    def frozen(self, on=True):
        """Emit a sub-network forward-only inside a training plan: a backbone whose parameters have
        requires_grad=False and whose input needs no gradient (audio_cues_video/models/early_fusion_mobile.py:100-103,
        131-133).  BatchNorm keeps its train-mode batch statistics and running-stat updates."""
        old = self.with_backward
        if on:
            self.with_backward = False
        try:
            yield
        finally:
            self.with_backward = old

    def dummy_stats(self):
        if not hasattr(self, "_dummy"):
            self._dummy = self.alloc(4096, torch.float64)
        return self._dummy

    # ---- MobileNetV3 -----------------------------------------------------------------------------------
    def mbv3_features(self, feats, frames, layout, scale, B, T, H, W):
        """torchvision `features` Sequential on B*T frames read in the caller's layout -> last activation T2."""
        stem = feats[0]
        conv, bn, act = stem[0], stem[1], _act_code(stem[2])
        assert conv.kernel_size == (3, 3) and conv.stride == (2, 2) and conv.out_channels == 16 and conv.in_channels == 3
        F = B * T
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        if self.tc and os.environ.get("LIPREAD_STEM", "gemm") == "gemm":
            # tensor-core mode: the stem as patch matrix + tcgen05 GEMM (its weight gradient is a split-K GEMM over
            # 1.8 M pixels: ~100 us against 540 us for the direct kernel); the direct kernels stay the fp32 path
            raw = self.dense_conv(None, conv, frames=(frames, layout, scale))
            if self.with_backward:
                self.dense_conv_bwd(raw)
        else:
            raw = T2(self, F, Ho, Wo, 16, h=False)
            raw.stat_slot = self.stat_slot(16)
            st = (lambda s=raw.stat_slot: s["fwd"]) if self.training else self.dummy_stats()
            self.fwd.add("lr_stem_conv_fwd", frames, *layout, float(scale), conv.weight, raw.val, st)
            if self.with_backward:
                g = self.bgroup()
                g.add("lr_stem_conv_wgrad", frames, *layout, float(scale), raw.grad, self.flat.g(conv.weight), leaf=True)
        cur = T2(self, F, Ho, Wo, 16, h=raw.h)
        self.bn_act(raw, bn, act, cur)
        for blk in feats[1:]:
            if hasattr(blk, "block"):
                cur = self.inverted_residual(blk, cur)
            else:
                cur = self.pw_bn_act(cur, blk[0], blk[1], _act_code(blk[2]))
        return cur

    @staticmethod
    def _ir_layers(blk):
        """torchvision InvertedResidual (MobileNetV3: blk.block, MobileNetV2: blk.conv) -> [("conv", conv, bn, act) |
        ("se", module)], use_res."""
        out = []
        mods = list(blk.block) if hasattr(blk, "block") else list(blk.conv)
        i = 0
        while i < len(mods):
            m = mods[i]
            if type(m).__name__ == "SqueezeExcitation":
                out.append(("se", m))
            elif isinstance(m, nn.Conv2d):                       # MobileNetV2 tail: bare Conv2d followed by BatchNorm2d
                out.append(("conv", m, mods[i + 1], ACT_NONE))
                i += 1
            else:                                                # Conv2dNormActivation
                out.append(("conv", m[0], m[1], _act_code(m[2]) if len(m) > 2 else ACT_NONE))
            i += 1
        return out, bool(blk.use_res_connect)

    def inverted_residual(self, blk, x):
        """torchvision InvertedResidual: [1x1 expand+BN+act] -> depthwise+BN+act -> [SE] -> 1x1 project+BN (+x)."""
        layers, use_res = self._ir_layers(blk)
        cur, deferred = x, []
        for li, layer in enumerate(layers):
            last = li == len(layers) - 1
            if layer[0] == "se":
                cur = self.squeeze_excite(layer[1], cur)
                continue
            _, conv, bn, act = layer
            if conv.groups == 1:
                assert conv.kernel_size == (1, 1)
                raw = self.pw_conv(cur, conv, cur.F, cur.H, cur.W)
                if self.with_backward:
                    # Backward groups run in reverse registration order, so the conv's group is registered here
                    # (before its BatchNorm's) and filled in once the block output -- whose gradient the residual
                    # branch adds to the block input's -- exists.
                    deferred.append((self.bgroup(), cur, conv, raw, (cur is x) and use_res))
            else:
                assert conv.groups == conv.in_channels
                if (cur is x) and use_res:
                    raise NotImplementedError("residual block without an expansion conv")
                raw = self.dw_conv(cur, conv)
            out = T2(self, raw.F, raw.H, raw.W, raw.C, h=raw.h)
            self.bn_act(raw, bn, act, out, residual=x if (last and use_res) else None)
            cur = out
        for g, inp, conv, raw, add_res in deferred:
            Cout, Cin = conv.out_channels, conv.in_channels
            self.linear_bwd(g, inp.val, Cin, inp.rows, conv.weight, None, raw.grad, Cout, dx=inp.grad, ldx=Cin,
                            dx_residual=(cur.grad if add_res else 0), ldr=Cin, h=inp.h)
        return cur

    def pw_bn_act(self, cur, conv, bn, act):
        """Stand-alone 1x1 Conv2dNormActivation (the last `features` stage of MobileNetV2 / V3)."""
        assert conv.kernel_size == (1, 1) and conv.groups == 1
        raw = self.pw_conv(cur, conv, cur.F, cur.H, cur.W)
        if self.with_backward:
            self.linear_bwd(self.bgroup(), cur.val, conv.in_channels, cur.rows, conv.weight, None, raw.grad,
                            conv.out_channels, dx=cur.grad, ldx=conv.in_channels, h=cur.h)
        out = T2(self, cur.F, cur.H, cur.W, conv.out_channels, h=raw.h)
        self.bn_act(raw, bn, act, out)
        return out

    def mbv2_features(self, feats, frames):
        """torchvision mobilenet_v2 `features` on frames = (tensor, layout, scale) -> last activation T2
        (audio_cues_video/models/late_fusion_mobile.py:33-40, video/models/mobilenet_lstm.py:29-39)."""
        stem = feats[0]
        raw = self.dense_conv(None, stem[0], frames=frames)
        if self.with_backward:
            self.dense_conv_bwd(raw)
        cur = T2(self, raw.F, raw.H, raw.W, raw.C, h=raw.h)
        self.bn_act(raw, stem[1], _act_code(stem[2]), cur)
        self.trace = [cur]
        for blk in feats[1:]:
            if hasattr(blk, "conv"):
                cur = self.inverted_residual(blk, cur)
            else:
                cur = self.pw_bn_act(cur, blk[0], blk[1], _act_code(blk[2]))
            self.trace.append(cur)
        return cur

    # ---- torchvision ShuffleNetV2 -----------------------------------------------------------------------------
    def _pw_strided(self, xval, lda, xgrad, ldg, x, conv, bn, act, accumulate=False):
        """1x1 conv + BN + act whose input is a column slice (row stride lda) of a wider activation; its input gradient
        goes to the matching slice of that activation's gradient (written, or accumulated when another branch wrote)."""
        Cout, Cin = conv.out_channels, conv.in_channels
        assert conv.kernel_size == (1, 1) and conv.groups == 1 and conv.bias is None
        if x.h:
            raise NotImplementedError("ShuffleNetV2 units address column slices of fp32 rows (fp32 storage only)")
        raw = T2(self, x.F, x.H, x.W, Cout, h=False)
        raw.stat_slot = self.stat_slot(Cout)
        st = (lambda sl=raw.stat_slot: sl["fwd"]) if self.training else 0
        self.gemm_auto(self.fwd, xval, lda, 0, conv.weight, Cin, 0, raw.val, Cout, x.rows, Cout, Cin, stats=st)
        if self.with_backward:
            self.linear_bwd(self.bgroup(), xval, lda, x.rows, conv.weight, None, raw.grad, Cout, dx=xgrad, ldx=ldg,
                            dx_residual=(xgrad if accumulate else 0), ldr=ldg)
        out = T2(self, x.F, x.H, x.W, Cout, h=raw.h)
        self.bn_act(raw, bn, act, out)
        return out

    def _dw_bn(self, x, conv, bn):
        raw = self.dw_conv(x, conv)
        out = T2(self, raw.F, raw.H, raw.W, raw.C, h=raw.h)
        self.bn_act(raw, bn, ACT_NONE, out)
        return out

    def shuffle_unit(self, blk, x):
        """torchvision shufflenetv2.InvertedResidual: stride 1: cat(x1, branch2(x2)); stride 2: cat(branch1(x),
        branch2(x)); then channel_shuffle(., 2) (an interleave of the two halves in channels-last rows)."""
        wb = self.with_backward
        b2 = blk.branch2
        if blk.stride == 1:
            C = x.C
            bf = C // 2
            xg = (x.grad.data_ptr() + 4 * bf) if wb else None
            h = self._pw_strided(x.val.data_ptr() + 4 * bf, C, xg, C, x, b2[0], b2[1], ACT_RELU)
            h = self._dw_bn(h, b2[3], b2[4])
            right = self.pw_bn_act(h, b2[5], b2[6], ACT_RELU)
            out = T2(self, x.F, x.H, x.W, C, h=x.h)
            self.fwd.add("lr_shuffle2_fwd", x.val, C, right.val, bf, out.val, x.rows, bf)
            if wb:
                self.bgroup().add("lr_shuffle2_bwd", out.grad, x.grad, C, right.grad, bf, x.rows, bf)
            return out
        # stride 2: branch2 is registered first so that, in the reversed backward, branch1's depthwise dgrad WRITES
        # x.grad and branch2's first 1x1 conv then accumulates onto it
        b1 = blk.branch1
        h = self._pw_strided(x.val, x.C, x.grad, x.C, x, b2[0], b2[1], ACT_RELU, accumulate=True)
        h = self._dw_bn(h, b2[3], b2[4])
        right = self.pw_bn_act(h, b2[5], b2[6], ACT_RELU)
        l = self._dw_bn(x, b1[0], b1[1])
        left = self.pw_bn_act(l, b1[2], b1[3], ACT_RELU)
        bf = left.C
        out = T2(self, left.F, left.H, left.W, 2 * bf, h=left.h)
        self.fwd.add("lr_shuffle2_fwd", left.val, bf, right.val, bf, out.val, left.rows, bf)
        if wb:
            self.bgroup().add("lr_shuffle2_bwd", out.grad, left.grad, bf, right.grad, bf, left.rows, bf)
        return out

    def shufflenet_features(self, seq, frames):
        """Sequential(conv1, maxpool, stage2, stage3, stage4, conv5) of torchvision shufflenet_v2_x0_5 / x1_0 on
        frames = (tensor, layout, scale) -> last activation T2 (video/models/shufflenet_lstm.py:47-55)."""
        conv1, mp, stages, conv5 = seq[0], seq[1], (seq[2], seq[3], seq[4]), seq[5]
        # the split / shuffle kernels address column slices of fp32 rows: this trunk keeps fp32 storage in every mode
        raw = self.dense_conv(None, conv1[0], frames=frames, h=False)
        if self.with_backward:
            self.dense_conv_bwd(raw)
        cur = T2(self, raw.F, raw.H, raw.W, raw.C, h=raw.h)
        self.bn_act(raw, conv1[1], ACT_RELU, cur)
        cur = self.maxpool(cur, mp.kernel_size, mp.stride, mp.padding)
        for stage in stages:
            for blk in stage:
                cur = self.shuffle_unit(blk, cur)
        return self.pw_bn_act(cur, conv5[0], conv5[1], ACT_RELU)

    def squeeze_excite(self, se, a):
        """SqueezeExcitation: b = a * hardsigmoid(fc2(relu(fc1(mean_hw(a)))))."""
        F, HW, C = a.F, a.H * a.W, a.C
        Cs = se.fc1.out_channels
        p, h1, s = self.alloc(F * C), self.alloc(F * Cs), self.alloc(F * C)      # per-frame vectors: always fp32
        out = T2(self, a.F, a.H, a.W, C, h=a.h)
        sfx = "_h" if a.h else ""
        a1, a2 = _act_code(se.activation), _act_code(se.scale_activation)
        # fc1 -> act -> fc2 -> scale_act on the pooled vectors as ONE launch (csrc/se.cu), and one for the chain
        # dz2 -> dz1 -> dp of the backward; LIPREAD_SE_FUSED=0 keeps the four / six separate GEMM / activation launches
        fused = (os.environ.get("LIPREAD_SE_FUSED", "1") == "1" and C % 4 == 0 and Cs % 4 == 0 and C <= 2048
                 and se.fc1.bias is not None and se.fc2.bias is not None
                 and all(x in (ACT_NONE, ACT_RELU, ACT_RELU6, ACT_HSIGMOID) for x in (a1, a2)))
        self.fwd.add("lr_frame_reduce" + sfx, a.val, 0, p, F, HW, C, 0)
        # (forward of the widest gates: 308 MFLOP of fp32 FMA at C = 576 take 31 us in the row-block kernel, the two
        # tensor-core GEMM launches 20; the backward chain wins at every width: 19 us against four launches / 37 us)
        if fused and C * Cs <= int(os.environ.get("LIPREAD_SE_FWD_MAX", 40000)):
            self.fwd.add("lr_se_fc_fwd", p, se.fc1.weight, se.fc1.bias, se.fc2.weight, se.fc2.bias, h1, s, F, C, Cs, a1, a2)
        else:
            self.linear(p, C, F, se.fc1.weight, se.fc1.bias, h1, Cs, act=a1)
            self.linear(h1, Cs, F, se.fc2.weight, se.fc2.bias, s, C, act=a2)
        self.fwd.add("lr_frame_scale" + sfx, a.val, s, 0, out.val, F, HW, C)
        if self.with_backward:
            ds, dh1, dp = self.alloc(F * C), self.alloc(F * Cs), self.alloc(F * C)
            g = self.bgroup()
            g.add("lr_frame_reduce" + sfx, out.grad, a.val, ds, F, HW, C, 1)
            if fused:
                g.add("lr_se_fc_bwd", ds, s, h1, se.fc1.weight, se.fc2.weight, dh1, dp, F, C, Cs, a1, a2)
                self.linear_bwd(g, h1, Cs, F, se.fc2.weight, se.fc2.bias, ds, C)        # weight gradients: leaf launches
                self.linear_bwd(g, p, C, F, se.fc1.weight, se.fc1.bias, dh1, Cs)
            else:
                g.add("lr_act_bwd", ds, s, F * C, a2)
                self.linear_bwd(g, h1, Cs, F, se.fc2.weight, se.fc2.bias, ds, C, dx=dh1, ldx=Cs)
                g.add("lr_act_bwd", dh1, h1, F * Cs, a1)
                self.linear_bwd(g, p, C, F, se.fc1.weight, se.fc1.bias, dh1, Cs, dx=dp, ldx=C)
            g.add("lr_frame_scale" + sfx, out.grad, s, dp, a.grad, F, HW, C)
        return out

    def avgpool(self, a):
        """AdaptiveAvgPool2d(1) + flatten: [F, HW, C] -> feat [F, C] (value, gradient)."""
        F, HW, C = a.F, a.H * a.W, a.C
        feat, dfeat = self.alloc(F * C), (self.alloc(F * C) if self.with_backward else None)
        sfx = "_h" if a.h else ""
        self.fwd.add("lr_frame_reduce" + sfx, a.val, 0, feat, F, HW, C, 0)
        if self.with_backward:
            g = self.bgroup()
            g.add("lr_frame_scale" + sfx, 0, 0, dfeat, a.grad, F, HW, C)
        return feat, dfeat

    # ---- dense convolutions (ResNet-18, AudioEncoder, MobileNetV2 stem) ----------------------------------
    def dense_conv(self, x, conv, frames=None, need_dx=True, act=ACT_NONE, with_stats=True, h=None):
        """nn.Conv2d with groups == 1 on a channels-last T2 (or, for a stem, on the raw frames in the caller's
        layout: frames = (tensor, (is_u8, B, T, H, W, sb, st, sc, sh, sw), scale)) as im2col + GEMM.  Returns the
        raw output T2 (BatchNorm statistic slot attached).  The backward group is reserved here (its position fixes
        when it runs) and filled by dense_conv_bwd once the caller knows what accumulates into x.grad.
        Storage type: that of x; for a stem the plan's (h overrides it: audio encoders whose consumers read fp32)."""
        Cout, Cin = conv.out_channels, conv.in_channels
        kh, kw, st_, pd, pw = _conv_geometry(conv)
        hh = x.h if x is not None else (self.h if h is None else bool(h))
        if frames is not None:
            ft, (is_u8, B, T, Hs, Ws, sb, stt, sc, sh, sw), scale = frames
            F = B * T
            src = (int(is_u8), float(scale), F, T, sb, stt, sc, sh, sw)
            xptr = ft
            need_dx = False
        else:
            F, Hs, Ws = x.F, x.H, x.W
            src = (0, 1.0, F, 1, Hs * Ws * Cin, 0, 1, Ws * Cin, Cin)
            xptr = x.val
        q = 8 if hh else 4                                     # row pitches: 16-byte multiples for TMA
        if hh and Cout % 8:
            raise NotImplementedError(f"bf16 storage needs Cout % 8 == 0 (conv with {Cout} output channels)")
        Ho, Wo = (Hs + 2 * pd - kh) // st_ + 1, (Ws + 2 * pw - kw) // st_ + 1
        rows = F * Ho * Wo
        K = Cin * kh * kw
        pointwise = frames is None and kh == 1 and kw == 1 and st_ == 1 and pd == 0 and Cin % q == 0
        # tap-major patch matrix (shifted vector copies) for channels-last inputs; torch-order gather for the stems
        tap = frames is None and not pointwise and Cin % q == 0 and Cout % 4 == 0
        assert tap or pw == pd, "unequal paddings need the tap-major path (channels-last input, C % 4 == 0)"
        if hh and frames is None and not (pointwise or tap):
            raise NotImplementedError("bf16 storage: dense conv on a channels-last input needs Cin % 8 == 0")
        ldk = K if (pointwise or tap) else (K + q - 1) // q * q
        adt = torch.bfloat16 if hh else torch.float32
        sfx = "_h" if hh else ""
        # implicit GEMM (csrc/conv_igemm.cu): 3x3 / stride 1 / pad 1 on bf16 storage -- the tap's shifted tile comes
        # straight from the activation through a 4-D TMA box, no patch matrix is written or read
        igemm = (hh and tap and kh == 3 and kw == 3 and st_ == 1 and pd == 1 and pw == 1 and Cin % 64 == 0
                 and Cout % 64 == 0 and Ws <= 128 and os.environ.get("LIPREAD_IGEMM", "1") == "1")
        # the same for stride 2 on even input sizes (BasicBlock conv1 of layer2 / 3 / 4): forward and weight gradient read
        # the input through the 5-D parity view, the input gradient goes through the zero-stuffed dy
        igemm_s2 = (hh and tap and kh == 3 and kw == 3 and st_ == 2 and pd == 1 and pw == 1 and Cin % 64 == 0
                    and Cout % 64 == 0 and Hs % 2 == 0 and Ws % 2 == 0 and Ws <= 256 and conv.bias is None and act == ACT_NONE
                    and os.environ.get("LIPREAD_IGEMM", "1") == "1" and os.environ.get("LIPREAD_IGEMM_S2", "1") == "1")
        if pointwise:
            col = xptr
        elif igemm or igemm_s2:
            col = None
        elif tap:
            col = self.alloc(rows * K, adt)
            self.fwd.add("lr_im2col_tap" + sfx, xptr, F, Hs, Ws, Cin, kh, kw, st_, pd, pw, 0, Ho, Wo, col)
        else:
            col = self.alloc(rows * ldk, adt)
            self.fwd.add("lr_im2col" + sfx, xptr, *src, Hs, Ws, Cin, kh, kw, st_, pd, 0, Ho, Wo, col, ldk)
        wop_tap = None
        if tap and hh:                                         # tap-major bf16 operand in one launch
            wmat = None
            wop_tap = self.alloc(Cout * K, torch.bfloat16)
            self.weight_taps.append((conv.weight, wop_tap, Cout, Cin, kh * kw, 0))     # one batched launch per step
        elif tap:
            wmat = self.alloc(Cout * K)
            self.fwd.add("lr_weight_tap", conv.weight, wmat, Cout, Cin, kh * kw, 0)
        elif ldk != K:
            wmat = self.alloc(Cout * ldk)
            wmat.zero_()
            self.fwd.add("lr_copy2d", wmat, ldk, conv.weight, K, Cout, K)
        else:
            wmat = conv.weight
        if wop_tap is not None:
            wop = wop_tap
        elif hh:                                               # the GEMM's B operand in bf16
            wop = self.wh(conv.weight) if wmat is conv.weight else self.cast_h(self.fwd, wmat)
        else:
            wop = wmat
        y = T2(self, F, Ho, Wo, Cout, h=hh)
        stt_ = 0
        if with_stats:
            y.stat_slot = self.stat_slot(Cout)
            stt_ = (lambda s=y.stat_slot: s["fwd"]) if self.training else 0
        y._act = act
        if igemm and (conv.bias is not None or act != ACT_NONE):
            raise NotImplementedError("implicit-GEMM conv: bias / fused activation (the ResNet convs have neither)")
        if igemm:
            self.fwd.add("lr_conv3x3_bf16", xptr, wop, y.val, 0, stt_, F, Hs, Ws, Cin, Cout, 0)
        elif igemm_s2:
            self.fwd.add("lr_conv3x3s2_bf16", xptr, wop, y.val, stt_, F, Hs, Ws, Cin, Cout)
        else:
            self.gemm_auto(self.fwd, col, ldk, 0, wop, ldk, 0, y.val, Cout, rows, Cout, ldk,
                           bias=(conv.bias if conv.bias is not None else 0), act=act, stats=stt_, h=hh)
        if self.with_backward:
            g = self.bgroup()
            y._conv_bwd = (g, x, conv, col, ldk, wmat, src, (F, Hs, Ws, Ho, Wo), need_dx and frames is None, tap)
            y._igemm = igemm
            y._igemm_s2 = igemm_s2
        return y

    def dense_conv_bwd(self, y, dx_residual=0):
        """Emit the backward of a dense_conv output y (after everything that writes y.grad has been registered):
        dW += dy^T col, db += colsum(dy), x.grad = im2col_T(dy) . Wt^T (+ dx_residual)."""
        g, x, conv, col, ldk, wmat, src, (F, Hs, Ws, Ho, Wo), need_dx, tap = y._conv_bwd
        Cout, Cin = conv.out_channels, conv.in_channels
        kh, kw, st_, pd, pw = _conv_geometry(conv)
        K = Cin * kh * kw
        rows = F * Ho * Wo
        hh = y.h
        sfx = "_h" if hh else ""
        dw = self.flat.g(conv.weight)
        if getattr(y, "_act", ACT_NONE) != ACT_NONE:          # fused activation: dy *= act'(y) first
            g.add("lr_act_bwd" + sfx, y.grad, y.val, rows * Cout, y._act)
        igemm = getattr(y, "_igemm", False)
        if tap:
            dwp = self.alloc(Cout * K)                              # gradient in the tap-major layout, then back to torch's
            g.add("lr_memset", dwp, Cout * K * 4, leaf=True)
            if igemm:
                g.add("lr_conv3x3_wgrad_bf16", y.grad, x.val, dwp, F, Hs, Ws, Cin, Cout, leaf=True)
            elif getattr(y, "_igemm_s2", False):
                g.add("lr_conv3x3s2_wgrad_bf16", y.grad, x.val, dwp, F, Hs, Ws, Cin, Cout, leaf=True)
            else:
                self.gemm_auto(g, y.grad, Cout, 1, col, K, 1, dwp, K, Cout, K, rows, split_ok=True, h=hh)
            g.add("lr_weight_tap", dwp, dw, Cout, Cin, kh * kw, 2, leaf=True)
        elif ldk != K:
            dwp = self.alloc(Cout * ldk)
            g.add("lr_memset", dwp, Cout * ldk * 4, leaf=True)
            self.gemm_auto(g, y.grad, Cout, 1, col, ldk, 1, dwp, ldk, Cout, ldk, rows, split_ok=True, h=hh)
            g.add("lr_copy2d", dw, K, dwp, ldk, Cout, K, leaf=True)
        else:
            self.gemm_auto(g, y.grad, Cout, 1, col, ldk, 1, dw, ldk, Cout, ldk, rows, split_ok=True, h=hh)
        if conv.bias is not None:
            g.add("lr_colsum" + sfx, y.grad, Cout, rows, Cout, self.flat.g(conv.bias), leaf=True)
        if not need_dx:
            return
        rows_in = F * Hs * Ws
        if kh == 1 and kw == 1 and st_ == 1 and pd == 0:
            self.gemm_auto(g, y.grad, Cout, 0, (self.wh(conv.weight) if hh else conv.weight), Cin, 1, x.grad, Cin, rows_in, Cin,
                           Cout, R=dx_residual, ldr=Cin, h=hh)
            return
        Kt = Cout * kh * kw
        if tap:
            if hh:
                wt_op = self.alloc(Cin * Kt, torch.bfloat16)
                self.weight_taps.append((conv.weight, wt_op, Cout, Cin, kh * kw, 1))
            else:
                wt_op = self.alloc(Cin * Kt)
                g.add("lr_weight_tap", conv.weight, wt_op, Cout, Cin, kh * kw, 1)
            if igemm:                                             # dx = conv(dy, mirrored taps) (+ residual): same kernel
                g.add("lr_conv3x3_bf16", y.grad, wt_op, x.grad, dx_residual, 0, F, Hs, Ws, Cout, Cin, 1)
                return
            if (hh and kh == 3 and kw == 3 and st_ == 2 and pd == 1 and pw == 1 and Cin % 64 == 0 and Cout % 64 == 0
                    and Ws <= 128 and os.environ.get("LIPREAD_IGEMM", "1") == "1"):
                # stride 2: dy zero-stuffed onto the input grid, then the same stride-1 implicit-GEMM kernel
                up = self.workspace(rows_in * Cout, h=True)
                g.add("lr_zero_stuff2_h", y.grad, up, F, Ho, Wo, Hs, Ws, Cout)
                g.add("lr_conv3x3_bf16", up, wt_op, x.grad, dx_residual, 0, F, Hs, Ws, Cout, Cin, 1)
                return
            colT = self.workspace(rows_in * Kt, h=hh)
            g.add("lr_im2col_tap" + sfx, y.grad, F, Ho, Wo, Cout, kh, kw, st_, pd, pw, 1, Hs, Ws, colT)
            self.gemm_auto(g, colT, Kt, 0, wt_op, Kt, 0, x.grad, Cin, rows_in, Cin, Kt, R=dx_residual, ldr=Cin, h=hh)
            return
        if hh:
            raise NotImplementedError("bf16 storage: input gradient of a dense conv needs the tap-major path")
        ldt = (Kt + 3) // 4 * 4
        wt = self.alloc(Cin * ldt)
        wt.zero_()
        g.add("lr_weight_transpose", conv.weight, wt, Cout, Cin, kh * kw, ldt)
        colT = self.workspace(rows_in * ldt)
        dsrc = (0, 1.0, F, 1, Ho * Wo * Cout, 0, 1, Wo * Cout, Cout)
        g.add("lr_im2col", y.grad, *dsrc, Ho, Wo, Cout, kh, kw, st_, pd, 1, Hs, Ws, colT, ldt)
        self.gemm_auto(g, colT, ldt, 0, wt, ldt, 0, x.grad, Cin, rows_in, Cin, ldt, R=dx_residual, ldr=Cin)

    def maxpool(self, x, k, stride, pad):
        Ho, Wo = (x.H + 2 * pad - k) // stride + 1, (x.W + 2 * pad - k) // stride + 1
        y = T2(self, x.F, Ho, Wo, x.C, h=x.h)
        arg = self.alloc(y.rows * x.C, torch.uint8)
        sfx = "_h" if x.h else ""
        self.fwd.add("lr_maxpool_fwd" + sfx, x.val, y.val, arg, x.F, x.H, x.W, x.C, k, stride, pad)
        if self.with_backward:
            self.bgroup().add("lr_maxpool_bwd" + sfx, y.grad, arg, x.grad, x.F, x.H, x.W, x.C, k, stride, pad)
        return y

    def dropout_active(self, p):
        return self.training and p > 0.0

    def dropout(self, x, dx, n, p, dy=None):
        """nn.Dropout(p) on a flat buffer of n floats (training plans only; identity otherwise).
        Returns (y, dy): dy is the buffer the consumer's backward must write (allocated here unless given)."""
        if not self.dropout_active(p):
            return x, dx
        if self.rng_step is None:
            # ONE device counter per model (shared by all its plans, saved in checkpoints): a new batch shape or a
            # resumed run continues the mask sequence instead of restarting it
            if self.flat.rng_step is None:
                self.flat.rng_step = torch.zeros(1, dtype=torch.int64, device=self.dev)
            self.rng_step = self.flat.rng_step
        self._n_dropout += 1
        y, mask = self.alloc(n), self.alloc(n, torch.uint8)
        self.fwd.add("lr_dropout_fwd", x, y, mask, n, float(p), dropout_seed(self._n_dropout), self.rng_step)
        if self.with_backward:
            dy = self.alloc(n) if dy is None else dy
            if dx is not None:
                self.bgroup().add("lr_dropout_bwd", dy, mask, dx, n, float(p))
        else:
            dy = None
        return y, dy

    def cnn_sequential(self, mods, frames, h=None):
        """nn.Sequential of Conv2d(3x3, padding 1) [+ BatchNorm2d] + ReLU, MaxPool2d(2) and a closing
        AdaptiveAvgPool2d on frames = (tensor, layout, scale): the audio encoders of audio_video/models/*.py, VGGLite
        (video/models/vgg_lstm.py:21-41), torchvision vgg*_bn features (audio/models/vgg_model.py:11-13).
        Returns ("pooled", feat, dfeat, C) after AdaptiveAvgPool2d(1) or ("map", T2) when it ends on a feature map."""
        cur, pooled = None, None
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv2d):
                has_bn = isinstance(mods[i + 1], nn.BatchNorm2d)
                if has_bn:
                    assert isinstance(mods[i + 2], nn.ReLU)
                    raw = self.dense_conv(cur, m, frames=frames if cur is None else None, h=h)
                    if self.with_backward:
                        self.dense_conv_bwd(raw)
                    a = T2(self, raw.F, raw.H, raw.W, raw.C, h=raw.h)
                    self.bn_act(raw, mods[i + 1], ACT_RELU, a)
                    cur = a
                    i += 3                                       # conv, bn, relu
                else:
                    assert isinstance(mods[i + 1], nn.ReLU)
                    cur = self.dense_conv(cur, m, frames=frames if cur is None else None, act=ACT_RELU, with_stats=False, h=h)
                    if self.with_backward:
                        self.dense_conv_bwd(cur)
                    i += 2                                       # conv, relu
            elif isinstance(m, nn.MaxPool2d):
                k = m.kernel_size if isinstance(m.kernel_size, int) else m.kernel_size[0]
                st = m.stride if isinstance(m.stride, int) else m.stride[0]
                pd = m.padding if isinstance(m.padding, int) else m.padding[0]
                cur = self.maxpool(cur, k, st, pd)
                i += 1
            elif isinstance(m, nn.AdaptiveAvgPool2d):
                osz = m.output_size if isinstance(m.output_size, tuple) else (m.output_size, m.output_size)
                if tuple(osz) == (1, 1):
                    pooled = self.avgpool(cur)
                elif tuple(osz) != (cur.H, cur.W):
                    raise NotImplementedError(f"AdaptiveAvgPool2d{tuple(osz)} on a {cur.H}x{cur.W} map")
                i += 1                                           # output size == input size: identity
            else:
                raise NotImplementedError(type(m).__name__)
        if pooled is not None:
            return "pooled", pooled[0], pooled[1], cur.C
        return "map", cur

    # ---- nn.LSTM (batch_first, bidirectional) with an out[:, -1] head -------------------------------------
    def bilstm_last(self, x, dx, I, B, T, lstm, out, ldo, dout):
        """x: [B*T, I] features (dx: its gradient buffer, written here).  Writes out[b, 0:2H] (row stride ldo) =
        lstm(x)[:, -1]: all layers below the top run both directions over the whole sequence; the top layer runs
        its forward direction over T steps and its reverse direction for ONE step.
        dout: gradient of that row (same stride), read by the backward."""
        assert lstm.bidirectional and lstm.batch_first
        H, L = lstm.hidden_size, lstm.num_layers
        F, G4 = B * T, 4 * H
        p_drop = float(lstm.dropout)

        def par(name, layer, rev):
            return getattr(lstm, f"{name}_l{layer}{'_reverse' if rev else ''}")

        cur, dcur, Icur = x, dx, I
        for l in range(L - 1):
            seq, dseq = self._bilstm_full_layer(cur, dcur, Icur, B, T, lstm, l)
            cur, dcur = self.dropout(seq, dseq, F * 2 * H, p_drop)
            Icur = 2 * H
        # ---- top layer
        l = L - 1
        xp_f, hs_f = self.alloc(F * G4), self.alloc(F * H)
        gates_f, c_f, hp_f = self.alloc(F * G4), self.alloc(F * H), self.alloc(F * H)
        optr = out if isinstance(out, int) else out.data_ptr()
        xp_r, gates_r, c_r = self.alloc(B * G4), self.alloc(B * G4), self.alloc(B * H)
        cur_last = (cur if isinstance(cur, int) else cur.data_ptr()) + 4 * (T - 1) * Icur
        # the reverse direction's single step depends on the features only: a side branch next to the forward walk
        with self.fwd.side_branch():
            self.linear(cur_last, T * Icur, B, par("weight_ih", l, 1), par("bias_ih", l, 1), xp_r, G4)
            self.fwd.add("lr_lstm_fwd", xp_r, G4, par("bias_hh", l, 1), par("weight_hh", l, 1), optr + 4 * H, ldo, gates_r, c_r, 0,
                         B, 1, H, 1, 1)
        self.linear(cur, Icur, F, par("weight_ih", l, 0), par("bias_ih", l, 0), xp_f, G4)
        self.fwd.add(self.lstm_op("fwd", H, T), xp_f, G4, par("bias_hh", l, 0), par("weight_hh", l, 0), hs_f, H, gates_f, c_f, hp_f,
                     B, T, H, T, 0)
        self.fwd.add("lr_copy2d", optr, ldo, hs_f.data_ptr() + 4 * (T - 1) * H, T * H, B, H)
        self.fwd.join()
        if self.with_backward:
            dptr = dout if isinstance(dout, int) else dout.data_ptr()
            dg_f, dg_r = self.alloc(F * G4), self.alloc(B * G4)
            g = self.bgroup()
            g.add(self.lstm_op("bwd", H, T), dptr, ldo, T - 1, gates_f, c_f, par("weight_hh", l, 0), dg_f, B, T, H, T, 0)
            self.linear_bwd(g, hp_f, H, F, par("weight_hh", l, 0), par("bias_hh", l, 0), dg_f, G4)
            self.linear_bwd(g, cur, Icur, F, par("weight_ih", l, 0), par("bias_ih", l, 0), dg_f, G4, dx=dcur, ldx=Icur)
            # reverse direction: one step from the zero state (W_hh_reverse gets no gradient)
            g.add("lr_lstm_bwd", dptr + 4 * H, ldo, 0, gates_r, c_r, par("weight_hh", l, 1), dg_r, B, 1, H, 1, 1)
            g.add("lr_colsum", dg_r, G4, B, G4, self.flat.g(par("bias_hh", l, 1)), leaf=True)
            if dcur is None:                                   # the sequence input needs no gradient (raw features)
                self.linear_bwd(g, cur_last, T * Icur, B, par("weight_ih", l, 1), par("bias_ih", l, 1), dg_r, G4)
            else:
                dcur_last = (dcur if isinstance(dcur, int) else dcur.data_ptr()) + 4 * (T - 1) * Icur
                self.linear_bwd(g, cur_last, T * Icur, B, par("weight_ih", l, 1), par("bias_ih", l, 1), dg_r, G4,
                                dx=dcur_last, ldx=T * Icur, dx_residual=dcur_last, ldr=T * Icur)

    def _bilstm_full_layer(self, cur, dcur, Icur, B, T, lstm, l):
        """Layer l of a bidirectional nn.LSTM over the whole sequence, both directions -> (seq [B*T, 2H], dseq)."""
        H = lstm.hidden_size
        F, G4 = B * T, 4 * H

        def par(name, rev):
            return getattr(lstm, f"{name}_l{l}{'_reverse' if rev else ''}")
        seq = self.alloc(F * 2 * H)
        dseq = self.alloc(F * 2 * H) if self.with_backward else None
        saved = []
        # the two directions are independent walks of T sequential steps: the reverse one runs as a side branch of the
        # step graph, concurrently with the forward one (each is latency-bound on a fraction of the SMs)
        for rev in (1, 0):
            xp, gates, cst, hp = self.alloc(F * G4), self.alloc(F * G4), self.alloc(F * H), self.alloc(F * H)
            with (self.fwd.side_branch() if rev else contextlib.nullcontext()):
                self.linear(cur, Icur, F, par("weight_ih", rev), par("bias_ih", rev), xp, G4)
                self.fwd.add(self.lstm_op("fwd", H, T), xp, G4, par("bias_hh", rev), par("weight_hh", rev),
                             seq.data_ptr() + 4 * H * rev, 2 * H, gates, cst, hp, B, T, H, T, rev)
            saved.append((gates, cst, hp))
        self.fwd.join()
        saved.reverse()                                          # index by direction again
        if self.with_backward:
            g = self.bgroup()
            dgs = [self.alloc(F * G4), self.alloc(F * G4)]
            for rev in (1, 0):
                gates, cst, hp = saved[rev]
                with (g.side_branch() if rev else contextlib.nullcontext()):
                    g.add(self.lstm_op("bwd", H, T), dseq.data_ptr() + 4 * H * rev, 2 * H, -1, gates, cst, par("weight_hh", rev), dgs[rev],
                          B, T, H, T, rev)
            g.join()
            for rev in (0, 1):
                gates, cst, hp = saved[rev]
                self.linear_bwd(g, hp, H, F, par("weight_hh", rev), par("bias_hh", rev), dgs[rev], G4)
                self.linear_bwd(g, cur, Icur, F, par("weight_ih", rev), par("bias_ih", rev), dgs[rev], G4,
                                dx=dcur, ldx=Icur, dx_residual=(dcur if rev else 0), ldr=Icur)
        return seq, dseq

    def bilstm_seq(self, x, dx, I, B, T, lstm):
        """Every layer of a bidirectional nn.LSTM over the whole sequence -> (out [B*T, 2H], dout): the heads that read
        all time steps (audio/models/lstm_resnet_attn_model.py:78-81)."""
        assert lstm.bidirectional and lstm.batch_first
        H, L = lstm.hidden_size, lstm.num_layers
        cur, dcur, Icur = x, dx, I
        for l in range(L):
            cur, dcur = self._bilstm_full_layer(cur, dcur, Icur, B, T, lstm, l)
            if l < L - 1:
                cur, dcur = self.dropout(cur, dcur, B * T * 2 * H, float(lstm.dropout))
            Icur = 2 * H
        return cur, dcur

    def bilstm_hn(self, x, dx, I, B, T, lstm, out, ldo, dout):
        """Single-layer bidirectional nn.LSTM whose head is cat(h_n[0], h_n[1]) (audio_video/models/late_fusion.py:61-62,
        early_fusion_fast.py:53-54): BOTH directions walk all T steps; h_n[0] = forward output at t = T-1, h_n[1] =
        reverse output at t = 0.  Writes out[b, 0:2H] (row stride ldo); dout is the gradient of that row."""
        assert lstm.bidirectional and lstm.batch_first and lstm.num_layers == 1
        H = lstm.hidden_size
        F, G4 = B * T, 4 * H

        def par(name, rev):
            return getattr(lstm, f"{name}_l0{'_reverse' if rev else ''}")

        optr = out if isinstance(out, int) else out.data_ptr()
        saved = []
        for rev in (0, 1):
            xp, hs = self.alloc(F * G4), self.alloc(F * H)
            gates, cst, hp = self.alloc(F * G4), self.alloc(F * H), self.alloc(F * H)
            self.linear(x, I, F, par("weight_ih", rev), par("bias_ih", rev), xp, G4)
            self.fwd.add(self.lstm_op("fwd", H, T), xp, G4, par("bias_hh", rev), par("weight_hh", rev), hs, H, gates, cst, hp,
                         B, T, H, T, rev)
            t_final = 0 if rev else T - 1
            self.fwd.add("lr_copy2d", optr + 4 * H * rev, ldo, hs.data_ptr() + 4 * t_final * H, T * H, B, H)
            saved.append((gates, cst, hp, t_final))
        if self.with_backward:
            dptr = dout if isinstance(dout, int) else dout.data_ptr()
            g = self.bgroup()
            for rev in (0, 1):
                gates, cst, hp, t_final = saved[rev]
                dg = self.alloc(F * G4)
                g.add(self.lstm_op("bwd", H, T), dptr + 4 * H * rev, ldo, t_final, gates, cst, par("weight_hh", rev), dg, B, T, H, T, rev)
                self.linear_bwd(g, hp, H, F, par("weight_hh", rev), par("bias_hh", rev), dg, G4)
                self.linear_bwd(g, x, I, F, par("weight_ih", rev), par("bias_ih", rev), dg, G4, dx=dx, ldx=I,
                                dx_residual=(dx if rev else 0), ldr=I)

    # ---- nn.MultiheadAttention (self-attention over time) ---------------------------------------------------
    def multihead_attention(self, x, dx, B, T, mha, dx_residual=0, dout=None):
        """nn.MultiheadAttention(embed_dim, heads, dropout, batch_first=True)(x, x, x)[0] on x [B*T, E] -> (out, dout).
        In-projection GEMM -> lr_mha_scores_fwd -> [dropout on the weights] -> lr_mha_apply_fwd -> out-projection GEMM
        (video/models/resnet_attn.py:23-35)."""
        E, heads = mha.embed_dim, mha.num_heads
        assert mha.batch_first and mha._qkv_same_embed_dim and mha.in_proj_bias is not None
        F, wb = B * T, self.with_backward
        qkv = self.alloc(F * 3 * E)
        dqkv = self.alloc(F * 3 * E) if wb else None
        self.linear(x, E, F, mha.in_proj_weight, mha.in_proj_bias, qkv, 3 * E)
        if wb:
            self.linear_bwd(self.bgroup(), x, E, F, mha.in_proj_weight, mha.in_proj_bias, dqkv, 3 * E, dx=dx, ldx=E,
                            dx_residual=dx_residual, ldr=E)
        n_p = B * heads * T * T
        P = self.alloc(n_p)
        dP = self.alloc(n_p) if wb else None
        self.fwd.add("lr_mha_scores_fwd", qkv, 3 * E, P, B, T, E, heads)
        if wb:
            self.bgroup().add("lr_mha_scores_bwd", P, dP, qkv, 3 * E, dqkv, B, T, E, heads)
        Pd, dPd = self.dropout(P, dP, n_p, float(mha.dropout))
        O = self.alloc(F * E)
        dO = self.alloc(F * E) if wb else None
        self.fwd.add("lr_mha_apply_fwd", Pd, qkv, 3 * E, O, B, T, E, heads)
        if wb:
            self.bgroup().add("lr_mha_apply_bwd", dO, Pd, qkv, 3 * E, dPd, dqkv, B, T, E, heads)
        out = self.alloc(F * E)
        dout = (self.alloc(F * E) if dout is None else dout) if wb else None
        self.linear(O, E, F, mha.out_proj.weight, mha.out_proj.bias, out, E)
        if wb:
            self.linear_bwd(self.bgroup(), O, E, F, mha.out_proj.weight, mha.out_proj.bias, dout, E, dx=dO, ldx=E)
        return out, dout

    # ---- nn.TransformerEncoderLayer (post-norm, ReLU) ---------------------------------------------------------
    def layer_norm(self, a, b, ln, rows, ds):
        """y = LayerNorm(a + b); the backward writes the gradient of the sum (of a and of b alike) into `ds`."""
        (D,) = ln.normalized_shape
        y, sm, st = self.alloc(rows * D), self.alloc(rows * D), self.alloc(rows * 2)
        dy = self.alloc(rows * D) if self.with_backward else None
        self.fwd.add("lr_layernorm_fwd", a, (b if b is not None else 0), ln.weight, ln.bias, float(ln.eps), y, sm, st, rows, D)
        if self.with_backward:
            self.bgroup().add("lr_layernorm_bwd", dy, sm, st, ln.weight, ds, self.flat.g(ln.weight), self.flat.g(ln.bias), rows, D)
        return y, dy

    def transformer_encoder_layer(self, x, dx, B, T, layer):
        """nn.TransformerEncoderLayer(batch_first=True, norm_first=False, activation=relu) on x [B*T, E]:
             h = norm1(x + dropout1(self_attn(x)));  y = norm2(h + dropout2(linear2(dropout(relu(linear1(h))))))
        (video/models/resnet_trans.py:96-103).  dx receives dL/dx.  Returns (y, dy)."""
        if layer.norm_first or getattr(layer.activation, "__name__", "") != "relu":
            raise NotImplementedError("only the post-norm ReLU TransformerEncoderLayer the reference builds")
        E, Dff = layer.self_attn.embed_dim, layer.linear1.out_features
        F, wb = B * T, self.with_backward
        # -- self-attention block: ds1 is the gradient of (x + dropout1(sa)), shared by the residual and the branch
        ds1 = self.alloc(F * E) if wb else None
        p1 = float(layer.dropout1.p)
        dsa = (self.alloc(F * E) if self.dropout_active(p1) else ds1) if wb else None
        sa, _ = self.multihead_attention(x, dx, B, T, layer.self_attn, dx_residual=(ds1 if wb else 0), dout=dsa)
        sa_d, _ = self.dropout(sa, dsa, F * E, p1, dy=ds1)
        h, dh = self.layer_norm(x, sa_d, layer.norm1, F, ds1)
        # -- feed-forward block: ds2 is the gradient of (h + dropout2(ff))
        ds2 = self.alloc(F * E) if wb else None
        f1 = self.alloc(F * Dff)
        df1 = self.alloc(F * Dff) if wb else None
        self.linear(h, E, F, layer.linear1.weight, layer.linear1.bias, f1, Dff, act=ACT_RELU)
        if wb:
            g = self.bgroup()
            g.add("lr_act_bwd", df1, f1, F * Dff, ACT_RELU)
            self.linear_bwd(g, h, E, F, layer.linear1.weight, layer.linear1.bias, df1, Dff, dx=dh, ldx=E, dx_residual=ds2, ldr=E)
        f1d, df1d = self.dropout(f1, df1, F * Dff, float(layer.dropout.p))
        p2 = float(layer.dropout2.p)
        f2 = self.alloc(F * E)
        df2 = (self.alloc(F * E) if self.dropout_active(p2) else ds2) if wb else None
        self.linear(f1d, Dff, F, layer.linear2.weight, layer.linear2.bias, f2, E)
        if wb:
            self.linear_bwd(self.bgroup(), f1d, Dff, F, layer.linear2.weight, layer.linear2.bias, df2, E, dx=df1d, ldx=Dff)
        f2d, _ = self.dropout(f2, df2, F * E, p2, dy=ds2)
        return self.layer_norm(h, f2d, layer.norm2, F, ds2)

    # ---- torchvision ResNet (BasicBlock / Bottleneck) ------------------------------------------------------
    def resnet_features(self, net, frames, x=None):
        """torchvision resnet18/34/50 children()[:-2] (conv1, bn1, relu, maxpool, layer1..4) on frames given as
        (tensor, layout, scale) -> last activation T2.  video/models/resnet_lstm.py:90-93,
        audio/models/resnet_model.py:12-17.  x: a T2 input instead of raw frames when the image is itself computed and
        needs a gradient (audio/models/lstm_resnet_model.py:47-49)."""
        raw = self.dense_conv(x, net.conv1, frames=frames if x is None else None)
        if self.with_backward:
            self.dense_conv_bwd(raw)
        a = T2(self, raw.F, raw.H, raw.W, raw.C, h=raw.h)
        self.bn_act(raw, net.bn1, ACT_RELU, a)
        mp = net.maxpool
        k = mp.kernel_size if isinstance(mp.kernel_size, int) else mp.kernel_size[0]
        s_ = mp.stride if isinstance(mp.stride, int) else mp.stride[0]
        p_ = mp.padding if isinstance(mp.padding, int) else mp.padding[0]
        cur = self.maxpool(a, k, s_, p_)
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            for blk in layer:
                cur = self.bottleneck_block(blk, cur) if type(blk).__name__ == "Bottleneck" else self.basic_block(blk, cur)
        return cur

    def bottleneck_block(self, blk, x):
        """torchvision Bottleneck (resnet50, video/models/resnet_lstm.py:84): relu(bn3(conv3(relu(bn2(conv2(relu(bn1(
        conv1(x)))))))) + identity), conv1 / conv3 1x1, conv2 3x3 carrying the stride, identity = x or bn_d(conv_d(x))."""
        ident, dmask, ds_raw = x, None, None
        # backward groups run in reverse registration order: bn3, conv3, bn2, conv2, bn1, conv1, [bn_d, conv_d]
        if blk.downsample is not None:
            dconv, dbn = blk.downsample[0], blk.downsample[1]
            ds_raw = self.dense_conv(x, dconv)
            ident = T2(self, ds_raw.F, ds_raw.H, ds_raw.W, ds_raw.C, h=ds_raw.h)
            self.bn_act(ds_raw, dbn, ACT_NONE, ident)
        raw1 = self.dense_conv(x, blk.conv1)
        h1 = T2(self, raw1.F, raw1.H, raw1.W, raw1.C, h=raw1.h)
        self.bn_act(raw1, blk.bn1, ACT_RELU, h1)
        raw2 = self.dense_conv(h1, blk.conv2)
        h2 = T2(self, raw2.F, raw2.H, raw2.W, raw2.C, h=raw2.h)
        self.bn_act(raw2, blk.bn2, ACT_RELU, h2)
        raw3 = self.dense_conv(h2, blk.conv3)
        out = T2(self, raw3.F, raw3.H, raw3.W, raw3.C, h=raw3.h)
        if self.with_backward:
            dmask = ident.grad if ds_raw is not None else self.alloc_like(out, out.rows * out.C)
        self.bn_act(raw3, blk.bn3, ACT_RELU, out, residual=ident, res_pre=True, dres=dmask)
        if self.with_backward:
            self.dense_conv_bwd(raw3)
            self.dense_conv_bwd(raw2)
            self.dense_conv_bwd(raw1, dx_residual=(0 if ds_raw is not None else dmask))
            if ds_raw is not None:
                self.dense_conv_bwd(ds_raw, dx_residual=x.grad)
        return out

    def basic_block(self, blk, x):
        """torchvision BasicBlock: relu(bn2(conv2(relu(bn1(conv1(x))))) + identity), identity = x or
        bn_d(conv_d(x)) when the shape changes."""
        if type(blk).__name__ != "BasicBlock":
            raise NotImplementedError(f"{type(blk).__name__} blocks have no lipread_b200 plan")
        ident, dmask = x, None
        ds_raw = None
        # Backward groups run in reverse registration order: bn2, conv2, bn1, conv1, [bn_d, conv_d].  conv1 writes
        # x.grad (plus the identity gradient when there is no downsample); conv_d then accumulates onto it.
        if blk.downsample is not None:
            dconv, dbn = blk.downsample[0], blk.downsample[1]
            ds_raw = self.dense_conv(x, dconv)
            ident = T2(self, ds_raw.F, ds_raw.H, ds_raw.W, ds_raw.C, h=ds_raw.h)
            self.bn_act(ds_raw, dbn, ACT_NONE, ident)
        raw1 = self.dense_conv(x, blk.conv1)
        h1 = T2(self, raw1.F, raw1.H, raw1.W, raw1.C, h=raw1.h)
        self.bn_act(raw1, blk.bn1, ACT_RELU, h1)
        raw2 = self.dense_conv(h1, blk.conv2)
        out = T2(self, raw2.F, raw2.H, raw2.W, raw2.C, h=raw2.h)
        if self.with_backward:
            dmask = ident.grad if ds_raw is not None else self.alloc_like(out, out.rows * out.C)
        self.bn_act(raw2, blk.bn2, ACT_RELU, out, residual=ident, res_pre=True, dres=dmask)
        if self.with_backward:
            self.dense_conv_bwd(raw2)
            self.dense_conv_bwd(raw1, dx_residual=(0 if ds_raw is not None else dmask))
            if ds_raw is not None:
                self.dense_conv_bwd(ds_raw, dx_residual=x.grad)
        return out
