"""torch.library custom ops (`torch.ops.lipread.*`) over the C ABI.  CUDA only: calling an op with
CPU tensors raises NotImplementedError from the dispatcher -- there is no CPU kernel to fall to."""
import torch

from . import _lib
from ._lib import lib, check


def _ptr(t):
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need(t, dtype, name):
    if not t.is_cuda:
        raise _lib.LipreadError(f"{name} must be a CUDA tensor (lipread_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.LipreadError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.LipreadError(f"{name} must be contiguous")


# ---------------------------------------------------------------- K1 log-mel
@torch.library.custom_op("lipread::logmel_plan", mutates_args=(), device_types="cuda")
def logmel_plan(window: torch.Tensor, fb: torch.Tensor) -> torch.Tensor:
    _need(window, torch.float32, "window")
    _need(fb, torch.float32, "fb")
    if window.numel() != 400 or tuple(fb.shape) != (201, 80):
        raise _lib.LipreadError("logmel_plan expects window[400] and fb[201,80]")
    n = lib.lr_logmel_plan_bytes()
    plan = torch.empty(n, dtype=torch.uint8, device=window.device)
    check(lib.lr_logmel_plan_init(_ptr(window), _ptr(fb), _ptr(plan), n, _stream()))
    return plan


@logmel_plan.register_fake
def _(window, fb):
    return window.new_empty(lib.lr_logmel_plan_bytes(), dtype=torch.uint8)


@torch.library.custom_op("lipread::logmel", mutates_args=(), device_types="cuda")
def logmel(wav: torch.Tensor, plan: torch.Tensor, n_out: int, mode: int) -> torch.Tensor:
    """wav (B, 20000) f32 -> (B, 80, n_out) f32.  mode 0: log-mel + normalise + crop; 1: raw log-mel."""
    _need(wav, torch.float32, "wav")
    if wav.dim() != 2 or wav.shape[1] != 20000:
        raise _lib.LipreadError(f"wav must be (B, 20000), got {tuple(wav.shape)}")
    out = torch.empty(wav.shape[0], 80, n_out, dtype=torch.float32, device=wav.device)
    check(lib.lr_logmel_fwd(_ptr(wav), _ptr(plan), _ptr(out), wav.shape[0], n_out, mode, _stream()))
    return out


@logmel.register_fake
def _(wav, plan, n_out, mode):
    return wav.new_empty(wav.shape[0], 80, n_out)


@torch.library.custom_op("lipread::normalize", mutates_args=(), device_types="cuda")
def normalize(x: torch.Tensor) -> torch.Tensor:
    """Row-wise (x - mean) / (std_unbiased + 1e-9) over all but the first dimension."""
    _need(x, torch.float32, "x")
    out = torch.empty_like(x)
    b = x.shape[0]
    check(lib.lr_normalize_fwd(_ptr(x), _ptr(out), b, x.numel() // max(b, 1), _stream()))
    return out


@normalize.register_fake
def _(x):
    return torch.empty_like(x)


# ---------------------------------------------------------------- PCM ingestion (in front of K1)
@torch.library.custom_op("lipread::pcm_ingest", mutates_args=(), device_types="cuda")
def pcm_ingest(pcm: torch.Tensor, offset: torch.Tensor, n_frames: torch.Tensor, channels: torch.Tensor, scale: float,
               target: int) -> torch.Tensor:
    """Packed int16 PCM (ragged clips, interleaved channels) -> (B, target) f32: mono mean, truncate / right zero-pad
    (audio/utils/audio_processor.py:29,37,40-44).  channels: int32 [B], or an empty tensor for all-mono."""
    _need(pcm, torch.int16, "pcm")
    _need(offset, torch.int64, "offset")
    _need(n_frames, torch.int32, "n_frames")
    B = offset.numel()
    if n_frames.numel() != B or channels.numel() not in (0, B):
        raise _lib.LipreadError("pcm_ingest: offset, n_frames and channels must have one entry per clip")
    if channels.numel():
        _need(channels, torch.int32, "channels")
    wav = torch.empty(B, target, dtype=torch.float32, device=pcm.device)
    check(lib.lr_pcm_ingest(_ptr(pcm), _ptr(offset), _ptr(n_frames), _ptr(channels) if channels.numel() else None,
                            scale, _ptr(wav), B, target, _stream()))
    return wav


@pcm_ingest.register_fake
def _(pcm, offset, n_frames, channels, scale, target):
    return pcm.new_empty(offset.numel(), target, dtype=torch.float32)


# ---------------------------------------------------------------- nn.Linear (+ ReLU) with autograd
@torch.library.custom_op("lipread::linear", mutates_args=(), device_types="cuda")
def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, act: int) -> torch.Tensor:
    """y = act(x @ weight^T + bias) for x [M, K], weight [N, K], bias [N] (empty tensor: no bias); act 0 none, 1 ReLU
    -- nn.Linear / a 1x1 convolution on channels-last rows, e.g. the classifier of
    audio_video/models/middle_fusion_fast.py:20-25.  One lr_gemm launch with the bias / activation in its epilogue."""
    from . import kernels as K
    for t, n in ((x, "x"), (weight, "weight")):
        _need(t, torch.float32, n)
    if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1] or act not in (0, 1):
        raise _lib.LipreadError(f"linear: x {tuple(x.shape)}, weight {tuple(weight.shape)}, act {act}")
    out = torch.empty(x.shape[0], weight.shape[0], dtype=torch.float32, device=x.device)
    K.linear_fwd(x, weight, out, bias=bias if bias.numel() else None, act=act)
    return out


@linear.register_fake
def _(x, weight, bias, act):
    return x.new_empty(x.shape[0], weight.shape[0])


@torch.library.custom_op("lipread::linear_bwd", mutates_args=(), device_types="cuda")
def linear_bwd(dy: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, y: torch.Tensor, act: int,
               has_bias: bool) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(dx, dweight, dbias) of lipread::linear: lr_act_bwd through the output, then the dgrad / wgrad GEMMs and the
    bias column sum -- the launches the step graphs use for every Linear."""
    from . import kernels as K
    g = dy.contiguous().clone()
    if act:
        K.act_bwd(g, y, g.numel(), act)
    dx = torch.empty_like(x)
    K.linear_dgrad(g, weight, dx)
    dw = torch.zeros_like(weight)
    K.linear_wgrad(g, x, dw)
    db = torch.zeros(weight.shape[0] if has_bias else 0, dtype=torch.float32, device=x.device)
    if has_bias:
        K.colsum(g, g.shape[1], g.shape[0], g.shape[1], db)
    return dx, dw, db


@linear_bwd.register_fake
def _(dy, x, weight, y, act, has_bias):
    return torch.empty_like(x), torch.empty_like(weight), x.new_empty(weight.shape[0] if has_bias else 0)


def _linear_setup(ctx, inputs, output):
    x, weight, bias, act = inputs
    ctx.save_for_backward(x, weight, output)
    ctx.act, ctx.has_bias = act, bool(bias.numel())


def _linear_backward(ctx, dy):
    x, weight, y = ctx.saved_tensors
    dx, dw, db = linear_bwd(dy, x, weight, y, ctx.act, ctx.has_bias)
    return dx, dw, (db if ctx.has_bias else None), None


linear.register_autograd(_linear_backward, setup_context=_linear_setup)
