"""Per-kernel parity on the GPU, through the C ABI, against the torch CPU ops the reference dispatches to
(nn.Conv2d / BatchNorm2d / LSTM / Linear / CrossEntropyLoss / Adam in fp32).  Tolerances are fp32 round-off:
the kernels compute in fp32 with a different summation order than MKL/oneDNN."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu

ACTS = {0: lambda u: u, 1: Fn.relu, 2: Fn.hardswish, 3: Fn.hardsigmoid}


@pytest.fixture(scope="module")
def K(cuda_device):
    from multimodal_lipread_b200 import kernels
    return kernels


def _close(a, b, rtol=1e-4, atol=None):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    scale = b.abs().max().item() + 1e-30
    atol = rtol * scale if atol is None else atol
    err = (a - b).abs().max().item()
    assert err <= atol, f"max err {err:.3e} > {atol:.3e} (scale {scale:.3e})"


def _cl(x):   # NCHW -> channels-last rows [F*H*W, C]
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("M,N,K_,at,bt", [(300, 72, 16, 0, 0), (129, 40, 96, 0, 1), (88, 24, 1000, 1, 1),
                                          (64, 64, 64, 0, 0), (7, 5, 3, 0, 0), (33, 130, 50, 1, 0)])
def test_gemm_layouts(K, M, N, K_, at, bt):
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K_, generator=g)
    B = torch.randn(N, K_, generator=g)
    ref = A @ B.t()
    Ad = (A.t().contiguous() if at else A).cuda()
    Bd = (B.t().contiguous() if bt else B).cuda()
    C = torch.empty(M, N, device="cuda")
    K.gemm(Ad, M if at else K_, at, Bd, N if bt else K_, bt, C, N, M, N, K_)
    _close(C, ref)


def test_gemm_epilogues_and_splitk(K):
    g = torch.Generator().manual_seed(3)
    M, N, Kd = 500, 88, 24
    A, B, bias, R = torch.randn(M, Kd, generator=g), torch.randn(N, Kd, generator=g), torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    for act in (0, 1, 3):
        C = torch.empty(M, N, device="cuda")
        stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
        K.gemm(A.cuda(), Kd, 0, B.cuda(), Kd, 0, C, N, M, N, Kd, bias=bias.cuda(), act=act, R=R.cuda(), ldr=N, stats=stats)
        ref = ACTS[act](A @ B.t() + bias) + R
        _close(C, ref)
        _close(stats[:N], ref.double().sum(0), rtol=1e-5)
        _close(stats[N:], (ref.double() ** 2).sum(0), rtol=1e-5)
    # accumulate (R aliases C) and strided C
    Cbig = torch.randn(M, 2 * N, generator=g).cuda()
    base = Cbig.clone()
    view = Cbig[:, N:]
    K.gemm(A.cuda(), Kd, 0, B.cuda(), Kd, 0, view, 2 * N, M, N, Kd, R=view, ldr=2 * N)
    _close(Cbig[:, N:], base[:, N:].cpu() + A @ B.t())
    assert torch.equal(Cbig[:, :N], base[:, :N])
    # split-K: wgrad-shaped (long reduction) accumulating atomically onto an existing value
    Mr = 5000
    dY, X = torch.randn(Mr, 40, generator=g), torch.randn(Mr, 96, generator=g)
    dW0 = torch.randn(40, 96, generator=g)
    dW = dW0.clone().cuda()
    K.gemm(dY.cuda(), 40, 1, X.cuda(), 96, 1, dW, 96, 40, 96, Mr, ksplit=13)
    _close(dW, dW0 + dY.t() @ X)
    # split-K with bias (audio_fc-shaped)
    x, w, b = torch.randn(8, 3000, generator=g), torch.randn(128, 3000, generator=g), torch.randn(128, generator=g)
    out = torch.zeros(8, 128, device="cuda")
    K.linear_fwd(x.cuda(), w.cuda(), out, bias=b.cuda(), ksplit=7)
    _close(out, x @ w.t() + b)


@pytest.mark.parametrize("u8,size", [(True, 44), (False, 44), (True, 30)])
def test_stem_conv(K, u8, size):
    from multimodal_lipread_b200 import synthetic
    B, T = 2, 5
    lips = synthetic.make_lips_u8(B, size=size)[:, :T].contiguous()            # (B,T,H,W,3) u8
    video = (lips.float() / 255.0).permute(0, 4, 1, 2, 3).contiguous()        # (B,3,T,H,W) f32 (reference input)
    conv = nn.Conv2d(3, 16, 3, 2, 1, bias=False)
    frames = video.permute(0, 2, 1, 3, 4).reshape(B * T, 3, size, size)
    ref = conv(frames)
    Ho = ref.shape[2]
    if u8:
        x = lips.cuda()
        layout = (1, B, T, size, size, T * size * size * 3, size * size * 3, 1, size * 3, 3)
        scale = 1.0 / 255.0
    else:
        x = video.cuda()
        layout = (0, B, T, size, size, 3 * T * size * size, size * size, T * size * size, size, 1)
        scale = 1.0
    y = torch.empty(B * T, Ho, Ho, 16, device="cuda")
    stats = torch.zeros(32, dtype=torch.float64, device="cuda")
    K.stem_conv_fwd(x, layout, conv.weight.detach().cuda(), y, stats, scale)
    _close(y, _cl(ref), rtol=2e-5)
    _close(stats[:16], ref.double().sum((0, 2, 3)), rtol=1e-5)
    _close(stats[16:], (ref.double() ** 2).sum((0, 2, 3)), rtol=1e-5)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    dw = torch.zeros(16, 3, 3, 3, device="cuda")
    K.stem_conv_wgrad(x, layout, _cl(dy).cuda(), dw, scale)
    _close(dw, conv.weight.grad)


@pytest.mark.parametrize("C,k,s,H", [(16, 3, 2, 22), (72, 3, 2, 11), (88, 3, 1, 6), (96, 5, 2, 6), (240, 5, 1, 3),
                                     (576, 5, 1, 2), (288, 5, 2, 3), (96, 5, 2, 22),
                                     # torchvision mobilenet_v2 (3x3 only) at 44 / 88 px lip frames
                                     (32, 3, 1, 22), (96, 3, 2, 22), (144, 3, 1, 11), (144, 3, 2, 11), (192, 3, 1, 6),
                                     (192, 3, 2, 6), (384, 3, 1, 3), (576, 3, 2, 3), (960, 3, 1, 2), (32, 3, 1, 44),
                                     (96, 3, 2, 44), (144, 3, 1, 22), (960, 3, 1, 3), (576, 3, 2, 6),
                                     # small-image kernels (dwconv_small.cuh): whole map per thread
                                     (240, 5, 1, 6), (120, 5, 1, 6), (288, 5, 2, 6), (576, 5, 1, 3), (72, 5, 2, 3), (40, 3, 1, 6)])
def test_dwconv(K, C, k, s, H):
    F = 6 if C != 120 else 37
    g = torch.Generator().manual_seed(C + k)
    conv = nn.Conv2d(C, C, k, s, k // 2, groups=C, bias=False)
    x = torch.randn(F, C, H, H, generator=g, requires_grad=True)
    ref = conv(x)
    Ho = ref.shape[2]
    xd = _cl(x.detach()).cuda()
    w = conv.weight.detach().cuda()
    y = torch.empty(F, Ho, Ho, C, device="cuda")
    stats = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    K.dwconv_fwd(xd, w, y, stats, F, H, H, C, k, s)
    _close(y, _cl(ref), rtol=2e-5)
    _close(stats[:C], ref.double().sum((0, 2, 3)), rtol=1e-5, atol=1e-4)
    _close(stats[C:], (ref.double() ** 2).sum((0, 2, 3)), rtol=1e-5)
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    dyd = _cl(dy).cuda()
    dx = torch.empty(F, H, H, C, device="cuda")
    K.dwconv_dgrad(dyd, w, dx, F, H, H, C, k, s)
    _close(dx, _cl(x.grad), rtol=2e-5)
    dw = torch.zeros_like(w)
    K.dwconv_wgrad(dyd, xd, dw, F, H, H, C, k, s)
    _close(dw, conv.weight.grad, rtol=5e-5)


@pytest.mark.parametrize("C,act,res,training", [(16, 2, False, True), (72, 1, False, True), (24, 0, True, True),
                                                (576, 2, False, True), (40, 0, True, False), (88, 1, False, False)])
def test_bn_act(K, C, act, res, training):
    F, H = 5, 7
    g = torch.Generator().manual_seed(C)
    bn = nn.BatchNorm2d(C, eps=1e-3, momentum=0.01)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    bn.train(training)
    x = (torch.randn(F, C, H, H, generator=g) * 2 + 0.3).requires_grad_(True)
    r = torch.randn(F, C, H, H, generator=g, requires_grad=True) if res else None
    import copy
    bnd = copy.deepcopy(bn).cuda()
    z_ref = ACTS[act](bn(x))
    if res:
        z_ref = z_ref + r
    rows = F * H * H
    xd = _cl(x.detach()).cuda()
    x2 = xd.view(rows, C).double()
    stats = torch.stack([x2.sum(0), (x2 ** 2).sum(0)]).reshape(-1).contiguous()
    z = torch.empty(rows, C, device="cuda")
    K.bn_act_fwd(xd, stats if training else None, bnd, act, training, z, rows, C,
                 residual=_cl(r.detach()).cuda() if res else None)
    _close(z.view(F, H, H, C), _cl(z_ref), rtol=2e-5)
    _close(bnd.running_mean, bn.running_mean, rtol=1e-5)
    _close(bnd.running_var, bn.running_var, rtol=1e-5)
    assert int(bnd.num_batches_tracked) == int(bn.num_batches_tracked)
    dz = torch.randn(F, C, H, H, generator=g)
    z_ref.backward(dz)
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    dx = torch.empty(rows, C, device="cuda")
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    K.bn_act_bwd(xd, stats if training else None, bnd, act, training, _cl(dz).cuda(), sums, dx, dgamma, dbeta, rows, C)
    _close(dx.view(F, H, H, C), _cl(x.grad), rtol=1e-4)
    _close(dgamma, bn.weight.grad, rtol=1e-4)
    _close(dbeta, bn.bias.grad, rtol=1e-4)


def test_frame_ops_and_small_kernels(K):
    g = torch.Generator().manual_seed(1)
    F, HW, C = 7, 9, 72
    a, gt = torch.randn(F, HW, C, generator=g), torch.randn(F, HW, C, generator=g)
    s, dp = torch.rand(F, C, generator=g), torch.randn(F, C, generator=g)
    p = torch.empty(F, C, device="cuda")
    K.frame_reduce(a.cuda(), None, p, F, HW, C, 0)
    _close(p, a.mean(1))
    K.frame_reduce(a.cuda(), gt.cuda(), p, F, HW, C, 1)
    _close(p, (a * gt).sum(1))
    out = torch.empty(F, HW, C, device="cuda")
    K.frame_scale(a.cuda(), s.cuda(), None, out, F, HW, C)
    _close(out, a * s[:, None, :])
    K.frame_scale(a.cuda(), s.cuda(), dp.cuda(), out, F, HW, C)
    _close(out, a * s[:, None, :] + dp[:, None, :] / HW)
    K.frame_scale(None, None, dp.cuda(), out, F, HW, C)
    _close(out, (dp[:, None, :] / HW).expand(F, HW, C))
    y = torch.randn(1000, generator=g)
    for act in (1, 3):
        yo = ACTS[act](y)
        dy = torch.randn(1000, generator=g)
        d = dy.clone().cuda()
        K.act_bwd(d, yo.cuda(), 1000, act)
        yy = y.clone().requires_grad_(True)
        ACTS[act](yy).backward(dy)
        _close(d, yy.grad)
    dY = torch.randn(777, 300, generator=g)
    db = torch.ones(300, device="cuda")
    K.colsum(dY.cuda(), 300, 777, 300, db)
    _close(db, 1 + dY.sum(0))
    src = torch.randn(10, 50, generator=g).cuda()
    dst = torch.zeros(10, 30, device="cuda")
    K.copy2d(dst[:, 5:], 30, src[:, 20:], 50, 10, 20)
    assert torch.equal(dst[:, 5:25], src[:, 20:40]) and dst[:, :5].abs().sum() == 0 and dst[:, 25:].abs().sum() == 0


@pytest.mark.parametrize("M,N,K_,act,strided", [(32, 256, 384, 1, False), (32, 40, 256, 0, False), (32, 512, 576, 0, True),
                                                (5, 13, 36, 3, False), (1, 8, 1600, 2, False), (17, 100, 132, 0, True)])
def test_small_row_linear_forward(K, M, N, K_, act, strided):
    """lr_linear_small_fwd against nn.Linear (+ activation) on <= 32 rows, contiguous and strided (x[:, -1] of a
    [B, T, I] sequence) inputs; bit-reproducible."""
    g = torch.Generator().manual_seed(M * 31 + N)
    T = 3 if strided else 1
    xs = torch.randn(M, T, K_, generator=g)
    w, b = torch.randn(N, K_, generator=g) / K_ ** 0.5, torch.randn(N, generator=g)
    ref = ACTS[act](xs[:, -1] @ w.t() + b)
    xd = xs.cuda()
    y = torch.full((M, N + 3), 7.0, device="cuda")
    xv = xd[:, -1]
    K.linear_small_fwd(xv, T * K_, w.cuda(), b.cuda(), y, N + 3, M, N, K_, act)
    _close(y[:, :N], ref, rtol=2e-5)
    assert (y[:, N:] == 7.0).all()
    y2 = torch.empty_like(y)
    K.linear_small_fwd(xv, T * K_, w.cuda(), b.cuda(), y2, N + 3, M, N, K_, act)
    assert torch.equal(y[:, :N], y2[:, :N])


@pytest.mark.parametrize("M,N,K_,res", [(32, 256, 384, False), (32, 40, 256, False), (32, 512, 576, True), (7, 12, 64, True),
                                        (1, 1088, 32, False)])
def test_small_row_linear_input_gradient(K, M, N, K_, res):
    g = torch.Generator().manual_seed(M + N + K_)
    dy, w, r = torch.randn(M, N, generator=g), torch.randn(N, K_, generator=g), torch.randn(M, K_, generator=g)
    ref = dy @ w + (r if res else 0)
    dx = r.clone().cuda() if res else torch.empty(M, K_, device="cuda")
    K.linear_small_dgrad(dy.cuda(), N, w.cuda(), dx, K_, M, N, K_, r=(dx if res else None), ldr=(K_ if res else 0))   # r aliases dx
    _close(dx, ref, rtol=2e-5)


@pytest.mark.parametrize("F,C,Cs", [(928, 576, 144), (928, 16, 8), (61, 240, 64), (7, 96, 24), (100, 288, 72), (9, 120, 32)])
def test_se_gate_single_launch_kernels(K, F, C, Cs):
    """lr_se_fc_fwd / lr_se_fc_bwd against torchvision's SqueezeExcitation arithmetic (fc1 -> ReLU -> fc2 -> Hardsigmoid
    on the pooled vectors) and torch autograd; the forward is bit-reproducible."""
    g = torch.Generator().manual_seed(F + C)
    p = torch.randn(F, C, generator=g)
    w1 = torch.randn(Cs, C, generator=g) / C ** 0.5
    b1 = torch.randn(Cs, generator=g) * 0.1
    w2 = torch.randn(C, Cs, generator=g) * (3.0 / Cs ** 0.5)          # pre-activations spread over the hard-sigmoid's kinks
    b2 = torch.randn(C, generator=g) * 0.5
    ds = torch.randn(F, C, generator=g)
    pr, w1r, b1r, w2r, b2r = (t.double().requires_grad_(True) for t in (p, w1, b1, w2, b2))
    h1_ref = Fn.relu(pr @ w1r.t() + b1r)
    s_ref = Fn.hardsigmoid(h1_ref @ w2r.t() + b2r)
    s_ref.backward(ds.double())
    h1, s = torch.empty(F, Cs, device="cuda"), torch.empty(F, C, device="cuda")
    args = [t.cuda() for t in (p, w1, b1, w2, b2)]
    K.se_fc_fwd(*args, h1, s, F, C, Cs)
    _close(h1, h1_ref, rtol=1e-5)
    _close(s, s_ref, rtol=1e-5)
    h1b, sb = torch.empty_like(h1), torch.empty_like(s)
    K.se_fc_fwd(*args, h1b, sb, F, C, Cs)
    assert torch.equal(h1, h1b) and torch.equal(s, sb)
    # backward on the reference's own h1 / s so that no kink is crossed by forward round-off
    h1d, sd = h1_ref.detach().float().cuda(), s_ref.detach().float().cuda()
    dsd = ds.clone().cuda()
    dz1, dp = torch.empty(F, Cs, device="cuda"), torch.empty(F, C, device="cuda")
    K.se_fc_bwd(dsd, sd, h1d, args[1], args[3], dz1, dp, F, C, Cs)
    inside = ((s_ref > 0) & (s_ref < 1)).detach()
    _close(dsd, ds.double() * inside / 6.0, rtol=1e-6)
    _close(dp, pr.grad, rtol=2e-5)
    # the weight gradients the caller forms from dz2 / dz1
    _close(dsd.t() @ h1d, w2r.grad, rtol=1e-4)
    _close(dz1.t() @ args[0], w1r.grad, rtol=1e-4)
    _close(dz1.sum(0), b1r.grad, rtol=1e-4)


@pytest.mark.parametrize("H,I,B,T", [(128, 576, 6, 29), (32, 20, 5, 4), (256, 64, 33, 7), (128, 48, 32, 5), (512, 32, 3, 4)])
def test_lstm_direction_kernels(K, H, I, B, T):
    """Both directions of a bidirectional nn.LSTM layer, full sequence with external gradients at every t."""
    torch.manual_seed(H)
    lstm = nn.LSTM(I, H, 1, batch_first=True, bidirectional=True)
    x = torch.randn(B, T, I, requires_grad=True)
    out, _ = lstm(x)
    dout = torch.randn_like(out)
    out.backward(dout)
    xd = x.detach().cuda()
    for d, sfx in enumerate(("", "_reverse")):
        wih, whh = getattr(lstm, "weight_ih_l0" + sfx).detach().cuda(), getattr(lstm, "weight_hh_l0" + sfx).detach().cuda()
        bih, bhh = getattr(lstm, "bias_ih_l0" + sfx).detach().cuda(), getattr(lstm, "bias_hh_l0" + sfx).detach().cuda()
        xproj = torch.empty(B * T, 4 * H, device="cuda")
        K.linear_fwd(xd.view(B * T, I), wih, xproj, bias=bih)
        o = torch.zeros(B, T, 2 * H, device="cuda")
        gates, cst, hprev = torch.empty(B, T, 4 * H, device="cuda"), torch.empty(B, T, H, device="cuda"), torch.empty(B, T, H, device="cuda")
        K.lstm_fwd(xproj, 4 * H, bhh, whh, o[:, :, d * H:], 2 * H, gates, cst, hprev, B, T, H, T, d)
        _close(o[:, :, d * H:(d + 1) * H], out[:, :, d * H:(d + 1) * H], rtol=2e-5)
        dg = torch.zeros(B, T, 4 * H, device="cuda")
        dod = dout.contiguous().cuda()          # nn.LSTM(batch_first) hands back a transposed view
        K.lstm_bwd(dod[:, :, d * H:], 2 * H, -1, gates, cst, whh, dg, B, T, H, T, d)
        dwih, dwhh = torch.zeros_like(wih), torch.zeros_like(whh)
        K.linear_wgrad(dg.view(B * T, 4 * H), xd.view(B * T, I), dwih)
        K.linear_wgrad(dg.view(B * T, 4 * H), hprev.view(B * T, H), dwhh)
        db = torch.zeros(4 * H, device="cuda")
        K.colsum(dg, 4 * H, B * T, 4 * H, db)
        _close(dwih, getattr(lstm, "weight_ih_l0" + sfx).grad, rtol=1e-4)
        _close(dwhh, getattr(lstm, "weight_hh_l0" + sfx).grad, rtol=1e-4)
        _close(db, getattr(lstm, "bias_ih_l0" + sfx).grad, rtol=1e-4)
        dx = torch.empty(B * T, I, device="cuda")
        K.linear_dgrad(dg.view(B * T, 4 * H), wih, dx)
        if d == 0:
            dx_total = dx.clone()
        else:
            dx_total += dx
    _close(dx_total.view(B, T, I), x.grad, rtol=1e-4)


@pytest.mark.parametrize("H,I,B,T", [(128, 64, 32, 29), (128, 48, 6, 5), (256, 64, 33, 7), (512, 32, 32, 29), (512, 32, 3, 4)])
def test_lstm_tensor_core_kernels(K, H, I, B, T):
    """csrc/lstm_tc.cu (precision "bf16"): W_hh / h and W_hh^T / dgates rounded to bf16, tcgen05 products, fp32 state.
    Against torch's fp32 nn.LSTM: outputs within 2e-2 of max|out| (bf16 operand rounding through T recurrent steps; the
    fp32 kernels sit at 2e-5), dgates-derived weight / input gradients within 5e-2 norm-wise; and against the fp32
    kernel driven with the SAME bf16-rounded W_hh, where only the rounding of h / dgates remains."""
    torch.manual_seed(H + B)
    lstm = nn.LSTM(I, H, 1, batch_first=True, bidirectional=True)
    x = torch.randn(B, T, I, requires_grad=True)
    out, _ = lstm(x)
    dout = torch.randn_like(out)
    out.backward(dout)
    xd = x.detach().cuda()
    dod = dout.contiguous().cuda()
    for d, sfx in enumerate(("", "_reverse")):
        wih, whh = getattr(lstm, "weight_ih_l0" + sfx).detach().cuda(), getattr(lstm, "weight_hh_l0" + sfx).detach().cuda()
        bih, bhh = getattr(lstm, "bias_ih_l0" + sfx).detach().cuda(), getattr(lstm, "bias_hh_l0" + sfx).detach().cuda()
        xproj = torch.empty(B * T, 4 * H, device="cuda")
        K.linear_fwd(xd.view(B * T, I), wih, xproj, bias=bih)
        res = {}
        for kind in ("tc", "fp32_rounded_w"):
            w_used = whh if kind == "tc" else whh.to(torch.bfloat16).float()
            o = torch.zeros(B, T, 2 * H, device="cuda")
            gates, cst, hprev = (torch.full((B, T, 4 * H), float("nan"), device="cuda"), torch.full((B, T, H), float("nan"), device="cuda"),
                                 torch.full((B, T, H), float("nan"), device="cuda"))
            dg = torch.zeros(B, T, 4 * H, device="cuda")
            fwd, bwd = (K.lstm_fwd_tc, K.lstm_bwd_tc) if kind == "tc" else (K.lstm_fwd, K.lstm_bwd)
            fwd(xproj, 4 * H, bhh, w_used, o[:, :, d * H:], 2 * H, gates, cst, hprev, B, T, H, T, d)
            bwd(dod[:, :, d * H:], 2 * H, -1, gates, cst, w_used, dg, B, T, H, T, d)
            torch.cuda.synchronize()
            assert torch.isfinite(gates).all() and torch.isfinite(cst).all() and torch.isfinite(hprev).all()
            res[kind] = (o[:, :, d * H:(d + 1) * H].clone(), dg, hprev)
        o_tc, dg_tc, hp_tc = res["tc"]
        o_32, dg_32, _ = res["fp32_rounded_w"]
        ref = out[:, :, d * H:(d + 1) * H].detach().cuda()
        scale = ref.abs().max().item()
        assert (o_tc - ref).abs().max().item() <= 2e-2 * scale
        assert (o_tc - o_32).abs().max().item() <= 1e-2 * scale          # same rounded weights: only h's rounding differs
        assert (dg_tc - dg_32).abs().max().item() <= 3e-2 * dg_32.abs().max().item()
        dwih, dwhh = torch.zeros_like(wih), torch.zeros_like(whh)
        K.linear_wgrad(dg_tc.view(B * T, 4 * H), xd.view(B * T, I), dwih)
        K.linear_wgrad(dg_tc.view(B * T, 4 * H), hp_tc.view(B * T, H), dwhh)
        for got, want in ((dwih, getattr(lstm, "weight_ih_l0" + sfx).grad), (dwhh, getattr(lstm, "weight_hh_l0" + sfx).grad)):
            want = want.cuda()
            assert (got - want).abs().max().item() <= 5e-2 * want.abs().max().item()


def test_lstm_last_step_only(K):
    """out[:, -1] head (middle_fusion_fast.py:36): forward direction with a gradient at t = T-1 only, and the
    reverse direction reduced to its first step."""
    torch.manual_seed(5)
    H, I, B, T = 128, 576, 5, 29
    lstm = nn.LSTM(I, H, 1, batch_first=True, bidirectional=True)
    x = torch.randn(B, T, I, requires_grad=True)
    out, _ = lstm(x)
    last = out[:, -1]
    dlast = torch.randn(last.shape)
    last.backward(dlast)
    xd, dl = x.detach().cuda(), dlast.cuda()
    P = {n: p.detach().cuda() for n, p in lstm.named_parameters()}
    fused = torch.zeros(B, 2 * H, device="cuda")
    # reverse direction: one step on x[:, T-1]
    xr = torch.empty(B, 4 * H, device="cuda")
    K.linear_fwd(xd[:, T - 1], P["weight_ih_l0_reverse"], xr, bias=P["bias_ih_l0_reverse"], M=B, lda=T * I)
    gr, cr = torch.empty(B, 1, 4 * H, device="cuda"), torch.empty(B, 1, H, device="cuda")
    K.lstm_fwd(xr, 4 * H, P["bias_hh_l0_reverse"], P["weight_hh_l0_reverse"], fused[:, H:], 2 * H, gr, cr, None, B, 1, H, 1, 1)
    _close(fused[:, H:], last[:, H:], rtol=2e-5)
    dgr = torch.empty(B, 1, 4 * H, device="cuda")
    K.lstm_bwd(dl[:, H:], 2 * H, 0, gr, cr, P["weight_hh_l0_reverse"], dgr, B, 1, H, 1, 1)
    dwih_r = torch.zeros_like(P["weight_ih_l0_reverse"])
    K.linear_wgrad(dgr.view(B, 4 * H), xd[:, T - 1], dwih_r, M=B, ldx=T * I)
    _close(dwih_r, lstm.weight_ih_l0_reverse.grad, rtol=1e-4)
    assert lstm.weight_hh_l0_reverse.grad.abs().max() == 0
    # forward direction: all steps, gradient only at the last
    xp = torch.empty(B * T, 4 * H, device="cuda")
    K.linear_fwd(xd.view(B * T, I), P["weight_ih_l0"], xp, bias=P["bias_ih_l0"])
    hs = torch.empty(B, T, H, device="cuda")
    gf, cf, hp = torch.empty(B, T, 4 * H, device="cuda"), torch.empty(B, T, H, device="cuda"), torch.empty(B, T, H, device="cuda")
    K.lstm_fwd(xp, 4 * H, P["bias_hh_l0"], P["weight_hh_l0"], hs, H, gf, cf, hp, B, T, H, T, 0)
    K.copy2d(fused, 2 * H, hs[:, T - 1], T * H, B, H)
    _close(fused[:, :H], last[:, :H], rtol=2e-5)
    dgf = torch.empty(B, T, 4 * H, device="cuda")
    K.lstm_bwd(dl, 2 * H, T - 1, gf, cf, P["weight_hh_l0"], dgf, B, T, H, T, 0)
    dx = torch.empty(B * T, I, device="cuda")
    K.linear_dgrad(dgf.view(B * T, 4 * H), P["weight_ih_l0"], dx)
    K.linear_dgrad(dgr.view(B, 4 * H), P["weight_ih_l0_reverse"], dx.view(B, T, I)[:, T - 1], M=B, ldx=T * I, accumulate=True)
    _close(dx.view(B, T, I), x.grad, rtol=1e-4)
    dwhh = torch.zeros_like(P["weight_hh_l0"])
    K.linear_wgrad(dgf.view(B * T, 4 * H), hp.view(B * T, H), dwhh)
    _close(dwhh, lstm.weight_hh_l0.grad, rtol=1e-4)


def test_audio_conv_ce_adam(K):
    torch.manual_seed(2)
    B = 3
    conv = nn.Conv2d(1, 16, 3, padding=1)
    mel = torch.randn(B, 80, 117)
    a_ref = Fn.max_pool2d(Fn.relu(conv(mel.unsqueeze(1))), 2).flatten(1)
    out = torch.zeros(B, 37120 + 8, device="cuda")
    arg = torch.empty(B * 37120, dtype=torch.uint8, device="cuda")
    K.audio_conv_fwd(mel.cuda(), conv.weight.detach().cuda(), conv.bias.detach().cuda(), out, 37128, arg, B, 80, 117)
    _close(out[:, :37120], a_ref, rtol=2e-5)
    dA = torch.randn_like(a_ref)
    a_ref.backward(dA)
    dw, db = torch.zeros(16, 1, 3, 3, device="cuda"), torch.zeros(16, device="cuda")
    K.audio_conv_bwd(mel.cuda(), dA.cuda(), 37120, arg, dw, db, B, 80, 117)
    _close(dw, conv.weight.grad, rtol=1e-4)
    _close(db, conv.bias.grad, rtol=1e-4)
    # cross entropy
    for C in (4, 40, 500):
        logits = torch.randn(9, C, requires_grad=True)
        labels = torch.randint(0, C, (9,))
        loss_ref = Fn.cross_entropy(logits, labels)
        loss_ref.backward()
        loss = torch.zeros(1, device="cuda")
        dl = torch.empty(9, C, device="cuda")
        correct = torch.zeros(1, dtype=torch.int32, device="cuda")
        K.ce_loss(logits.detach().cuda(), labels.cuda(), loss, dl, correct, 9, C, 1.0 / 9)
        _close(loss, loss_ref.reshape(1), rtol=1e-5)
        _close(dl, logits.grad, rtol=1e-4)
        assert int(correct) == int((logits.argmax(1) == labels).sum())
    # Adam, three steps, with and without coupled weight decay
    from multimodal_lipread_b200 import _lib
    for wd in (0.0, 1e-4):
        p = torch.randn(1000, requires_grad=True)
        opt = torch.optim.Adam([p], lr=3e-4, weight_decay=wd)
        pd = p.detach().clone().cuda()
        m, v = torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
        state = torch.tensor([0.0, 0.0, 0.0, 3e-4], device="cuda")
        assert _lib.lib.lr_adam_state_bytes() == 16
        for i in range(3):
            gr = torch.randn(1000)
            p.grad = gr.clone()
            opt.step()
            K.adam_step(pd, (gr * 4).cuda(), m, v, state, 1000, 0.9, 0.999, 1e-8, wd, 0.25)
        _close(pd, p, rtol=1e-6)
        assert float(state[0]) == 3.0


@pytest.mark.parametrize("B,T,E,heads", [(3, 29, 512, 4), (2, 6, 256, 4), (1, 1, 64, 8), (2, 64, 128, 1), (2, 10, 256, 2)])
def test_multihead_attention_kernels(K, B, T, E, heads):
    """lr_mha_* against nn.MultiheadAttention's own forward / autograd (video/models/resnet_attn.py:23-35): attention
    weights, concatenated heads and the gradient of the packed in-projection."""
    from multimodal_lipread_b200._lib import lib, check
    torch.manual_seed(T * 7 + E)
    mha = nn.MultiheadAttention(E, heads, dropout=0.0, batch_first=True)
    x = torch.randn(B, T, E)
    qkv_ref = (x @ mha.in_proj_weight.T + mha.in_proj_bias).detach().requires_grad_(True)         # [B, T, 3E]
    q, k, v = qkv_ref.chunk(3, dim=-1)
    d = E // heads
    split = lambda t: t.view(B, T, heads, d).transpose(1, 2)                                      # [B, h, T, d]
    P_ref = torch.softmax((split(q) * (1.0 / d) ** 0.5) @ split(k).transpose(-1, -2), dim=-1)
    O_ref = (P_ref @ split(v)).transpose(1, 2).reshape(B, T, E)
    out_ref, w_ref = mha(x, x, x)                                                                 # the module's own answer
    assert torch.allclose(out_ref, O_ref @ mha.out_proj.weight.T + mha.out_proj.bias, atol=1e-5)
    assert torch.allclose(w_ref, P_ref.mean(dim=1), atol=1e-6)
    dO_ref = torch.randn(B, T, E)
    O_ref.backward(dO_ref)
    dev = "cuda"
    qkv = qkv_ref.detach().reshape(B * T, 3 * E).to(dev).contiguous()
    P = torch.empty(B, heads, T, T, device=dev)
    O = torch.empty(B * T, E, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    check(lib.lr_mha_scores_fwd(qkv.data_ptr(), 3 * E, P.data_ptr(), B, T, E, heads, s))
    check(lib.lr_mha_apply_fwd(P.data_ptr(), qkv.data_ptr(), 3 * E, O.data_ptr(), B, T, E, heads, s))
    _close(P.cpu(), P_ref.detach(), rtol=2e-5)
    _close(O.cpu().view(B, T, E), O_ref.detach(), rtol=2e-5)
    dO = dO_ref.reshape(B * T, E).to(dev).contiguous()
    dP = torch.empty_like(P)
    dqkv = torch.full_like(qkv, float("nan"))                       # every element must be written
    check(lib.lr_mha_apply_bwd(dO.data_ptr(), P.data_ptr(), qkv.data_ptr(), 3 * E, dP.data_ptr(), dqkv.data_ptr(), B, T, E, heads, s))
    check(lib.lr_mha_scores_bwd(P.data_ptr(), dP.data_ptr(), qkv.data_ptr(), 3 * E, dqkv.data_ptr(), B, T, E, heads, s))
    _close(dqkv.cpu().view(B, T, 3 * E), qkv_ref.grad, rtol=5e-5)
    # argument errors do not launch
    assert lib.lr_mha_scores_fwd(qkv.data_ptr(), 3 * E, P.data_ptr(), B, 65, E, heads, s) == -1
    assert lib.lr_mha_scores_fwd(qkv.data_ptr(), 3 * E, P.data_ptr(), B, T, E, 3 if E % 3 else 7, s) == -1


@pytest.mark.parametrize("rows,D,residual", [(290, 256, True), (7, 1024, False), (33, 100, True), (1, 32, True)])
def test_layernorm_residual_and_broadcast_add(K, rows, D, residual):
    """lr_layernorm_fwd / _bwd against nn.LayerNorm(a + b) and its autograd (norm1 / norm2 of
    nn.TransformerEncoderLayer, video/models/resnet_trans.py:96-103); lr_add_bcast against repeat + positional add."""
    from multimodal_lipread_b200._lib import lib, check
    torch.manual_seed(rows + D)
    ln = nn.LayerNorm(D)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.5, 0.5)
    a = torch.randn(rows, D, requires_grad=True)
    b = torch.randn(rows, D, requires_grad=True) if residual else None
    y_ref = ln(a + b if residual else a)
    dy_ref = torch.randn(rows, D)
    y_ref.backward(dy_ref)
    dev, s = "cuda", torch.cuda.current_stream().cuda_stream
    ad, bd = a.detach().to(dev), (b.detach().to(dev) if residual else None)
    g, be = ln.weight.detach().to(dev), ln.bias.detach().to(dev)
    y, sm, st = torch.empty(rows, D, device=dev), torch.empty(rows, D, device=dev), torch.empty(rows, 2, device=dev)
    check(lib.lr_layernorm_fwd(ad.data_ptr(), bd.data_ptr() if residual else None, g.data_ptr(), be.data_ptr(), ln.eps,
                               y.data_ptr(), sm.data_ptr(), st.data_ptr(), rows, D, s))
    _close(y.cpu(), y_ref, rtol=2e-5)
    ds = torch.full((rows, D), float("nan"), device=dev)
    dg, db = torch.ones(D, device=dev), torch.ones(D, device=dev)          # accumulated onto
    dyd = dy_ref.to(dev)
    check(lib.lr_layernorm_bwd(dyd.data_ptr(), sm.data_ptr(), st.data_ptr(), g.data_ptr(), ds.data_ptr(),
                               dg.data_ptr(), db.data_ptr(), rows, D, s))
    _close(ds.cpu(), a.grad, rtol=5e-5)
    if residual:
        assert torch.equal(a.grad, b.grad)
    _close(dg.cpu() - 1.0, ln.weight.grad, rtol=5e-5)
    _close(db.cpu() - 1.0, ln.bias.grad, rtol=5e-5)
    assert lib.lr_layernorm_fwd(ad.data_ptr(), None, g.data_ptr(), be.data_ptr(), ln.eps, y.data_ptr(), sm.data_ptr(),
                                st.data_ptr(), rows, 1025, s) == -1
    # broadcast add
    F, T = 5, 10
    x, r = torch.randn(F, D), torch.randn(T, D)
    xd, rd = x.to(dev), r.to(dev)
    out = torch.empty(F, T, D, device=dev)
    check(lib.lr_add_bcast(xd.data_ptr(), rd.data_ptr(), out.data_ptr(), F, T, D, s))
    assert torch.equal(out.cpu(), x.unsqueeze(1).repeat(1, T, 1) + r)
    check(lib.lr_add_bcast(xd.data_ptr(), None, out.data_ptr(), F, T, D, s))
    assert torch.equal(out.cpu(), x.unsqueeze(1).repeat(1, T, 1))


def test_channel_shuffle_interleave(K):
    """lr_shuffle2_fwd / _bwd against torchvision's channel_shuffle(cat(x1, branch), 2) on an NCHW tensor, with x1 a
    column slice left in place (x.chunk(2, dim=1))."""
    from torchvision.models.shufflenetv2 import channel_shuffle
    from multimodal_lipread_b200._lib import lib, check
    torch.manual_seed(3)
    F, H, W, C = 3, 5, 4, 48
    bf = C // 2
    x = torch.randn(F, C, H, W)
    right = torch.randn(F, bf, H, W)
    ref = channel_shuffle(torch.cat((x[:, :bf], right), dim=1), 2)
    dev, s = "cuda", torch.cuda.current_stream().cuda_stream
    xr, rr = _cl(x).reshape(-1, C).to(dev), _cl(right).reshape(-1, bf).to(dev)
    out = torch.empty(F * H * W, C, device=dev)
    check(lib.lr_shuffle2_fwd(xr.data_ptr(), C, rr.data_ptr(), bf, out.data_ptr(), F * H * W, bf, s))
    assert torch.equal(out.cpu(), _cl(ref).reshape(-1, C))
    dout = torch.randn(F * H * W, C, device=dev)
    dx = torch.zeros(F * H * W, C, device=dev)
    dr = torch.empty(F * H * W, bf, device=dev)
    check(lib.lr_shuffle2_bwd(dout.data_ptr(), dx.data_ptr(), C, dr.data_ptr(), bf, F * H * W, bf, s))
    assert torch.equal(dx[:, :bf].cpu(), dout[:, 0::2].cpu()) and torch.equal(dr.cpu(), dout[:, 1::2].cpu())
    assert (dx[:, bf:] == 0).all()                                     # the other half belongs to branch2's 1x1 conv


@pytest.mark.parametrize("M,K_,N,act,bias", [(37, 384, 256, 1, True), (5, 128, 40, 0, True), (64, 100, 24, 1, False)])
def test_custom_op_linear_with_autograd(K, M, K_, N, act, bias):
    """torch.ops.lipread.linear: forward and register_autograd backward against nn.functional.linear (+ ReLU), and the
    fake (meta) kernel the dispatcher uses for shape inference."""
    from multimodal_lipread_b200 import ops
    torch.manual_seed(M)
    x = torch.randn(M, K_, device="cuda", requires_grad=True)
    w = torch.randn(N, K_, device="cuda", requires_grad=True)
    b = torch.randn(N, device="cuda", requires_grad=True) if bias else torch.empty(0, device="cuda")
    y = torch.ops.lipread.linear(x, w, b, act)
    ref = Fn.linear(x.detach().cpu().double(), w.detach().cpu().double(), b.detach().cpu().double() if bias else None)
    ref_y = torch.relu(ref) if act else ref
    _close(y, ref_y, rtol=1e-5)
    dy = torch.randn(M, N, device="cuda")
    y.backward(dy)
    xr, wr = x.detach().cpu().double().requires_grad_(True), w.detach().cpu().double().requires_grad_(True)
    br = b.detach().cpu().double().requires_grad_(True) if bias else None
    yr = Fn.linear(xr, wr, br)
    (torch.relu(yr) if act else yr).backward(dy.cpu().double())
    _close(x.grad, xr.grad, rtol=2e-5)
    _close(w.grad, wr.grad, rtol=2e-5)
    if bias:
        _close(b.grad, br.grad, rtol=2e-5)
    with torch._subclasses.fake_tensor.FakeTensorMode():
        fy = torch.ops.lipread.linear(torch.empty(M, K_, device="cuda"), torch.empty(N, K_, device="cuda"),
                                      torch.empty(N, device="cuda"), act)
        assert tuple(fy.shape) == (M, N)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.lipread.linear(x.detach().cpu(), w.detach().cpu(), b.detach().cpu(), act)       # no CPU kernel
