"""bf16-storage kernels (precision "bf16": the `_h` entry points of include/lipread_b200.h and lr_gemm_bf16) through the
C ABI.  Two kinds of reference:
  * the tensor-core GEMM against a float64 matmul of the SAME bf16-rounded operands (fp32 accumulation: 1e-5 of
    max|ref| for an fp32 result; one bf16 rounding, 2^-8 relative, for a bf16 result);
  * every streaming kernel against its own fp32 namesake run on the bf16-rounded inputs: identical arithmetic, so the
    results agree to the final bf16 rounding of the output (and statistics / reductions, which stay fp32 / double, to
    1e-5)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from multimodal_lipread_b200 import _lib as L
    return L


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _bf(x):
    return x.to(torch.bfloat16)


def _gemm_bf16(A, lda, at, B, ldb, bt, C, ldc, M, N, K, bias=None, act=0, R=None, ldr=0, stats=None, ksplit=1):
    L = _lib()
    L.check(L.lib.lr_gemm_bf16(_p(A), lda, at, _p(B), ldb, bt, _p(C), ldc, int(C.dtype == torch.bfloat16), M, N, K,
                               _p(bias), act, _p(R), ldr, _p(stats), ksplit, _s()))


SHAPES = [(4000, 72, 16), (1000, 24, 72), (129, 16, 16), (5000, 96, 24), (777, 40, 96), (640, 240, 40),
          (300, 576, 96), (333, 96, 576), (2000, 288, 48), (128, 16, 8), (64, 144, 40), (1, 24, 88), (3000, 16, 32),
          (500, 512, 1152), (260, 64, 200)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("out", ["bf16", "f32"])
def test_gemm_bf16_forward_layout(cuda_device, M, N, K, out):
    g = torch.Generator().manual_seed(M + N + K)
    A, B = _bf(torch.randn(M, K, generator=g)).cuda(), _bf(torch.randn(N, K, generator=g)).cuda()
    ref = A.double().cpu() @ B.double().cpu().t()
    C = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16 if out == "bf16" else torch.float32)
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    _gemm_bf16(A, K, 0, B, K, 0, C, N, M, N, K, stats=stats)
    torch.cuda.synchronize()
    assert torch.isfinite(C.float()).all()
    err = (C.double().cpu() - ref).abs().max().item()
    tol = (2.0 ** -8 if out == "bf16" else 1e-5) * ref.abs().max().item()
    assert err <= tol, (err, tol)
    Cd = C.double()
    assert torch.allclose(stats[:N], Cd.sum(0), rtol=1e-5, atol=1e-3)          # statistics of the STORED values
    assert torch.allclose(stats[N:], (Cd * Cd).sum(0), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("M,N,K,out", [(460000, 16, 32, "bf16"), (460000, 16, 32, "f32"), (230000, 72, 16, "bf16"),
                                       (115100, 24, 24, "bf16"), (230100, 32, 64, "bf16"), (120000, 96, 24, "f32")])
def test_gemm_bf16_tile_batching(cuda_device, M, N, K, out):
    """Tall single-k-block problems run several 128-row tiles per CTA (gemm_tc.cu, P::mt): tails (a last CTA with fewer
    tiles, a last tile with fewer rows), bias + activation + residual, and the once-per-CTA statistics."""
    g = torch.Generator().manual_seed(M + N + K)
    A, B = _bf(torch.randn(M, K, generator=g)).cuda(), _bf(torch.randn(N, K, generator=g)).cuda()
    dt = torch.bfloat16 if out == "bf16" else torch.float32
    ref = A.double() @ B.double().t()
    C = torch.full((M, N), float("nan"), device="cuda", dtype=dt)
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    _gemm_bf16(A, K, 0, B, K, 0, C, N, M, N, K, stats=stats)
    torch.cuda.synchronize()
    tol = (2.0 ** -8 if out == "bf16" else 1e-5) * ref.abs().max().item()
    assert torch.isfinite(C.float()).all() and (C.double() - ref).abs().max().item() <= tol
    Cd = C.double()
    assert torch.allclose(stats[:N], Cd.sum(0), rtol=1e-5, atol=1e-2)
    assert torch.allclose(stats[N:], (Cd * Cd).sum(0), rtol=1e-5, atol=1e-2)
    if out == "bf16":
        bias, R = torch.randn(N, generator=g).cuda(), _bf(torch.randn(M, N, generator=g)).cuda()
        C2 = torch.full((M, N), float("nan"), device="cuda", dtype=dt)
        _gemm_bf16(A, K, 0, B, K, 0, C2, N, M, N, K, bias=bias, act=1, R=R, ldr=N)
        ref2 = torch.relu(ref + bias.double()) + R.double()
        assert (C2.double() - ref2).abs().max().item() <= 2.0 ** -8 * ref2.abs().max().item()


def test_gemm_bf16_epilogue_residual_and_strides(cuda_device):
    g = torch.Generator().manual_seed(0)
    M, N, K = 1500, 88, 24
    A, B = _bf(torch.randn(M, K + 8, generator=g)).cuda(), _bf(torch.randn(N, K, generator=g)).cuda()
    bias, R = torch.randn(N, generator=g).cuda(), _bf(torch.randn(M, N + 8, generator=g)).cuda()
    Cbig = torch.zeros(M, N + 16, device="cuda", dtype=torch.bfloat16)
    fns = {0: lambda u: u, 1: torch.relu, 2: torch.nn.functional.hardswish}
    for act, fn in fns.items():
        _gemm_bf16(A, K + 8, 0, B, K, 0, Cbig[:, 8:], N + 16, M, N, K, bias=bias, act=act, R=R, ldr=N + 8)
        ref = fn(A[:, :K].double() @ B.double().t() + bias.double()) + R[:, :N].double()
        err = (Cbig[:, 8:8 + N].double() - ref).abs().max().item()
        assert err <= 2.0 ** -8 * ref.abs().max().item(), (act, err)
        assert Cbig[:, :8].float().abs().sum().item() == 0 and Cbig[:, 8 + N:].float().abs().sum().item() == 0


# dgrad: dX[M, Cin] = dY[M, Cout] @ W[Cout, Cin] (+ residual gradient): B stored [K][N], MN-major bf16 operand
@pytest.mark.parametrize("M,Cout,Cin", [(4000, 72, 16), (1000, 24, 72), (300, 576, 96), (333, 96, 576), (2000, 288, 48),
                                        (129, 16, 16), (640, 40, 240), (500, 88, 24), (700, 512, 256)])
def test_gemm_bf16_dgrad_layout(cuda_device, M, Cout, Cin):
    g = torch.Generator().manual_seed(M + Cout)
    dY, W = _bf(torch.randn(M, Cout, generator=g)).cuda(), _bf(torch.randn(Cout, Cin, generator=g)).cuda()
    R = _bf(torch.randn(M, Cin, generator=g)).cuda()
    ref = dY.double() @ W.double() + R.double()
    dX = torch.full((M, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    _gemm_bf16(dY, Cout, 0, W, Cin, 1, dX, Cin, M, Cin, Cout, R=R, ldr=Cin)
    err = (dX.double() - ref).abs().max().item()
    assert err <= 2.0 ** -8 * ref.abs().max().item(), (err, ref.abs().max().item())


# wgrad: dW[Cout, Cin] (fp32) += dY[M, Cout]^T @ X[M, Cin]: both operands MN-major bf16, split over the M rows
@pytest.mark.parametrize("M,Cout,Cin,ks", [(40000, 72, 16, 37), (9000, 24, 72, 8), (3000, 576, 96, 5), (3333, 96, 576, 3),
                                           (2000, 288, 48, 1), (129, 16, 16, 1), (5000, 40, 240, 11), (70, 88, 24, 2),
                                           (9000, 128, 1152, 4)])
def test_gemm_bf16_wgrad_layout(cuda_device, M, Cout, Cin, ks):
    g = torch.Generator().manual_seed(M + Cin)
    dY, X = _bf(torch.randn(M, Cout, generator=g)).cuda(), _bf(torch.randn(M, Cin, generator=g)).cuda()
    W0 = torch.randn(Cout, Cin, generator=g)
    ref = W0.double() + (dY.double().t() @ X.double()).cpu()
    dW = W0.clone().cuda()
    if ks > 1:
        _gemm_bf16(dY, Cout, 1, X, Cin, 1, dW, Cin, Cout, Cin, M, ksplit=ks)
    else:
        _gemm_bf16(dY, Cout, 1, X, Cin, 1, dW, Cin, Cout, Cin, M, R=dW, ldr=Cin)
    err = (dW.cpu().double() - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item(), (err, ref.abs().max().item())


def test_cast_bf16(cuda_device):
    L = _lib()
    for n in (1, 3, 4, 1023, 100001):
        x = torch.randn(n, device="cuda")
        y = torch.zeros(n + 4, device="cuda", dtype=torch.bfloat16)
        L.check(L.lib.lr_cast_bf16(x.data_ptr(), y.data_ptr(), n, _s()))
        assert torch.equal(y[:n], x.to(torch.bfloat16)) and y[n:].float().abs().sum().item() == 0


def _agree(h, f, what):
    """bf16 result `h` against the fp32 kernel's result `f` on the same (bf16-rounded) inputs: one output rounding."""
    f = f.double()
    err = (h.double() - f).abs().max().item()
    assert err <= 2.0 ** -8 * f.abs().max().item() + 1e-30, (what, err, f.abs().max().item())


@pytest.mark.parametrize("rows,C,act", [(5000, 16, 2), (3001, 72, 1), (777, 576, 2), (64, 96, 0), (12000, 24, 4)])
def test_bn_act_forward_backward_h(cuda_device, rows, C, act):
    from multimodal_lipread_b200 import kernels as Kn
    L = _lib()
    g = torch.Generator().manual_seed(rows + C)
    bn = torch.nn.BatchNorm2d(C).cuda()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5); bn.bias.copy_(torch.randn(C, generator=g))
    xh = _bf(torch.randn(rows, C, generator=g) * 2 + 0.5).cuda()
    rh = _bf(torch.randn(rows, C, generator=g)).cuda()
    dzh = _bf(torch.randn(rows, C, generator=g)).cuda()
    xf, rf, dzf = xh.float(), rh.float(), dzh.float()
    stats = torch.cat([xf.double().sum(0), (xf.double() ** 2).sum(0)])
    bn2 = torch.nn.BatchNorm2d(C).cuda()
    bn2.load_state_dict(bn.state_dict())
    zf, zh = torch.empty_like(xf), torch.empty_like(xh)
    Kn.bn_act_fwd(xf, stats, bn, act, True, zf, rows, C, residual=rf)
    L.check(L.lib.lr_bn_act_fwd_h(_p(xh), _p(stats), _p(bn2.weight), _p(bn2.bias), _p(bn2.running_mean), _p(bn2.running_var),
                                  _p(bn2.num_batches_tracked), bn2.eps, bn2.momentum, act, 1, _p(rh), 0, _p(zh), rows, C, _s()))
    _agree(zh, zf, "bn fwd")
    assert torch.equal(bn.running_mean, bn2.running_mean) and torch.equal(bn.running_var, bn2.running_var)
    dxf, dxh = torch.empty_like(xf), torch.empty_like(xh)
    sums_f, sums_h = (torch.zeros(2 * C, dtype=torch.float64, device="cuda") for _ in range(2))
    dg_f, db_f, dg_h, db_h = (torch.zeros(C, device="cuda") for _ in range(4))
    Kn.bn_act_bwd(xf, stats, bn, act, True, dzf, sums_f, dxf, dg_f, db_f, rows, C)
    L.check(L.lib.lr_bn_act_bwd_h(_p(xh), _p(stats), _p(bn2.weight), _p(bn2.bias), _p(bn2.running_mean), _p(bn2.running_var),
                                  bn2.eps, act, 1, _p(dzh), 0, 0, _p(sums_h), _p(dxh), _p(dg_h), _p(db_h), rows, C, _s()))
    _agree(dxh, dxf, "bn bwd")
    assert torch.allclose(dg_h, dg_f, rtol=1e-4, atol=1e-3) and torch.allclose(db_h, db_f, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("F,H,C,k,s", [(58, 44, 16, 3, 2), (58, 22, 72, 3, 2), (40, 11, 88, 3, 1), (40, 11, 96, 5, 2),
                                       (64, 6, 240, 5, 1), (64, 6, 288, 5, 2), (64, 3, 576, 5, 1), (9, 44, 32, 3, 1)])
def test_depthwise_h(cuda_device, F, H, C, k, s):
    from multimodal_lipread_b200 import kernels as Kn
    L = _lib()
    g = torch.Generator().manual_seed(F + H + C)
    Ho = (H + 2 * (k // 2) - k) // s + 1
    xh = _bf(torch.randn(F, H, H, C, generator=g)).cuda()
    w = torch.randn(C, 1, k, k, generator=g).cuda()
    dyh = _bf(torch.randn(F, Ho, Ho, C, generator=g)).cuda()
    xf, dyf = xh.float(), dyh.float()
    yf, yh = torch.empty(F, Ho, Ho, C, device="cuda"), torch.empty(F, Ho, Ho, C, device="cuda", dtype=torch.bfloat16)
    st_f, st_h = (torch.zeros(2 * C, dtype=torch.float64, device="cuda") for _ in range(2))
    Kn.dwconv_fwd(xf, w, yf, st_f, F, H, H, C, k, s)
    L.check(L.lib.lr_dwconv_fwd_h(_p(xh), _p(w), _p(yh), _p(st_h), F, H, H, C, k, s, _s()))
    _agree(yh, yf, "dw fwd")
    yd = yh.double().reshape(-1, C)
    assert torch.allclose(st_h[:C], yd.sum(0), rtol=1e-5, atol=1e-3) and torch.allclose(st_h[C:], (yd * yd).sum(0), rtol=1e-5, atol=1e-3)
    dxf, dxh = torch.empty_like(xf), torch.empty_like(xh)
    Kn.dwconv_dgrad(dyf, w, dxf, F, H, H, C, k, s)
    L.check(L.lib.lr_dwconv_dgrad_h(_p(dyh), _p(w), _p(dxh), F, H, H, C, k, s, _s()))
    _agree(dxh, dxf, "dw dgrad")
    dwf, dwh = torch.zeros_like(w), torch.zeros_like(w)
    Kn.dwconv_wgrad(dyf, xf, dwf, F, H, H, C, k, s)
    L.check(L.lib.lr_dwconv_wgrad_h(_p(dyh), _p(xh), _p(dwh), F, H, H, C, k, s, _s()))
    assert (dwh - dwf).abs().max().item() <= 1e-4 * dwf.abs().max().item()


def test_frame_pool_scale_act_colsum_h(cuda_device):
    from multimodal_lipread_b200 import kernels as Kn
    L = _lib()
    g = torch.Generator().manual_seed(5)
    F, HW, C = 58, 36, 240
    ah, gh = _bf(torch.randn(F * HW, C, generator=g)).cuda(), _bf(torch.randn(F * HW, C, generator=g)).cuda()
    s, dp = torch.rand(F, C, generator=g).cuda(), torch.randn(F, C, generator=g).cuda()
    for mode in (0, 1):
        pf, ph = torch.empty(F, C, device="cuda"), torch.empty(F, C, device="cuda")
        Kn.frame_reduce(ah.float(), gh.float() if mode else None, pf, F, HW, C, mode)
        L.check(L.lib.lr_frame_reduce_h(_p(ah), _p(gh) if mode else 0, _p(ph), F, HW, C, mode, _s()))
        assert torch.equal(pf, ph)                                   # same fp32 arithmetic in the same order
    of, oh = torch.empty(F * HW, C, device="cuda"), torch.empty(F * HW, C, device="cuda", dtype=torch.bfloat16)
    Kn.frame_scale(ah.float(), s, dp, of, F, HW, C)
    L.check(L.lib.lr_frame_scale_h(_p(ah), _p(s), _p(dp), _p(oh), F, HW, C, _s()))
    _agree(oh, of, "frame_scale")
    dyh = gh.clone()
    dyf = gh.float()
    Kn.act_bwd(dyf, torch.relu(ah.float()), F * HW * C, 1)
    L.check(L.lib.lr_act_bwd_h(_p(dyh), _p(torch.relu(ah)), F * HW * C, 1, _s()))
    assert torch.equal(dyh.float(), dyf)
    cf, ch = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    Kn.colsum(gh.float(), C, F * HW, C, cf)
    L.check(L.lib.lr_colsum_h(_p(gh), C, F * HW, C, _p(ch), _s()))
    assert torch.allclose(cf, ch, rtol=1e-4, atol=1e-3)


def test_im2col_and_maxpool_h(cuda_device):
    from multimodal_lipread_b200 import kernels as Kn, synthetic
    from multimodal_lipread_b200.model_base import video_layout
    L = _lib()
    lips = synthetic.make_lips_u8(2, size=44)[:, :5].contiguous().cuda()
    layout, scale = video_layout(lips)
    is_u8, B, T, H, W, sb, st, sc, sh, sw = layout
    F, Ho = B * T, 22
    src = (int(is_u8), float(scale), F, T, sb, st, sc, sh, sw)
    colf = torch.empty(F * Ho * Ho, 28, device="cuda")
    colh = torch.full((F * Ho * Ho, 32), float("nan"), device="cuda", dtype=torch.bfloat16)
    Kn.im2col(lips, src, H, W, 3, 3, 3, 2, 1, 0, Ho, Ho, colf, 28)
    L.check(L.lib.lr_im2col_h(_p(lips), *src, H, W, 3, 3, 3, 2, 1, 0, Ho, Ho, _p(colh), 32, _s()))
    assert torch.equal(colh[:, :27], colf[:, :27].to(torch.bfloat16)) and colh[:, 27:].float().abs().sum().item() == 0
    g = torch.Generator().manual_seed(2)
    xh = _bf(torch.randn(F, 22, 22, 64, generator=g)).cuda()
    yf, yh = torch.empty(F, 11, 11, 64, device="cuda"), torch.empty(F, 11, 11, 64, device="cuda", dtype=torch.bfloat16)
    af, ah = (torch.empty(F, 11, 11, 64, device="cuda", dtype=torch.uint8) for _ in range(2))
    Kn.maxpool_fwd(xh.float(), yf, af, F, 22, 22, 64, 3, 2, 1)
    L.check(L.lib.lr_maxpool_fwd_h(_p(xh), _p(yh), _p(ah), F, 22, 22, 64, 3, 2, 1, _s()))
    assert torch.equal(yh.float(), yf) and torch.equal(af, ah)
    dyh = _bf(torch.randn(F, 11, 11, 64, generator=g)).cuda()
    dxf, dxh = torch.empty(F, 22, 22, 64, device="cuda"), torch.empty(F, 22, 22, 64, device="cuda", dtype=torch.bfloat16)
    Kn.maxpool_bwd(dyh.float(), af, dxf, F, 22, 22, 64, 3, 2, 1)
    L.check(L.lib.lr_maxpool_bwd_h(_p(dyh), _p(ah), _p(dxh), F, 22, 22, 64, 3, 2, 1, _s()))
    _agree(dxh, dxf, "maxpool bwd")
    colf2 = torch.empty(F * 22 * 22, 9 * 64, device="cuda")
    colh2 = torch.empty(F * 22 * 22, 9 * 64, device="cuda", dtype=torch.bfloat16)
    Kn.im2col_tap(xh.float(), F, 22, 22, 64, 3, 3, 1, 1, 0, 22, 22, colf2)
    L.check(L.lib.lr_im2col_tap_h(_p(xh), F, 22, 22, 64, 3, 3, 1, 1, 1, 0, 22, 22, _p(colh2), _s()))
    assert torch.equal(colh2.float(), colf2)
    # the stride-2 3x3 and 1x1 (downsample) windows, forward and transposed (dgrad) forms: the 16-byte kernels
    for k, pad, tr in ((3, 1, 0), (1, 0, 0), (1, 0, 1), (3, 1, 1)):
        Hd = 22 if tr else 11                                  # transposed: destination = the stride-2 conv's input grid
        Hs = 11 if tr else 22
        src = _bf(torch.randn(F, Hs, Hs, 64, generator=g)).cuda()
        cf = torch.empty(F * Hd * Hd, k * k * 64, device="cuda")
        ch = torch.empty(F * Hd * Hd, k * k * 64, device="cuda", dtype=torch.bfloat16)
        Kn.im2col_tap(src.float(), F, Hs, Hs, 64, k, k, 2, pad, bool(tr), Hd, Hd, cf)
        L.check(L.lib.lr_im2col_tap_h(_p(src), F, Hs, Hs, 64, k, k, 2, pad, pad, tr, Hd, Hd, _p(ch), _s()))
        assert torch.equal(ch.float(), cf), (k, tr)
