"""Implicit-GEMM 3x3 / stride 1 / pad 1 convolution (csrc/conv_igemm.cu: lr_conv3x3_bf16, lr_conv3x3_wgrad_bf16) against
torch.nn.functional.conv2d in float64 on the SAME bf16-rounded operands: forward (+ BatchNorm statistics of the
stored values), input gradient (+ residual) and weight gradient, on the spatial sizes / channel counts of the ResNet-18
trunk at 88 px and 44 px (22, 11, 6, 3 and non-square audio maps), including frame counts that leave ragged tiles."""
import pytest
import torch
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu

SHAPES = [  # F, H, W, Cin, Cout
    (9, 22, 22, 64, 64), (11, 11, 11, 128, 128), (23, 6, 6, 256, 256), (61, 3, 3, 512, 512), (5, 11, 11, 64, 128),
    (4, 20, 30, 64, 64), (3, 10, 15, 128, 128), (7, 5, 8, 256, 256), (2, 44, 44, 64, 64), (30, 2, 2, 128, 64),
]


def _lib():
    from multimodal_lipread_b200 import _lib as L
    return L


def _s():
    return torch.cuda.current_stream().cuda_stream


def _tap(w):  # [Cout, Cin, 3, 3] -> tap-major [Cout][9*Cin]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


@pytest.mark.parametrize("F,H,W,Cin,Cout", SHAPES)
def test_conv3x3_forward_dgrad_wgrad(cuda_device, F, H, W, Cin, Cout):
    L = _lib()
    g = torch.Generator().manual_seed(F * 131 + H * 7 + Cin)
    x = torch.randn(F, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).to(torch.bfloat16).cuda()
    dy = torch.randn(F, H, W, Cout, generator=g).to(torch.bfloat16).cuda()
    res = torch.randn(F, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    xd, wd, dyd = x.double().permute(0, 3, 1, 2), w.double(), dy.double().permute(0, 3, 1, 2)
    # ---- forward + statistics
    y = torch.full((F, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    wt = _tap(w)
    L.check(L.lib.lr_conv3x3_bf16(x.data_ptr(), wt.data_ptr(), y.data_ptr(), 0, stats.data_ptr(), F, H, W, Cin, Cout, 0, _s()))
    ref = Fn.conv2d(xd, wd, padding=1).permute(0, 2, 3, 1)
    assert torch.isfinite(y.float()).all()
    err = (y.double() - ref).abs().max().item()
    assert err <= 2.0 ** -8 * ref.abs().max().item(), (err, ref.abs().max().item())
    yd = y.double().reshape(-1, Cout)
    assert torch.allclose(stats[:Cout], yd.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(stats[Cout:], (yd * yd).sum(0), rtol=1e-5, atol=1e-3)
    # ---- input gradient (+ residual): weights as [Cin][9*Cout], taps mirrored inside the kernel
    wtd = w.permute(1, 2, 3, 0).reshape(Cin, -1).contiguous()
    dx = torch.full((F, H, W, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.lr_conv3x3_bf16(dy.data_ptr(), wtd.data_ptr(), dx.data_ptr(), res.data_ptr(), 0, F, H, W, Cout, Cin, 1, _s()))
    dref = Fn.conv_transpose2d(dyd, wd, padding=1).permute(0, 2, 3, 1) + res.double()
    err = (dx.double() - dref).abs().max().item()
    assert err <= 2.0 ** -8 * dref.abs().max().item(), (err, dref.abs().max().item())
    # ---- weight gradient, accumulated onto an existing fp32 buffer
    g0 = torch.randn(Cout, 9 * Cin, generator=g).cuda()
    dwp = g0.clone()
    L.check(L.lib.lr_conv3x3_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dwp.data_ptr(), F, H, W, Cin, Cout, _s()))
    wref = torch.nn.grad.conv2d_weight(xd, (Cout, Cin, 3, 3), dyd, padding=1)             # [Cout, Cin, 3, 3]
    wref = g0.double() + _tap(wref)
    err = (dwp.double() - wref).abs().max().item()
    assert err <= 3e-5 * wref.abs().max().item() + 1e-4, (err, wref.abs().max().item())


@pytest.mark.parametrize("F,Hi,Wi,Cin,Cout", [(9, 22, 22, 64, 128), (11, 11, 11, 128, 256), (23, 6, 6, 256, 512), (3, 10, 15, 64, 64),
                                              (4, 21, 30, 64, 128)])
def test_stride2_input_gradient_through_zero_stuffing(cuda_device, F, Hi, Wi, Cin, Cout):
    """The input gradient of a 3x3 / stride-2 / pad-1 convolution (BasicBlock conv1 of layer2 / 3 / 4) as
    lr_zero_stuff2_h + lr_conv3x3_bf16(flip = 1) against conv_transpose2d in float64 on the same bf16 operands, even and
    odd input sizes, with a residual."""
    L = _lib()
    g = torch.Generator().manual_seed(F * 17 + Hi + Cin)
    Ho, Wo = (Hi - 1) // 2 + 1, (Wi - 1) // 2 + 1
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).to(torch.bfloat16).cuda()
    dy = torch.randn(F, Ho, Wo, Cout, generator=g).to(torch.bfloat16).cuda()
    res = torch.randn(F, Hi, Wi, Cin, generator=g).to(torch.bfloat16).cuda()
    up = torch.full((F, Hi, Wi, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.lr_zero_stuff2_h(dy.data_ptr(), up.data_ptr(), F, Ho, Wo, Hi, Wi, Cout, _s()))
    want_up = torch.zeros(F, Hi, Wi, Cout, dtype=torch.bfloat16, device="cuda")
    want_up[:, 0:2 * Ho:2, 0:2 * Wo:2] = dy
    assert torch.equal(up, want_up)
    wtd = w.permute(1, 2, 3, 0).reshape(Cin, -1).contiguous()
    dx = torch.full((F, Hi, Wi, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.lib.lr_conv3x3_bf16(up.data_ptr(), wtd.data_ptr(), dx.data_ptr(), res.data_ptr(), 0, F, Hi, Wi, Cout, Cin, 1, _s()))
    x0 = torch.zeros(F, Cin, Hi, Wi, dtype=torch.float64, device="cuda", requires_grad=True)
    Fn.conv2d(x0, w.double(), stride=2, padding=1).backward(dy.double().permute(0, 3, 1, 2))
    dref = x0.grad.permute(0, 2, 3, 1) + res.double()
    err = (dx.double() - dref).abs().max().item()
    assert err <= 2.0 ** -8 * dref.abs().max().item(), (err, dref.abs().max().item())


@pytest.mark.parametrize("F,Hi,Wi,Cin,Cout", [(9, 22, 22, 64, 128), (23, 12, 12, 128, 256), (61, 6, 6, 256, 512), (3, 10, 16, 64, 64),
                                              (2, 44, 44, 64, 128), (5, 20, 30, 64, 64), (40, 4, 4, 128, 64)])
def test_stride2_forward_and_weight_gradient_without_patch_matrix(cuda_device, F, Hi, Wi, Cin, Cout):
    """lr_conv3x3s2_bf16 / lr_conv3x3s2_wgrad_bf16 (the 5-D parity view of the input: every tap of the stride-2 window is
    a dense TMA box) against conv2d(stride=2, padding=1) and its weight gradient in float64 on the same bf16 operands."""
    L = _lib()
    g = torch.Generator().manual_seed(F * 19 + Hi + Cin)
    Ho, Wo = Hi // 2, Wi // 2
    x = torch.randn(F, Hi, Wi, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).to(torch.bfloat16).cuda()
    dy = torch.randn(F, Ho, Wo, Cout, generator=g).to(torch.bfloat16).cuda()
    xd, wd, dyd = x.double().permute(0, 3, 1, 2), w.double(), dy.double().permute(0, 3, 1, 2)
    y = torch.full((F, Ho, Wo, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    wt = _tap(w)
    L.check(L.lib.lr_conv3x3s2_bf16(x.data_ptr(), wt.data_ptr(), y.data_ptr(), stats.data_ptr(), F, Hi, Wi, Cin, Cout, _s()))
    ref = Fn.conv2d(xd, wd, stride=2, padding=1).permute(0, 2, 3, 1)
    assert ref.shape[1:3] == (Ho, Wo) and torch.isfinite(y.float()).all()
    err = (y.double() - ref).abs().max().item()
    assert err <= 2.0 ** -8 * ref.abs().max().item(), (err, ref.abs().max().item())
    yd = y.double().reshape(-1, Cout)
    assert torch.allclose(stats[:Cout], yd.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(stats[Cout:], (yd * yd).sum(0), rtol=1e-5, atol=1e-3)
    g0 = torch.randn(Cout, 9 * Cin, generator=g).cuda()
    dwp = g0.clone()
    L.check(L.lib.lr_conv3x3s2_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dwp.data_ptr(), F, Hi, Wi, Cin, Cout, _s()))
    wref = torch.nn.grad.conv2d_weight(xd, (Cout, Cin, 3, 3), dyd, stride=2, padding=1)
    wref = g0.double() + _tap(wref)
    err = (dwp.double() - wref).abs().max().item()
    assert err <= 3e-5 * wref.abs().max().item() + 1e-4, (err, wref.abs().max().item())
