"""Per-kernel parity of the dense-convolution building blocks (csrc/conv2d.cu, the res_pre BatchNorm mode) against
the torch CPU ops they replace."""
import pytest
import torch
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu


def _cl(x):
    """NCHW -> channels-last rows [F, H, W, C] contiguous."""
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("Cin,Cout,k,stride,pad,H,W", [
    (8, 16, 3, 1, 1, 9, 7), (8, 12, 3, 2, 1, 11, 10), (3, 8, 7, 2, 3, 20, 22), (1, 4, 3, 1, 1, 6, 5),
    (16, 8, 1, 2, 0, 9, 9),
])
def test_im2col_conv_fwd_dgrad_wgrad(cuda_device, Cin, Cout, k, stride, pad, H, W):
    from multimodal_lipread_b200 import kernels as K
    torch.manual_seed(0)
    F = 3
    x = torch.randn(F, Cin, H, W, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, requires_grad=True)
    y = Fn.conv2d(x, w, stride=stride, padding=pad)
    dy = torch.randn_like(y)
    y.backward(dy)
    Ho, Wo = y.shape[2], y.shape[3]
    Kd = Cin * k * k
    ldk = (Kd + 3) // 4 * 4
    xd = _cl(x.detach()).cuda()
    col = torch.full((F * Ho * Wo, ldk), 7.0, device="cuda")
    K.im2col(xd, K.nhwc_layout(F, H, W, Cin), H, W, Cin, k, k, stride, pad, False, Ho, Wo, col, ldk)
    wp = torch.zeros(Cout, ldk, device="cuda")
    wp[:, :Kd] = w.detach().reshape(Cout, Kd).cuda()
    yd = torch.empty(F * Ho * Wo, Cout, device="cuda")
    K.gemm(col, ldk, 0, wp, ldk, 0, yd, Cout, F * Ho * Wo, Cout, ldk)
    torch.testing.assert_close(yd.cpu().view(F, Ho, Wo, Cout), _cl(y.detach()), rtol=1e-4, atol=1e-4)
    assert torch.all(col[:, Kd:] == 0)
    # wgrad: dW = dy^T col
    dyd = _cl(dy).cuda().view(F * Ho * Wo, Cout)
    dw = torch.zeros(Cout, ldk, device="cuda")
    K.gemm(dyd, Cout, 1, col, ldk, 1, dw, ldk, Cout, ldk, F * Ho * Wo, R=dw, ldr=ldk)
    torch.testing.assert_close(dw[:, :Kd].cpu().view_as(w), w.grad, rtol=1e-4, atol=2e-4)
    # dgrad: dx = im2col_T(dy) . Wt^T
    Kt = Cout * k * k
    ldt = (Kt + 3) // 4 * 4
    wt = torch.zeros(Cin, ldt, device="cuda")
    K.weight_transpose(w.detach().cuda().contiguous(), wt, Cout, Cin, k * k, ldt)
    colT = torch.empty(F * H * W, ldt, device="cuda")
    K.im2col(dyd, K.nhwc_layout(F, Ho, Wo, Cout), Ho, Wo, Cout, k, k, stride, pad, True, H, W, colT, ldt)
    dx = torch.empty(F * H * W, Cin, device="cuda")
    K.gemm(colT, ldt, 0, wt, ldt, 0, dx, Cin, F * H * W, Cin, ldt)
    torch.testing.assert_close(dx.cpu().view(F, H, W, Cin), _cl(x.grad), rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("k,stride,pad", [(7, 2, 3), (3, 2, 1), (3, 1, 1)])
def test_im2col_reads_uint8_frames_in_place(cuda_device, k, stride, pad):
    """The stem reads (B,T,H,W,3) uint8 frames with /255 and the TimeDistributed reshape folded into the addressing
    (3x3 windows take the one-thread-per-pixel fast path)."""
    from multimodal_lipread_b200 import kernels as K
    from multimodal_lipread_b200.model_base import video_layout
    torch.manual_seed(0)
    B, T, H, W = 2, 3, 10, 12
    lips = torch.randint(0, 256, (B, T, H, W, 3), dtype=torch.uint8)
    x = (lips.float() / 255.0).permute(0, 1, 4, 2, 3).reshape(B * T, 3, H, W)
    ref = Fn.unfold(x, k, padding=pad, stride=stride)                       # [F, 3*k*k, L]
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    (is_u8, _, _, _, _, sb, st, sc, sh, sw), scale = video_layout(lips.cuda())
    Kd, ldk = 3 * k * k, (3 * k * k + 3) // 4 * 4
    col = torch.empty(B * T * Ho * Wo, ldk, device="cuda")
    K.im2col(lips.cuda(), (is_u8, scale, B * T, T, sb, st, sc, sh, sw), H, W, 3, k, k, stride, pad, False, Ho, Wo, col, ldk)
    got = col[:, :Kd].cpu().view(B * T, Ho * Wo, Kd).permute(0, 2, 1)
    torch.testing.assert_close(got, ref, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("H,W,dt", [(88, 88, torch.float32), (88, 88, torch.bfloat16), (21, 30, torch.float32), (44, 44, torch.bfloat16)])
def test_stem_patch_rows_7x7_stride2(cuda_device, H, W, dt):
    """The ResNet stem's patch matrix (7x7 / stride 2 / pad 3 on uint8 frames, row pitch 152 = 147 columns + zero tail)
    takes the staged one-block-per-output-row kernel; fp32 and bf16 rows, odd sizes, against unfold."""
    from multimodal_lipread_b200 import _lib as L
    from multimodal_lipread_b200.model_base import video_layout
    torch.manual_seed(H)
    B, T = 2, 3
    lips = torch.randint(0, 256, (B, T, H, W, 3), dtype=torch.uint8)
    x = (lips.float() / 255.0).permute(0, 1, 4, 2, 3).reshape(B * T, 3, H, W)
    ref = Fn.unfold(x, 7, padding=3, stride=2)
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    d = lips.cuda()
    (is_u8, _, _, _, _, sb, st, sc, sh, sw), scale = video_layout(d)
    col = torch.full((B * T * Ho * Wo, 152), float("nan"), device="cuda", dtype=dt)
    fn = L.lib.lr_im2col if dt == torch.float32 else L.lib.lr_im2col_h
    L.check(fn(d.data_ptr(), int(is_u8), float(scale), B * T, T, sb, st, sc, sh, sw, H, W, 3, 7, 7, 2, 3, 0, Ho, Wo,
               col.data_ptr(), 152, torch.cuda.current_stream().cuda_stream))
    got = col[:, :147].float().cpu().view(B * T, Ho * Wo, 147).permute(0, 2, 1)
    want = ref if dt == torch.float32 else ref.to(torch.bfloat16).float()
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
    assert col[:, 147:].float().abs().sum().item() == 0


@pytest.mark.parametrize("k,stride,pad,H,W", [(2, 2, 0, 8, 11), (3, 2, 1, 9, 10), (3, 2, 1, 22, 22)])
def test_maxpool_fwd_bwd(cuda_device, k, stride, pad, H, W):
    from multimodal_lipread_b200 import kernels as K
    torch.manual_seed(1)
    F, C = 3, 8
    x = torch.relu(torch.randn(F, C, H, W)).requires_grad_(True)          # many exact ties at 0 (post-ReLU input)
    y = Fn.max_pool2d(x, k, stride, pad)
    dy = torch.randn_like(y)
    y.backward(dy)
    Ho, Wo = y.shape[2], y.shape[3]
    yd = torch.empty(F, Ho, Wo, C, device="cuda")
    arg = torch.empty(F, Ho, Wo, C, dtype=torch.uint8, device="cuda")
    K.maxpool_fwd(_cl(x.detach()).cuda(), yd, arg, F, H, W, C, k, stride, pad)
    assert torch.equal(yd.cpu(), _cl(y.detach()))
    dx = torch.empty(F, H, W, C, device="cuda")
    K.maxpool_bwd(_cl(dy).cuda(), arg, dx, F, H, W, C, k, stride, pad)
    torch.testing.assert_close(dx.cpu(), _cl(x.grad), rtol=1e-6, atol=1e-6)


def test_dropout_kernels(cuda_device):
    from multimodal_lipread_b200 import kernels as K
    n, p = 1 << 20, 0.3
    x = torch.randn(n, device="cuda")
    y, mask = torch.empty_like(x), torch.empty(n, dtype=torch.uint8, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    K.dropout_fwd(x, y, mask, n, p, 17, step)
    keep = mask.float().mean().item()
    assert abs(keep - (1 - p)) < 5e-3
    torch.testing.assert_close(y, torch.where(mask.bool(), x / (1 - p), torch.zeros_like(x)))
    y2, mask2 = torch.empty_like(x), torch.empty_like(mask)
    K.dropout_fwd(x, y2, mask2, n, p, 17, step)
    assert torch.equal(mask, mask2)                        # same (seed, step): same mask
    K.rng_tick(step)
    K.dropout_fwd(x, y2, mask2, n, p, 17, step)
    assert int(step.item()) == 1 and not torch.equal(mask, mask2)
    agree = (mask == mask2).float().mean().item()
    assert abs(agree - ((1 - p) ** 2 + p ** 2)) < 5e-3     # independent draws
    dy, dx = torch.randn_like(x), torch.empty_like(x)
    K.dropout_bwd(dy, mask, dx, n, p)
    torch.testing.assert_close(dx, torch.where(mask.bool(), dy / (1 - p), torch.zeros_like(dy)))
    K.dropout_fwd(x, y, mask, n, 0.0, 17, step)
    assert torch.equal(y, x) and bool(mask.all())


@pytest.mark.parametrize("training", [True, False])
def test_bn_residual_before_activation(cuda_device, training):
    """BasicBlock tail: z = relu(bn(x) + identity); backward through the output."""
    from multimodal_lipread_b200 import kernels as K
    torch.manual_seed(2)
    F, C, H, W = 4, 16, 5, 6
    bn = torch.nn.BatchNorm2d(C)
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_()
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
    bn.train(training)
    bnd = torch.nn.BatchNorm2d(C)
    bnd.load_state_dict(bn.state_dict())
    bnd.cuda()
    x = torch.randn(F, C, H, W, requires_grad=True)
    r = torch.randn(F, C, H, W, requires_grad=True)
    z = torch.relu(bn(x) + r)
    dz = torch.randn_like(z)
    z.backward(dz)
    rows = F * H * W
    xd, rd = _cl(x.detach()).cuda(), _cl(r.detach()).cuda()
    stats = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    xf = xd.view(rows, C).double()
    stats[:C], stats[C:] = xf.sum(0), (xf * xf).sum(0)
    zd = torch.empty_like(xd)
    K.bn_act_fwd(xd, stats if training else None, bnd, K.ACT_RELU, training, zd, rows, C, residual=rd, res_pre=True)
    torch.testing.assert_close(zd.cpu(), _cl(z.detach()), rtol=1e-5, atol=1e-5)
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    dx, dres = torch.empty_like(xd), torch.empty_like(xd)
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    K.bn_act_bwd(xd, stats if training else None, bnd, K.ACT_RELU, training, _cl(dz).cuda(), sums, dx, dgamma, dbeta,
                 rows, C, z_out=zd, dres=dres)
    torch.testing.assert_close(dres.cpu(), _cl(r.grad), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dx.cpu(), _cl(x.grad), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dgamma.cpu(), bn.weight.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dbeta.cpu(), bn.bias.grad, rtol=1e-4, atol=1e-4)


def test_relu6_and_act_fwd(cuda_device):
    from multimodal_lipread_b200 import kernels as K
    from multimodal_lipread_b200._lib import lib, check
    x = torch.linspace(-8, 8, 1001, device="cuda")
    y = torch.empty_like(x)
    check(lib.lr_act_fwd(x.data_ptr(), y.data_ptr(), x.numel(), K.ACT_RELU6, torch.cuda.current_stream().cuda_stream))
    torch.testing.assert_close(y, Fn.relu6(x))
    dy = torch.ones_like(x)
    K.act_bwd(dy, y, x.numel(), K.ACT_RELU6)
    assert torch.equal(dy, ((y > 0) & (y < 6)).float())


@pytest.mark.parametrize("Cin,Cout,k,stride,pad,H,W", [(8, 16, 3, 1, 1, 9, 7), (8, 12, 3, 2, 1, 11, 10), (16, 8, 1, 2, 0, 9, 9),
                                                        (64, 64, 3, 1, 1, 11, 11), (32, 64, 3, 2, 1, 6, 6)])
def test_tap_major_conv_fwd_dgrad_wgrad(cuda_device, Cin, Cout, k, stride, pad, H, W):
    """The ResNet path: tap-major patch matrix (lr_im2col_tap) with the weight layouts of lr_weight_tap."""
    from multimodal_lipread_b200 import kernels as K
    torch.manual_seed(0)
    F = 3
    x = torch.randn(F, Cin, H, W, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k, requires_grad=True)
    y = Fn.conv2d(x, w, stride=stride, padding=pad)
    dy = torch.randn_like(y)
    y.backward(dy)
    Ho, Wo = y.shape[2], y.shape[3]
    Kd, kk = Cin * k * k, k * k
    xd, wd_ = _cl(x.detach()).cuda(), w.detach().cuda().contiguous()
    col = torch.full((F * Ho * Wo, Kd), 7.0, device="cuda")
    K.im2col_tap(xd, F, H, W, Cin, k, k, stride, pad, False, Ho, Wo, col)
    wp = torch.empty(Cout, Kd, device="cuda")
    K.weight_tap(wd_, wp, Cout, Cin, kk, 0)
    assert torch.equal(wp.cpu().view(Cout, kk, Cin), w.detach().reshape(Cout, Cin, kk).permute(0, 2, 1))
    yd = torch.empty(F * Ho * Wo, Cout, device="cuda")
    K.gemm(col, Kd, 0, wp, Kd, 0, yd, Cout, F * Ho * Wo, Cout, Kd)
    torch.testing.assert_close(yd.cpu().view(F, Ho, Wo, Cout), _cl(y.detach()), rtol=1e-4, atol=1e-4)
    dyd = _cl(dy).cuda().view(F * Ho * Wo, Cout)
    dwp = torch.zeros(Cout, Kd, device="cuda")
    K.gemm(dyd, Cout, 1, col, Kd, 1, dwp, Kd, Cout, Kd, F * Ho * Wo, R=dwp, ldr=Kd)
    dw = torch.empty(Cout, Cin, k, k, device="cuda")
    K.weight_tap(dwp, dw, Cout, Cin, kk, 2)
    torch.testing.assert_close(dw.cpu(), w.grad, rtol=1e-4, atol=3e-4)
    Kt = Cout * kk
    wt = torch.empty(Cin, Kt, device="cuda")
    K.weight_tap(wd_, wt, Cout, Cin, kk, 1)
    colT = torch.empty(F * H * W, Kt, device="cuda")
    K.im2col_tap(dyd, F, Ho, Wo, Cout, k, k, stride, pad, True, H, W, colT)
    dx = torch.empty(F * H * W, Cin, device="cuda")
    K.gemm(colT, Kt, 0, wt, Kt, 0, dx, Cin, F * H * W, Cin, Kt)
    torch.testing.assert_close(dx.cpu().view(F, H, W, Cin), _cl(x.grad), rtol=1e-4, atol=3e-4)
