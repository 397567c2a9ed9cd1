"""tcgen05 / TMA / TMEM GEMM (lr_gemm_tf32) against fp64 matmul on the shapes of the MobileNetV3 1x1 convolutions.
TF32 keeps 10 mantissa bits of each operand (truncation) and accumulates in fp32: tolerance 2e-3 of max|ref|
(bf16 would be 8e-3); the statistics epilogue is checked against the kernel's own output (exact up to fp32 sums)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(4000, 72, 16), (1000, 24, 72), (129, 16, 16), (5000, 96, 24), (777, 40, 96), (640, 240, 40),
          (300, 576, 96), (333, 96, 576), (2000, 288, 48), (128, 16, 8), (64, 144, 40), (1, 24, 88)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tf32_shapes(cuda_device, M, N, K):
    from multimodal_lipread_b200 import kernels as Kn
    g = torch.Generator().manual_seed(M + N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().t()
    C = torch.full((M, N), float("nan"), device="cuda")
    stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
    Kn.gemm_tf32(A.cuda(), K, 0, B.cuda(), K, 0, C, N, M, N, K, stats=stats)
    torch.cuda.synchronize()
    assert torch.isfinite(C).all()
    err = (C.cpu().double() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item(), (err, ref.abs().max().item())
    Cd = C.double()
    assert torch.allclose(stats[:N], Cd.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(stats[N:], (Cd * Cd).sum(0), rtol=1e-5, atol=1e-3)


def test_gemm_tf32_epilogue_and_strides(cuda_device):
    from multimodal_lipread_b200 import kernels as Kn
    g = torch.Generator().manual_seed(0)
    M, N, K = 1500, 88, 24
    A, B = torch.randn(M, K + 8, generator=g), torch.randn(N, K, generator=g)
    bias, R = torch.randn(N, generator=g), torch.randn(M, N + 4, generator=g)
    Cbig = torch.zeros(M, N + 12, device="cuda")
    for act, fn in ((0, lambda u: u), (1, torch.relu), (2, torch.nn.functional.hardswish)):
        Kn.gemm_tf32(A.cuda(), K + 8, 0, B.cuda(), K, 0, Cbig[:, 4:], N + 12, M, N, K, bias=bias.cuda(), act=act,
                     R=R.cuda(), ldr=N + 4)
        ref = fn(A[:, :K].double() @ B.double().t() + bias.double()) + R[:, :N].double()
        err = (Cbig[:, 4:4 + N].cpu().double() - ref).abs().max().item()
        assert err <= 2e-3 * ref.abs().max().item(), (act, err)
        assert Cbig[:, :4].abs().sum().item() == 0 and Cbig[:, 4 + N:].abs().sum().item() == 0


# dgrad shapes: dX[M, Cin] = dY[M, Cout] @ W[Cout, Cin]  (B stored [K][N]: MN-major operand)
@pytest.mark.parametrize("M,Cout,Cin", [(4000, 72, 16), (1000, 24, 72), (300, 576, 96), (333, 96, 576), (2000, 288, 48),
                                        (129, 16, 16), (640, 40, 240), (500, 88, 24)])
def test_gemm_tf32_dgrad_layout(cuda_device, M, Cout, Cin):
    from multimodal_lipread_b200 import kernels as Kn
    g = torch.Generator().manual_seed(M + Cout)
    dY, W, R = torch.randn(M, Cout, generator=g), torch.randn(Cout, Cin, generator=g), torch.randn(M, Cin, generator=g)
    ref = dY.double() @ W.double() + R.double()
    dX = torch.full((M, Cin), float("nan"), device="cuda")
    Kn.gemm_tf32(dY.cuda(), Cout, 0, W.cuda(), Cin, 1, dX, Cin, M, Cin, Cout, R=R.cuda(), ldr=Cin)
    err = (dX.cpu().double() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item(), (err, ref.abs().max().item())


# wgrad shapes: dW[Cout, Cin] += dY[M, Cout]^T @ X[M, Cin]  (both operands MN-major, split over the M rows)
@pytest.mark.parametrize("M,Cout,Cin,ks", [(40000, 72, 16, 37), (9000, 24, 72, 8), (3000, 576, 96, 5), (3333, 96, 576, 3),
                                           (2000, 288, 48, 1), (129, 16, 16, 1), (5000, 40, 240, 11), (70, 88, 24, 2)])
def test_gemm_tf32_wgrad_layout(cuda_device, M, Cout, Cin, ks):
    from multimodal_lipread_b200 import kernels as Kn
    g = torch.Generator().manual_seed(M + Cin)
    dY, X, W0 = torch.randn(M, Cout, generator=g), torch.randn(M, Cin, generator=g), torch.randn(Cout, Cin, generator=g)
    ref = W0.double() + dY.double().t() @ X.double()
    dW = W0.clone().cuda()
    if ks > 1:
        Kn.gemm_tf32(dY.cuda(), Cout, 1, X.cuda(), Cin, 1, dW, Cin, Cout, Cin, M, ksplit=ks)
    else:
        Kn.gemm_tf32(dY.cuda(), Cout, 1, X.cuda(), Cin, 1, dW, Cin, Cout, Cin, M, R=dW, ldr=Cin)
    err = (dW.cpu().double() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item(), (err, ref.abs().max().item())
