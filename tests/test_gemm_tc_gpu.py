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
    Kn.gemm_tf32(A.cuda(), K, B.cuda(), K, C, N, M, N, K, stats=stats)
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
        Kn.gemm_tf32(A.cuda(), K + 8, B.cuda(), K, Cbig[:, 4:], N + 12, M, N, K, bias=bias.cuda(), act=act,
                     R=R.cuda(), ldr=N + 4)
        ref = fn(A[:, :K].double() @ B.double().t() + bias.double()) + R[:, :N].double()
        err = (Cbig[:, 4:4 + N].cpu().double() - ref).abs().max().item()
        assert err <= 2e-3 * ref.abs().max().item(), (act, err)
        assert Cbig[:, :4].abs().sum().item() == 0 and Cbig[:, 4 + N:].abs().sum().item() == 0
