"""Stand-alone launches of the trunk's streaming kernels at the headline step's largest shapes (for ncu and for
CUDA-event timing after an L2 flush).  usage: python tools/microbench.py [bn|dw|gemm|all] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_lipread_b200 import _lib
L = _lib.lib
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timed(name, fn, nbytes):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{name:44s} {best*1e3:8.1f} us  {nbytes/1e6:8.1f} MB  {nbytes/1e9/(best/1e3):7.0f} GB/s", flush=True)

p = lambda t: 0 if t is None else t.data_ptr()
for rows, C in ((449152, 72), (1796608, 16), (112288, 96), (33408, 240)):
    if which not in ("bn", "all"): break
    bn = torch.nn.BatchNorm2d(C).to(dev)
    for dt, sfx, es in ((torch.float32, "", 4), (torch.bfloat16, "_h", 2)):
        x = torch.randn(rows, C, device=dev).to(dt); dz = torch.randn(rows, C, device=dev).to(dt)
        z = torch.empty_like(x); dx = torch.empty_like(x)
        stats = torch.cat([x.double().sum(0), (x.double() ** 2).sum(0)])
        sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        fwd = getattr(L, "lr_bn_act_fwd" + sfx); bwd = getattr(L, "lr_bn_act_bwd" + sfx)
        timed(f"bn_act_fwd{sfx} [{rows}x{C}]", lambda: _lib.check(fwd(p(x), p(stats), p(bn.weight), p(bn.bias), p(bn.running_mean), p(bn.running_var), p(bn.num_batches_tracked), 1e-3, 0.01, 2, 1, 0, 0, p(z), rows, C, s)), 2 * es * rows * C)
        timed(f"bn_act_bwd{sfx} [{rows}x{C}]", lambda: (sums.zero_(), _lib.check(bwd(p(x), p(stats), p(bn.weight), p(bn.bias), p(bn.running_mean), p(bn.running_var), 1e-3, 2, 1, p(dz), 0, 0, p(sums), p(dx), p(dg), p(db), rows, C, s))), 3 * es * rows * C)
for F, H, C, k, st in ((928, 22, 72, 3, 2), (928, 44, 16, 3, 2), (928, 11, 88, 3, 1), (928, 6, 240, 5, 1)):
    if which not in ("dw", "all"): break
    Ho = (H + 2 * (k // 2) - k) // st + 1
    w = torch.randn(C, 1, k, k, device=dev)
    for dt, sfx, es in ((torch.float32, "", 4), (torch.bfloat16, "_h", 2)):
        x = torch.randn(F, H, H, C, device=dev).to(dt); y = torch.empty(F, Ho, Ho, C, device=dev, dtype=dt)
        dy = torch.randn(F, Ho, Ho, C, device=dev).to(dt); dx = torch.empty_like(x); dw = torch.zeros_like(w)
        stt = torch.zeros(2 * C, dtype=torch.float64, device=dev)
        nb = es * F * C * (H * H + Ho * Ho)
        timed(f"dwconv_fwd{sfx} [{F},{H},{C},k{k}s{st}]", lambda: _lib.check(getattr(L, "lr_dwconv_fwd" + sfx)(p(x), p(w), p(y), p(stt), F, H, H, C, k, st, s)), nb)
        timed(f"dwconv_dgrad{sfx} [{F},{H},{C},k{k}s{st}]", lambda: _lib.check(getattr(L, "lr_dwconv_dgrad" + sfx)(p(dy), p(w), p(dx), F, H, H, C, k, st, s)), nb)
        timed(f"dwconv_wgrad{sfx} [{F},{H},{C},k{k}s{st}]", lambda: _lib.check(getattr(L, "lr_dwconv_wgrad" + sfx)(p(dy), p(x), p(dw), F, H, H, C, k, st, s)), nb)
for M, N, K in ((1796608, 16, 32), (449152, 72, 16), (449152, 16, 72), (112288, 88, 24), (33408, 240, 40), (8352, 576, 96)):
    if which not in ("gemm", "all"): break
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    Cc = torch.empty(M, N, device=dev, dtype=torch.bfloat16); stt = torch.zeros(2 * N, dtype=torch.float64, device=dev)
    timed(f"gemm_bf16 fwd+stats [{M}x{N}x{K}]", lambda: _lib.check(L.lr_gemm_bf16(p(A), K, 0, p(B), K, 0, p(Cc), N, 1, M, N, K, 0, 0, 0, 0, p(stt), 1, s)), 2 * (M * K + N * K + M * N))
    timed(f"gemm_bf16 fwd       [{M}x{N}x{K}]", lambda: _lib.check(L.lr_gemm_bf16(p(A), K, 0, p(B), K, 0, p(Cc), N, 1, M, N, K, 0, 0, 0, 0, 0, 1, s)), 2 * (M * K + N * K + M * N))
for H in (128, 256, 512):
    if which not in ("lstm", "all"): break
    B, T = 32, 29
    xp = torch.randn(B * T, 4 * H, device=dev) * 0.5; whh = torch.randn(4 * H, H, device=dev) * (H ** -0.5); bhh = torch.zeros(4 * H, device=dev)
    out = torch.zeros(B, T, H, device=dev); gates = torch.zeros(B, T, 4 * H, device=dev); cst = torch.zeros(B, T, H, device=dev); hp = torch.zeros(B, T, H, device=dev)
    dout = torch.randn(B, T, H, device=dev); dg = torch.zeros(B, T, 4 * H, device=dev)
    for sfx in ("", "_tc"):
        f, b = getattr(L, "lr_lstm_fwd" + sfx), getattr(L, "lr_lstm_bwd" + sfx)
        timed(f"lstm_fwd{sfx} [B{B} T{T} H{H}]", lambda: _lib.check(f(p(xp), 4 * H, p(bhh), p(whh), p(out), H, p(gates), p(cst), p(hp), B, T, H, T, 0, s)), 4 * B * T * 11 * H)
        timed(f"lstm_bwd{sfx} [B{B} T{T} H{H}]", lambda: _lib.check(b(p(dout), H, -1, p(gates), p(cst), p(whh), p(dg), B, T, H, T, 0, s)), 4 * B * T * 11 * H)
