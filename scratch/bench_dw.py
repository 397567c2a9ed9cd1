import sys, torch
sys.path.insert(0, '.')
from multimodal_lipread_b200 import kernels as K
F = 928
layers = [(16,3,2,44),(72,3,2,22),(88,3,1,11),(96,5,2,11),(240,5,1,6),(120,5,1,6),(144,5,1,6),(288,5,2,6),(576,5,1,3)]
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
tot = [0, 0, 0]
for (C, k, s, H) in layers:
    Ho = (H + 2*(k//2) - k)//s + 1
    x = torch.randn(F, H, H, C, device='cuda'); w = torch.randn(C, 1, k, k, device='cuda')
    y = torch.empty(F, Ho, Ho, C, device='cuda'); dy = torch.randn_like(y); dx = torch.empty_like(x); dw = torch.zeros_like(w)
    st = torch.zeros(2*C, dtype=torch.float64, device='cuda')
    t1 = timeit(lambda: K.dwconv_fwd(x, w, y, st, F, H, H, C, k, s))
    t2 = timeit(lambda: K.dwconv_dgrad(dy, w, dx, F, H, H, C, k, s))
    t3 = timeit(lambda: K.dwconv_wgrad(dy, x, dw, F, H, H, C, k, s))
    mb = (x.numel() + y.numel()) * 4 / 1e6
    print(f"C={C:4d} k={k} s={s} H={H:3d}: fwd {t1:7.1f}us dgrad {t2:7.1f}us wgrad {t3:7.1f}us   ({mb:6.1f} MB in+out -> {mb/6544.7*1e3:6.1f}us at HBM peak)")
    for i, t in enumerate((t1, t2, t3)): tot[i] += t * (2 if (C, H) in ((240, 6), (576, 3)) else 1)
print("step totals (us): fwd %.0f dgrad %.0f wgrad %.0f" % tuple(tot))
