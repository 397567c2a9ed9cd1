export PYTHONPATH=$PWD
python -m pytest tests/test_conv2d_gpu.py tests/test_bf16_kernels_gpu.py -q -m gpu -k "maxpool" 2>&1 | tail -5 > gpurun_out/r2_pool.log
python - >> gpurun_out/r2_pool.log 2>&1 <<'PY'
import torch
from multimodal_lipread_b200 import _lib as L
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * n) * 1e3
F, H, C = 928, 44, 64
x = torch.randn(F, H, H, C, device="cuda").bfloat16()
y = torch.empty(F, 22, 22, C, device="cuda", dtype=torch.bfloat16)
arg = torch.empty(F, 22, 22, C, device="cuda", dtype=torch.uint8)
dy = torch.randn(F, 22, 22, C, device="cuda").bfloat16()
dx = torch.empty_like(x)
s = lambda: torch.cuda.current_stream().cuda_stream
f = t(lambda: L.check(L.lib.lr_maxpool_fwd_h(x.data_ptr(), y.data_ptr(), arg.data_ptr(), F, H, H, C, 3, 2, 1, s())))
b = t(lambda: L.check(L.lib.lr_maxpool_bwd_h(dy.data_ptr(), arg.data_ptr(), dx.data_ptr(), F, H, H, C, 3, 2, 1, s())))
fb = (F * H * H * C * 2 + F * 22 * 22 * C * 3)
print(f"maxpool 3/2/1 928x44x44x64 bf16: fwd {f:.1f} us ({fb / f / 1e3:.0f} GB/s)  bwd {b:.1f} us ({fb / b / 1e3:.0f} GB/s)")
PY
python bench.py --workload video_resnet_lstm --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config 2', round(d['value']), round(d['ms_per_step'],3))" >> gpurun_out/r2_pool.log
