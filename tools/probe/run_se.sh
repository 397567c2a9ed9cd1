python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "se_gate" 2>&1 | tail -15 > gpurun_out/r2_se_test.log
python -m pytest tests/test_midfusion_gpu.py tests/test_models_gpu.py -q -m gpu -x -k "midfusion or mobilenet or fast" 2>&1 | tail -15 >> gpurun_out/r2_se_test.log
python - > gpurun_out/r2_se_micro.log 2>&1 <<'PY'
import torch
from multimodal_lipread_b200 import kernels as K
def t(fn, n=20):
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
F = 928
for C, Cs in ((576, 144), (288, 72), (240, 64), (144, 40), (120, 32), (96, 24), (16, 8)):
    p = torch.randn(F, C, device="cuda"); w1 = torch.randn(Cs, C, device="cuda") * 0.05; b1 = torch.zeros(Cs, device="cuda")
    w2 = torch.randn(C, Cs, device="cuda") * 0.1; b2 = torch.zeros(C, device="cuda")
    h1 = torch.empty(F, Cs, device="cuda"); s = torch.empty(F, C, device="cuda")
    ds = torch.randn(F, C, device="cuda"); dz1 = torch.empty(F, Cs, device="cuda"); dp = torch.empty(F, C, device="cuda")
    f = t(lambda: K.se_fc_fwd(p, w1, b1, w2, b2, h1, s, F, C, Cs))
    b = t(lambda: K.se_fc_bwd(ds, s, h1, w1, w2, dz1, dp, F, C, Cs))
    print(f"C {C} Cs {Cs}: fwd {f:.1f} us  bwd {b:.1f} us (graph of 20 launches, warm L2)")
PY
for f in 0 1; do LIPREAD_SE_FUSED=$f python bench.py --steps 30 --warmup 5 --no-sub-records 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('se_fused $f', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])" >> gpurun_out/r2_se_micro.log; done
python -m pytest tests/test_models_gpu.py -q -m gpu -k "resnet34" -s 2>&1 | grep -v "^$" | tail -12 >> gpurun_out/r2_se_test.log
