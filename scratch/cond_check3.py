import sys, os, torch, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import av_models as O
from oracle.frontend import lips_u8_to_model_input
from multimodal_lipread_b200 import synthetic
for B, T, size in [(2, 3, 88), (2, 4, 88), (3, 3, 88)]:
    lips = synthetic.make_lips_u8(B, size=size)[:, :T].contiguous()
    video, labels = lips_u8_to_model_input(lips), synthetic.make_labels(B, 40)
    torch.manual_seed(0)
    base = O.ResNet2DBiLSTMOracle(40, O.DictConfig({"model": {"dropout": 0.0}})).train()
    def grads(v):
        m = copy.deepcopy(base)
        torch.nn.functional.cross_entropy(m(v), labels).backward()
        return {n: p.grad.clone() for n, p in m.named_parameters()}
    g0 = grads(video)
    gen = torch.Generator().manual_seed(5)
    out = []
    for i in range(6):
        v = video * (1 + 2e-7 * torch.randn(video.shape, generator=gen))
        g = grads(v)
        worst = max(((g[n] - g0[n]).abs().max().item() / (g0[n].abs().max().item() + 3e-5), n) for n in g0)
        out.append(f"{worst[0]:.1e}")
    print(B, T, size, out)
