import sys, os, torch, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_models_gpu as T
from oracle.frontend import lips_u8_to_model_input
name = "video_resnet_lstm"
B, TT, size = 2, 3, 88
ref0, ours, C = T._case(name)
wav, mel, lips, labels = T._data(B, size, TT, C)
video = lips_u8_to_model_input(lips)
gen = torch.Generator().manual_seed(5)
vs = [video] + [video * (1 + 2e-7 * torch.randn(video.shape, generator=gen)) for _ in range(3)]
v = vs[2]
ref = copy.deepcopy(ref0).train()
torch.nn.functional.cross_entropy(ref(v), labels).backward()
prev = None
for rep in range(3):
    ours.train(); ours.configure_optimizer(lr=0.0)
    loss, logits = ours.train_step(v.cuda(), labels.cuda(), use_graph=False)
    flat = ours._flat
    g = flat.grad.clone()
    if prev is not None:
        print("rep diff", (g - prev).abs().max().item())
    prev = g
    rows = [(T._grad_err(flat.g(p), q.grad, 3e-3), n, q.grad.abs().max().item(), (flat.g(p).cpu() - q.grad).abs().max().item()) for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters())]
    for e, n, m, d in rows:
        if e > 5e-4:
            print(f"  rep{rep} {e:.2e} {n} max|g| {m:.2e} maxdiff {d:.2e}")
# where is the difference located inside the worst tensor?
for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
    if n == "cnn_features.6.1.conv1.weight":
        d = (flat.g(p).cpu() - q.grad).abs()
        idx = d.flatten().topk(8).indices
        print("top diffs at", [tuple(int(x) for x in torch.unravel_index(i, d.shape)) for i in idx], d.flatten()[idx])
        print("per-out-channel max diff top", d.amax(dim=(1, 2, 3)).topk(5))
        print("per-in-channel max diff top", d.amax(dim=(0, 2, 3)).topk(5))
