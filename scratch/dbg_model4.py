import sys, os, torch, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_models_gpu as T
from oracle.frontend import lips_u8_to_model_input
name = "video_resnet_lstm"
for B, TT, size in [(2, 3, 88), (3, 3, 88), (2, 3, 64), (2, 6, 88)]:
    ref0, ours, C = T._case(name)
    wav, mel, lips, labels = T._data(B, size, TT, C)
    video = lips_u8_to_model_input(lips)
    gen = torch.Generator().manual_seed(5)
    for i in range(4):
        v = video if i == 0 else video * (1 + 2e-7 * torch.randn(video.shape, generator=gen))
        ref = copy.deepcopy(ref0).train()
        torch.nn.functional.cross_entropy(ref(v), labels).backward()
        ours.train(); ours.configure_optimizer(lr=0.0)
        loss, logits = ours.train_step(v.cuda(), labels.cuda(), use_graph=False)
        flat = ours._flat
        rows = sorted(((T._grad_err(flat.g(p), q.grad, 3e-3), n) for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters())), reverse=True)
        print(f"B{B} T{TT} s{size} pert{i}: worst {rows[0][0]:.2e} {rows[0][1]}; n>3e-3: {sum(1 for r in rows if r[0] > 3e-3)}")
