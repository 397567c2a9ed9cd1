import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_midfusion_gpu import _pair, _inputs, _grad_err, lips_u8_to_model_input
ref, ours = _pair(precision=sys.argv[1] if len(sys.argv) > 1 else "tf32")
B, size = 4, 88
wav, mel, lips, labels = _inputs(B, size)
ref.train(); ours.train()
logits_ref = ref(mel, lips_u8_to_model_input(lips))
torch.nn.functional.cross_entropy(logits_ref, labels).backward()
ours.configure_optimizer(lr=3e-4)
loss, logits = ours.train_step(mel.cuda(), lips.cuda(), labels.cuda(), use_graph=False)
flat = ours._flat
errs = sorted(((_grad_err(flat.g(p), q.grad), n, q.grad.abs().max().item()) for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters())), reverse=True)
for e, n, m in errs[:25]:
    print(f"{e:.3e} {n} max|g|={m:.3e}")
