import sys, os, torch, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_models_gpu as T
name = "acv_late_fusion_mobile"
for B, TT, size in [(3, 6, 44), (4, 8, 44), (2, 4, 88)]:
    ref, ours, C = T._case(name)
    wav, mel, lips, labels = T._data(B, size, TT, C)
    ref_in, our_in = T._inputs_for(name, mel, lips)
    ref.train(); ours.train()
    logits_ref = ref(*ref_in)
    torch.nn.functional.cross_entropy(logits_ref, labels).backward()
    ours.configure_optimizer(lr=0.0)
    loss, logits = ours.train_step(*our_in, labels.cuda(), use_graph=False)
    flat = ours._flat
    print(f"== B{B} T{TT} s{size}: logits {T._rel(logits, logits_ref):.2e} launches {ours.launches_per_step()}")
    rows = [(T._grad_err(flat.g(p), q.grad, 3e-3), n, q.grad.abs().max().item()) for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters())]
    print("  n bad", sum(1 for r in rows if r[0] > 3e-3), "of", len(rows))
    for e, n, m in rows:
        if e > 2e-3 and (n.endswith("weight") and ("conv.0.0" in n or "conv.1.0" in n or "conv.2" in n or "lstm" in n or n.count(".") < 3)):
            print(f"   {e:.2e} {n} max|g| {m:.1e}")
