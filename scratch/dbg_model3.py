import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_models_gpu as T
def run(name, B, TT, size, poison=None):
    if poison is not None:
        junk = [torch.full((64 << 20,), poison, device="cuda") for _ in range(8)]
        del junk
    ref, ours, C = T._case(name)
    wav, mel, lips, labels = T._data(B, size, TT, C)
    ref_in, our_in = T._inputs_for(name, mel, lips)
    ref.train(); ours.train()
    logits_ref = ref(*ref_in)
    torch.nn.functional.cross_entropy(logits_ref, labels).backward()
    ours.configure_optimizer()
    loss, logits = ours.train_step(*our_in, labels.cuda(), use_graph=False)
    flat = ours._flat
    rows = [(T._grad_err(flat.g(p), q.grad, 3e-3), n) for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()) if not (n.startswith("audio_encoder.cnn") and n.endswith("bias"))]
    print(f"== {name} B{B} T{TT} s{size} poison={poison}: logits {T._rel(logits, logits_ref):.2e}; {sum(1 for r in rows if r[0] > 3e-3)} bad; worst {max(rows)}")
    return flat.grad.clone()
g1 = run("early_fusion_mobilenet", 3, 7, 44)
g2 = run("early_fusion_mobilenet", 3, 7, 44)
print("run-to-run max diff", (g1 - g2).abs().max().item())
g3 = run("early_fusion_mobilenet", 3, 7, 44, poison=float("nan"))
g4 = run("early_fusion_mobilenet", 3, 7, 44, poison=1000.0)
print("poison diffs", (g1 - g3).abs().max().item(), (g1 - g4).abs().max().item())
g5 = run("early_fusion_mobilenet", 3, 8, 44, poison=float("nan"))
