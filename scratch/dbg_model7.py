import sys, os, torch, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_models_gpu as T
name = "acv_late_fusion_mobile"
B, TT, size = 3, 6, 44
ref, ours, C = T._case(name)
wav, mel, lips, labels = T._data(B, size, TT, C)
ref_in, our_in = T._inputs_for(name, mel, lips)
ref.train(); ours.train()
feats = ref.video.cnn[0]
acts, grads = {}, {}
for i, blk in enumerate(feats):
    blk.register_forward_hook(lambda m, inp, out, i=i: acts.__setitem__(i, out.detach()))
    blk.register_full_backward_hook(lambda m, gi, go, i=i: grads.__setitem__(i, go[0].detach()))
logits_ref = ref(*ref_in)
torch.nn.functional.cross_entropy(logits_ref, labels).backward()
ours.configure_optimizer(lr=0.0)
loss, logits = ours.train_step(*our_in, labels.cuda(), use_graph=False)
plan = next(iter(ours._plans.values()))
for i, t in enumerate(plan.trace):
    v = t.val.view(t.F, t.H, t.W, t.C).permute(0, 3, 1, 2).cpu()
    g = t.grad.view(t.F, t.H, t.W, t.C).permute(0, 3, 1, 2).cpu()
    ev = (v - acts[i]).abs().max().item() / acts[i].abs().max().item()
    eg = (g - grads[i]).abs().max().item() / grads[i].abs().max().item()
    print(f"features[{i}] out {tuple(v.shape)} val err {ev:.2e} grad err {eg:.2e}")
