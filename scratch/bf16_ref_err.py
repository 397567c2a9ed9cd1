import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_midfusion_gpu import _inputs, _grad_err, lips_u8_to_model_input, MidFusionFastOracle, C, _rel
torch.manual_seed(0); ref = MidFusionFastOracle(C).train()
torch.manual_seed(0); lo = MidFusionFastOracle(C).train()
B, size = 4, 88
wav, mel, lips, labels = _inputs(B, size)
video = lips_u8_to_model_input(lips)
out = ref(mel, video); torch.nn.functional.cross_entropy(out, labels).backward()
with torch.autocast("cpu", dtype=torch.bfloat16):
    out2 = lo(mel, video)
    loss2 = torch.nn.functional.cross_entropy(out2.float(), labels)
loss2.backward()
print("logits rel err bf16 autocast:", _rel(out2.float(), out))
errs = sorted(((_grad_err(q.grad, p.grad), n) for (n, p), (_, q) in zip(ref.named_parameters(), lo.named_parameters())), reverse=True)
print("worst:", errs[:8]); import statistics
print("median:", statistics.median(e for e, _ in errs))
