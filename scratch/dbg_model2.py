import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_models_gpu as T
from oracle.av_models import MidFusionFastOracle
from multimodal_lipread_b200.audio_video_models import MidFusionFast
def run(name, B, TT, size):
    if name == "mid":
        torch.manual_seed(0); ref = MidFusionFastOracle(40); torch.manual_seed(0); ours = MidFusionFast(40, precision="fp32").cuda(); C = 40
    else:
        ref, ours, C = T._case(name)
    wav, mel, lips, labels = T._data(B, size, TT, C)
    ref_in, our_in = T._inputs_for("early_fusion" if name == "mid" else name, mel, lips)
    ref.train(); ours.train()
    logits_ref = ref(*ref_in)
    torch.nn.functional.cross_entropy(logits_ref, labels).backward()
    ours.configure_optimizer()
    loss, logits = ours.train_step(*our_in, labels.cuda(), use_graph=False)
    flat = ours._flat
    rows = [(T._grad_err(flat.g(p), q.grad, 3e-3), n) for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters())]
    print(f"== {name} B{B} T{TT} s{size}: logits {T._rel(logits, logits_ref):.2e}; {sum(1 for r in rows if r[0] > 3e-3)} bad; worst {max(rows)}")
for a in [("mid", 3, 7, 44), ("mid", 4, 29, 44), ("early_fusion_mobilenet", 4, 29, 44), ("early_fusion_mobilenet", 3, 7, 44), ("early_fusion_mobilenet", 3, 8, 44), ("early_fusion_mobilenet", 4, 7, 44), ("early_fusion_mobilenet", 3, 7, 88)]:
    run(*a)
