"""Per-op device times of one MidFusionFast train step (graph-replayed, no host overhead) -> JSON."""
import sys, os, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_lipread_b200 import synthetic, engine
from multimodal_lipread_b200.audio_video_models import MidFusionFast
B, size = int(sys.argv[1]), int(sys.argv[2])
out = sys.argv[3]
torch.manual_seed(0)
m = MidFusionFast(40).cuda().train()
m.configure_optimizer(lr=3e-4)
wav = synthetic.make_waveforms(B).cuda()
lips = synthetic.make_lips_u8(B, size=size, grayscale=True).cuda()
lab = synthetic.make_labels(B, 40).cuda()
for _ in range(2):
    m.train_step(wav, lips, lab, use_graph=False)
torch.cuda.synchronize()
plan = next(p for p in m._plans.values() if p.with_backward)
rows = []
for phase, ops in (("fwd", plan.fwd), ("bwd", plan.bwd)):
    for name, args, ms in engine.profile_ops_graph(ops, reps=20):
        rows.append({"phase": phase, "op": name, "ms": ms, "bytes": engine.op_algorithmic_bytes(name, args),
                     "ints": [a for a in args if isinstance(a, int) and 0 <= a < (1 << 31)]})
json.dump(rows, open(out, "w"))
tot = sum(r["ms"] for r in rows)
by = {}
for r in rows:
    by[r["op"]] = by.get(r["op"], 0) + r["ms"]
print("sum of ops ms", tot)
for k, v in sorted(by.items(), key=lambda kv: -kv[1]):
    print(f"{k:22s} {v*1000:9.1f} us")
