import sys, os, torch, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import av_models as O
from oracle.frontend import AudioProcessorPort, lips_u8_to_model_input
from multimodal_lipread_b200 import synthetic
B, T, size, C = 3, 6, 44, 40
wav = synthetic.make_waveforms(B, pad_fraction=0.5)
lips = synthetic.make_lips_u8(B, size=size)[:, :T].contiguous()
mel, video, labels, cue = AudioProcessorPort().batch_frontend_loop(wav), lips_u8_to_model_input(lips), synthetic.make_labels(B, C), synthetic.make_cues(B)
torch.manual_seed(0)
m32 = O.LateFusionMobileOracle(C, lstm_dropout=0.0).train()
m64 = copy.deepcopy(m32).double()
torch.nn.functional.cross_entropy(m32(mel, cue, video), labels).backward()
torch.nn.functional.cross_entropy(m64(mel.double(), cue.double(), video.double()), labels).backward()
rows = []
for (n, p), (_, q) in zip(m32.named_parameters(), m64.named_parameters()):
    e = (p.grad.double() - q.grad).abs().max().item() / (q.grad.abs().max().item() + 1e-7 / 3e-3)
    rows.append((e, n))
print("n > 3e-3:", sum(1 for r in rows if r[0] > 3e-3), "of", len(rows))
rows.sort(reverse=True)
for e, n in rows[:12]:
    print(f"{e:.2e} {n}")
