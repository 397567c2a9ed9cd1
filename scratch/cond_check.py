import sys, os, torch, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import av_models as O
from oracle.frontend import AudioProcessorPort, lips_u8_to_model_input
from multimodal_lipread_b200 import synthetic
def data(B, size, T, C):
    wav = synthetic.make_waveforms(B, pad_fraction=0.5)
    lips = synthetic.make_lips_u8(B, size=size)[:, :T].contiguous()
    return AudioProcessorPort().batch_frontend_loop(wav), lips_u8_to_model_input(lips), synthetic.make_labels(B, C)
for B, T in [(3, 7), (3, 8), (4, 7)]:
    torch.manual_seed(0)
    m32 = O.EarlyFusionMobileNetOracle(40, lstm_dropout=0.0, head_dropout=0.0).train()
    m64 = copy.deepcopy(m32).double()
    mel, video, labels = data(B, 44, T, 40)
    torch.nn.functional.cross_entropy(m32(mel, video), labels).backward()
    torch.nn.functional.cross_entropy(m64(mel.double(), video.double()), labels).backward()
    rows = []
    for (n, p), (_, q) in zip(m32.named_parameters(), m64.named_parameters()):
        e = (p.grad.double() - q.grad).abs().max().item() / (q.grad.abs().max().item() + 1e-7 / 3e-3)
        rows.append((e, n))
    rows.sort(reverse=True)
    print(B, T, [(f"{e:.2e}", n) for e, n in rows[:4]])
