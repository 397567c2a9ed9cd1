import sys, os, torch, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import av_models as O
from oracle.frontend import AudioProcessorPort, lips_u8_to_model_input
from multimodal_lipread_b200 import synthetic
B, T = 3, 7
wav = synthetic.make_waveforms(B, pad_fraction=0.5)
lips = synthetic.make_lips_u8(B, size=44)[:, :T].contiguous()
mel, video, labels = AudioProcessorPort().batch_frontend_loop(wav), lips_u8_to_model_input(lips), synthetic.make_labels(B, 40)
torch.manual_seed(0)
base = O.EarlyFusionMobileNetOracle(40, lstm_dropout=0.0, head_dropout=0.0).train()
def grads(v):
    m = copy.deepcopy(base)
    torch.nn.functional.cross_entropy(m(mel, v), labels).backward()
    return {n: p.grad.clone() for n, p in m.named_parameters()}
g0 = grads(video)
gen = torch.Generator().manual_seed(5)
for i in range(8):
    v = video * (1 + 2e-7 * torch.randn(video.shape, generator=gen))
    g = grads(v)
    worst = max(((g[n] - g0[n]).abs().max().item() / (g0[n].abs().max().item() + 3e-5), n) for n in g0 if "video_encoder.cnn" in n)
    print(i, f"{worst[0]:.3e}", worst[1])
