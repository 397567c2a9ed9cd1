import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_lipread_b200 import synthetic
from multimodal_lipread_b200.audio_video_models import MidFusionFast
B, T, size = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(0)
m = MidFusionFast(40).cuda().train()
m.configure_optimizer(lr=3e-4)
wav = synthetic.make_waveforms(B).cuda()
lips = synthetic.make_lips_u8(B, size=size)[:, :T].contiguous().cuda()
lab = synthetic.make_labels(B, 40).cuda()
loss, _ = m.train_step(wav, lips, lab, use_graph=False)
torch.cuda.synchronize()
print("loss", loss.item())
