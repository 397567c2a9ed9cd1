"""One tiny train step of MidFusionFast on the GPU, checked against the oracle (called by __graft_entry__.smoke).
Lives at the repo root, outside the product package: nothing under multimodal_lipread_b200/ may import oracle/."""
import torch


def run(dev):
    from multimodal_lipread_b200 import synthetic
    from multimodal_lipread_b200.audio_video_models import MidFusionFast
    from oracle.av_models import MidFusionFastOracle          # checker only (smoke() is allowed to use the oracle)
    from oracle.frontend import AudioProcessorPort, lips_u8_to_model_input
    B, C, size, T = 2, 40, 44, 8
    wav = synthetic.make_waveforms(B, seed=11)
    lips = synthetic.make_lips_u8(B, size=size, seed=12)[:, :T].contiguous()
    labels = synthetic.make_labels(B, C, seed=13)
    torch.manual_seed(0)
    ref = MidFusionFastOracle(C).train()
    mel = AudioProcessorPort().batch_frontend_loop(wav)
    logits_ref = ref(mel, lips_u8_to_model_input(lips))
    loss_ref = torch.nn.functional.cross_entropy(logits_ref, labels)
    for precision, tol in (("fp32", 1e-4), ("tf32", 5e-3)):
        torch.manual_seed(0)
        model = MidFusionFast(C, precision=precision).to(dev).train()
        model.configure_optimizer(lr=3e-4)
        loss, logits = model.train_step(wav.to(dev), lips.to(dev), labels.to(dev), use_graph=False)   # raw inputs: log-mel on the GPU
        err = (logits.cpu() - logits_ref).abs().max().item() / logits_ref.abs().max().item()
        assert err <= tol, f"{precision}: logits deviate from the oracle by {err:.3e}"
        assert abs(loss.item() - loss_ref.item()) <= tol * max(1.0, abs(loss_ref.item()))
        print(f"smoke: MidFusionFast train step [{precision}] logits rel err {err:.2e}, loss {loss.item():.5f} (oracle {loss_ref.item():.5f}), "
              f"{model.launches_per_step()} kernels")
