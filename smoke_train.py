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
    # bf16: bf16 activation storage + tcgen05 kind::f16.  Its tolerance is MEASURED on this very batch, as in
    # tests/test_midfusion_gpu.py: the deviation of the reference's own model under torch.autocast(bfloat16) from its
    # fp32 run (3.8e-2 norm-wise at this 2-clip / 8-frame shape, whose BatchNorm statistics rest on 16 frames), x 1.5
    with torch.no_grad():
        torch.manual_seed(0)
        low = MidFusionFastOracle(C).train()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            logits_low = low(mel, lips_u8_to_model_input(lips)).float()
    bf16_bar = 1.5 * (logits_low - logits_ref).abs().max().item() / logits_ref.abs().max().item()
    for precision, tol in (("fp32", 1e-4), ("tf32", 5e-3), ("bf16", max(3e-2, bf16_bar))):
        torch.manual_seed(0)
        model = MidFusionFast(C, precision=precision).to(dev).train()
        model.configure_optimizer(lr=3e-4)
        loss, logits = model.train_step(wav.to(dev), lips.to(dev), labels.to(dev), use_graph=False)   # raw inputs: log-mel on the GPU
        err = (logits.cpu() - logits_ref).abs().max().item() / logits_ref.abs().max().item()
        assert err <= tol, f"{precision}: logits deviate from the oracle by {err:.3e}"
        assert abs(loss.item() - loss_ref.item()) <= tol * max(1.0, abs(loss_ref.item()))
        print(f"smoke: MidFusionFast train step [{precision}] logits rel err {err:.2e}, loss {loss.item():.5f} (oracle {loss_ref.item():.5f}), "
              f"{model.launches_per_step()} kernels")

    # the ResNet-18 video model in bf16: exercises the implicit-GEMM 3x3 convolutions (csrc/conv_igemm.cu)
    from multimodal_lipread_b200.video_models import ResNet2DBiLSTM
    from multimodal_lipread_b200.model_base import Cfg
    from oracle import av_models as O
    torch.manual_seed(0)
    ref = O.ResNet2DBiLSTMOracle(C, O.DictConfig({"model": {"dropout": 0.0}})).train()
    with torch.no_grad():
        logits_ref = ref(lips_u8_to_model_input(lips))
    torch.manual_seed(0)
    model = ResNet2DBiLSTM(C, Cfg({"model.dropout": 0.0}), precision="bf16").to(dev).train()
    model.configure_optimizer(lr=0.0)
    loss, logits = model.train_step(lips.to(dev), labels.to(dev), use_graph=False)
    err = (logits.cpu() - logits_ref).abs().max().item() / logits_ref.abs().max().item()
    assert err <= 3e-2, f"resnet bf16: logits deviate from the oracle by {err:.3e}"
    print(f"smoke: ResNet2DBiLSTM train step [bf16, implicit-GEMM convs] logits rel err {err:.2e}, {model.launches_per_step()} kernels")
