"""The reference's YAML-driven train flow on lipread_b200 models.

Mirrors the model selection and the epoch loops of the reference's four train scripts -- same `model.name` /
`train.model_name` strings, same `ValueError` for an unknown name, same loss / accuracy bookkeeping -- with the loop
body replaced by the fused `model.train_step`:

  audio_video/train.py:57-75 (train_epoch), :78-90 (validate), :112-127 (model selection)
  video/train.py:85-114, :189-204        audio/train.py:59-84, :118-134        audio_cues_video/train.py:52-81, :144-155

Model names without a launch plan yet raise NotImplementedError (they are listed in DESIGN.md section 7), so a config
that selects one fails loudly instead of silently running something else.  Datasets / DataLoaders stay the caller's
(the reference's own classes work unchanged: their items are (mel, lips, label) tuples etc.)."""
import os

import torch
import yaml

from .model_base import Cfg


class Config(Cfg):
    """YAML config with dotted `get` (reference `config/config.py:9-61`): FileNotFoundError for a missing file."""

    def __init__(self, config_path):
        if not os.path.exists(config_path):
            raise FileNotFoundError(f"Config file not found: {config_path}")
        with open(config_path, "r") as f:
            super().__init__(yaml.safe_load(f) or {})
        self.config_path = config_path

    def get_all(self):
        return self.values


AV_MODELS = ("early_fusion_resnet", "early_fusion_mobilenet", "late_fusion_mobilenet", "middle_fusion_mobilenet",
             "early_fusion_fast", "late_fusion_fast", "middle_fusion_fast")                       # av_config.yaml:10
VIDEO_MODELS = ("vgg_lstm", "resnet_lstm", "shufflenet_lstm", "mobilenet_lstm", "resnet_attn", "cnn", "resnet_trans")
AUDIO_MODELS = ("resnet", "resnet_lstm", "vgg", "vgg_lstm", "lstm_resnet", "lstm_resnet_attn", "lstm_resnet_trans")
ACV_MODELS = ("early_fusion_mobile", "middle_fusion_mobile", "late_fusion_mobile", "early_fusion_resnet",
              "middle_fusion_resnet", "late_fusion_resnet")


def _no_plan(name):
    raise NotImplementedError(f"model {name!r} of the reference has no lipread_b200 launch plan yet (DESIGN.md section 7)")


def create_av_model(model_name, num_classes, config):
    """audio_video/train.py:112-127."""
    from . import audio_video_models as M
    if model_name == "early_fusion_resnet":
        return M.create_early_fusion_resnet_model(num_classes, config)
    if model_name == "early_fusion_mobilenet":
        return M.create_early_fusion_mobilenet_model(num_classes, config)
    if model_name == "middle_fusion_fast":
        return M.create_mid_fusion_fast(num_classes, config)
    if model_name == "late_fusion_mobilenet":
        return M.create_late_fusion_mobilenet_model(num_classes, config)
    if model_name == "middle_fusion_mobilenet":
        return M.create_mid_fusion_mobilenet_model(num_classes, config)
    if model_name == "early_fusion_fast":
        return M.create_early_fusion_fast(num_classes, config)
    if model_name == "late_fusion_fast":
        return M.create_late_fusion_fast(num_classes, config)
    raise ValueError(f"Unknown model name: {model_name}")


def create_video_model(model_name, num_classes, config):
    """video/train.py:189-204."""
    from . import video_models as M
    if model_name == "resnet_lstm":
        return M.ResNet2DBiLSTM(num_classes=num_classes, config=config)
    if model_name == "mobilenet_lstm":
        return M.MobileNetLSTM(num_classes=num_classes, config=config)
    if model_name == "vgg_lstm":
        return M.VGGLSTM(num_classes=num_classes, config=config)
    if model_name == "cnn":
        return M.CNNOnly(num_classes=num_classes, config=config)
    if model_name in VIDEO_MODELS:
        _no_plan(model_name)
    raise ValueError(f"Unknown model: {model_name}")


def create_audio_model(model_name, num_classes, input_size=117, version=None):
    """audio/train.py:118-134 (get_model)."""
    from . import audio_models as M
    if model_name == "resnet":
        return M.AudioResNet(num_classes=num_classes)
    if model_name == "resnet_lstm":
        return M.AudioResNetLSTM(num_classes=num_classes)
    if model_name == "vgg":
        return M.VGGAudioClassifier(num_classes=num_classes, version=version or 11)
    if model_name == "vgg_lstm":
        return M.VGGWithLSTMClassifier(num_classes=num_classes, version=version or 11)
    if model_name == "lstm_resnet":
        return M.LSTMResNet(num_classes=num_classes, input_size=input_size)
    if model_name in AUDIO_MODELS:
        _no_plan(model_name)
    raise ValueError(f"Invalid model name: {model_name}")


def create_acv_model(model_name, num_classes, cue_dim=768, video_cfg=None):
    """audio_cues_video/train.py:144-155."""
    from . import audio_cues_video_models as M
    if model_name == "late_fusion_mobile":
        return M.MultimodalAttentionLate(num_classes, cue_dim=cue_dim, video_cfg=video_cfg, pretrained=False)
    if model_name == "late_fusion_resnet":
        return M.MultimodalAttentionLateResNet(num_classes, cue_dim=cue_dim, video_cfg=video_cfg, pretrained=False)
    if model_name in ACV_MODELS:
        _no_plan(model_name)
    raise ValueError(f"Unknown model name: {model_name}")


def _to_dev(t, device):
    return t.to(device, non_blocking=True)


def train_epoch(model, loader, device, batch_to_inputs=None):
    """One epoch with the fused step.  Returns (mean of the per-batch mean losses, accuracy %) exactly as
    audio_video/train.py:57-75 reports them (sum(batch_mean_loss) / len(loader)).  `batch_to_inputs(batch)` maps a
    DataLoader batch to (inputs tuple, labels); the default handles the reference's tuple / dict items."""
    model.train()
    loss_sum = torch.zeros((), device=device)
    correct = torch.zeros((), dtype=torch.int64, device=device)
    total, n_batches = 0, 0
    for batch in loader:
        inputs, labels = (batch_to_inputs or default_batch_to_inputs)(batch)
        inputs = tuple(_to_dev(t, device) for t in inputs)
        labels = _to_dev(labels, device)
        loss, logits = model.train_step(*inputs, labels)
        loss_sum += loss.reshape(())                         # device-side accumulation: no host sync per step
        correct += (logits.argmax(1) == labels).sum()
        total += labels.numel()
        n_batches += 1
    return (loss_sum / max(n_batches, 1)).item(), 100.0 * correct.item() / max(total, 1)


@torch.no_grad()
def validate(model, loader, device, batch_to_inputs=None):
    """audio_video/train.py:78-90: eval mode, mean CE per batch, accuracy %."""
    model.eval()
    loss_sum, correct, total, n_batches = 0.0, 0, 0, 0
    for batch in loader:
        inputs, labels = (batch_to_inputs or default_batch_to_inputs)(batch)
        inputs = tuple(_to_dev(t, device) for t in inputs)
        labels = _to_dev(labels, device)
        out = model(*inputs)
        loss_sum += torch.nn.functional.cross_entropy(out, labels).item()
        correct += (out.argmax(1) == labels).sum().item()
        total += labels.numel()
        n_batches += 1
    return loss_sum / max(n_batches, 1), 100.0 * correct / max(total, 1)


def default_batch_to_inputs(batch):
    """(mel, lips, label) [audio_video/data_utils/dataset_av.py:77], {"lip_regions", "label"}
    [video/data_utils/dataset_loader.py:98-101], (mel, label) [audio/data_utils/dataset.py:52],
    (mel, cue, lips, label) [audio_cues_video/data_utils/dataset.py:273]."""
    if isinstance(batch, dict):
        return (batch["lip_regions"],), batch["label"]
    *inputs, labels = batch
    return tuple(inputs), labels
