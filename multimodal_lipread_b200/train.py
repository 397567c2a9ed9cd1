"""The reference's YAML-driven train flow on lipread_b200 models.

Mirrors the model selection and the epoch loops of the reference's four train scripts -- same `model.name` /
`train.model_name` strings, same `ValueError` for an unknown name, same loss / accuracy bookkeeping -- with the loop
body replaced by the fused `model.train_step`:

  audio_video/train.py:57-75 (train_epoch), :78-90 (validate), :112-127 (model selection)
  video/train.py:85-114, :189-204        audio/train.py:59-84, :118-134        audio_cues_video/train.py:52-81, :144-155

Every model name of the four train scripts has a launch plan; an unknown name raises the reference's ValueError.  Datasets / DataLoaders stay the caller's
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


def create_av_model(model_name, num_classes, config, pretrained_state_dict=None):
    """audio_video/train.py:112-127.  Every AV model's video trunk is ImageNet-initialised in the reference
    (weights=...IMAGENET1K_V1, e.g. middle_fusion_fast.py:15): offline the checkpoint is `pretrained_state_dict` or the
    YAML key model.pretrained_weights; with neither the factory warns and keeps the seeded random init."""
    from . import audio_video_models as M
    from .model_base import pretrained_weights
    classes = {"early_fusion_resnet": M.EarlyFusionAV, "early_fusion_mobilenet": M.EarlyFusionAVMobileNet,
               "middle_fusion_fast": M.MidFusionFast, "late_fusion_mobilenet": M.LateFusionAVMobileNet,
               "middle_fusion_mobilenet": M.MidFusionAVMobileNet, "early_fusion_fast": M.EarlyFusionFast,
               "late_fusion_fast": M.LateFusionFast}
    if model_name not in classes:
        raise ValueError(f"Unknown model name: {model_name}")
    sd = pretrained_weights(config, pretrained_state_dict, f"audio_video {model_name}")
    return classes[model_name](num_classes, config, pretrained_state_dict=sd)


def create_video_model(model_name, num_classes, config, pretrained_state_dict=None):
    """video/train.py:189-204.  resnet / mobilenet / shufflenet trunks are ImageNet-initialised in the reference
    (video/models/resnet_lstm.py:80-84): see create_av_model for how the checkpoint gets here offline; vgg_lstm (VGGLite)
    and cnn have no pretrained part."""
    from . import video_models as M
    from .model_base import pretrained_weights
    if model_name == "vgg_lstm":
        return M.VGGLSTM(num_classes=num_classes, config=config)
    if model_name == "cnn":
        return M.CNNOnly(num_classes=num_classes, config=config)
    classes = {"resnet_lstm": M.ResNet2DBiLSTM, "mobilenet_lstm": M.MobileNetLSTM, "resnet_attn": M.ResNet2DAttention,
               "shufflenet_lstm": M.ShuffleNet2DBiLSTM, "resnet_trans": M.ResNet2DTransformer}
    if model_name in classes:
        sd = pretrained_weights(config, pretrained_state_dict, f"video {model_name}")
        return classes[model_name](num_classes=num_classes, config=config, pretrained_state_dict=sd)
    if model_name in VIDEO_MODELS:
        _no_plan(model_name)
    raise ValueError(f"Unknown model: {model_name}")


def create_audio_model(model_name, num_classes, input_size=117, version=None, pretrained_state_dict=None, config=None):
    """audio/train.py:118-134 (get_model).  The resnet18 / vgg trunks are ImageNet-initialised in the reference
    (audio/models/resnet_model.py:12): see create_av_model for how the checkpoint gets here offline."""
    from . import audio_models as M
    from .model_base import pretrained_weights
    if model_name in AUDIO_MODELS and model_name != "lstm_resnet_attn":
        pretrained_state_dict = pretrained_weights(config, pretrained_state_dict, f"audio {model_name}")
    kw = {"pretrained_state_dict": pretrained_state_dict}
    if model_name == "resnet":
        return M.AudioResNet(num_classes=num_classes, **kw)
    if model_name == "resnet_lstm":
        return M.AudioResNetLSTM(num_classes=num_classes, **kw)
    if model_name == "vgg":
        return M.VGGAudioClassifier(num_classes=num_classes, version=version or 11, **kw)
    if model_name == "vgg_lstm":
        return M.VGGWithLSTMClassifier(num_classes=num_classes, version=version or 11, **kw)
    if model_name == "lstm_resnet":
        return M.LSTMResNet(num_classes=num_classes, input_size=input_size, **kw)
    if model_name == "lstm_resnet_attn":
        return M.DeepAudioNetWithAttention(num_classes=num_classes, input_size=input_size)
    if model_name == "lstm_resnet_trans":
        return M.LSTMResNetWithTransformer(num_classes=num_classes, input_size=input_size, **kw)
    if model_name in AUDIO_MODELS:
        _no_plan(model_name)
    raise ValueError(f"Invalid model name: {model_name}")


def create_acv_model(model_name, num_classes, cue_dim=768, video_cfg=None, pretrained_state_dicts=None):
    """audio_cues_video/train.py:144-155.  The reference builds every ACV model with pretrained=True (ImageNet resnet18
    audio encoder, mobilenet_v2 / resnet18 video trunk; FROZEN in the early / middle-resnet variants).  Offline the
    checkpoints are pretrained_state_dicts = {"audio": ..., "video": ...} (torchvision key names); without them the
    models warn that their trunks start from random init."""
    from . import audio_cues_video_models as M
    cls = {"late_fusion_mobile": M.MultimodalAttentionLate, "late_fusion_resnet": M.MultimodalAttentionLateResNet,
           "early_fusion_mobile": M.MultimodalAttentionEarly, "middle_fusion_mobile": M.MultimodalAttentionMiddle,
           "early_fusion_resnet": M.MultimodalAttentionEarlyResNet,
           "middle_fusion_resnet": M.MultimodalAttentionMiddleResNet}.get(model_name)
    if cls is not None:
        return cls(num_classes, cue_dim=cue_dim, video_cfg=video_cfg, pretrained=True,
                   pretrained_state_dicts=pretrained_state_dicts)
    if model_name in ACV_MODELS:
        _no_plan(model_name)
    raise ValueError(f"Unknown model name: {model_name}")


def _to_dev(t, device):
    return t.to(device, non_blocking=True)


def train_epoch(model, loader, device, batch_to_inputs=None, per_sample_loss=False, grad_allreduce=None, world=1):
    """One epoch with the fused step.  Returns (mean of the per-batch mean losses, accuracy %) exactly as
    audio_video/train.py:57-75, video/train.py:85-114 and audio/train.py:59-84 report them
    (sum(batch_mean_loss) / len(loader)); per_sample_loss=True weights every batch by its size instead, as
    audio_cues_video/train.py:52-81 does (sum(loss * n) / total).  `batch_to_inputs(batch)` maps a DataLoader batch
    to (inputs tuple, labels); the default handles the reference's tuple / dict items.  Data parallel: hand in
    dp.GradAllReduce() and the world size (the loader then shards by rank, data.DeviceBatchLoader(rank=, world=));
    loss / accuracy are this rank's."""
    model.train()
    loss_sum = torch.zeros((), device=device)
    correct = torch.zeros((), dtype=torch.int64, device=device)
    total, n_batches = 0, 0
    for batch in loader:
        inputs, labels = (batch_to_inputs or default_batch_to_inputs)(batch)
        inputs = tuple(_to_dev(t, device) for t in inputs)
        labels = _to_dev(labels, device)
        loss, logits = (model.train_step(*inputs, labels) if grad_allreduce is None else
                        model.train_step(*inputs, labels, grad_allreduce=grad_allreduce, world=world))
        # device-side accumulation: no host sync per step
        loss_sum += loss.reshape(()) * (labels.numel() if per_sample_loss else 1)
        correct += (logits.argmax(1) == labels).sum()
        total += labels.numel()
        n_batches += 1
    denom = max(total, 1) if per_sample_loss else max(n_batches, 1)
    return (loss_sum / denom).item(), 100.0 * correct.item() / max(total, 1)


@torch.no_grad()
def validate(model, loader, device, batch_to_inputs=None, per_sample_loss=False):
    """audio_video/train.py:78-90: eval mode, mean CE per batch, accuracy %  (per_sample_loss: the size-weighted mean
    of audio_cues_video/train.py:52-81)."""
    model.eval()
    loss_sum = torch.zeros((), device=device)
    correct = torch.zeros((), dtype=torch.int64, device=device)
    total, n_batches = 0, 0
    for batch in loader:
        inputs, labels = (batch_to_inputs or default_batch_to_inputs)(batch)
        inputs = tuple(_to_dev(t, device) for t in inputs)
        labels = _to_dev(labels, device)
        loss, n_ok, _ = model.eval_step(*inputs, labels)          # lr_ce_loss on the device: no host sync per batch
        loss_sum += loss.reshape(()) * (labels.numel() if per_sample_loss else 1)
        correct += n_ok.reshape(())
        total += labels.numel()
        n_batches += 1
    denom = max(total, 1) if per_sample_loss else max(n_batches, 1)
    return (loss_sum / denom).item(), 100.0 * correct.item() / max(total, 1)


def default_batch_to_inputs(batch):
    """(mel, lips, label) [audio_video/data_utils/dataset_av.py:77], {"lip_regions", "label"}
    [video/data_utils/dataset_loader.py:98-101], (mel, label) [audio/data_utils/dataset.py:52],
    (mel, cue, lips, label) [audio_cues_video/data_utils/dataset.py:273]."""
    if isinstance(batch, dict):
        return (batch["lip_regions"],), batch["label"]
    *inputs, labels = batch
    return tuple(inputs), labels


# ----------------------------------------------------------------------------------------------------------------
# Run bookkeeping around the epochs: log files, plateau schedule, checkpoints.  Formats are the reference's so that
# downstream plotting and resume keep working after the switch.
# ----------------------------------------------------------------------------------------------------------------
LOG_HEADER = ["epoch", "train_loss", "train_acc", "val_loss", "val_acc", "test_loss", "test_acc"]


def init_log_files(model_name, out_dir="./metrics"):
    """<out_dir>/<model>_training_log.csv with the reference's header row and the .txt companion
    (video/train.py:33-53, audio_video/train.py:21-33, audio_cues_video/train.py:25-35)."""
    import csv
    os.makedirs(out_dir, exist_ok=True)
    csv_path = os.path.join(out_dir, f"{model_name}_training_log.csv")
    txt_path = os.path.join(out_dir, f"{model_name}_training_log.txt")
    if not os.path.exists(csv_path):
        with open(csv_path, "w", newline="") as f:
            csv.writer(f).writerow(LOG_HEADER)
    if not os.path.exists(txt_path):
        with open(txt_path, "w") as f:
            f.write("Training Log\n\n")
    return csv_path, txt_path


def log_to_files(model_name, epoch, train_loss, train_acc, val_loss, val_acc, test_loss, test_acc, out_dir="./metrics"):
    """One CSV row + one text block per epoch (audio_video/train.py:36-48)."""
    import csv
    with open(os.path.join(out_dir, f"{model_name}_training_log.csv"), "a", newline="") as f:
        csv.writer(f).writerow([epoch, train_loss, train_acc, val_loss, val_acc, test_loss, test_acc])
    with open(os.path.join(out_dir, f"{model_name}_training_log.txt"), "a") as f:
        f.write(f"Epoch {epoch}\n"
                f"  Train Loss: {train_loss:.4f}, Train Acc: {train_acc:.2f}%\n"
                f"  Val Loss:   {val_loss:.4f}, Val Acc:   {val_acc:.2f}%\n"
                f"  Test Loss:  {test_loss:.4f}, Test Acc:  {test_acc:.2f}%\n\n")


def log_final_results(model_name, test_loss, test_acc, out_dir="./metrics"):
    """audio_video/train.py:51-53."""
    with open(os.path.join(out_dir, f"{model_name}_training_log.txt"), "a") as f:
        f.write(f"Final Test Loss: {test_loss:.4f}, Final Test Acc: {test_acc:.2f}%\n")


class ReduceLROnPlateau:
    """optim.lr_scheduler.ReduceLROnPlateau as the reference configures it -- mode "max" on val_acc, factor 0.5,
    patience 5 (video/train.py:213-215); mode "min" on val_loss (audio/train.py:156, patience 5;
    audio_cues_video/train.py:163, patience 3) -- with torch's defaults for the rest (threshold 1e-4 relative,
    cooldown 0, min_lr 0, eps 1e-8).  The new rate goes to the device-side Adam state through model.set_lr."""

    def __init__(self, model, mode="min", factor=0.5, patience=5, threshold=1e-4, min_lr=0.0, eps=1e-8):
        if mode not in ("min", "max"):
            raise ValueError(f"mode {mode} is unknown!")
        if factor >= 1.0:
            raise ValueError("Factor should be < 1.0.")
        self.model, self.mode, self.factor, self.patience = model, mode, factor, patience
        self.threshold, self.min_lr, self.eps = threshold, min_lr, eps
        self.best = float("inf") if mode == "min" else -float("inf")
        self.num_bad_epochs = 0
        self.last_epoch = 0

    def is_better(self, a):
        if self.mode == "min":
            return a < self.best * (1.0 - self.threshold)
        return a > self.best * (self.threshold + 1.0)

    def step(self, metric):
        current = float(metric)
        self.last_epoch += 1
        if self.is_better(current):
            self.best, self.num_bad_epochs = current, 0
        else:
            self.num_bad_epochs += 1
        if self.num_bad_epochs > self.patience:
            old = self.model._opt["lr"]
            new = max(old * self.factor, self.min_lr)
            if old - new > self.eps:
                self.model.set_lr(new)
            self.num_bad_epochs = 0

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "model"}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)


def make_checkpoint(model, epoch, best_val_acc):
    """video/train.py:246-251 / audio_cues_video/train.py:178-183: {"epoch": next epoch, "state_dict", "optimizer",
    "best_val_acc"}; the optimizer entry has torch.optim.Adam's state_dict layout."""
    return {"epoch": epoch + 1, "state_dict": model.state_dict(), "optimizer": model.optimizer_state_dict(),
            "best_val_acc": best_val_acc, "rng_step": model.rng_step()}          # extra key: dropout mask counter


def resume(model, path):
    """video/train.py:221-227 -> (start_epoch, best_val_acc).  Also reads audio/train.py:174-179 checkpoints
    ({"model_state_dict", "optimizer_state_dict", "val_acc", "epoch"}) and bare state_dicts
    (audio_video/train.py:150,153)."""
    ckpt = torch.load(path, map_location="cpu")
    if "state_dict" in ckpt:
        model.load_state_dict(ckpt["state_dict"])
        model.load_optimizer_state_dict(ckpt["optimizer"])
        if ckpt.get("rng_step") and next(model.parameters()).device.type == "cuda":
            model.set_rng_step(ckpt["rng_step"])                 # absent in checkpoints written by the reference
        return ckpt["epoch"], ckpt["best_val_acc"]
    if "model_state_dict" in ckpt:
        model.load_state_dict(ckpt["model_state_dict"])
        model.load_optimizer_state_dict(ckpt["optimizer_state_dict"])
        return ckpt["epoch"] + 1, ckpt["val_acc"]
    model.load_state_dict(ckpt)
    return 1, 0.0


def fit(model, model_name, loaders, device, epochs, save_dir, out_dir="./metrics", schedule=None, resume_from=None,
        batch_to_inputs=None, log=print, per_sample_loss=False, grad_allreduce=None, world=1):
    """The epoch loop of the reference's main(): train, validate, test every epoch, log, keep
    `<model>_checkpoint.pth` and `model_best.pth`, reload the best weights for the final test and write
    test_results.txt (video/train.py:232-283; audio_cues_video/train.py:166-207).
    loaders = (train, val, test); schedule = None | ("max" | "min", patience) for ReduceLROnPlateau on
    val_acc | val_loss.  Returns {"best_val_acc", "test_loss", "test_acc"}."""
    train_loader, val_loader, test_loader = loaders
    os.makedirs(save_dir, exist_ok=True)
    init_log_files(model_name, out_dir)
    sched = ReduceLROnPlateau(model, mode=schedule[0], factor=0.5, patience=schedule[1]) if schedule else None
    start_epoch, best_val_acc = (1, 0.0) if not resume_from else resume(model, resume_from)
    for epoch in range(start_epoch, epochs + 1):
        train_loss, train_acc = train_epoch(model, train_loader, device, batch_to_inputs, per_sample_loss, grad_allreduce, world)
        val_loss, val_acc = validate(model, val_loader, device, batch_to_inputs, per_sample_loss)
        if sched:
            sched.step(val_acc if sched.mode == "max" else val_loss)
        test_loss, test_acc = validate(model, test_loader, device, batch_to_inputs, per_sample_loss)
        log(f"Epoch {epoch}/{epochs}  Train Loss: {train_loss:.4f} | Train Acc: {train_acc:.2f}%  "
            f"Val Loss: {val_loss:.4f} | Val Acc: {val_acc:.2f}%  Test Loss: {test_loss:.4f} | Test Acc: {test_acc:.2f}%")
        log_to_files(model_name, epoch, train_loss, train_acc, val_loss, val_acc, test_loss, test_acc, out_dir)
        is_best = val_acc > best_val_acc or not os.path.exists(os.path.join(save_dir, "model_best.pth"))
        best_val_acc = max(best_val_acc, val_acc)
        ckpt = make_checkpoint(model, epoch, best_val_acc)
        torch.save(ckpt, os.path.join(save_dir, f"{model_name}_checkpoint.pth"))
        if is_best:
            torch.save(ckpt, os.path.join(save_dir, "model_best.pth"))
    best = torch.load(os.path.join(save_dir, "model_best.pth"), map_location="cpu")
    model.load_state_dict(best["state_dict"])
    test_loss, test_acc = validate(model, test_loader, device, batch_to_inputs, per_sample_loss)
    log_final_results(model_name, test_loss, test_acc, out_dir)
    with open(os.path.join(save_dir, "test_results.txt"), "w") as f:
        f.write(f"Final Test Loss: {test_loss:.4f}\n")
        f.write(f"Final Test Acc: {test_acc:.2f}%\n")
        f.write(f"Best Val Acc: {best_val_acc:.2f}%\n")
    return {"best_val_acc": best_val_acc, "test_loss": test_loss, "test_acc": test_acc}
