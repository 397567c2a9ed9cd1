"""Import and run the UNMODIFIED reference modules (SURVEY.md 8(c) recipe).  TEST INFRASTRUCTURE ONLY.

Where the reference lives: oracle/_ref/ (staged by oracle/stage_ref.py; travels to the GPU box) or, in the build
container, /root/reference itself.  Used by bench.py's CPU arm (the reference's own train loop on the host cores) and
by tests/golden/make_golden.py.  Never imported by the product package.

The recipe: (1) `librosa` / `pydub` are imported at module top by audio_processor.py:4,6 but unused on this path ->
empty stub modules; (2) the ImageNet weight download (weights=...IMAGENET1K_V1, middle_fusion_fast.py:15) becomes a
no-op -> seeded random init; (3) the packages import their siblings by bare names (`from models.x import ...`) -> one
package at a time on sys.path, the shared top-level names purged in between.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_SHARED = ("models", "config", "configs", "utils", "data_utils", "train")
_CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference")


def ref_root():
    """Directory holding the reference packages, or None."""
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, "audio_video", "models")):
            return c
    return None


def _stub_unused_imports():
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    if "pydub" not in sys.modules:
        pd = types.ModuleType("pydub")
        pd.AudioSegment = object
        sys.modules["pydub"] = pd


_patched = False


def _offline_weights():
    global _patched
    if _patched:
        return
    import torch
    from torchvision.models import _api
    _api.WeightsEnum.get_state_dict = lambda self, *a, **k: None
    orig = torch.nn.Module.load_state_dict

    def load_state_dict(self, sd, *a, **k):
        if sd is None:
            return None
        return orig(self, sd, *a, **k)
    torch.nn.Module.load_state_dict = load_state_dict
    _patched = True


def load(pkg, module):
    """importlib.import_module(module) with reference package `pkg` (audio | video | audio_video | audio_cues_video)
    first on sys.path."""
    root = ref_root()
    if root is None:
        raise ImportError("reference not staged: run `python oracle/stage_ref.py` where /root/reference exists")
    _stub_unused_imports()
    _offline_weights()
    for name in list(sys.modules):
        if name.split(".")[0] in _SHARED:
            del sys.modules[name]
    sys.path[:] = [p for p in sys.path if not any(p.startswith(c) for c in _CANDIDATES)]
    sys.path.insert(0, os.path.join(root, pkg))
    return importlib.import_module(module)


class Cfg:
    """Any object with .get(key, default) is a valid reference config (config/config.py:41-61)."""

    def __init__(self, d=None):
        self.d = d or {}

    def get(self, key, default=None):
        return self.d.get(key, default)


class _NoBar:
    """tqdm stand-in: the reference's loops wrap their loader in tqdm(...) and call set_postfix on it."""

    def __init__(self, it, *a, **k):
        self.it = it

    def __iter__(self):
        return iter(self.it)

    def __len__(self):
        return len(self.it)

    def set_postfix(self, *a, **k):
        pass


def train_loop(workload):
    """-> (pkg, make_model(num_classes), run(model, optimizer, batch)) for a bench workload: `run` drives ONE batch
    through the reference's OWN epoch function (audio_video/train.py:57-75 train_epoch, video/train.py:85-114,
    audio/train.py:59-84, audio_cues_video/train.py:52-81 run_epoch) with a one-element list as the loader.  The train
    scripts import their dataset modules at the top; when that import chain cannot be satisfied here (hard-coded
    /home paths, cv2, ...) the loop body is restated (same statements) around the reference's model."""
    import torch
    crit = torch.nn.CrossEntropyLoss()
    spec = {
        "mid_fusion_fast": ("audio_video", "models.middle_fusion_fast", lambda m, C: m.create_mid_fusion_fast(C, Cfg())),
        "early_fusion_mobilenet": ("audio_video", "models.early_fusion", lambda m, C: m.create_early_fusion_mobilenet_model(C, Cfg())),
        "early_fusion_resnet": ("audio_video", "models.ef_cnn_lstm_resnet", lambda m, C: m.create_early_fusion_resnet_model(C, Cfg())),
        "video_resnet_lstm": ("video", "models.resnet_lstm", lambda m, C: m.ResNet2DBiLSTM(num_classes=C, config=Cfg({"model.feature_dim": 1024, "model.dropout": 0.5}))),
        "audio_resnet": ("audio", "models.resnet_model", lambda m, C: m.AudioResNet(num_classes=C)),
        "acv_late_fusion_mobile": ("audio_cues_video", "models.late_fusion_mobile", lambda m, C: m.MultimodalAttentionLate(C, pretrained=False)),
    }[workload]
    pkg, modname, ctor = spec
    mod = load(pkg, modname)

    def make_model(C):
        return ctor(mod, C)

    epoch_fn = None
    try:
        tr = load(pkg, "train")
        tr.tqdm = _NoBar
        if pkg == "audio_cues_video":
            tr.device = "cpu"
            epoch_fn = lambda model, opt, batch: tr.run_epoch(model, [batch], crit, opt, train=True)
        elif pkg == "video":
            epoch_fn = lambda model, opt, batch: tr.train_epoch(model, [batch], crit, opt, "cpu", 1, 1)
        else:
            epoch_fn = lambda model, opt, batch: tr.train_epoch(model, [batch], crit, opt, "cpu")
        how = f"{pkg}/train.py epoch function"
    except Exception as e:                                   # noqa: BLE001 -- any import-chain failure of the train script
        how = f"{pkg}/train.py loop body restated around the reference model (train.py import failed: {type(e).__name__})"

        def epoch_fn(model, opt, batch):
            model.train()
            if isinstance(batch, dict):
                inputs, labels = (batch["lip_regions"],), batch["label"]
            else:
                *inputs, labels = batch
            opt.zero_grad()
            out = model(*inputs)
            loss = crit(out, labels)
            loss.backward()
            opt.step()
            return loss.item(), 100.0 * out.max(1)[1].eq(labels).sum().item() / labels.size(0)
        # the model module must be the live one again (load() purged it when importing train failed half-way)
        mod2 = load(pkg, modname)
        make_model = lambda C: ctor(mod2, C)                 # noqa: E731
    return pkg, make_model, epoch_fn, how


def audio_processor():
    """The reference's AudioProcessor (audio_video/utils/audio_processor.py)."""
    return load("audio_video", "utils.audio_processor").AudioProcessor()
