"""Test scaffolding: run the reference's OWN trainer (`train.py`, class BaseTrainer) unchanged on top of the drop-in.

`train.py` imports its data pipeline, config manager, TensorBoard writer and debug plotting at module level
(train.py:1-16); none of them is on the hot path and most need packages that are absent here (zarr, albumentations,
tensorboard, ...).  This module installs minimal stand-ins for exactly those imports - a synthetic dataset with the
reference dataset's dictionary format (dataset.py: {"image": [C, D, H, W] float32, "<task>": ...}), a config object with
the attributes train.py reads, no-op writer / plotting - registers the drop-in as `builders`
(`resenc_b200.install_as_builders()`), and loads train.py from the reference tree (/root/reference in the build
container, its byte-for-byte copy under oracle/_ref/ on the GPU box).  The model, the losses
(training/losses/losses.py, unmodified) and the whole loop body are the reference's.
"""
import importlib.util
import os
import sys
import types
from pathlib import Path

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    for cand in (os.environ.get("RESENC_REFERENCE_ROOT", "/root/reference"), os.path.join(ROOT, "oracle", "_ref", "reference")):
        if cand and os.path.exists(os.path.join(cand, "train.py")) and os.path.isdir(os.path.join(cand, "training")):
            return cand
    return None


class SyntheticSegmentationDataset(torch.utils.data.Dataset):
    """Stands in for dataloading.dataset.ZarrSegmentationDataset3D (dataset.py:128-131 scales images to [0, 1])."""

    def __init__(self, mgr):
        self.mgr = mgr
        self.n = int(getattr(mgr, "synthetic_dataset_len", 8))

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(1000 + i)
        p = list(self.mgr.train_patch_size)
        item = {"image": torch.rand(self.mgr.in_channels, *p, generator=g)}
        for t, info in self.mgr.tasks.items():
            if t == "normals":
                item[t] = torch.nn.functional.normalize(torch.randn(3, *p, generator=g), dim=0)
            else:
                item[t] = (torch.rand(info["channels"], *p, generator=g) > 0.8).float()
        return item


class _Writer:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, tag, value, step):
        self.scalars.append((tag, float(value), int(step)))


def make_config(tmp_dir, patch=(16, 16, 16), tasks=None, **over):
    tasks = tasks or {"sheet": {"channels": 1, "activation": "sigmoid", "loss_fn": "BCEDiceLoss",
                                "loss_kwargs": {"alpha": 0.5, "beta": 0.5}},
                      "normals": {"channels": 3, "activation": "none", "loss_fn": "MaskedCosineLoss"}}
    cfg = dict(tasks=tasks, train_patch_size=list(patch), train_batch_size=2, in_channels=1, vram_max=16.0,
               autoconfigure=True, model_config={}, model_name="Model", optimizer="AdamW", initial_lr=1e-3,
               weight_decay=1e-4, max_epoch=1, max_steps_per_epoch=3, max_val_steps_per_epoch=1, gradient_accumulation=1,
               tr_val_split=0.75, train_num_dataloader_workers=0, checkpoint_path=None, load_weights_only=False,
               ckpt_out_base=Path(tmp_dir) / "ckpt", tensorboard_log_dir=str(Path(tmp_dir) / "tb"), synthetic_dataset_len=8)
    cfg.update(over)
    return types.SimpleNamespace(**cfg)


def load_reference_trainer(config):
    """Import the reference's train.py with the stand-ins in place; returns (module, BaseTrainer instance)."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("no reference tree (neither /root/reference nor oracle/_ref/reference)")
    import resenc_b200
    resenc_b200.install_as_builders()

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod("dataloading")
    mod("dataloading.dataset", ZarrSegmentationDataset3D=SyntheticSegmentationDataset)
    mod("training")
    mod("training.visualization")
    mod("training.visualization.plotting", save_debug_gif=lambda **k: None, export_data_dict_as_tif=lambda **k: None)
    mod("training.losses")
    spec = importlib.util.spec_from_file_location("training.losses.losses", os.path.join(root, "training", "losses", "losses.py"))
    losses = importlib.util.module_from_spec(spec)
    sys.modules["training.losses.losses"] = losses
    spec.loader.exec_module(losses)
    mod("configuration")
    mod("configuration.config_manager", ConfigManager=lambda path: config)
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:
        mod("torch.utils.tensorboard", SummaryWriter=_Writer)
    spec = importlib.util.spec_from_file_location("reference_train", os.path.join(root, "train.py"))
    train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(train)
    train.SummaryWriter = _Writer            # no event files from a test
    return train, train.BaseTrainer("unused.yaml")


def uninstall():
    for name in list(sys.modules):
        if name == "builders" or name.startswith("builders.") or name.split(".")[0] in ("dataloading", "configuration") \
                or name == "training" or name.startswith("training.") or name == "reference_train":
            sys.modules.pop(name, None)
