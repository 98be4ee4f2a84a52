"""Checkpoint I/O in the reference's own file formats.

The reference writes three ``torch.save`` dictionaries, all keyed around ``model_state_dict``:
  * ``best_wakeword_model.pth``   -- wakeword_training_script.py:326-334 (epoch, optimizer state, accuracies, losses)
  * ``final_wakeword_model.pth``  -- wakeword_training_script.py:479-488 (config dicts, best_val_acc, device)
  * the deployment package        -- wakeword_training.ipynb:951-977 (state dict + configs + class names)
``WakewordModel`` here has the identical module tree, so the tensors load with ``load_state_dict``; these helpers
only unwrap / wrap the surrounding dictionaries.  Weights reach the device buffers of the C ABI at the next call
(``Engine.sync_module`` -> ``ww_set_weights``); nothing else of the checkpoint is interpreted."""
from __future__ import annotations

import torch

CLASS_NAMES = ["negative", "wakeword"]          # wakeword_training.ipynb:973


def extract_state_dict(obj):
    """Accept a bare state_dict or any of the reference's checkpoint dictionaries."""
    if isinstance(obj, dict) and "model_state_dict" in obj:
        return obj["model_state_dict"]
    return obj


def load_checkpoint(path, model, map_location="cpu"):
    """Load ``best_wakeword_model.pth`` / ``final_wakeword_model.pth`` / a deployment package / a bare state_dict into
    ``model``; returns the full dictionary (metadata untouched)."""
    obj = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(extract_state_dict(obj))
    return obj


def _cfg_dict(cls):
    return {k: v for k, v in vars(cls).items() if k.isupper()}


def save_best_checkpoint(path, model, epoch, val_acc, train_acc, train_loss, val_loss, optimizer_state=None):
    """Same keys as wakeword_training_script.py:326-334."""
    torch.save({"epoch": epoch, "model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                "optimizer_state_dict": optimizer_state or {}, "val_acc": val_acc, "train_acc": train_acc,
                "train_loss": train_loss, "val_loss": val_loss}, path)


def save_final_checkpoint(path, model, best_val_acc, device, training_config=None):
    """Same keys as wakeword_training_script.py:479-488."""
    from .config import TrainingConfig
    torch.save({"model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                "config": {"model": _cfg_dict(model.config), "audio": _cfg_dict(model.audio_config),
                           "training": _cfg_dict(training_config or TrainingConfig)},
                "best_val_acc": best_val_acc, "device": str(device)}, path)
