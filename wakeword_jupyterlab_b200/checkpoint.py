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


def save_deployment_package(path, model, checkpoint=None, device="cuda", architecture_path=None):
    """The notebook's deployment package (wakeword_training.ipynb:951-977): state dict + the model / audio configuration
    values + training info + class names, same keys and nesting.  ``checkpoint`` is the dictionary of
    ``best_wakeword_model.pth`` (``val_acc`` / ``epoch`` are read with the notebook's defaults).  When
    ``architecture_path`` is given, the plain-text summary of :980-991 (``model_architecture.txt``) is written too."""
    mc, ac = model.config, model.audio_config
    checkpoint = checkpoint or {}
    pkg = {
        "model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
        "model_config": {k: getattr(mc, k) for k in ("HIDDEN_SIZE", "NUM_LAYERS", "DROPOUT", "NUM_CLASSES")},
        "audio_config": {k: getattr(ac, k) for k in ("SAMPLE_RATE", "DURATION", "N_MELS", "N_FFT", "HOP_LENGTH", "FMIN", "FMAX")},
        "training_info": {"best_val_accuracy": checkpoint.get("val_acc", 0), "epoch": checkpoint.get("epoch", 0) + 1,
                          "device": str(device)},
        "classes": list(CLASS_NAMES),
    }
    torch.save(pkg, path)
    if architecture_path:
        with open(architecture_path, "w") as f:
            f.write("Wakeword Detection Model Architecture\n")
            f.write("================================\n\n")
            f.write("Model Type: CNN + LSTM\n")
            f.write(f"Input Shape: (1, {ac.N_MELS}, 31)\n")
            f.write(f"Hidden Size: {mc.HIDDEN_SIZE}\n")
            f.write(f"Number of Layers: {mc.NUM_LAYERS}\n")
            f.write(f"Dropout: {mc.DROPOUT}\n")
            f.write(f"Number of Classes: {mc.NUM_CLASSES}\n")
            f.write(f"Parameters: {sum(p.numel() for p in model.parameters()):,}\n")
            f.write(f"Device: {device}\n")
    return pkg


def load_deployment_package(path, device="cuda"):
    """Rebuild a ``WakewordModel`` from a deployment package: configuration classes are reconstructed from the stored
    values (the notebook's consumer does the same by hand) and the weights loaded."""
    from .config import AudioConfig, ModelConfig
    from .model import WakewordModel
    pkg = torch.load(path, map_location="cpu", weights_only=False)
    MC = type("ModelConfig", (ModelConfig,), dict(pkg.get("model_config", {})))
    AC = type("AudioConfig", (AudioConfig,), dict(pkg.get("audio_config", {})))
    model = WakewordModel(MC, AC)
    model.load_state_dict(pkg["model_state_dict"])
    return model.to(device).eval(), pkg
