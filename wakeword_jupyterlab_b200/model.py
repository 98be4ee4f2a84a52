"""``WakewordModel`` -- same module tree, parameter names and ``forward`` contract as the reference
(/root/reference/wakeword_training_script.py:141-184), so reference checkpoints load with
``load_state_dict``; ``forward`` runs the hand-written sm_100a kernels through the C ABI."""
from __future__ import annotations

import itertools
import os

import torch
import torch.nn as nn

from . import processor as _p
from .config import AudioConfig, ModelConfig
from .engine import get_engine


_uid = itertools.count(1)


class WakewordModel(nn.Module):
    def __init__(self, config=ModelConfig, audio_config=AudioConfig):
        super().__init__()
        self._ww_uid = next(_uid)          # identity token for the engine's weight cache
        self.config = config
        self.audio_config = audio_config
        self.mel_height = audio_config.N_MELS
        self.mel_width = int(audio_config.SAMPLE_RATE * audio_config.DURATION / audio_config.HOP_LENGTH) + 1
        # Parameter containers only (names/shapes/initialisation identical to the reference);
        # their torch forward() is never called.
        self.conv1 = nn.Conv2d(1, 32, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(64, 128, kernel_size=3, padding=1)
        self.pool = nn.AdaptiveAvgPool2d((1, 1))
        self.cnn_output_size = 128
        self.lstm = nn.LSTM(input_size=self.cnn_output_size, hidden_size=config.HIDDEN_SIZE,
                            num_layers=config.NUM_LAYERS, batch_first=True,
                            dropout=config.DROPOUT if config.NUM_LAYERS > 1 else 0)
        self.dropout = nn.Dropout(config.DROPOUT)
        self.fc = nn.Linear(config.HIDDEN_SIZE, config.NUM_CLASSES)
        self.threshold = 0.8
        self.conv_mode = None     # None -> WW_CONV_MODE env or package default

    def mark_weights_dirty(self):
        """Call after modifying parameters through ``.data`` (which does not bump tensor versions)."""
        self._ww_uid = next(_uid)

    def _conv_mode(self):
        return self.conv_mode or os.environ.get("WW_CONV_MODE", _p.DEFAULT_CONV_MODE)

    def engine(self, device=None, width=None):
        dev = device if device is not None else next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("WakewordModel (B200) runs on CUDA only: move the module with .to('cuda') "
                               "(there is no CPU fallback)")
        ac = self.audio_config
        n_samples = None
        if width is not None and width != self.mel_width:
            n_samples = (width - 1) * ac.HOP_LENGTH        # any clip length whose frame count is `width`
        eng = get_engine(ac, self.config, dev.index or 0, threshold=self.threshold,
                         conv_mode=self._conv_mode(), n_samples=n_samples)
        eng.sync_module(self)
        return eng

    def forward(self, x):
        """x [B, 1, N_MELS, W] float32 CUDA tensor -> logits [B, NUM_CLASSES]."""
        if self.training and (self.config.DROPOUT > 0):
            raise NotImplementedError(
                "train-mode forward with dropout runs inside WakewordTrainer.train_step (forward + backward + Adam in "
                "one C-ABI call, no autograd graph); call model.eval() for scoring")
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (there is no CPU fallback)")
        eng = self.engine(x.device, width=x.shape[-1])
        return eng.forward(x.float())
