"""Scoring entry points: the reference's ``predict_wakeword`` (wakeword_training.ipynb:871-893) and its
batched / streaming generalisations (BASELINE.json configs 3 and 4)."""
from __future__ import annotations

import numpy as np
import torch

from .engine import AugBatch


def predict_wakeword(audio_file_path, model, processor, device, threshold=0.8):
    """Predict if audio contains wakeword -> (is_wakeword: bool, wakeword_prob: float)."""
    model.eval()
    mel_spec = processor.process_audio_file(audio_file_path, augment=False)
    if mel_spec is None:
        print(f"Error processing audio file: {audio_file_path}")
        return False, 0.0
    mel_tensor = torch.as_tensor(np.asarray(mel_spec), dtype=torch.float32).unsqueeze(0).unsqueeze(0).to(device)
    with torch.no_grad():
        output = model(mel_tensor)
        probabilities = torch.softmax(output, dim=1)
        wakeword_prob = probabilities[0][1].item()
    is_wakeword = wakeword_prob >= threshold
    return is_wakeword, wakeword_prob


def score_clips(clips, model, aug: AugBatch = None, noise_bank=None, normalize=True, threshold=0.8):
    """Fused (augment) -> log-mel -> CNN+LSTM -> softmax -> threshold over a batch [B, n_samples].
    Returns device tensors (logits [B,C], prob1 [B], decision [B] uint8)."""
    model.eval()
    model.threshold = threshold
    eng = model.engine()
    return eng.score(clips, aug=aug, noise_bank=noise_bank, normalize=normalize)


def score_stream(audio, model, hop_samples=160, threshold=0.8):
    """predict_wakeword's maths on every 1 s window at ``hop_samples`` (10 ms default) of a long signal."""
    model.eval()
    model.threshold = threshold
    return model.engine().score_stream(audio, hop_samples)
