"""The callers either side of the hot path.

``WakewordDataset`` mirrors the reference class (/root/reference/wakeword_training_script.py:187-216): same constructor,
same ``(FloatTensor[1, N_MELS, W], LongTensor[1])`` item, same ``zeros(N_MELS, 31)`` stand-in for a file that fails to
load, so ``torch.utils.data.DataLoader(WakewordDataset(...))`` keeps working as the reference's ``main()`` wires it
(:452-463).  Every item is one ``process_audio_file`` call = several small GPU launches; it is the drop-in, not the fast
path.

``DeviceFeatureLoader`` is the fast path: the files are decoded ONCE into 16-bit PCM, and every batch goes
host int16 PCM -> ``ww_augment_pcm16`` (peak normalise + the host-drawn augmentation of that batch) -> ``ww_logmel``
without leaving the device.  It yields ``(features [B, 1, N_MELS, W] CUDA float32, labels [B, 1] CUDA int64)`` batches,
i.e. what ``WakewordTrainer.train_epoch`` / ``validate`` consume, in place of the reference's
``DataLoader(..., num_workers=2)`` + per-item CPU librosa pipeline (the measured 453 clips/s bottleneck)."""
from __future__ import annotations

import random
import wave

import numpy as np
import torch

from . import _lib
from .config import AugmentationConfig
from .engine import AugBatch


class WakewordDataset(torch.utils.data.Dataset):
    def __init__(self, wakeword_files, negative_files, processor, augment=False):
        self.wakeword_files = wakeword_files
        self.negative_files = negative_files
        self.processor = processor
        self.augment = augment
        self.files = list(wakeword_files) + list(negative_files)
        self.labels = [1] * len(wakeword_files) + [0] * len(negative_files)
        print(f"Dataset created with {len(self.files)} samples")
        print(f"Wakeword samples: {len(wakeword_files)}")
        print(f"Negative samples: {len(negative_files)}")

    def __len__(self):
        return len(self.files)

    def __getitem__(self, idx):
        mel_spec = self.processor.process_audio_file(self.files[idx], augment=self.augment)
        if mel_spec is None:
            mel_spec = np.zeros((self.processor.config.N_MELS, 31))            # the reference's stand-in width (:211)
        return torch.FloatTensor(np.asarray(mel_spec)).unsqueeze(0), torch.LongTensor([self.labels[idx]])


def _read_pcm16(path, sample_rate):
    """16-bit mono PCM of a WAV file as int16 (what is on disk), or None with the reference's printed error."""
    try:
        with wave.open(path, "rb") as w:
            if w.getframerate() != sample_rate:
                raise ValueError(f"sample rate {w.getframerate()} != {sample_rate}")
            if w.getsampwidth() != 2:
                raise ValueError(f"sample width {w.getsampwidth()} (16-bit PCM expected)")
            a = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
            if w.getnchannels() > 1:
                a = a.reshape(-1, w.getnchannels()).mean(axis=1).round().astype(np.int16)
            return np.ascontiguousarray(a, dtype=np.int16)
    except Exception as e:       # noqa: BLE001 -- the reference's convention: print and carry on (:65-71)
        print(f"Error loading {path}: {e}")
        return None


class DeviceFeatureLoader:
    """Batched, GPU-fed replacement of ``DataLoader(WakewordDataset(...))``.

    * decode: every WAV once, kept as int16 PCM in host memory (2 bytes per sample);
    * per batch: crop offsets (``random.randint``, like ``pad_or_truncate`` :78-83) and augmentation parameters
      (``AudioProcessor.draw_augmentation``, the reference's draw order) are drawn on the host, the PCM rows are packed
      into one pinned int16 matrix, and the device does normalise -> augment -> log-mel;
    * a file that failed to load yields an all-zero feature image, like the reference's ``zeros(N_MELS, 31)`` item.

    Peak normalisation: the reference normalises the whole file and then crops (:130-131).  For files no longer than
    ``DURATION`` the two orders agree and the device normalises the packed clip; longer files are scaled on the host by
    their whole-file peak first (fp32 rows), so the result equals the reference's in both cases."""

    def __init__(self, wakeword_files, negative_files, processor, batch_size=16, shuffle=False, augment=False,
                 aug_config=AugmentationConfig, drop_last=False):
        self.processor, self.batch_size, self.shuffle, self.augment = processor, int(batch_size), shuffle, augment
        self.aug_config, self.drop_last = aug_config, drop_last
        self.files = list(wakeword_files) + list(negative_files)
        self.labels = np.array([1] * len(wakeword_files) + [0] * len(negative_files), np.int64)
        sr = processor.config.SAMPLE_RATE
        self.N = processor.target_length
        self.pcm = [_read_pcm16(f, sr) for f in self.files]
        self.failed = np.array([p is None or len(p) == 0 for p in self.pcm])
        self.long = any(p is not None and len(p) > self.N for p in self.pcm)
        self.peaks = np.array([float(np.abs(p.astype(np.int32)).max()) / 32768.0 if p is not None and len(p) else 0.0
                               for p in self.pcm], np.float32)
        self._eng = None

    def __len__(self):
        n = len(self.files)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _engine(self):
        if self._eng is None:
            self._eng = self.processor._engine(n_samples=self.N)
        return self._eng

    def _pack(self, idx):
        """Rows of one batch: int16 PCM (device normalises) or, with over-long files, fp32 already divided by the file peak."""
        B, N = len(idx), self.N
        rows = np.zeros((B, N), np.float32 if self.long else np.int16)
        for r, i in enumerate(idx):
            p = self.pcm[i]
            if p is None or len(p) == 0:
                continue
            if len(p) > N:
                off = random.randint(0, len(p) - N)
                p = p[off:off + N]
            if self.long:
                x = p.astype(np.float32) / np.float32(32768.0)
                rows[r, :len(p)] = x / self.peaks[i] if self.peaks[i] > 0 else x
            else:
                rows[r, :len(p)] = p
        return rows

    def __iter__(self):
        order = list(range(len(self.files)))
        if self.shuffle:
            random.shuffle(order)
        eng = self._engine()
        for b0 in range(0, len(order), self.batch_size):
            idx = order[b0:b0 + self.batch_size]
            if self.drop_last and len(idx) < self.batch_size:
                break
            rows = self._pack(idx)
            B = len(idx)
            if self.augment:
                aug = self.processor.draw_augmentation(B, self.aug_config, n_samples=self.N)
            else:
                z = np.zeros(B, np.int32)
                aug = AugBatch(np.zeros(B, np.uint32), z, z + 100, z + 100, z, z, z, np.zeros(B, np.float32),
                               np.ones(B, np.float32))
            if not self.long:
                aug.flags = aug.flags | np.uint32(_lib.AUG_NORM_IN)
            clips = eng.augment(rows, aug, self.processor.noise_bank)           # int16 rows -> ww_augment_pcm16
            feats = eng.logmel(clips, normalize=False)
            bad = self.failed[idx]
            if bad.any():
                feats[torch.from_numpy(bad).to(feats.device)] = 0.0
            labels = torch.from_numpy(self.labels[idx]).to(feats.device).reshape(-1, 1)
            yield feats, labels
