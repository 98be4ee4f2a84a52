"""``AudioProcessor`` -- same class, methods and error conventions as the reference
(/root/reference/wakeword_training_script.py:61-138); the arithmetic runs in
libwakeword_b200.so on the GPU.  Batched device-side variants are added next to each method."""
from __future__ import annotations

import os
import random
import wave

import numpy as np
import torch

from . import _lib
from .config import AudioConfig, AugmentationConfig, ModelConfig
from .engine import AugBatch, get_engine


def _resample_plan_out_len(orig, new, n):
    import math
    g = math.gcd(orig, new)
    o, nn = orig // g, new // g
    return -(-nn * n // o)


class AudioProcessor:
    def __init__(self, config=AudioConfig, device=0, noise_bank=None, stage_set="north_star"):
        self.config = config
        self.device = device
        self.noise_bank = noise_bank          # optional [M, L] float32 array/tensor for the SNR-mix stage
        # "north_star": {shift, polyphase speed, SNR noise mix} (BASELINE.json); "reference": the reference's own four
        # stages {shift, phase-vocoder pitch shift, phase-vocoder time stretch, Gaussian noise} (:103-123)
        assert stage_set in ("north_star", "reference")
        self.stage_set = stage_set
        self.target_length = int(config.SAMPLE_RATE * config.DURATION)

    # -- engines are created lazily so that constructing the processor never touches the GPU
    def _engine(self, n_samples=None):
        return get_engine(self.config, ModelConfig, self.device, n_samples=n_samples,
                          conv_mode=os.environ.get("WW_CONV_MODE", DEFAULT_CONV_MODE))

    # ------------------------------------------------------------------ reference API
    def load_audio(self, file_path):
        """WAV (PCM16/PCM32/float32) loader; other codecs and sample-rate conversion are out of scope
        (reference: librosa.load, :65-71).  Same error convention: print and return None."""
        try:
            with wave.open(file_path, "rb") as w:
                sr, ch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
                raw = w.readframes(n)
            if sw == 2:
                a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
            elif sw == 4:
                a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
            else:
                raise ValueError(f"unsupported sample width {sw}")
            if ch > 1:
                a = a.reshape(-1, ch).mean(axis=1)
            if sr != self.config.SAMPLE_RATE:
                raise ValueError(f"sample rate {sr} != {self.config.SAMPLE_RATE} (resampling on load is out of scope)")
            return a
        except Exception as e:
            print(f"Error loading {file_path}: {e}")
            return None

    def normalize_audio(self, audio):
        if len(audio) == 0:
            return audio
        return self._engine().normalize(np.asarray(audio, dtype=np.float32)).cpu().numpy()

    def pad_or_truncate(self, audio, target_length):
        if len(audio) > target_length:
            start_idx = random.randint(0, len(audio) - target_length)
            return audio[start_idx:start_idx + target_length]
        return np.pad(audio, (0, target_length - len(audio)), mode="constant")

    def audio_to_mel(self, audio):
        """float[n] -> float32 ndarray [N_MELS, 1 + n // HOP_LENGTH] (dB, per-clip max = 0)."""
        if len(audio) == 0:
            return np.zeros((self.config.N_MELS,
                             int(self.config.SAMPLE_RATE * self.config.DURATION / self.config.HOP_LENGTH) + 1))
        a = np.asarray(audio, dtype=np.float32)[None, :]
        eng = self._engine(n_samples=a.shape[1])
        return eng.logmel(a)[0, 0].cpu().numpy()

    def draw_augmentation(self, n, config=AugmentationConfig, n_samples=None):
        """Host-side parameter draws in the reference's stage order (augment_audio :103-123):
        Bernoulli(p) shift, [pitch: not in the north-star stage set], Bernoulli(p) speed (+ crop offset),
        Bernoulli(p) noise.  Uses the global ``random`` module like the reference."""
        N = self.target_length if n_samples is None else n_samples
        f = np.zeros(n, np.uint32); shift = np.zeros(n, np.int32)
        ro = np.full(n, 100, np.int32); rn = np.full(n, 100, np.int32); crop = np.zeros(n, np.int32)
        ni = np.zeros(n, np.int32); no = np.zeros(n, np.int32); snr = np.zeros(n, np.float32)
        bank = self.noise_bank
        for b in range(n):
            if random.random() < config.AUGMENTATION_PROB:
                f[b] |= _lib.AUG_SHIFT
                shift[b] = int(random.uniform(-config.TIME_SHIFT_MAX, config.TIME_SHIFT_MAX) * self.config.SAMPLE_RATE)
            if random.random() < config.AUGMENTATION_PROB:
                s = int(round(100 * random.uniform(config.SPEED_CHANGE_MIN, config.SPEED_CHANGE_MAX)))
                if s != 100:
                    f[b] |= _lib.AUG_SPEED
                    ro[b] = s
                    out_len = _resample_plan_out_len(s, 100, N)
                    if out_len > N:
                        crop[b] = random.randint(0, out_len - N)
            if random.random() < config.AUGMENTATION_PROB and bank is not None:
                f[b] |= _lib.AUG_NOISE
                ni[b] = random.randrange(bank.shape[0])
                no[b] = random.randint(0, bank.shape[1] - N)
                snr[b] = random.choice(getattr(config, "SNR_GRID_DB", (0.0, 10.0, 20.0, 30.0, 40.0)))
        return AugBatch(f, shift, ro, rn, crop, ni, no, snr, np.ones(n, np.float32))

    def augment_audio_reference(self, audio, config=AugmentationConfig):
        """The reference's stage set and draw order (:103-123): Bernoulli(p) np.roll; Bernoulli(p) pitch shift by
        U(-PITCH_SHIFT_MAX, +) semitones; Bernoulli(p) time stretch by U(SPEED_CHANGE_MIN, MAX) + pad_or_truncate (its crop
        offset drawn with random.randint like the reference); Bernoulli(p) Gaussian noise of std NOISE_FACTOR (values from a
        device Philox stream seeded from np.random, the stream the reference draws its noise from)."""
        a = np.asarray(audio, dtype=np.float32)
        n = len(a)
        eng = self._engine(n_samples=n)
        x = torch.from_numpy(np.ascontiguousarray(a)).to(eng.device)[None, :]
        if random.random() < config.AUGMENTATION_PROB:
            shift = int(random.uniform(-config.TIME_SHIFT_MAX, config.TIME_SHIFT_MAX) * self.config.SAMPLE_RATE)
            x = torch.roll(x, shift, dims=1)
        if random.random() < config.AUGMENTATION_PROB:
            n_steps = random.uniform(-config.PITCH_SHIFT_MAX, config.PITCH_SHIFT_MAX)
            x = eng.pitch_shift(x.contiguous(), [n_steps])
        if random.random() < config.AUGMENTATION_PROB:
            rate = random.uniform(config.SPEED_CHANGE_MIN, config.SPEED_CHANGE_MAX)
            L = int(round(n / rate))
            crop = random.randint(0, L - n) if L > n else 0
            x = eng.time_stretch(x.contiguous(), [rate], [crop])
        if random.random() < config.AUGMENTATION_PROB:
            x = eng.add_gaussian_noise(x.contiguous().clone(), config.NOISE_FACTOR, int(np.random.randint(0, 2 ** 31 - 1)))
        return x[0].cpu().numpy()

    def augment_audio(self, audio, config=AugmentationConfig):
        if self.stage_set == "reference":
            return self.augment_audio_reference(audio, config)
        a = np.asarray(audio, dtype=np.float32)[None, :]
        params = self.draw_augmentation(1, config, n_samples=a.shape[1])
        eng = self._engine(n_samples=a.shape[1])
        return eng.augment(a, params, self.noise_bank)[0].cpu().numpy()

    def process_audio_file(self, file_path, augment=False):
        audio = self.load_audio(file_path)
        if audio is None:
            return None
        audio = self.normalize_audio(audio)
        audio = self.pad_or_truncate(audio, self.target_length)
        if augment:
            audio = self.augment_audio(audio)
        return self.audio_to_mel(audio)

    # ------------------------------------------------------------------ batched device-side variants
    def audio_to_mel_batch(self, clips, normalize=False):
        """[B, n_samples] (tensor or ndarray) -> CUDA tensor [B, 1, N_MELS, W]."""
        return self._engine(n_samples=clips.shape[1]).logmel(clips, normalize=normalize)

    def augment_batch(self, clips, params: AugBatch):
        return self._engine(n_samples=clips.shape[1]).augment(clips, params, self.noise_bank)


DEFAULT_CONV_MODE = "split2"
