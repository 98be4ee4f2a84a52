"""Numpy restatement of the reference's phase-vocoder augmentations (TEST INFRASTRUCTURE).

``augment_audio`` (/root/reference/wakeword_training_script.py:110-121) calls
  * ``librosa.effects.pitch_shift(y, sr=16000, n_steps)``  (:112)
  * ``librosa.effects.time_stretch(y, rate)`` + ``pad_or_truncate`` (:116-117)
  * ``y + np.random.normal(0, NOISE_FACTOR, len(y))``          (:120-121)
The arithmetic of the first two lives in librosa 0.10.1 (pinned README.md:386, not vendored, not installed here), so
this file restates librosa's published algorithm with its defaults at those call sites:
  stft / istft: n_fft 2048, hop 512, periodic Hann, center=True with pad_mode="constant", window sum-of-squares
                normalisation in the inverse, trimmed to ``length``;
  phase_vocoder: time steps arange(0, T, rate), linear magnitude interpolation between the two neighbouring frames,
                phase advance = wrapped phase difference + expected advance pi * hop * k / (n_bins - 1), accumulated
                from the phase of frame 0;
  time_stretch = istft(phase_vocoder(stft(y), rate), length=round(len(y) / rate));
  pitch_shift  = fix_length(resample(time_stretch(y, 2^(-n_steps/12)), orig = sr / rate, target = sr), len(y)).
"parity unpinned" for this file against librosa itself; it is cross-checked against independent implementations that
ARE importable here (tests/test_oracle.py): torch.stft / torch.istft and torchaudio.functional.phase_vocoder, which
documents itself as a port of the same librosa routine.  librosa resamples with soxr_hq (not available); the oracle and
the CUDA path use the polyphase windowed-sinc resampler of oracle/augment.py with the rate rounded to 1/1000, so pitch
shift is compared oracle-to-kernel exactly and to librosa only in distribution (pitch within 0.1 %)."""
from __future__ import annotations

import math

import numpy as np

from . import augment as A

N_FFT, HOP = 2048, 512


def hann(n=N_FFT):
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)


def stft(y, n_fft=N_FFT, hop=HOP):
    """complex64 [n_fft/2 + 1, 1 + len(y) // hop] (librosa.stft, center=True, pad_mode='constant', float32 input)."""
    y = np.asarray(y, dtype=np.float32)
    ypad = np.pad(y, (n_fft // 2, n_fft // 2))
    n_frames = 1 + (len(ypad) - n_fft) // hop
    w = hann(n_fft)
    frames = np.lib.stride_tricks.sliding_window_view(ypad, n_fft)[::hop][:n_frames]
    return np.fft.rfft(frames.astype(np.float64) * w, axis=1).T.astype(np.complex64)


def istft(D, length, n_fft=N_FFT, hop=HOP):
    """float32 [length] (librosa.istft with center=True and ``length``)."""
    D = np.asarray(D)
    n_frames = D.shape[1]
    padded = length + 2 * (n_fft // 2)
    n_frames = min(n_frames, int(math.ceil(padded / hop)))
    w = hann(n_fft).astype(np.float64)
    total = n_fft + hop * (n_frames - 1)
    y = np.zeros(total, np.float64)
    wss = np.zeros(total, np.float64)
    frames = np.fft.irfft(D[:, :n_frames].astype(np.complex128), n=n_fft, axis=0)
    for t in range(n_frames):
        y[t * hop:t * hop + n_fft] += w * frames[:, t]
        wss[t * hop:t * hop + n_fft] += w * w
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    y = y[n_fft // 2:]
    if len(y) >= length:
        y = y[:length]
    else:
        y = np.pad(y, (0, length - len(y)))
    return y.astype(np.float32)


def phase_vocoder(D, rate, hop=HOP):
    D = np.asarray(D)
    n_bins, T = D.shape
    steps = np.arange(0, T, rate, dtype=np.float64)
    out = np.zeros((n_bins, len(steps)), dtype=D.dtype)
    phi_advance = np.linspace(0, np.pi * hop, n_bins)
    phase_acc = np.angle(D[:, 0]).astype(np.float64)
    Dp = np.pad(D, ((0, 0), (0, 2)))
    for t, step in enumerate(steps):
        c0, c1 = Dp[:, int(step)], Dp[:, int(step) + 1]
        alpha = np.mod(step, 1.0)
        mag = (1.0 - alpha) * np.abs(c0) + alpha * np.abs(c1)
        out[:, t] = (mag * np.exp(1j * phase_acc)).astype(D.dtype)
        dphase = np.angle(c1).astype(np.float64) - np.angle(c0).astype(np.float64) - phi_advance
        dphase = dphase - 2.0 * np.pi * np.round(dphase / (2.0 * np.pi))
        phase_acc = phase_acc + phi_advance + dphase
    return out


def stretch_len(n, rate):
    return int(round(n / rate))


def time_stretch(y, rate):
    y = np.asarray(y, dtype=np.float32)
    return istft(phase_vocoder(stft(y), rate), stretch_len(len(y), rate))


def pitch_ratio(n_steps):
    """(rate, rs_orig, rs_new): librosa's rate 2^(-n/12) and the rational stand-in for sr/rate -> sr (1/1000 grid)."""
    rate = 2.0 ** (-float(n_steps) / 12.0)
    return rate, int(round(1000.0 / rate)), 1000


def pitch_shift(y, n_steps):
    y = np.asarray(y, dtype=np.float32)
    rate, o, n = pitch_ratio(n_steps)
    z = time_stretch(y, rate)
    if o != n:                      # librosa.resample returns its input when orig_sr == target_sr
        z = A.resample(z, o, n)
    return A.pad_or_truncate(z, len(y), 0).astype(np.float32)


def stretch_and_fit(y, rate, crop_off=0):
    """time_stretch followed by the reference's pad_or_truncate (:116-117), crop offset host-drawn."""
    return A.pad_or_truncate(time_stretch(y, rate), len(y), crop_off).astype(np.float32)
