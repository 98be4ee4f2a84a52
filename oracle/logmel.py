"""Numpy restatement of the reference's log-mel front end (TEST INFRASTRUCTURE).

Follows the call site ``AudioProcessor.audio_to_mel``
(/root/reference/wakeword_training_script.py:85-101), i.e.
``librosa.feature.melspectrogram(y, sr, n_mels, n_fft, hop_length, win_length,
fmin, fmax)`` followed by ``librosa.power_to_db(S, ref=np.max)``.  librosa
(pinned 0.10.1, README.md:386) is a third-party dependency that is NOT present
under /root/reference nor installed here; its published algorithm is restated
below and cross-checked against torchaudio 2.11 in
``tests/golden/make_golden.py``.

dtype flow restated from librosa 0.10 ``stft``: window is float64, the framed
signal times window is float64, ``rfft`` runs in float64, the result is stored
into a complex64 matrix for float32 input; ``abs()**2``, the mel projection and
the dB conversion then run in float32.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------
# Slaney mel scale (librosa.hz_to_mel / mel_to_hz with htk=False)
# ----------------------------------------------------------------------------
_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    mel = f / _F_SP
    log_t = f >= _MIN_LOG_HZ
    safe = np.where(log_t, f, _MIN_LOG_HZ)
    return np.where(log_t, _MIN_LOG_MEL + np.log(safe / _MIN_LOG_HZ) / _LOGSTEP, mel)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f = _F_SP * m
    log_t = m >= _MIN_LOG_MEL
    return np.where(log_t, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), f)


def mel_filterbank(sr=16000, n_fft=2048, n_mels=80, fmin=0.0, fmax=8000.0):
    """``librosa.filters.mel(htk=False, norm='slaney', dtype=float32)`` -> f32[n_mels, 1+n_fft//2]."""
    n_bins = 1 + n_fft // 2
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def hann_periodic(win_length):
    """``scipy.signal.get_window('hann', M, fftbins=True)`` (float64)."""
    n = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def num_frames(n_samples, n_fft=2048, hop=512, center=True):
    if center:
        return 1 + n_samples // hop
    return 1 + (n_samples - n_fft) // hop


def frame_index(t, n, n_fft=2048, hop=512):
    """Source sample index of tap ``n`` of frame ``t`` (centre padding n_fft//2); <0 or >=N means zero."""
    return hop * t + n - n_fft // 2


def stft_power(y, n_fft=2048, hop=512, win_length=2048, high_precision=False):
    """|STFT|^2, centre zero padding (librosa>=0.10 ``pad_mode='constant'``) -> [1+n_fft//2, W]."""
    y = np.asarray(y)
    in_dtype = y.dtype if y.dtype in (np.float32, np.float64) else np.float32
    y = y.astype(in_dtype, copy=False)
    win = hann_periodic(win_length)
    if win_length < n_fft:  # librosa util.pad_center
        lpad = (n_fft - win_length) // 2
        win = np.pad(win, (lpad, n_fft - win_length - lpad))
    ypad = np.pad(y, (n_fft // 2, n_fft // 2), mode="constant")
    n_fr = 1 + (len(ypad) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_fr)[None, :]
    frames = ypad[idx]                                   # [n_fft, W]
    spec = np.fft.rfft(win[:, None] * frames, axis=0)    # float64 product -> complex128
    if high_precision or in_dtype == np.float64:
        return (np.abs(spec) ** 2.0)
    spec = spec.astype(np.complex64)
    return np.abs(spec) ** np.float32(2.0)


def power_to_db(S, amin=1e-10, top_db=80.0):
    """``librosa.power_to_db(S, ref=np.max)``."""
    S = np.asarray(S)
    magnitude = np.abs(S)
    ref_value = np.max(magnitude)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def melspectrogram(y, sr=16000, n_fft=2048, hop=512, win_length=2048, n_mels=80,
                   fmin=0.0, fmax=8000.0, high_precision=False):
    S = stft_power(y, n_fft, hop, win_length, high_precision=high_precision)
    fb = mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    if high_precision:
        return fb.astype(np.float64) @ S.astype(np.float64)
    return np.einsum("ft,mf->mt", S, fb.astype(S.dtype), optimize=True)


def audio_to_mel(y, sr=16000, n_fft=2048, hop=512, win_length=2048, n_mels=80,
                 fmin=0.0, fmax=8000.0, duration=1.0, high_precision=False):
    """Restatement of ``AudioProcessor.audio_to_mel`` (wakeword_training_script.py:85-101)."""
    if len(y) == 0:
        return np.zeros((n_mels, int(sr * duration / hop) + 1))
    return power_to_db(melspectrogram(y, sr, n_fft, hop, win_length, n_mels, fmin, fmax,
                                      high_precision=high_precision))


def audio_to_mel_batch(clips, **kw):
    return np.stack([audio_to_mel(c, **kw) for c in clips]).astype(np.float32)
