"""Numpy restatement of the augmentation / conditioning stages (TEST INFRASTRUCTURE).

Stage semantics and the reference lines they follow:
  * peak normalise   -- ``AudioProcessor.normalize_audio`` wakeword_training_script.py:73-76
  * crop / zero pad  -- ``AudioProcessor.pad_or_truncate``  wakeword_training_script.py:78-83
  * circular shift   -- ``np.roll`` in ``augment_audio``    wakeword_training_script.py:106-108
  * SNR noise mix    -- ``snr_mixer`` stock/ms_snsd/MS-SNSD/audiolib.py:55-71 (cannot be
                        imported: audiolib.py:7 imports soundfile) -- restated line by line,
                        including its ``sqrt`` quirk (realised SNR = snr/2 dB)
  * speed change     -- polyphase windowed-sinc resampling with the structure of
                        ``torchaudio.functional.resample`` (sinc_interp_hann, width 6,
                        rolloff 0.99), then crop/pad.

All random quantities (shift, crop offset, noise index/offset, SNR, speed) are
HOST-drawn integers/floats handed in as parameters, so the device path can be
bit-exact on every index.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# flag bits of the per-clip augmentation word (mirrors include/wakeword_b200.h)
F_NORM_IN = 1 << 0    # peak normalise first (process_audio_file order, :130)
F_SHIFT = 1 << 1      # np.roll by `shift`
F_SPEED = 1 << 2      # resample rs_orig -> rs_new, then crop(crop_off)/pad to N
F_NOISE = 1 << 3      # snr_mixer with bank[noise_idx, noise_off : noise_off+N]
F_GAIN = 1 << 4       # multiply by `gain`
F_NORM_OUT = 1 << 5   # peak normalise last


def normalize_audio(a):
    a = np.asarray(a)
    if len(a) == 0:
        return a
    return a / np.max(np.abs(a))


def pad_or_truncate(a, target, crop_off=0):
    """crop_off plays the role of ``random.randint(0, len-target)`` (inclusive bounds)."""
    a = np.asarray(a)
    if len(a) > target:
        assert 0 <= crop_off <= len(a) - target
        return a[crop_off:crop_off + target]
    return np.pad(a, (0, target - len(a)), mode="constant")


def time_shift(a, shift):
    """``np.roll``: out[i] = a[(i - shift) mod N]."""
    return np.roll(a, shift)


def shift_source_index(i, shift, n):
    return (i - shift) % n


def snr_mixer(clean, noise, snr):
    clean = np.asarray(clean)
    noise = np.asarray(noise)
    rmsclean = (clean ** 2).mean() ** 0.5
    scalarclean = 10 ** (-25 / 20) / rmsclean
    clean = clean * scalarclean
    rmsclean = (clean ** 2).mean() ** 0.5
    rmsnoise = (noise ** 2).mean() ** 0.5
    scalarnoise = 10 ** (-25 / 20) / rmsnoise
    noise = noise * scalarnoise
    rmsnoise = (noise ** 2).mean() ** 0.5
    noisescalar = np.sqrt(rmsclean / (10 ** (snr / 20)) / rmsnoise)
    noisenewlevel = noise * noisescalar
    return clean, noisenewlevel, clean + noisenewlevel


# ----------------------------------------------------------------------------
# polyphase resampling
# ----------------------------------------------------------------------------
def resample_plan(orig, new, n_in, lowpass_filter_width=6, rolloff=0.99):
    """Integer structure of the resampler: (o, n, width, taps, out_len).  Bit-exact contract."""
    g = math.gcd(int(orig), int(new))
    o, n = int(orig) // g, int(new) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    taps = 2 * width + o
    out_len = int(math.ceil(n * n_in / o))
    return o, n, width, taps, out_len


def resample_kernel(orig, new, lowpass_filter_width=6, rolloff=0.99):
    """float32 [n phases, taps] windowed-sinc table (computed in float64 like torchaudio)."""
    g = math.gcd(int(orig), int(new))
    o, n = int(orig) // g, int(new) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    t = np.arange(0, -n, -1, dtype=np.float64)[:, None] / n + idx
    t = t * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / o
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * scale
    return k.astype(np.float32), width


def resample_tap_range(j, o, n, width):
    """Output sample j reads xpad[q*o + k], k in [0, 2*width+o)  ==  x[q*o + k - width]."""
    q, p = divmod(j, n)
    return q, p, q * o - width


def resample(x, orig, new):
    x = np.asarray(x, dtype=np.float32)
    o, n, width, taps, out_len = resample_plan(orig, new, len(x))
    kern, _ = resample_kernel(orig, new)
    xpad = np.pad(x, (width, width + o))
    n_q = (len(xpad) - taps) // o + 1
    win = np.lib.stride_tricks.sliding_window_view(xpad, taps)[::o][:n_q]   # [q, taps]
    y = (win.astype(np.float32) @ kern.T.astype(np.float32)).reshape(-1)    # j = q*n + p
    return y[:out_len]


def speed_change(x, rs_orig, rs_new, crop_off=0):
    n = len(x)
    return pad_or_truncate(resample(x, rs_orig, rs_new), n, crop_off)


# ----------------------------------------------------------------------------
# the batched stage set the CUDA augment kernel implements
# ----------------------------------------------------------------------------
@dataclass
class AugParams:
    """SoA per-clip parameters (host-drawn).  dtypes mirror ``ww_aug`` in include/wakeword_b200.h."""
    flags: np.ndarray      # u32
    shift: np.ndarray      # i32
    rs_orig: np.ndarray    # i32
    rs_new: np.ndarray     # i32
    crop_off: np.ndarray   # i32
    noise_idx: np.ndarray  # i32
    noise_off: np.ndarray  # i32
    snr_db: np.ndarray     # f32
    gain: np.ndarray       # f32

    def __len__(self):
        return len(self.flags)


def augment_clip(x, flags, shift=0, rs_orig=100, rs_new=100, crop_off=0, noise_seg=None,
                 snr_db=0.0, gain=1.0):
    x = np.asarray(x, dtype=np.float32)
    if flags & F_NORM_IN:
        x = normalize_audio(x).astype(np.float32)
    if flags & F_SHIFT:
        x = time_shift(x, int(shift))
    if flags & F_SPEED:
        x = speed_change(x, int(rs_orig), int(rs_new), int(crop_off)).astype(np.float32)
    if flags & F_NOISE:
        x = snr_mixer(x.astype(np.float32), np.asarray(noise_seg, dtype=np.float32), np.float32(snr_db))[2]
        x = x.astype(np.float32)
    if flags & F_GAIN:
        x = (x * np.float32(gain)).astype(np.float32)
    if flags & F_NORM_OUT:
        x = normalize_audio(x).astype(np.float32)
    return x


def augment_batch(clips, bank, p: AugParams):
    n = clips.shape[1]
    out = np.empty_like(clips, dtype=np.float32)
    for b in range(len(clips)):
        seg = None
        if p.flags[b] & F_NOISE:
            seg = bank[p.noise_idx[b], p.noise_off[b]:p.noise_off[b] + n]
        out[b] = augment_clip(clips[b], int(p.flags[b]), p.shift[b], p.rs_orig[b], p.rs_new[b],
                              p.crop_off[b], seg, p.snr_db[b], p.gain[b])
    return out
