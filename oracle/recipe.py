"""Deterministic synthetic inputs (TEST INFRASTRUCTURE).

The reference's own synthetic distribution, ``create_sample_data``
(/root/reference/wakeword_training_script.py:350-393), made deterministic as
SURVEY.md section 8(d) prescribes: clip i is "positive" if i % 3 == 0
(``0.1 N(0,1) + 0.3 sin(2 pi 200 t) + 0.2 sin(2 pi 400 t)``, :362-369), else
"negative" (``0.2 N(0,1)``, :377); noise bank rows are ``0.1 N(0,1)`` (:386).
Augmentation parameters are drawn on the host with ``random.Random(seed)`` in the
reference's draw order (``augment_audio`` :103-123).
"""
from __future__ import annotations

import random

import numpy as np

from . import augment as A

SR = 16000


def make_clips(n, seed=1234, n_samples=SR):
    rng = np.random.default_rng(seed)
    t = np.linspace(0, n_samples / SR, n_samples)
    tone = 0.3 * np.sin(2 * np.pi * 200 * t) + 0.2 * np.sin(2 * np.pi * 400 * t)
    out = np.empty((n, n_samples), dtype=np.float32)
    for i in range(n):
        noise = rng.standard_normal(n_samples)
        if i % 3 == 0:
            out[i] = (0.1 * noise + tone).astype(np.float32)
        else:
            out[i] = (0.2 * noise).astype(np.float32)
    return out


def make_labels(n):
    return (np.arange(n) % 3 == 0).astype(np.int64)


def make_noise_bank(m=20, length=5 * SR, seed=4321):
    rng = np.random.default_rng(seed)
    return (0.1 * rng.standard_normal((m, length))).astype(np.float32)


def draw_aug_params(n, n_samples=SR, bank_shape=(20, 5 * SR), seed=2024, prob=0.8,
                    shift_max=0.3, speed_grid=(80, 120), snr_grid=(0.0, 10.0, 20.0, 30.0, 40.0),
                    norm_in=True, norm_out=True) -> A.AugParams:
    """Host draws in the reference's stage order: shift, (pitch: skipped), speed(+crop), noise."""
    r = random.Random(seed)
    flags = np.zeros(n, np.uint32)
    shift = np.zeros(n, np.int32)
    rs_orig = np.full(n, 100, np.int32)
    rs_new = np.full(n, 100, np.int32)
    crop = np.zeros(n, np.int32)
    nidx = np.zeros(n, np.int32)
    noff = np.zeros(n, np.int32)
    snr = np.zeros(n, np.float32)
    gain = np.ones(n, np.float32)
    for b in range(n):
        f = 0
        if norm_in:
            f |= A.F_NORM_IN
        if r.random() < prob:                      # :106
            f |= A.F_SHIFT
            shift[b] = int(r.uniform(-shift_max, shift_max) * SR)   # :107 (truncation toward zero)
        if r.random() < prob:                      # :114 (speed; pitch stage :110 is out of scope)
            s = r.randint(speed_grid[0], speed_grid[1])
            if s != 100:
                f |= A.F_SPEED
                rs_orig[b], rs_new[b] = s, 100
                _, _, _, _, out_len = A.resample_plan(s, 100, n_samples)
                if out_len > n_samples:
                    crop[b] = r.randint(0, out_len - n_samples)    # :80 inclusive
        if r.random() < prob:                      # :119 (noise; SNR mixer replaces Gaussian)
            f |= A.F_NOISE
            nidx[b] = r.randrange(bank_shape[0])
            noff[b] = r.randint(0, bank_shape[1] - n_samples)
            snr[b] = r.choice(snr_grid)
        if norm_out:
            f |= A.F_NORM_OUT
        flags[b] = f
    return A.AugParams(flags, shift, rs_orig, rs_new, crop, nidx, noff, snr, gain)


def seeded_state_dict(hidden=256, layers=2, n_classes=2, seed=0):
    """Default-init-shaped weights (U(-1/sqrt(fan_in), 1/sqrt(fan_in)))
    from a numpy PCG64 stream, so fixtures need not store them."""
    rng = np.random.default_rng(seed)
    sd = {}

    def u(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    for name, cout, cin in (("conv1", 32, 1), ("conv2", 64, 32), ("conv3", 128, 64)):
        sd[f"{name}.weight"] = u((cout, cin, 3, 3), cin * 9)
        sd[f"{name}.bias"] = u((cout,), cin * 9)
    for l in range(layers):
        in_sz = 128 if l == 0 else hidden
        sd[f"lstm.weight_ih_l{l}"] = u((4 * hidden, in_sz), hidden)
        sd[f"lstm.weight_hh_l{l}"] = u((4 * hidden, hidden), hidden)
        sd[f"lstm.bias_ih_l{l}"] = u((4 * hidden,), hidden)
        sd[f"lstm.bias_hh_l{l}"] = u((4 * hidden,), hidden)
    sd["fc.weight"] = u((n_classes, hidden), hidden)
    sd["fc.bias"] = u((n_classes,), hidden)
    return sd
