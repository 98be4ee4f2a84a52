"""CPU restatement of ``WakewordModel.forward`` + ``predict_wakeword`` (TEST INFRASTRUCTURE).

Follows /root/reference/wakeword_training_script.py:167-184 (conv1/2/3 + ReLU,
global mean, 2-layer LSTM fed a length-1 sequence with zero state, Linear) and
wakeword_training.ipynb:886-891 (softmax, ``prob >= threshold``).

With T = 1 and h0 = c0 = 0 the LSTM reduces per layer to
    g = W_ih x + b_ih + b_hh ;  c = sigmoid(g_i) * tanh(g_g) ;  h = sigmoid(g_o) * tanh(c)
(gate row order i, f, g, o; ``weight_hh`` and the forget rows never reach the
output -- SURVEY.md section 0 trap 2).  ``tests/golden/make_golden.py`` pins this
closed form against the UNMODIFIED reference module (eval mode).

Two implementations: ``forward_numpy`` (float64 or float32, im2col + matmul;
the checker) and ``forward_torch_cpu`` (torch CPU fp32, multi-threaded; the port
timed as the CPU baseline, because the reference's own model is torch-on-CPU).
"""
from __future__ import annotations

import numpy as np


def _conv3x3_relu(x, w, b):
    """x [B,Cin,H,W], w [Cout,Cin,3,3], zero padding 1, stride 1 -> relu(conv) [B,Cout,H,W]."""
    B, C, H, W = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    cols = np.empty((B, C, 3, 3, H, W), dtype=x.dtype)
    for ky in range(3):
        for kx in range(3):
            cols[:, :, ky, kx] = xp[:, :, ky:ky + H, kx:kx + W]
    out = np.einsum("bckhw,ock->bohw", cols.reshape(B, C, 9, H, W),
                    w.reshape(w.shape[0], C, 9).astype(x.dtype), optimize=True)
    out += b.astype(x.dtype)[None, :, None, None]
    return np.maximum(out, 0)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def pooled_features(x, sd, dtype=np.float64):
    x = np.asarray(x, dtype=dtype)
    for name in ("conv1", "conv2", "conv3"):
        x = _conv3x3_relu(x, np.asarray(sd[f"{name}.weight"]), np.asarray(sd[f"{name}.bias"]))
    return x.mean(axis=(2, 3))


def head(p, sd, dtype=np.float64):
    x = np.asarray(p, dtype=dtype)
    layer = 0
    while f"lstm.weight_ih_l{layer}" in sd:
        w = np.asarray(sd[f"lstm.weight_ih_l{layer}"], dtype=dtype)
        b = (np.asarray(sd[f"lstm.bias_ih_l{layer}"], dtype=dtype)
             + np.asarray(sd[f"lstm.bias_hh_l{layer}"], dtype=dtype))
        h = w.shape[0] // 4
        g = x @ w.T + b
        gi, gg, go = g[:, 0:h], g[:, 2 * h:3 * h], g[:, 3 * h:4 * h]
        c = _sigmoid(gi) * np.tanh(gg)
        x = _sigmoid(go) * np.tanh(c)
        layer += 1
    return x @ np.asarray(sd["fc.weight"], dtype=dtype).T + np.asarray(sd["fc.bias"], dtype=dtype)


def forward_numpy(x, sd, dtype=np.float64, chunk=8):
    """x [B,1,n_mels,W] -> logits [B,2] (eval mode: dropout is identity)."""
    outs = []
    for i in range(0, len(x), chunk):
        outs.append(head(pooled_features(x[i:i + chunk], sd, dtype), sd, dtype))
    return np.concatenate(outs)


def prob_and_decision(logits, threshold=0.8):
    """``softmax(logits)[:, 1]`` and ``prob >= threshold`` (wakeword_training.ipynb:888-891)."""
    z = np.asarray(logits, dtype=np.float64)
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    p1 = e[:, 1] / e.sum(axis=1)
    return p1, p1 >= threshold


def forward_torch_cpu(x, sd_t):
    """torch-CPU fp32 port used only for timing the CPU baseline (sd_t: dict of torch tensors)."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        x = F.relu(F.conv2d(x, sd_t["conv1.weight"], sd_t["conv1.bias"], padding=1))
        x = F.relu(F.conv2d(x, sd_t["conv2.weight"], sd_t["conv2.bias"], padding=1))
        x = F.relu(F.conv2d(x, sd_t["conv3.weight"], sd_t["conv3.bias"], padding=1))
        x = x.mean(dim=(2, 3))
        layer = 0
        while f"lstm.weight_ih_l{layer}" in sd_t:
            g = F.linear(x, sd_t[f"lstm.weight_ih_l{layer}"],
                         sd_t[f"lstm.bias_ih_l{layer}"] + sd_t[f"lstm.bias_hh_l{layer}"])
            h = g.shape[1] // 4
            c = torch.sigmoid(g[:, :h]) * torch.tanh(g[:, 2 * h:3 * h])
            x = torch.sigmoid(g[:, 3 * h:]) * torch.tanh(c)
            layer += 1
        return F.linear(x, sd_t["fc.weight"], sd_t["fc.bias"])
