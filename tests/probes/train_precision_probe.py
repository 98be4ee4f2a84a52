#!/usr/bin/env python3
"""CPU probe for the round-2 tensor-core TRAINING step: which operand of the conv backward GEMMs
needs how many fp16 terms to keep the gradients inside the 1e-4 gate of tests/test_train_gpu.py?

    python tests/probes/train_precision_probe.py

The conv stack's backward is six GEMMs: dgrad3 / dgrad2 (dY x W, per pixel: rounding errors of dY are independent from
pixel to pixel, those of W are the same everywhere) and wgrad3 / wgrad2 / wgrad1 (dY x A summed over ALL pixels and clips:
rounding errors of both operands average out).  Each scheme below rounds the named operands to fp16 (scaled by a power of
two so nothing lands in fp16's subnormals, as the inference kernels already do for the weights), accumulates in fp64
(a stand-in for fp32 accumulation of at most a few thousand terms per tile with fp32 partial sums) and compares every conv
gradient with float64 autograd (max |d| / max |ref| per tensor, the test's metric).  The head (LSTM cells, Linear, loss)
stays exact fp32 on CUDA cores and is not perturbed here.  Test infrastructure only (imports oracle/)."""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
from torch.nn.grad import conv2d_input, conv2d_weight

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from oracle import augment as A, logmel as LM, recipe as R   # noqa: E402


def p2scale(t, top=13):
    m = t.abs().max().item()
    return 1.0 if m == 0 else 2.0 ** (top - math.frexp(m)[1])


def r16(t, terms):
    """t -> sum of `terms` fp16 numbers (after a power-of-two scale), returned in float64."""
    if terms == 0:
        return t.double()
    s = p2scale(t)
    x = (t * s).float()
    hi = x.half().float()
    out = hi.double()
    if terms == 2:
        out = out + (x - hi).half().double()
    return out / s


def reference(x, y, sd):
    p = {k: v.clone().double().requires_grad_(k.startswith("conv")) for k, v in sd.items()}
    a = x.double()
    for i in (1, 2, 3):
        a = F.relu(F.conv2d(a, p[f"conv{i}.weight"], p[f"conv{i}.bias"], padding=1))
    pooled = a.mean(dim=(2, 3))
    pooled.retain_grad()
    h = pooled
    layer = 0
    while f"lstm.weight_ih_l{layer}" in p:
        g = F.linear(h, p[f"lstm.weight_ih_l{layer}"], p[f"lstm.bias_ih_l{layer}"] + p[f"lstm.bias_hh_l{layer}"])
        H = g.shape[1] // 4
        i_, g_, o_ = torch.sigmoid(g[:, :H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
        h = o_ * torch.tanh(i_ * g_)
        layer += 1
    loss = F.cross_entropy(F.linear(h, p["fc.weight"], p["fc.bias"]), y)
    loss.backward()
    return {k: p[k].grad for k in p if k.startswith("conv")}, pooled.grad.detach()


def emulated(x, sd, dpooled, terms):
    """terms: dict operand -> fp16 terms (0 = exact) for 'act' (saved activations), 'dy' (output gradients), 'w' (weights)."""
    W = {i: sd[f"conv{i}.weight"].double() for i in (1, 2, 3)}
    acts = [x.double()]
    for i in (1, 2, 3):     # forward exact here: the inference-path probe covers its error (1e-6 on the logits)
        acts.append(F.relu(F.conv2d(acts[-1], W[i], sd[f"conv{i}.bias"].double(), padding=1)))
    grads = {}
    hw = acts[3].shape[2] * acts[3].shape[3]
    dy = (dpooled[:, :, None, None] / hw) * (acts[3] > 0)
    for i in (3, 2, 1):
        dyr = r16(dy.float(), terms["dy"])
        ar = r16(acts[i - 1].float(), terms["act"])
        grads[f"conv{i}.weight"] = conv2d_weight(ar, W[i].shape, dyr, padding=1)
        grads[f"conv{i}.bias"] = dy.sum(dim=(0, 2, 3))                      # fp32 CUDA-core reduction of the fp32 dY
        if i > 1:
            dy = conv2d_input(acts[i - 1].shape, r16(W[i].float(), terms["w"]), dyr, padding=1) * (acts[i - 1] > 0)
    return grads


def main():
    torch.manual_seed(0)
    B = 24
    clips = R.make_clips(B, seed=1234)
    norm = np.stack([A.normalize_audio(c) for c in clips]).astype(np.float32)
    x = torch.from_numpy(LM.audio_to_mel_batch(norm)[:, None])
    y = torch.from_numpy(R.make_labels(B).astype(np.int64))
    sd = {k: torch.from_numpy(v) for k, v in R.seeded_state_dict(256, seed=0).items()}
    ref, dpooled = reference(x, y, sd)
    schemes = [("everything fp16, 1 term          ", dict(act=1, dy=1, w=1)),
               ("W hi+lo; dY, activations 1 term  ", dict(act=1, dy=1, w=2)),
               ("W, dY hi+lo; activations 1 term  ", dict(act=1, dy=2, w=2)),
               ("all operands hi+lo               ", dict(act=2, dy=2, w=2))]
    names = ["conv3.weight", "conv2.weight", "conv1.weight", "conv3.bias", "conv2.bias", "conv1.bias"]
    print(f"B = {B}; max |d| / max |ref| per gradient tensor against float64 autograd (gate: 1e-4)\n")
    print(" " * 36 + "  ".join(f"{n:>13s}" for n in names))
    for label, t in schemes:
        g = emulated(x, sd, dpooled, t)
        errs = [((g[n] - ref[n]).abs().max() / ref[n].abs().max()).item() for n in names]
        print(f"{label} |" + "  ".join(f"{e:13.2e}" for e in errs))


if __name__ == "__main__":
    main()
