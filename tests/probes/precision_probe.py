#!/usr/bin/env python3
"""CPU probe: which operand of conv2 / conv3 needs how many bits to stay inside the 1e-4 logit gate?

    python tests/probes/precision_probe.py

Emulates the tensor-core operand roundings in torch (fp64 accumulate) on the committed golden weights and the recipe
clips, and prints the logit error (max |d| / max |ref|) of every scheme in DESIGN.md section 4.  Result: activation
rounding averages out in the global mean, weight rounding does not, so only the weights need a hi + lo split and the lo
half (and the activation copy it multiplies) can be e4m3.  Test infrastructure only (imports oracle/)."""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from oracle import augment as A, logmel as LM, model as M, recipe as R   # noqa: E402

f16 = lambda x: x.to(torch.float16).to(torch.float32)                    # noqa: E731
bf16 = lambda x: x.to(torch.bfloat16).to(torch.float32)                  # noqa: E731
f8 = lambda x: x.to(torch.float8_e4m3fn).to(torch.float32)               # noqa: E731


def scale_of(w):
    return 2.0 ** (13 - math.frexp(w.abs().max().item())[1])


def conv(x, w, b, scheme):
    s = scale_of(w)
    ws = w * s
    if scheme == "bf16":
        y = F.conv2d(bf16(x).double(), bf16(ws).double(), None, padding=1)
    elif scheme == "bf16x3":
        xh, wh = bf16(x), bf16(ws)
        xl, wl = bf16(x - xh), bf16(ws - wh)
        y = (F.conv2d(xh.double(), wh.double(), None, padding=1) + F.conv2d(xh.double(), wl.double(), None, padding=1)
             + F.conv2d(xl.double(), wh.double(), None, padding=1))
    else:
        hi = f16(ws)
        lo = ws - hi
        y = F.conv2d(f16(x).double(), hi.double(), None, padding=1)
        if scheme == "fp16_hi_lo":
            y = y + F.conv2d(f16(x).double(), f16(lo).double(), None, padding=1)
        elif scheme == "fp16_hi_e4m3_lo":
            y = y + F.conv2d(f8(x).double(), f8(lo).double(), None, padding=1)
    return (y / s + b.double()[None, :, None, None]).float()


def pooled(x, sd, scheme):
    x = F.relu(F.conv2d(x.double(), sd["conv1.weight"].double(), sd["conv1.bias"].double(), padding=1)).float()
    if scheme == "exact":
        x = F.relu(F.conv2d(x.double(), sd["conv2.weight"].double(), sd["conv2.bias"].double(), padding=1))
        x = F.relu(F.conv2d(x, sd["conv3.weight"].double(), sd["conv3.bias"].double(), padding=1))
    else:
        x = F.relu(conv(x, sd["conv2.weight"], sd["conv2.bias"], scheme))
        x = F.relu(conv(x, sd["conv3.weight"], sd["conv3.bias"], scheme))
    return x.mean(dim=(2, 3)).double().numpy()


def main():
    torch.set_num_threads(8)
    golden = os.path.join(os.path.dirname(__file__), "..", "golden")
    clips = R.make_clips(32, seed=1234)
    norm = np.stack([A.normalize_audio(c) for c in clips]).astype(np.float32)
    feats = torch.from_numpy(LM.audio_to_mel_batch(norm)[:, None])
    for name in ("model_seeded", "model_trained"):
        z = np.load(os.path.join(golden, name + ".npz"))
        if name == "model_trained":
            sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
        else:
            sd = {k: torch.from_numpy(v) for k, v in R.seeded_state_dict(int(z["hidden"]), seed=int(z["weight_seed"])).items()}
        sdn = {k: v.numpy() for k, v in sd.items()}
        ref = M.head(pooled(feats, sd, "exact"), sdn)
        for scheme in ("bf16", "fp16", "fp16_hi_lo", "fp16_hi_e4m3_lo", "bf16x3"):
            got = M.head(pooled(feats, sd, scheme), sdn)
            print(f"{name:14s} {scheme:16s} logit error {np.abs(got - ref).max() / np.abs(ref).max():.3e}")


if __name__ == "__main__":
    main()
