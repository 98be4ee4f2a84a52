"""Probe (not a test): gradient error of torch's own GPU conv arithmetic (fp32 and the TF32 default of the reference's
training script) and of this library's two training kernels against float64 autograd, same inputs.  Prints max|d|/max|ref|."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import recipe as R          # noqa: E402
import wakeword_jupyterlab_b200 as ww   # noqa: E402


def torch_grads(sd, x, y, dtype, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    p = {k: torch.from_numpy(np.asarray(v)).cuda().to(dtype).requires_grad_(True) for k, v in sd.items()}
    h = x.to(dtype)
    for n in ("conv1", "conv2", "conv3"):
        h = F.relu(F.conv2d(h, p[n + ".weight"], p[n + ".bias"], padding=1))
    h = h.mean(dim=(2, 3))
    layer = 0
    while f"lstm.weight_ih_l{layer}" in p:
        gt = F.linear(h, p[f"lstm.weight_ih_l{layer}"], p[f"lstm.bias_ih_l{layer}"] + p[f"lstm.bias_hh_l{layer}"])
        H = gt.shape[1] // 4
        c = torch.sigmoid(gt[:, :H]) * torch.tanh(gt[:, 2 * H:3 * H])
        h = torch.sigmoid(gt[:, 3 * H:]) * torch.tanh(c)
        layer += 1
    loss = F.cross_entropy(F.linear(h, p["fc.weight"], p["fc.bias"]), y)
    loss.backward()
    return loss.item(), {k: v.grad.double().cpu().numpy() for k, v in p.items() if v.grad is not None}


def ours(sd, x, y, kernel):
    class MC(ww.ModelConfig):
        DROPOUT = 0.0
    net = ww.WakewordModel(MC, ww.AudioConfig).cuda().train()
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    tr = ww.WakewordTrainer(net, "cuda")
    os.environ["WW_TRAIN_KERNEL"] = kernel
    loss, _ = tr.train_step(x, y)
    torch.cuda.synchronize()
    return loss.item(), {k: v.double().cpu().numpy() for k, v in tr.gradients(net.engine()).items()}


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


for B in (1, 5, 70, 512):
    sd = R.seeded_state_dict(256, seed=2)
    rng = np.random.default_rng(100 + B)
    x = torch.from_numpy((rng.standard_normal((B, 1, 80, 32)) * 15 - 40).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.integers(0, 2, B).astype(np.int64)).cuda()
    l64, g64 = torch_grads(sd, x, y, torch.float64, False)
    rows = {"torch fp32": torch_grads(sd, x, y, torch.float32, False), "torch tf32": torch_grads(sd, x, y, torch.float32, True),
            "b200 fp32": ours(sd, x, y, "fp32"), "b200 tc": ours(sd, x, y, "tc")}
    for name, (l, g) in rows.items():
        print(f"B={B:4d} {name:11s} loss_rel={abs(l - l64) / abs(l64):.1e} " +
              " ".join(f"{k.replace('.weight', '.w').replace('.bias', '.b')}={rel(g[k], g64[k]):.1e}" for k in g64 if k.startswith("conv")))
