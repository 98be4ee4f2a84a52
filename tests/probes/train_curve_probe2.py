"""Probe: tensor-core training on the toy problem; at every step the exact kernels' gradient at the SAME weights."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import recipe as R          # noqa: E402
import wakeword_jupyterlab_b200 as ww   # noqa: E402


def trainer(sd, hidden=64):
    class MC(ww.ModelConfig):
        HIDDEN_SIZE = hidden
        DROPOUT = 0.0
    net = ww.WakewordModel(MC, ww.AudioConfig).cuda().train()
    net.load_state_dict({k: (v if torch.is_tensor(v) else torch.from_numpy(np.asarray(v))) for k, v in sd.items()})
    return net, ww.WakewordTrainer(net, "cuda")


sd = R.seeded_state_dict(64, seed=3)
rng = np.random.default_rng(0)
y_np = rng.integers(0, 2, 64)
x_np = (rng.standard_normal((64, 1, 80, 32)) * 5 - 40).astype(np.float32)
x_np[y_np == 1, :, 20:40, :] += 25.0
x, y = torch.from_numpy(x_np).cuda(), torch.from_numpy(y_np.astype(np.int64)).cuda()
lr = 3e-3
os.environ["WW_TRAIN_KERNEL"] = "tc"
net, tr = trainer(sd)
tr.lr = lr
for step in range(48):
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    os.environ["WW_TRAIN_KERNEL"] = "fp32"
    n2, t2 = trainer(state)
    l2 = t2.train_step(x, y)[0].item()
    g2 = {k: v.cpu().numpy().copy() for k, v in t2.gradients(n2.engine()).items()}
    os.environ["WW_TRAIN_KERNEL"] = "tc"
    loss = tr.train_step(x, y)[0].item()
    g1 = {k: v.cpu().numpy() for k, v in tr.gradients(net.engine()).items()}
    rel = {k: float(np.abs(g2[k] - g1[k]).max() / max(np.abs(g2[k]).max(), 1e-30)) for k in g1 if np.abs(g2[k]).max() > 0}
    worst = max(rel, key=rel.get)
    wmax = {k: float(v.abs().max()) for k, v in state.items() if k.endswith("weight") and k.startswith("conv")}
    print(f"step {step:2d} loss tc {loss:.4f} fp32@same {l2:.4f}  worst {worst} {rel[worst]:.1e}  conv3.w {rel['conv3.weight']:.1e} conv2.w {rel['conv2.weight']:.1e} conv1.w {rel['conv1.weight']:.1e} |gmax3| {np.abs(g2['conv3.weight']).max():.1e} wmax {wmax}")
