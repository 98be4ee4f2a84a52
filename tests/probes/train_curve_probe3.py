"""Probe: spikiness of the toy-problem training curve for several seeds, both backward kernels."""
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
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return net, ww.WakewordTrainer(net, "cuda")


for lr in (3e-3, 1e-3):
    for seed in range(5):
        sd = R.seeded_state_dict(64, seed=3 + seed)
        rng = np.random.default_rng(seed)
        y_np = rng.integers(0, 2, 64)
        x_np = (rng.standard_normal((64, 1, 80, 32)) * 5 - 40).astype(np.float32)
        x_np[y_np == 1, :, 20:40, :] += 25.0
        x, y = torch.from_numpy(x_np).cuda(), torch.from_numpy(y_np.astype(np.int64)).cuda()
        for kern in ("fp32", "tc"):
            os.environ["WW_TRAIN_KERNEL"] = kern
            net, tr = trainer(sd)
            tr.lr = lr
            c = np.array([tr.train_step(x, y)[0].item() for _ in range(80 if lr > 2e-3 else 160)])
            first = int(np.argmax(c < 0.1)) if (c < 0.1).any() else -1
            ups = int(((c[1:] - c[:-1]) > 0.1).sum())
            print(f"lr {lr:g} seed {seed} {kern:4s} first<0.1 at {first:3d}  jumps>0.1: {ups:2d}  max after: {c[first:].max() if first >= 0 else float('nan'):.3f}  final {c[-1]:.4f}")
