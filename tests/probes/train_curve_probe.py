"""Probe: loss curves of the toy problem of tests/test_train_tc_gpu.py with both backward kernels, and the gradient
difference between the two kernels ALONG the fp32 trajectory (same weights each step)."""
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


sd = R.seeded_state_dict(64, seed=3)
rng = np.random.default_rng(0)
y_np = rng.integers(0, 2, 64)
x_np = (rng.standard_normal((64, 1, 80, 32)) * 5 - 40).astype(np.float32)
x_np[y_np == 1, :, 20:40, :] += 25.0
x, y = torch.from_numpy(x_np).cuda(), torch.from_numpy(y_np.astype(np.int64)).cuda()
lr = float(sys.argv[1]) if len(sys.argv) > 1 else 3e-3
for kern, fast in (("fp32", "1"), ("tc", "1"), ("tc", "0")):
    os.environ["WW_TRAIN_KERNEL"], os.environ["WW_TRAIN_FAST"] = kern, fast
    net, tr = trainer(sd)
    tr.lr = lr
    c = [tr.train_step(x, y)[0].item() for _ in range(60)]
    print(kern, "fast" + fast, " ".join(f"{v:.3f}" for v in c))
# gradient difference along the fp32 trajectory
os.environ["WW_TRAIN_FAST"] = "1"
os.environ["WW_TRAIN_KERNEL"] = "fp32"
net, tr = trainer(sd)
tr.lr = lr
for step in range(45):
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    if step % 4 == 0:
        os.environ["WW_TRAIN_KERNEL"] = "tc"
        n2, t2 = trainer({k: v.cpu().numpy() for k, v in state.items()})
        l2 = t2.train_step(x, y)[0].item()
        g2 = {k: v.cpu().numpy().copy() for k, v in t2.gradients(n2.engine()).items()}
        os.environ["WW_TRAIN_KERNEL"] = "fp32"
    loss = tr.train_step(x, y)[0].item()
    if step % 4 == 0:
        g1 = {k: v.cpu().numpy() for k, v in tr.gradients(net.engine()).items()}
        rel = {k: float(np.abs(g2[k] - g1[k]).max() / max(np.abs(g1[k]).max(), 1e-30)) for k in g1 if k.startswith("conv") or k.startswith("fc")}
        print(f"step {step:2d} loss fp32 {loss:.4f} tc {l2:.4f} " + " ".join(f"{k.replace('.weight', '.w').replace('.bias', '.b')}={v:.1e}" for k, v in rel.items()))

# ---- torch's own arithmetic on the same problem: exact fp32 and the TF32 default of the reference's GPU training
import torch.nn.functional as F  # noqa: E402


def torch_curve(tf32, steps=60):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    p = {k: torch.from_numpy(np.asarray(v)).cuda().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(p.values()), lr=lr, weight_decay=1e-5)
    out = []
    for _ in range(steps):
        opt.zero_grad()
        h = x
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
        opt.step()
        out.append(loss.item())
    return out


for tf32 in (False, True):
    print("torch", "tf32" if tf32 else "fp32", " ".join(f"{v:.3f}" for v in torch_curve(tf32)))

# ---- elementwise relative error of the tensor-core gradients at the initial weights (elements above 1 % of the tensor maximum)
os.environ["WW_TRAIN_KERNEL"] = "fp32"
n1, t1 = trainer(sd)
t1.train_step(x, y)
g1 = {k: v.cpu().numpy().copy() for k, v in t1.gradients(n1.engine()).items()}
os.environ["WW_TRAIN_KERNEL"] = "tc"
n2, t2 = trainer(sd)
t2.train_step(x, y)
g2 = {k: v.cpu().numpy().copy() for k, v in t2.gradients(n2.engine()).items()}
for k in g1:
    if not k.startswith("conv"):
        continue
    a, b = g1[k].ravel(), g2[k].ravel()
    big = np.abs(a) > 1e-2 * np.abs(a).max()
    r = np.abs(b[big] - a[big]) / np.abs(a[big])
    print(f"{k:13s} elements {big.sum():6d}/{a.size:6d}  rel err median {np.median(r):.1e}  p90 {np.quantile(r, 0.9):.1e}  max {r.max():.1e}")
