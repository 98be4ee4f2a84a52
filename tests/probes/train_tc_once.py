"""One tensor-core training step on a tiny batch (for compute-sanitizer runs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import recipe as R          # noqa: E402
import wakeword_jupyterlab_b200 as ww   # noqa: E402

os.environ["WW_TRAIN_KERNEL"] = "tc"
B, W = int(sys.argv[1]) if len(sys.argv) > 1 else 3, int(sys.argv[2]) if len(sys.argv) > 2 else 32


class MC(ww.ModelConfig):
    DROPOUT = 0.0


sd = R.seeded_state_dict(256, seed=2)
net = ww.WakewordModel(MC, ww.AudioConfig).cuda().train()
net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
tr = ww.WakewordTrainer(net, "cuda")
rng = np.random.default_rng(1)
x = torch.from_numpy((rng.standard_normal((B, 1, 80, W)) * 15 - 40).astype(np.float32)).cuda()
y = torch.from_numpy(rng.integers(0, 2, B).astype(np.int64)).cuda()
for _ in range(2):
    loss, _ = tr.train_step(x, y)
torch.cuda.synchronize()
print("loss", loss.item())
