#!/usr/bin/env python3
"""Golden vectors for the training step: run in the authoring container only.

    python tests/golden/make_golden_train.py

Runs the UNMODIFIED reference ``WakewordTrainer.train_epoch`` (/root/reference/wakeword_training_script.py:238-267,
torch CPU fp32, imported through ``oracle.ref_shim``) for three one-batch epochs on seeded in-memory features and
records the loss of every step, the gradients left in ``.grad`` after the first step and the parameters after the
third.  Dropout is set to 0 through the reference's own config hook (its masks come from torch's CPU generator and
cannot be reproduced elsewhere); hidden size 32 keeps the fixture small.  Writes tests/golden/train_ref.npz."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import recipe as R           # noqa: E402
from oracle import ref_shim              # noqa: E402

HIDDEN, B, STEPS, SEED = 32, 6, 3, 77


def make_batches():
    rng = np.random.default_rng(SEED)
    xs = (rng.standard_normal((STEPS, B, 1, 80, 32)) * 15.0 - 40.0).astype(np.float32)      # dB-like features
    ys = rng.integers(0, 2, size=(STEPS, B, 1)).astype(np.int64)
    return xs, ys


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    ref = ref_shim.load_reference()

    class MC(ref.ModelConfig):
        HIDDEN_SIZE = HIDDEN
        DROPOUT = 0.0

    model = ref.WakewordModel(MC, ref.AudioConfig)
    sd0 = R.seeded_state_dict(HIDDEN, seed=5)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd0.items()})
    trainer = ref.WakewordTrainer(model, torch.device("cpu"))
    xs, ys = make_batches()
    out = {"hidden": HIDDEN, "batch": B, "steps": STEPS, "seed": SEED, "weight_seed": 5}
    losses = []
    for s in range(STEPS):
        loader = [(torch.from_numpy(xs[s]), torch.from_numpy(ys[s]))]
        loss, acc = trainer.train_epoch(loader)
        losses.append(loss)
        if s == 0:
            for name, p in model.named_parameters():
                out["grad1/" + name] = p.grad.detach().numpy().copy()
    out["losses"] = np.asarray(losses, np.float64)
    for name, p in model.state_dict().items():
        out["param3/" + name] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "train_ref.npz"), **out)
    print("losses", losses)
    print("wrote train_ref.npz", os.path.getsize(os.path.join(HERE, "train_ref.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
