"""``WakewordTrainer`` -- the reference's training class (/root/reference/wakeword_training_script.py:219-345) with
its batch step (:247-257) running as hand-written CUDA through the C ABI (``ww_train_backward`` / ``ww_train_apply``).

Same constructor and ``train_epoch`` / ``validate`` contract: ``CrossEntropyLoss`` (mean), Adam(lr = LEARNING_RATE,
weight_decay = 1e-5), ``clip_grad_norm_`` kept as the no-op it is in the reference (it runs before ``backward`` on zeroed
gradients).  Data parallel: when ``torch.distributed`` is initialised, every rank steps on its own shard of the batch
and the flat gradient buffer is sum-all-reduced (NCCL over NVLink) and scaled by 1 / world_size before Adam -- the only
collective in the package.  Dropout masks are drawn on the device by torch (seed-supplied inputs of the kernels, like
the augmentation parameters)."""
from __future__ import annotations

import ctypes as C

import torch

from .config import TrainingConfig
from .sharding import allreduce_mean_


class _DevArray:
    """Expose a raw device pointer as a 1-D float32 array to torch (zero-copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class WakewordTrainer:
    BETAS = (0.9, 0.999)
    EPS = 1e-8
    WEIGHT_DECAY = 1e-5          # wakeword_training_script.py:226

    def __init__(self, model, device, config=TrainingConfig):
        self.model = model
        self.device = torch.device(device)
        self.config = config
        self.lr = float(config.LEARNING_RATE)
        self.train_losses, self.val_losses, self.train_accuracies, self.val_accuracies = [], [], [], []
        self.patience = 10
        self.best_val_acc = 0.0
        self.epochs_no_improve = 0
        self._grad_view = None

    # ------------------------------------------------------------------ one optimisation step
    def _engine(self, width):
        eng = self.model.engine(self.device, width=width)
        return eng

    def grad_view(self, eng):
        """Flat fp32 gradient buffer of the context as a torch tensor (shared memory, not a copy)."""
        n = int(eng.lib.ww_train_n_params(eng._ctx))
        if n < 0:
            eng._chk(-1, "ww_train_n_params")
        ptr = eng.lib.ww_train_grad_buffer(eng._ctx)
        return torch.as_tensor(_DevArray(ptr, n), device=eng.device)

    def gradients(self, eng):
        """name -> gradient tensor (copies), for tests and diagnostics."""
        flat = self.grad_view(eng)
        out = {}
        for name, p in self.model.state_dict().items():
            off, cnt = C.c_int64(), C.c_int64()
            eng._chk(eng.lib.ww_train_param_range(eng._ctx, name.encode(), C.byref(off), C.byref(cnt)), "ww_train_param_range")
            out[name] = flat[off.value:off.value + cnt.value].clone().reshape(p.shape)
        return out

    def _dropout_masks(self, B, generator=None):
        mc = self.model.config
        p = float(mc.DROPOUT)
        if p <= 0.0 or not self.model.training:
            return None, None
        H, L = mc.HIDDEN_SIZE, mc.NUM_LAYERS
        keep = 1.0 - p

        def mask(*shape):
            return (torch.rand(shape, device=self.device, generator=generator) < keep).float() / keep
        return (mask(L - 1, B, H) if L > 1 else None), mask(B, H)

    def train_step(self, data, target, generator=None):
        """data [B,1,N_MELS,W] float32, target int64 [B] (both CUDA) -> (loss 0-dim tensor, train-mode logits [B,C])."""
        import torch.distributed as dist
        data = data.to(self.device, torch.float32).contiguous()
        target = target.to(self.device, torch.int64).reshape(-1).contiguous()
        B = data.shape[0]
        eng = self._engine(data.shape[-1])
        st = eng._stream()
        loss = torch.empty((), device=self.device, dtype=torch.float32)
        logits = torch.empty((B, eng.n_classes), device=self.device, dtype=torch.float32)
        m_lstm, m_out = self._dropout_masks(B, generator)
        ptr = (lambda t: C.c_void_p(t.data_ptr()) if t is not None else None)
        eng._chk(eng.lib.ww_train_backward(eng._ctx, ptr(data), ptr(target), B, ptr(m_lstm), ptr(m_out), ptr(loss),
                                           ptr(logits), st), "ww_train_backward")
        scale = 1.0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            scale = allreduce_mean_(self.grad_view(eng))                    # the package's only collective
        eng._chk(eng.lib.ww_train_apply(eng._ctx, self.lr, self.BETAS[0], self.BETAS[1], self.EPS, self.WEIGHT_DECAY,
                                        scale, st), "ww_train_apply")
        self._pull_weights(eng)
        return loss, logits

    def _pull_weights(self, eng):
        """Copy the updated master weights back into the module's parameter tensors (device to device)."""
        for name, p in self.model.state_dict(keep_vars=True).items():
            eng._chk(eng.lib.ww_get_weights(eng._ctx, name.encode(), C.c_void_p(p.data_ptr())), f"ww_get_weights({name})")

    # ------------------------------------------------------------------ the reference's epoch loops
    def train_epoch(self, train_loader):
        self.model.train()
        running_loss, correct, total = 0.0, 0, 0
        for data, target in train_loader:
            loss, output = self.train_step(data, target.squeeze())
            running_loss += loss.item()
            predicted = output.argmax(dim=1)
            tgt = target.to(self.device).reshape(-1)
            total += tgt.numel()
            correct += int((predicted == tgt).sum().item())
        return running_loss / max(len(train_loader), 1), 100.0 * correct / max(total, 1)

    def validate(self, val_loader):
        self.model.eval()
        running_loss, correct, total = 0.0, 0, 0
        with torch.no_grad():
            for data, target in val_loader:
                data, tgt = data.to(self.device), target.to(self.device).reshape(-1)
                output = self.model(data)
                running_loss += torch.nn.functional.cross_entropy(output, tgt).item()
                correct += int((output.argmax(dim=1) == tgt).sum().item())
                total += tgt.numel()
        return running_loss / max(len(val_loader), 1), 100.0 * correct / max(total, 1)
