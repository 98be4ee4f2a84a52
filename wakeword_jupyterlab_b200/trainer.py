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


class _PlateauScheduler:
    """``ReduceLROnPlateau(mode='max', factor=0.5, patience=5)`` of the reference (wakeword_training_script.py:228-230) with
    torch's defaults (threshold 1e-4 relative, cooldown 0, min_lr 0, eps 1e-8), acting on the trainer's host-side ``lr``
    (the learning rate is an argument of ``ww_train_apply``, there is no torch optimizer to mutate)."""

    def __init__(self, trainer, mode="max", factor=0.5, patience=5, threshold=1e-4, min_lr=0.0, eps=1e-8):
        assert mode in ("max", "min")
        self.trainer, self.mode, self.factor, self.patience = trainer, mode, factor, patience
        self.threshold, self.min_lr, self.eps = threshold, min_lr, eps
        self.best = -float("inf") if mode == "max" else float("inf")
        self.num_bad_epochs = 0
        self.last_epoch = 0

    def _is_better(self, a):
        if self.mode == "max":
            return a > self.best * (1.0 + self.threshold)
        return a < self.best * (1.0 - self.threshold)

    def step(self, metric):
        metric = float(metric)
        self.last_epoch += 1
        if self._is_better(metric):
            self.best, self.num_bad_epochs = metric, 0
        else:
            self.num_bad_epochs += 1
        if self.num_bad_epochs > self.patience:
            new_lr = max(self.trainer.lr * self.factor, self.min_lr)
            if self.trainer.lr - new_lr > self.eps:
                self.trainer.lr = new_lr
            self.num_bad_epochs = 0

    def state_dict(self):
        return {k: v for k, v in vars(self).items() if k != "trainer"}


class _AdamHandle:
    """What the reference reaches through ``trainer.optimizer``: ``state_dict()`` in torch.optim.Adam's format (so
    ``best_wakeword_model.pth['optimizer_state_dict']`` can resume an optimiser), ``load_state_dict`` and ``param_groups``.
    The moments themselves live in the ``ww_ctx`` while this trainer owns it (``WakewordTrainer._own``)."""

    def __init__(self, trainer):
        self._t = trainer

    @property
    def param_groups(self):
        t = self._t
        return [{"lr": t.lr, "betas": t.BETAS, "eps": t.EPS, "weight_decay": t.WEIGHT_DECAY, "amsgrad": False,
                 "params": list(range(len(list(t.model.state_dict()))))}]

    def state_dict(self):
        t = self._t
        names = list(t.model.state_dict().keys())
        state = {}
        st = t.export_state()
        if st is not None:
            for i, n in enumerate(names):
                state[i] = {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"][n].cpu(),
                            "exp_avg_sq": st["exp_avg_sq"][n].cpu()}
        return {"state": state, "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        t = self._t
        names = list(t.model.state_dict().keys())
        if sd.get("param_groups"):
            t.lr = float(sd["param_groups"][0].get("lr", t.lr))
        if sd.get("state"):
            step = int(max(float(v["step"]) for v in sd["state"].values()))
            t._saved = {"step": step,
                        "exp_avg": {names[int(i)]: v["exp_avg"].float() for i, v in sd["state"].items()},
                        "exp_avg_sq": {names[int(i)]: v["exp_avg_sq"].float() for i, v in sd["state"].items()}}
            t._needs_restore = True

    def zero_grad(self):
        pass                                    # ww_train_backward zeroes the flat gradient buffer itself


class WakewordTrainer:
    BETAS = (0.9, 0.999)
    EPS = 1e-8
    WEIGHT_DECAY = 1e-5          # wakeword_training_script.py:226

    def __init__(self, model, device, config=TrainingConfig):
        self.model = model
        self.device = torch.device(device)
        self.config = config
        self.lr = float(config.LEARNING_RATE)
        self.criterion = torch.nn.CrossEntropyLoss()
        self.optimizer = _AdamHandle(self)
        self.scheduler = _PlateauScheduler(self, mode="max", factor=0.5, patience=5)
        self.train_losses, self.val_losses, self.train_accuracies, self.val_accuracies = [], [], [], []
        self.patience = 10
        self.best_val_acc = 0.0
        self.epochs_no_improve = 0
        self.best_checkpoint_path = "best_wakeword_model.pth"        # where train() saves, like the reference
        self._grad_view = None
        self._eng = None                 # the engine whose context currently holds THIS trainer's Adam state
        self._saved = None               # Adam state parked on the torch side while another trainer owns the context
        self._needs_restore = False

    # ------------------------------------------------------------------ optimiser state ownership
    # The Adam moments live in the ww_ctx, and get_engine() shares one context per (device, configuration).  The
    # reference builds a fresh optim.Adam per trainer (:226), so each trainer must see its own state: a context records
    # its owning trainer; taking it over parks the previous owner's state on that trainer and installs ours (or zeros).
    def export_state(self):
        """-> {'step', 'exp_avg': {name: tensor}, 'exp_avg_sq': {...}} (device tensors) or None before the first step."""
        eng = self._eng
        if eng is None or getattr(eng, "_train_owner", None) is not self:
            return self._saved
        out = {"step": int(eng.lib.ww_train_get_step(eng._ctx)), "exp_avg": {}, "exp_avg_sq": {}}
        for name, p in self.model.state_dict().items():
            m = torch.empty(p.shape, device=eng.device, dtype=torch.float32)
            v = torch.empty_like(m)
            eng._chk(eng.lib.ww_train_get_moments(eng._ctx, name.encode(), C.c_void_p(m.data_ptr()),
                                                  C.c_void_p(v.data_ptr())), "ww_train_get_moments")
            out["exp_avg"][name], out["exp_avg_sq"][name] = m, v
        return out

    def _install_state(self, eng, st):
        eng._chk(eng.lib.ww_train_reset(eng._ctx), "ww_train_reset")
        if st is None:
            return
        for name in st["exp_avg"]:
            m = st["exp_avg"][name].to(eng.device, torch.float32).contiguous()
            v = st["exp_avg_sq"][name].to(eng.device, torch.float32).contiguous()
            eng._chk(eng.lib.ww_train_set_moments(eng._ctx, name.encode(), C.c_void_p(m.data_ptr()),
                                                  C.c_void_p(v.data_ptr())), "ww_train_set_moments")
        eng._chk(eng.lib.ww_train_set_step(eng._ctx, int(st["step"])), "ww_train_set_step")

    def _own(self, eng):
        owner = getattr(eng, "_train_owner", None)
        if owner is self and not self._needs_restore:
            return
        if not self._needs_restore and self._eng is not None and self._eng is not eng and \
                getattr(self._eng, "_train_owner", None) is self:
            self._saved = self.export_state()                 # our state sits in another context (other frame count)
            self._eng._train_owner = None
        if owner is not None and owner is not self:
            owner._saved = owner.export_state()               # park the previous owner's optimiser on its trainer
            owner._eng = None
        eng._train_owner = self
        self._install_state(eng, self._saved)
        self._eng, self._saved, self._needs_restore = eng, None, False

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
        self._own(eng)
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
        """wakeword_training_script.py:238-265.  An empty loader raises ZeroDivisionError like the reference."""
        self.model.train()
        running_loss, correct, total = 0.0, 0, 0
        for data, target in train_loader:
            loss, output = self.train_step(data, target.squeeze())
            running_loss += loss.item()
            predicted = output.argmax(dim=1)
            tgt = target.to(self.device).reshape(-1)
            total += tgt.numel()
            correct += int((predicted == tgt).sum().item())
        return running_loss / len(train_loader), 100.0 * correct / total

    def validate(self, val_loader):
        """wakeword_training_script.py:267-287."""
        self.model.eval()
        running_loss, correct, total = 0.0, 0, 0
        with torch.no_grad():
            for data, target in val_loader:
                data, tgt = data.to(self.device), target.to(self.device).reshape(-1)
                output = self.model(data)
                running_loss += self.criterion(output, tgt).item()
                correct += int((output.argmax(dim=1) == tgt).sum().item())
                total += tgt.numel()
        return running_loss / len(val_loader), 100.0 * correct / total

    def train(self, train_loader, val_loader, epochs):
        """The reference's epoch driver (wakeword_training_script.py:289-345): per epoch train + validate, plateau
        scheduler on the validation accuracy, best checkpoint (same dictionary keys, real Adam state), early stopping after
        ``self.patience`` epochs without improvement."""
        from .checkpoint import save_best_checkpoint
        print(f"Starting training for {epochs} epochs...")
        print(f"Using device: {self.device}")
        print(f"Learning rate: {self.config.LEARNING_RATE}")
        print(f"Batch size: {self.config.BATCH_SIZE}")
        self.best_val_acc = 0.0
        self.epochs_no_improve = 0
        for epoch in range(epochs):
            print(f"\nEpoch {epoch + 1}/{epochs}")
            print(f"GPU Memory: {torch.cuda.memory_allocated() / 1e6:.1f}MB allocated, "
                  f"{torch.cuda.memory_reserved() / 1e6:.1f}MB reserved")
            train_loss, train_acc = self.train_epoch(train_loader)
            val_loss, val_acc = self.validate(val_loader)
            self.train_losses.append(train_loss)
            self.val_losses.append(val_loss)
            self.train_accuracies.append(train_acc)
            self.val_accuracies.append(val_acc)
            print(f"Train Loss: {train_loss:.4f}, Train Acc: {train_acc:.2f}%")
            print(f"Val Loss: {val_loss:.4f}, Val Acc: {val_acc:.2f}%")
            self.scheduler.step(val_acc)
            if val_acc > self.best_val_acc:
                self.best_val_acc = val_acc
                self.epochs_no_improve = 0
                save_best_checkpoint(self.best_checkpoint_path, self.model, epoch, val_acc, train_acc, train_loss, val_loss,
                                     optimizer_state=self.optimizer.state_dict())
                print(f"New best model saved! Validation accuracy: {val_acc:.2f}%")
            else:
                self.epochs_no_improve += 1
            if self.epochs_no_improve >= self.patience:
                print(f"\nEarly stopping triggered! No improvement for {self.patience} epochs.")
                print(f"Best validation accuracy: {self.best_val_acc:.2f}%")
                break
        print("\nTraining completed!")
