"""CPU tests of the N > 1 host logic with a real world_size-2 gloo process group (no GPU needed)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wakeword_jupyterlab_b200.sharding import (allreduce_mean_, flat_layout, gather_to_rank0, shard_bounds,
                                               window_shards)


def test_shard_bounds_cover_exactly_once():
    for total in (0, 1, 7, 65536, 65537, 359901):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_window_shards_match_config4():
    T, N, hop = 57_600_000, 16000, 160                     # 1 h at 16 kHz, 1 s windows, 10 ms hop
    assert 1 + (T - N) // hop == 359_901
    for world in (1, 2, 8):
        sh = window_shards(T, N, hop, world)
        assert sum(s[1] for s in sh) == 359_901
        for (w0, nw, s0, ns) in sh:
            assert s0 == w0 * hop and ns == (nw - 1) * hop + N and s0 + ns <= T
        for a, b in zip(sh, sh[1:]):
            assert a[0] + a[1] == b[0]
            assert (a[2] + a[3]) - b[2] == N - hop           # overlap of raw audio between neighbours


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(total, world, rank)
        # stand-in "scores": a deterministic function of the global clip index (the GPU path is tested with -m gpu)
        idx = torch.arange(lo, hi, dtype=torch.float32)
        local = torch.stack([idx * 2.0, idx * 2.0 + 1.0], dim=1)
        full = gather_to_rank0(local, total)
        if rank == 0:
            want = torch.arange(total, dtype=torch.float32)
            ok = full.shape == (total, 2) and torch.equal(full[:, 0], want * 2) and torch.equal(full[:, 1], want * 2 + 1)
            q.put(bool(ok))
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_gather_two_ranks_gloo(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


# ------------------------------------------------------------------ data-parallel training step (config 5)
def _ref_like_grads(sd, x, y):
    """Gradients of the mean cross entropy of the oracle's torch-CPU model (the reference's forward restated in
    oracle.model.forward_torch_cpu) with respect to every parameter that receives a data gradient."""
    from oracle import model as M
    params = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        feats = torch.from_numpy(x)
        # forward_torch_cpu runs under no_grad for timing; restate it here with autograd enabled
        import torch.nn.functional as F
        h = F.relu(F.conv2d(feats, params["conv1.weight"], params["conv1.bias"], padding=1))
        h = F.relu(F.conv2d(h, params["conv2.weight"], params["conv2.bias"], padding=1))
        h = F.relu(F.conv2d(h, params["conv3.weight"], params["conv3.bias"], padding=1)).mean(dim=(2, 3))
        layer = 0
        while f"lstm.weight_ih_l{layer}" in params:
            g = F.linear(h, params[f"lstm.weight_ih_l{layer}"], params[f"lstm.bias_ih_l{layer}"] + params[f"lstm.bias_hh_l{layer}"])
            H = g.shape[1] // 4
            c = torch.sigmoid(g[:, :H]) * torch.tanh(g[:, 2 * H:3 * H])
            h = torch.sigmoid(g[:, 3 * H:]) * torch.tanh(c)
            layer += 1
        loss = F.cross_entropy(F.linear(h, params["fc.weight"], params["fc.bias"]), torch.from_numpy(y))
    loss.backward()
    return {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}


def _pack(grads, layout, total):
    flat = torch.zeros(total)
    for name, (off, cnt) in layout.items():
        flat[off:off + cnt] = grads[name].reshape(-1)
    return flat


def _train_worker(rank, world, port, q):
    import numpy as np
    from oracle import recipe as R
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        sd = R.seeded_state_dict(32, seed=5)
        rng = np.random.default_rng(3)
        B = 8
        x = (rng.standard_normal((B, 1, 80, 32)) * 15 - 40).astype(np.float32)
        y = rng.integers(0, 2, B).astype(np.int64)
        layout, total = flat_layout({k: v.shape for k, v in sd.items()})
        lo, hi = shard_bounds(B, world, rank)
        flat = _pack(_ref_like_grads(sd, x[lo:hi], y[lo:hi]), layout, total)       # this rank's ww_train_backward
        scale = allreduce_mean_(flat)                                                # NCCL on the GPU box, gloo here
        flat *= scale                                                                # ww_train_apply(grad_scale)
        if rank == 0:
            full = _pack(_ref_like_grads(sd, x, y), layout, total)                   # one process, whole batch
            q.put(float((flat - full).abs().max() / full.abs().max()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_flat_layout_matches_the_abi_rule():
    layout, total = flat_layout({"a": (3, 3), "b": (4,), "c": (2, 1, 3, 3)})
    assert layout == {"a": (0, 9), "b": (12, 4), "c": (16, 18)} and total == 36


def test_data_parallel_gradients_two_ranks_gloo():
    """Equal shards: the mean of the per-rank mean-loss gradients is the full-batch gradient."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=10) < 1e-5
