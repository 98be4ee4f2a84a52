"""Batch sharding across the GPUs of one box: every clip / streaming window is independent,
so ranks take contiguous slices and there is NO collective on the scoring path.  The only communication is the
optional gather of the (tiny) results to rank 0, over whatever backend the process group uses (nccl or gloo)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` items for `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def window_shards(n_total_samples: int, n_samples: int, hop: int, world: int) -> List[Tuple[int, int, int, int]]:
    """Config 4: split the sliding windows into `world` contiguous ranges.
    Returns per rank (first_window, n_windows, first_sample, n_audio_samples): consecutive ranks overlap by
    n_samples - hop raw samples so that every window is scored by exactly one rank."""
    n_win = 0 if n_total_samples < n_samples else 1 + (n_total_samples - n_samples) // hop
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_win, world, r)
        first = lo * hop
        n_audio = 0 if hi == lo else (hi - lo - 1) * hop + n_samples
        out.append((lo, hi - lo, first, n_audio))
    return out


def gather_to_rank0(local: torch.Tensor, total: int):
    """Gather per-rank result slices (shard_bounds order) on rank 0; returns the full tensor there, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, parts, dst=0)
    if rank != 0:
        return None
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


# ---------------------------------------------------------------------------------------------------------------
# data-parallel training (config 5): the flat gradient buffer of the C ABI and its all-reduce
def flat_layout(shapes) -> Tuple[dict, int]:
    """Offsets of the parameters inside the flat fp32 gradient buffer of ``ww_train_grad_buffer``: state_dict order,
    every entry padded to a multiple of 4 floats.  shapes: ordered mapping name -> shape.  Returns ({name: (offset,
    count)}, total floats)."""
    out, off = {}, 0
    for name, shape in shapes.items():
        cnt = 1
        for d in shape:
            cnt *= int(d)
        out[name] = (off, cnt)
        off += (cnt + 3) & ~3
    return out, off


def allreduce_mean_(flat: torch.Tensor) -> float:
    """Sum-all-reduce the flat gradient buffer in place and return the scale (1 / world_size) the optimiser applies
    (``ww_train_apply(grad_scale=...)``).  The package's only collective; a no-op without a process group."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 1.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return 1.0 / dist.get_world_size()
