"""CPU tests of the N > 1 host logic with a real world_size-2 gloo process group (no GPU needed)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wakeword_jupyterlab_b200.sharding import gather_to_rank0, shard_bounds, window_shards


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
