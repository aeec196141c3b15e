"""Multi-rank host logic on CPU: world_size 2 over gloo (the N>1 path of bench.py / dist.py)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolo2_b200.dist import gather_frames, shard_bounds, shard_counts


@pytest.mark.parametrize("n,world", [(1024, 8), (10, 4), (3, 8), (0, 2), (7, 1), (129, 2)])
def test_shard_bounds_partition(n, world):
    spans = [shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))           # contiguous, no gaps, no overlap
    c = shard_counts(n, world)
    assert sum(c) == n and max(c) - min(c) <= 1


def test_shard_bounds_rejects_bad_rank():
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _worker(rank, world, port, n_frames, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_bounds(n_frames, world, rank)
    # stand-in for the per-frame region tensor: row f is a function of the global frame index only
    local = torch.stack([torch.full((6,), float(f)) + torch.arange(6) for f in range(s, e)]) if e > s else torch.zeros((0, 6))
    full = gather_frames(local, n_frames, dst=0)
    if rank == 0:
        want = torch.stack([torch.full((6,), float(f)) + torch.arange(6) for f in range(n_frames)])
        out.put(bool(torch.equal(full, want)))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [8, 5])
def test_gather_two_ranks_gloo(n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + n_frames
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
