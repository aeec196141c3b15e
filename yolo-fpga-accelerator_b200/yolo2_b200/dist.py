"""Frame sharding across the GPUs of one box and the final gather of results.

The path shards naturally by frame: frames are independent, weights are
replicated, every rank runs the full network on its contiguous slice and there is NO collective
on the data path.  The only communication is the gather of the fixed-size per-frame region
tensors (or detections) at the end - NCCL on GPUs, gloo in the CPU tests.
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [start, end) of frames for `rank`; the first n_frames % world ranks get one more."""
    if world <= 0 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_counts(n_frames: int, world: int) -> List[int]:
    return [shard_bounds(n_frames, world, r)[1] - shard_bounds(n_frames, world, r)[0] for r in range(world)]


def gather_frames(local: torch.Tensor, n_frames: int, dst: int = 0):
    """Gathers every rank's [count_r, ...] result rows to `dst` in frame order. Returns the
    [n_frames, ...] tensor on `dst`, None elsewhere. Ragged shards are padded to the largest."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = shard_counts(n_frames, world)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} frames, expected {counts[rank]}")
    most = max(counts)
    padded = local
    if local.shape[0] < most:
        pad = torch.zeros((most - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], 0)
    padded = padded.contiguous()
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)
