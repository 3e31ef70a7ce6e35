"""Multi-GPU plumbing for the stream-parallel forward path.

Streams never interact (every op is per stream), so N GPUs run N independent shards of the
stream batch: weights replicated, per-shard caches resident on their GPU, NO collective on the
data path.  torch.distributed is used only for rendezvous, barriers, the max-over-ranks step time
and -- when a caller wants all posteriors in one place -- a concatenating gather.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def stream_slice(num_streams: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced partition [begin, end) of the stream batch for `rank` (the same formula the
    kernels use to spread streams over thread blocks)."""
    assert 0 <= rank < world
    return (num_streams * rank) // world, (num_streams * (rank + 1)) // world


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """Step time of the whole job = slowest rank."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_streams(local: torch.Tensor, num_streams: int) -> torch.Tensor:
    """Concatenates per-rank shards (dim 0 = streams) in rank order on every rank.  Shards may differ in
    size by one stream; they are padded to the largest for the all_gather and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [stream_slice(num_streams, r, world) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: e - b] for o, (b, e) in zip(out, sizes)], dim=0)
