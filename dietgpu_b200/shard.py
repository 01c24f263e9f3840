"""Multi-GPU plumbing: batch members are independent (own statistics, own archive), so the path
shards by member with NO data-path collective; the only exchange is the vector of archive sizes."""
from __future__ import annotations

from typing import List, Sequence


def shard_range(num_members: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of [0, num_members) owned by `rank` (first ranks take the remainder)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(num_members, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_members(members: Sequence, rank: int, world: int) -> List:
    r = shard_range(len(members), rank, world)
    return [members[i] for i in r]


def gather_sizes(local_sizes, num_members: int, group=None):
    """All ranks learn every member's archive size.  local_sizes: int32 tensor of this rank's shard
    (any device the process group supports).  Returns an int32 tensor [num_members] in member order."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    counts = [len(shard_range(num_members, r, world)) for r in range(world)]
    width = max(counts) if counts else 0
    padded = torch.zeros(width, dtype=torch.int32, device=local_sizes.device)
    padded[: local_sizes.numel()] = local_sizes.to(torch.int32)
    out = torch.empty(world * width, dtype=torch.int32, device=local_sizes.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = [out[r * width: r * width + counts[r]] for r in range(world)]
    return torch.cat(parts) if parts else out
