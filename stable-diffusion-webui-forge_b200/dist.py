"""Multi-GPU layer: replica data parallelism over independent txt2img requests.

One process per GPU (torchrun), each holding a full UNet + VAE replica; a global batch of `per_rank * world` images is
sharded contiguous-by-seed so that rank r generates seeds [base + r*per_rank, base + (r+1)*per_rank) — results do not
depend on the number of ranks.  There is no collective on the data path; the finished images are gathered to rank 0
with one NCCL (or, on CPU test rigs, gloo) gather.
"""
from __future__ import annotations

from typing import List, Optional

import torch


def shard_seeds(base_seed: int, per_rank: int, rank: int, world: int) -> List[int]:
    assert 0 <= rank < world
    return [base_seed + rank * per_rank + i for i in range(per_rank)]


def gather_to_rank0(t: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """Gather equally-shaped per-rank results; returns the concatenation [world*B, ...] on rank 0, None elsewhere."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, bufs, dst=0, group=group)
    return torch.cat(bufs, 0) if rank == 0 else None
