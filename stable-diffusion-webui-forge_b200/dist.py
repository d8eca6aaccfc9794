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


def gather_to_rank0(t: torch.Tensor, group=None, bufs: Optional[List[torch.Tensor]] = None) -> Optional[torch.Tensor]:
    """Gather equally-shaped per-rank results; returns the concatenation [world*B, ...] on rank 0, None elsewhere.
    `bufs` (rank 0): preallocated receive buffers reused across jobs; then the list itself is returned un-concatenated."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    own = bufs is None
    if rank == 0 and own:
        bufs = [torch.empty_like(t) for _ in range(world)]
    dist.gather(t, bufs if rank == 0 else None, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat(bufs, 0) if own else bufs


def gather_images_u8(img: torch.Tensor, group=None, bufs: Optional[List[torch.Tensor]] = None):
    """The result gather of the multi-GPU path: fp32 images in [0, 1] are converted to uint8 on the device (the conversion
    the reference does on the host, modules/processing.py:1039-1040) and gathered to rank 0 — a quarter of the fp32 bytes."""
    from . import ops
    return gather_to_rank0(ops.images_to_u8(img), group, bufs)
