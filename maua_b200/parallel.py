"""Frame sharding across the GPUs of one box: one process per GPU, frames are the shard unit.

The reference renders on a single device (maua/GAN/wrappers/__init__.py:52-99); its only data-parallel
precedent is maua/super/image/bulk.py:31-109 (DistributedSampler + a writer queue).  Here every rank
renders a contiguous frame range; generator weights and the [T,...] input tensors are broadcast once,
finished uint8 frames are gathered to the root (or streamed by each rank).  No per-step collective.
Works with backend "nccl" (GPU) and "gloo" (CPU tests of the host logic).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def frame_range(rank: int, world: int, n_frames: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first (n_frames % world) ranks get one extra frame."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_module_(module: torch.nn.Module, src: int = 0) -> None:
    """Broadcast every parameter and buffer from `src` (weights broadcast once)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


def shard_inputs(inputs: Dict[str, torch.Tensor], src: int = 0) -> Tuple[Dict[str, torch.Tensor], Tuple[int, int]]:
    """Broadcast the [T,...] input tensors from `src` and return this rank's contiguous slice."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        T = len(next(iter(inputs.values())))
        return inputs, (0, T)
    rank, world = dist.get_rank(), dist.get_world_size()
    out = {}
    rng = None
    for k in sorted(inputs.keys()):
        t = inputs[k].contiguous()
        dist.broadcast(t, src=src)
        rng = frame_range(rank, world, t.shape[0])
        out[k] = t[rng[0]:rng[1]]
    return out, rng


def gather_frames(local: torch.Tensor, n_frames: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank frame blocks [n_local, ...] (uint8 NHWC in the render loop) to `dst` in frame order.
    Ranks may hold different counts (see frame_range); returns the full [n_frames, ...] tensor on dst."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    rank, world = dist.get_rank(), dist.get_world_size()
    counts = [frame_range(r, world, n_frames) for r in range(world)]
    max_n = max(e - s for s, e in counts)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: e - s] for r, (s, e) in enumerate(counts)], dim=0)
