"""Image-tile split over N ranks (one process per GPU) and the end-of-frame gather.

The path has no exchange step: a pixel's primary and shadow rays stay on the rank that owns its tile, the
scene is replicated.  Tiles of the row-major tile grid are dealt round-robin (tile k -> rank k % N), which
balances load far better than the reference's contiguous row bands (main.cpp:371-393).  At the end of a
frame every rank's compact result block is gathered on rank 0 (NCCL over NVLink; gloo in the CPU tests)
and re-assembled into the row-major frame -- on the GPU by ``dodrt_frame_assemble_device``, in the CPU
tests by ``assemble_host`` with the pixel maps of ``dodrt_frame_pixel_map``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import capi


def rank_frame(width: int, height: int, classes: int, rank: int, world: int, tile=(32, 32)) -> capi.Frame:
    """dodrt_frame of rank `rank` in a `world`-way tile split (compact results when world > 1)."""
    return capi.Frame.make(width, height, classes=classes, tile=tile, first_tile=rank, tile_stride=world,
                           compact=1 if world > 1 else 0)


def slots_per_rank(width: int, height: int, world: int, tile=(32, 32)) -> int:
    """Result slots of rank 0 = the largest block; every rank pads to it so the gather is regular."""
    return capi.frame_local_pixels(capi.Frame.make(width, height, tile=tile, first_tile=0, tile_stride=world, compact=1))


def gather_to_rank0(local, world: int, rank: int, out=None):
    """torch.distributed gather of equally sized per-rank blocks to rank 0.  `local` is a torch tensor on the
    backend's device (cuda for nccl, cpu for gloo); returns [world, *local.shape] on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local.unsqueeze(0)
    if rank == 0:
        if out is None:
            out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        dist.gather(local, list(out.unbind(0)), dst=0)
        return out
    dist.gather(local, None, dst=0)
    return None


def assemble_host(width: int, height: int, world: int, tile, gathered_hits: np.ndarray,
                  gathered_vis: Optional[np.ndarray] = None) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """Host-side twin of dodrt_frame_assemble_device (pure indexing, used by tests and small tools):
    gathered_hits [world, slots_per_rank] dodrt_hit, gathered_vis [world, slots_per_rank] uint8."""
    hits = np.zeros(width * height, capi.HIT_DT)
    vis = np.zeros(width * height, np.uint8) if gathered_vis is not None else None
    for r in range(world):
        f = capi.Frame.make(width, height, tile=tile, first_tile=r, tile_stride=world, compact=1)
        m = capi.frame_pixel_map(f)
        ok = m != 0xFFFFFFFF
        hits[m[ok]] = gathered_hits[r][: len(m)][ok]
        if vis is not None:
            vis[m[ok]] = gathered_vis[r][: len(m)][ok]
    return hits, vis
