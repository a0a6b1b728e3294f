"""Multi-GPU plumbing: one process per GPU (torch.distributed over NCCL/NVLink; gloo in CPU tests).

The path partitions over independent IMAGES (each image -> T views): ranks take whole images, there is no
data-path collective, and the only exchange is the final decoded-frame gather (uint8 THWC, ~14 MB per image),
as BASELINE.json's north_star names.  View-sharding ONE image across ranks needs three exchanges per temporal
layer (K/V all-gather, 1-frame conv halos, 3-D GroupNorm statistics) and is the next row.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend: Optional[str] = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_images(n_images: int, rank: int, world: int) -> List[int]:
    """Contiguous block partition of image indices; earlier ranks take the remainder."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def gather_frames(frames: torch.Tensor, counts: List[int]) -> Optional[torch.Tensor]:
    """All ranks contribute [n_local, T, H, W, 3] uint8 frames; every rank receives [sum(counts), T, H, W, 3]
    in image order. Uneven counts are padded to max(counts) for the fixed-size all-gather."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return frames
    assert len(counts) == world
    nmax = max(counts)
    pad = frames.new_zeros((nmax,) + tuple(frames.shape[1:]))
    pad[: frames.shape[0]] = frames
    out = frames.new_empty((world * nmax,) + tuple(frames.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous())
    out = out.reshape(world, nmax, *frames.shape[1:])
    return torch.cat([out[r, : counts[r]] for r in range(world)], dim=0)


def max_over_ranks(seconds: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
