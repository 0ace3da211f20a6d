"""Data-parallel plumbing for independent images: every image (seed) is an independent two-stage
trajectory, so images are sharded round-robin over ranks with NO collective inside the denoising loop.  The only
collectives are one broadcast of the weights from rank 0 at start-up and one all-gather of the final latents
(128 KiB per image).  torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests) is the transport."""
from typing import Dict, List

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """image j -> rank j mod world."""
    return list(range(rank, n_items, world))


def broadcast_state_dict(sd: Dict[str, torch.Tensor], src: int = 0) -> Dict[str, torch.Tensor]:
    """In-place broadcast of every tensor (same keys/shapes on every rank), deterministic key order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sd
    for k in sorted(sd):
        dist.broadcast(sd[k], src=src)
    return sd


def gather_latents(local: torch.Tensor, n_items: int) -> torch.Tensor:
    """local: (n_local, ...) latents of this rank's images in shard order -> (n_items, ...) in image order on every
    rank.  Ranks may hold different counts (n_items not divisible by world): shards are padded to the maximum."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_items + world - 1) // world
    pad = torch.zeros((per, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.empty((n_items, *local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        out[idx] = bufs[r][: len(idx)]
    return out
