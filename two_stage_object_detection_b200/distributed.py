"""Multi-GPU plumbing for the path: images are independent, so a batch is split contiguously over the
ranks (one process per GPU) and nothing is exchanged on the data path; the only collective is one
all-gather of the fixed-shape detections at the end of a step (NCCL over NVLink on GPUs; any
torch.distributed backend works, which is how the CPU tests cover it with gloo)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split: rank r owns images [lo, hi); the first n_images % world ranks get one more."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    q, r = divmod(n_images, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def all_gather_detections(rois: torch.Tensor, n_keep: Optional[torch.Tensor] = None, group=None):
    """rois [B_local, n_post, 4] (+ n_keep [B_local]) from every rank -> ([B_total, n_post, 4], [B_total]).
    Ranks may hold different B_local (uneven split): shorter shards are padded for the collective and
    trimmed afterwards.  Without an initialised process group this is the identity."""
    if not (dist.is_available() and dist.is_initialized()):
        return rois, n_keep
    world = dist.get_world_size(group)
    if world == 1:
        return rois, n_keep
    b_local = torch.tensor([rois.shape[0]], dtype=torch.int64, device=rois.device)
    sizes = [torch.zeros_like(b_local) for _ in range(world)]
    dist.all_gather(sizes, b_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    b_max = max(sizes)

    def gather(t):
        pad = torch.zeros((b_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad, group=group)
        return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)

    return gather(rois), (gather(n_keep) if n_keep is not None else None)
