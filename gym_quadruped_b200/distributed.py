"""Env sharding across GPUs: contiguous env-index ranges per rank, no data-path collective; the
only optional collective is an all-gather of the per-rank observation rows (NCCL over NVLink on GPUs, gloo in CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(rank: int, world_size: int, total_envs: int) -> tuple[int, int]:
    """[start, stop) of the global env ids owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total_envs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather equally sized [n_local, D] row blocks into [world * n_local, D], rank-major (global env order)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out
