"""Multi-GPU host logic for the sampling path: scenes shard over ranks, no data-path collective.

One process per GPU (torch.distributed, NCCL on GPUs / gloo in the CPU tests).  The only communication is
control-plane: the max-over-ranks of the device-measured time, and an optional gather of finished scenes.
Per-scene Philox streams are keyed by the GLOBAL scene index (`scene_offset`), so the union of all shards is
bit-identical to a single-GPU run of the whole batch.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [offset, offset + count) of `total` scenes owned by `rank`; remainders go to the first ranks."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def max_over_ranks(value: float, device: torch.device) -> float:
    """Max of a per-rank scalar (elapsed device milliseconds) over all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_scenes(local: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather variable-sized scene shards [count_r, N, d] into [total, N, d] in global scene order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = [shard_range(total, r, world)[1] for r in range(world)]
    pad = max(counts)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out: List[torch.Tensor] = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
