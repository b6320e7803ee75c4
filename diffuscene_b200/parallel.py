"""Multi-GPU host logic: scenes shard over ranks.  Sampling has no data-path collective; data-parallel training
has exactly one (the gradient mean).

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


def allreduce_gradients(params, bucket_bytes: int = 64 << 20) -> int:
    """Data-parallel training: average the gradients over all ranks, in place.

    Gradients are packed into flat buckets of ~`bucket_bytes` (the whole model is 311 MB fp32, i.e. five buckets:
    NCCL sees a few NVLink-sized messages instead of 334 small ones) and reduced with one all-reduce per bucket;
    each rank must hold the same per-rank batch size so that the mean of means is the global mean.  Gradient
    clipping has to come AFTER this call (the clip uses the norm of the averaged gradient, like the reference's
    single-process `clip_grad_norm_`).  Returns the number of collectives issued (0 outside a process group).
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    n_coll = 0
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size, n_coll
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_coll += 1
        bucket, size = [], 0

    for g in grads:
        if bucket and (g.dtype != bucket[0].dtype or size + g.numel() * g.element_size() > bucket_bytes):
            flush()
        bucket.append(g)
        size += g.numel() * g.element_size()
    flush()
    return n_coll


def allreduce_flat(flat_grads: torch.Tensor, other_params=(), bucket_bytes: int = 64 << 20) -> int:
    """Data-parallel gradient all-reduce of the native training path: the denoiser's gradients already ARE one flat
    fp32 buffer (diffuscene_b200.engine.DenoiserEngine.flat_layout), so each ~64 MB bucket is a view -- no packing, no
    copy back.  The native backward pre-scales its gradients by 1 / world_size (ds_train_step grad_scale), so a SUM
    all-reduce yields the mean.  `other_params`: the few parameters outside the denoiser (positional embedding,
    condition MLPs), reduced together in one small extra bucket.  Returns the number of collectives issued."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    n_coll = 0
    step = max(1, bucket_bytes // 4)
    handles = []
    for off in range(0, flat_grads.numel() if flat_grads is not None else 0, step):
        handles.append(dist.all_reduce(flat_grads[off:off + step], op=dist.ReduceOp.SUM, async_op=True))
        n_coll += 1
    others = [p.grad for p in other_params if p.grad is not None]
    if others:
        buf = torch.cat([g.reshape(-1) for g in others])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)      # already scaled by 1 / world through d(context)
        off = 0
        for g in others:
            g.copy_(buf[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_coll += 1
    for hd in handles:
        hd.wait()
    return n_coll


class GradOverlap:
    """Bucketed gradient all-reduce overlapped with the native backward pass.

    The flat gradient buffer is cut into ~64 MB buckets.  ds_train_step records one CUDA event per bucket as soon as
    every gradient in it is final (the backward pass finalises the buffer from its end towards its start).  launch()
    is called right after ds_train_step returned -- its kernels are enqueued, not finished -- and enqueues, on a side
    stream, `wait(event_k)` + all-reduce(bucket k) for every bucket, last bucket first, so NCCL works on the finished
    tail of the buffer while the GPU still computes the gradients of the earlier layers.  wait() makes the current
    stream wait for all of them (called before the gradients are consumed)."""

    def __init__(self, engine, flat_grads: torch.Tensor, bucket_bytes: int = 64 << 20):
        self.flat = flat_grads
        n = flat_grads.numel()
        step = max(1, bucket_bytes // 4)
        self.bounds = list(range(0, n, step)) + [n]
        self.events = [torch.cuda.Event() for _ in range(len(self.bounds) - 1)]
        for e in self.events:
            e.record()                               # materialise the cudaEvent_t handles
        engine.train_set_buckets(self.bounds, self.events)
        self.stream = torch.cuda.Stream(device=flat_grads.device)
        self.works = []

    def launch(self):
        self.works = []
        with torch.cuda.stream(self.stream):
            for k in reversed(range(len(self.events))):
                self.stream.wait_event(self.events[k])
                self.works.append(dist.all_reduce(self.flat[self.bounds[k]:self.bounds[k + 1]], op=dist.ReduceOp.SUM, async_op=True))

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []
        torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)
