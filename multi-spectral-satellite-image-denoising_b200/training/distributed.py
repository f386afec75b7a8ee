"""Data-parallel plumbing of the training step (BASELINE config C4).  The reference has no distributed code
(its README only suggests nn.DataParallel); semantics = identical replicas, micro-batch per GPU, per-GPU BatchNorm
statistics, gradients averaged before the optimizer step.

The model's backward hands back every gradient at once (one library call), so there is nothing to overlap bucket by
bucket: the gradients are packed into a few large fp32 buckets in reverse-layer order (decoder first) and each bucket
is all-reduced once over NCCL / NVLink; ~54 M trainable parameters = 215 MB, a fraction of a millisecond on NVSwitch.
The global-norm clip that follows (trainer.py:188-193) then sees identical averaged gradients on every rank and needs
no further collective."""
from __future__ import annotations

from typing import Iterable, List

import torch


def _buckets(grads: List[torch.Tensor], bucket_bytes: int) -> List[List[torch.Tensor]]:
    out, cur, size = [], [], 0
    for g in grads:
        cur.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            out.append(cur)
            cur, size = [], 0
    if cur:
        out.append(cur)
    return out


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world_size: int | None = None, bucket_mb: int = 64,
                        group=None) -> int:
    """Average `.grad` of every parameter across the process group, in place.  Returns the number of collectives issued.
    Parameters without a gradient (frozen blocks) are skipped on every rank alike."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = world_size or dist.get_world_size(group)
    if world <= 1:
        return 0
    grads = [p.grad for p in reversed(list(params)) if p.grad is not None]
    n = 0
    for bucket in _buckets(grads, bucket_mb << 20):
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n += 1
    return n
