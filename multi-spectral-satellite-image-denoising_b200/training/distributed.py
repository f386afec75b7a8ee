"""Data-parallel plumbing of the training step (BASELINE config C4).  The reference has no distributed code
(its README only suggests nn.DataParallel); semantics = identical replicas, micro-batch per GPU, per-GPU BatchNorm
statistics, gradients averaged before the optimizer step.

Product path -- ``SatMAERestoration.enable_data_parallel()`` + ``reduce_gradient_groups``: the library's backward writes every
gradient into one flat fp32 arena in completion order and records a CUDA event per gradient group (decoder, then the encoder
blocks last to first, then the patch embedding).  Each group is all-reduced IN PLACE on a communication stream that waits for
the group's event only, so NCCL over NVLink / NVSwitch runs underneath the rest of the backward: no concatenation, no copy
back, and only the last (small) group is exposed.  The global-norm clip that follows (trainer.py:188-193) sees identical
averaged gradients on every rank and needs no further collective.

``allreduce_gradients`` is the generic after-the-fact fallback for parameters that did not come out of that backward
(bucketed, not overlapped)."""
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


_comm_streams = {}


def reduce_gradient_groups(arena: torch.Tensor, bounds, group_ids, group=None, wait_fn=None) -> int:
    """Average the slices ``arena[lo:hi]`` (one per gradient group, in completion order) over the process group, in place.

    CUDA tensors: every all-reduce is enqueued on a per-device communication stream after ``wait_fn(group_id, stream_ptr)`` made
    that stream wait for the group's gradients (``msd_model_grad_group_wait``); the current stream waits for the communication
    stream at the end.  CPU tensors (gloo, the unit test of this logic): plain in-order all-reduces.  Returns the number of
    collectives issued."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world <= 1:
        return 0
    avg = dist.get_backend(group) == "nccl"            # ReduceOp.AVG exists on NCCL only
    n = 0
    if arena.is_cuda:
        dev = arena.device
        comm = _comm_streams.get(dev)
        if comm is None:
            comm = _comm_streams[dev] = torch.cuda.Stream(device=dev)
        arena.record_stream(comm)
        main = torch.cuda.current_stream(dev)
        for (lo, hi), gid in zip(bounds, group_ids):
            if hi <= lo:
                continue
            if wait_fn is not None:
                wait_fn(gid, comm.cuda_stream)
            else:
                comm.wait_stream(main)
            with torch.cuda.stream(comm):
                piece = arena[lo:hi]
                dist.all_reduce(piece, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=group)
                if not avg:
                    piece.mul_(1.0 / world)
            n += 1
        main.wait_stream(comm)
        return n
    for lo, hi in bounds:
        if hi <= lo:
            continue
        piece = arena[lo:hi]
        dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=group)
        piece.mul_(1.0 / world)
        n += 1
    return n
