"""Drop-in mirror of the reference ``utils/gather.py`` (GatherLayer, :5-20): all_gather with autograd.

Forward: every rank contributes its tensor and receives the tuple of all ranks' tensors (callers:
utils/loss.py:37-38, 102-105, 232-233, then ``torch.cat(..., dim=0)``).  Backward keeps ONLY the local slice of
the incoming gradients - no cross-rank reduction (quirk Q17: not the mathematically complete gradient; reproduced).
One collective into one preallocated [world, ...] buffer (NCCL all_gather_into_tensor on GPUs, gloo on CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GatherLayer(torch.autograd.Function):
    """Gather tensors from all processes, supporting backward propagation."""

    @staticmethod
    def forward(ctx, input):
        world, rank = dist.get_world_size(), dist.get_rank()
        ctx.rank = rank
        ctx.shape = input.shape
        src = input.contiguous()
        buf = torch.empty((world,) + tuple(src.shape), dtype=src.dtype, device=src.device)
        if src.is_cuda:
            dist.all_gather_into_tensor(buf, src)
        else:   # gloo has no all_gather_into_tensor on every build: list form, same result
            dist.all_gather(list(buf.unbind(0)), src)
        return tuple(buf.unbind(0))

    @staticmethod
    def backward(ctx, *grads):
        g = grads[ctx.rank]
        if g is None:
            return None
        return g.reshape(ctx.shape).clone()
