"""Synchronised batch-norm statistics across the data-parallel ranks (north_star: one of the three NCCL uses).

The reference converts its model with ``torch.nn.SyncBatchNorm.convert_sync_batchnorm`` before wrapping it in DDP
(main.py:189,400); the only BatchNorm layers of the repository are the two ``BatchNorm1d(mmhid)`` of ``BilinearFusion``
(models/fusion.py:29,31, ``fusion_type: pofusion``) on ``[B, mmhid]`` vectors.  The exchange is tiny - per layer and step one
all_gather of (mean, biased variance, count) = 2 C + 1 floats per rank forward, one all_reduce of (sum dy, sum dy (x - mean))
backward - so it is plain ``torch.distributed`` on the job's process group (NCCL over NVLink on the GPUs, gloo in the CPU
tests); unequal per-rank batch sizes are handled (counts ride along), as ``torch.nn.SyncBatchNorm`` does.

Note (reference quirk): ``DeformPathomicNet`` itself cannot run with ``fusion_type != 'concat'`` - its forward uses
``self.classifier_tumor``, which that branch of ``__init__`` never creates (models/model.py:490-505,542) - so this layer serves
the reference's other BilinearFusion users (PathomicNet*, model.py:303,398) through ``convert_sync_batchnorm``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn


def _world(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


class SyncBatchNormFn(torch.autograd.Function):
    """y = (x - mean) / sqrt(var + eps) * w + b over the GLOBAL batch (all ranks of `group`); x [B, C] or [B, C, L]."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, group):
        red = [0] + list(range(2, x.dim()))
        C = x.shape[1]
        cnt = x.numel() // C
        xf = x.float()
        mean_l = xf.mean(red)
        var_l = xf.var(red, unbiased=False) if cnt > 1 else torch.zeros_like(mean_l)
        world = _world(group)
        if world > 1:
            packed = torch.cat([mean_l, var_l, torch.full((1,), float(cnt), device=x.device)])
            parts = [torch.empty_like(packed) for _ in range(world)]
            dist.all_gather(parts, packed, group=group)                        # 2 C + 1 floats per rank
            allp = torch.stack(parts)
            means, vars_, counts = allp[:, :C], allp[:, C:2 * C], allp[:, 2 * C:]
            total = counts.sum()
            mean = (means * counts).sum(0) / total
            var = ((vars_ + (means - mean) ** 2) * counts).sum(0) / total      # parallel-variance combination
        else:
            total = torch.tensor(float(cnt), device=x.device)
            mean, var = mean_l, var_l
        invstd = torch.rsqrt(var + eps)
        shape = [1, C] + [1] * (x.dim() - 2)
        xhat = (xf - mean.view(shape)) * invstd.view(shape)
        y = xhat * weight.view(shape) + bias.view(shape) if weight is not None else xhat
        ctx.save_for_backward(xhat, weight, invstd, total)
        ctx.group, ctx.red, ctx.shape = group, red, shape
        ctx.mark_non_differentiable(mean, var, total)
        return y.to(x.dtype), mean, var, total

    @staticmethod
    def backward(ctx, dy, _dm, _dv, _dt):
        xhat, weight, invstd, total = ctx.saved_tensors
        red, shape = ctx.red, ctx.shape
        dyf = dy.float()
        sum_dy = dyf.sum(red)
        sum_dy_xhat = (dyf * xhat).sum(red)
        dw = sum_dy_xhat if weight is not None else None        # LOCAL parameter gradients: the job's gradient all-reduce averages them
        db = sum_dy if weight is not None else None
        if _world(ctx.group) > 1:
            packed = torch.cat([sum_dy, sum_dy_xhat])
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=ctx.group)     # 2 C floats
            C = sum_dy.numel()
            sum_dy, sum_dy_xhat = packed[:C], packed[C:]
        w = weight.view(shape) if weight is not None else 1.0
        dx = (dyf - sum_dy.view(shape) / total - xhat * (sum_dy_xhat.view(shape) / total)) * invstd.view(shape) * w
        return dx.to(dy.dtype), dw, db, None, None


class SyncBatchNorm1d(nn.BatchNorm1d):
    """nn.BatchNorm1d whose training statistics span every rank of `process_group` (same parameters, buffers and
    state_dict keys, so BatchNorm1d checkpoints load unchanged).  Eval mode and world size 1 behave like BatchNorm1d."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, process_group=None):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        self.process_group = process_group

    def forward(self, x):
        if not self.training and self.track_running_stats:
            return super().forward(x)
        y, mean, var, total = SyncBatchNormFn.apply(x, self.weight, self.bias, self.eps, self.process_group)
        if self.training and self.track_running_stats:
            with torch.no_grad():
                self.num_batches_tracked += 1
                mom = self.momentum if self.momentum is not None else 1.0 / float(self.num_batches_tracked)
                unbiased = var * (total / (total - 1).clamp_min(1.0))
                self.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
                self.running_var.mul_(1 - mom).add_(unbiased * mom)
        return y


def convert_sync_batchnorm(module: nn.Module, process_group=None) -> nn.Module:
    """Replace every nn.BatchNorm1d below `module` by SyncBatchNorm1d (parameters and buffers are shared, not copied) -
    the role torch.nn.SyncBatchNorm.convert_sync_batchnorm plays at main.py:189,400."""
    if isinstance(module, nn.BatchNorm1d) and not isinstance(module, SyncBatchNorm1d):
        new = SyncBatchNorm1d(module.num_features, module.eps, module.momentum, module.affine, module.track_running_stats,
                              process_group)
        if module.affine:
            new.weight, new.bias = module.weight, module.bias
        if module.track_running_stats:
            new.running_mean, new.running_var, new.num_batches_tracked = (module.running_mean, module.running_var,
                                                                          module.num_batches_tracked)
        new.train(module.training)
        return new
    for name, child in module.named_children():
        setattr(module, name, convert_sync_batchnorm(child, process_group))
    return module
