"""Drop-in mirror of the reference ``utils/loss.py``: the batch / domain losses that consume the ``GatherLayer`` all_gather.

Same class names, constructor arguments and forward signatures.  The heavy part - the N x N similarity matrices over the
gathered attention maps (``att.view(N, -1) @ att.view(N, -1).t()``, K = 8 x 2 500 x 144 floats per row, loss.py:42-52,
118-127) - runs on ``dml_gram_fwd`` (csrc/gram.cu): one streaming pass over the maps where they lie (the per-head
``view(N, 8, -1).transpose(0, 1)`` is a stride, the gathered buffer is read in place), and its backward produces the gradient
of the LOCAL rows only (``GatherLayer.backward`` drops the other ranks' slices, utils/gather.py:16-20).  Everything of size
N x N (row norms, means, the squared difference) is plain torch under autograd.  ``BatchLoss`` works on [N, 128] / [8, N, 288]
tensors (a few hundred KB): plain torch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import call, ptr, stream
from .gather import GatherLayer


def _gather_rows(x: torch.Tensor, world_size: int):
    """[B, ...] -> contiguous [world * B, ...] (all_gather; rank order = row order, as torch.cat(GatherLayer.apply(x)))."""
    x = x.contiguous()
    if world_size <= 1:
        return x
    buf = torch.empty((world_size,) + tuple(x.shape), dtype=x.dtype, device=x.device)
    try:
        dist.all_gather_into_tensor(buf, x)
    except (RuntimeError, NotImplementedError):        # gloo builds without the tensor form
        dist.all_gather(list(buf.unbind(0)), x)
    return buf.view((world_size * x.shape[0],) + tuple(x.shape[1:]))


def _row_table(t: torch.Tensor, n_rows: int, row_stride: int) -> torch.Tensor:
    base = t.data_ptr()
    return torch.tensor([base + 4 * i * row_stride for i in range(n_rows)], dtype=torch.int64, device=t.device)


class _GatheredGramFn(torch.autograd.Function):
    """a, b: LOCAL [B, G, K] fp32 maps (b may be a itself) -> sim [G, N, N], N = world * B, sim[g] = A_all[g] B_all[g]^T over the
    rows of all ranks; backward returns the gradient of the local rows only."""

    @staticmethod
    def forward(ctx, a, b, world_size):
        same = b is a
        Bl, G, K = a.shape
        rank = dist.get_rank() if world_size > 1 else 0
        lib = _lib.load(check_device=True)
        A = _gather_rows(a.float(), world_size)
        Bm = A if same else _gather_rows(b.float(), world_size)
        N = A.shape[0]
        ta = _row_table(A, N, G * K)
        tb = ta if same else _row_table(Bm, N, G * K)
        nsplit = lib.dml_gram_splits(G, K, 0)
        part = torch.empty(G, nsplit, N, N, device=a.device, dtype=torch.float32)
        call("dml_gram_fwd", ptr(ta), K, ptr(tb), K, G, N, K, ptr(part), stream())
        ctx.save_for_backward(A, Bm, ta, tb)
        ctx.meta = (same, rank, Bl, G, K, N)
        return part.sum(1)

    @staticmethod
    def backward(ctx, dsim):
        A, Bm, ta, tb = ctx.saved_tensors
        same, rank, Bl, G, K, N = ctx.meta
        lo = rank * Bl
        dsim = dsim.contiguous().float()
        rows = dsim[:, lo:lo + Bl, :]                                  # d sim[g][i, :] for the local i
        cols = dsim[:, :, lo:lo + Bl].transpose(1, 2)                  # d sim[g][:, j] for the local j
        da = torch.empty(Bl, G, K, device=A.device, dtype=torch.float32)
        if same:
            W = (rows + cols).contiguous()
            call("dml_rows_mix", ptr(W), ptr(ta), K, G, Bl, N, K, ptr(da), K, G * K, stream())
            return da, None, None
        Wa, Wb = rows.contiguous(), cols.contiguous()
        db = torch.empty(Bl, G, K, device=A.device, dtype=torch.float32)
        call("dml_rows_mix", ptr(Wa), ptr(tb), K, G, Bl, N, K, ptr(da), K, G * K, stream())
        call("dml_rows_mix", ptr(Wb), ptr(ta), K, G, Bl, N, K, ptr(db), K, G * K, stream())
        return da, db, None


def gathered_gram(a: torch.Tensor, b: torch.Tensor, groups: int, world_size: int) -> torch.Tensor:
    """a, b [B, ...] local maps whose rows split into `groups` equal contiguous parts -> [groups, N, N]."""
    if not a.is_cuda:
        raise _lib.DmlError("dml_b200 ops need CUDA tensors (there is no CPU fallback)")
    Bl = a.shape[0]
    a3 = a.reshape(Bl, groups, -1)
    b3 = a3 if b is a else b.reshape(Bl, groups, -1)
    return _GatheredGramFn.apply(a3, b3, world_size)


def _row_normalised(sim: torch.Tensor) -> torch.Tensor:
    return sim / torch.norm(sim, 2, -1, keepdim=True)                  # loss.py:49-50: norm over dim 1 of each [N, N]


class DistillationLoss(nn.Module):
    """utils/loss.py:7-23."""

    def __init__(self, temperature=2.0):
        super().__init__()
        self.temperature = temperature

    def forward(self, student_logits, teacher_logits):
        soft_targets = F.softmax(teacher_logits / self.temperature, dim=1)
        soft_prob = F.log_softmax(student_logits / self.temperature, dim=1)
        return F.kl_div(soft_prob, soft_targets, reduction='batchmean') * (self.temperature ** 2)


class PathBatchLoss(nn.Module):
    """utils/loss.py:25-64: per-head similarity of the attention maps of two scales over the global batch."""

    def __init__(self, batch_size, world_size):
        super().__init__()
        self.batch_size = batch_size
        self.world_size = world_size

    def forward(self, att10, att20):
        N = self.batch_size * self.world_size
        mean10 = _row_normalised(gathered_gram(att10, att10, 8, self.world_size)).mean(0)      # Q18: 8 heads hard-coded
        mean20 = _row_normalised(gathered_gram(att20, att20, 8, self.world_size)).mean(0)
        assert mean10.shape == (N, N)
        return (mean10 - mean20) ** 2 / N


def low_rank_loss(tensor):
    """utils/loss.py:67-74."""
    u, s, v = torch.svd(tensor)
    return torch.sum(s[1:])


def diag_variance_loss(x, weight=1.0):
    """utils/loss.py:82-85."""
    return weight * torch.var(x.diagonal())


class OmicDomainScaleLoss(nn.Module):
    """utils/loss.py:90-143: cross-scale similarity of the whole maps, variance of its diagonal."""

    def __init__(self, batch_size, world_size):
        super().__init__()
        self.batch_size = batch_size
        self.world_size = world_size

    def forward(self, att1_tea10, att1_tea20, att2_tea10, att2_tea20):
        sim1 = _row_normalised(gathered_gram(att1_tea10, att1_tea20, 1, self.world_size)[0])
        sim2 = _row_normalised(gathered_gram(att2_tea10, att2_tea20, 1, self.world_size)[0])
        return diag_variance_loss(sim1, weight=10000) + diag_variance_loss(sim2, weight=10000)


def directional_consistency_loss(M, eps=1e-6):
    """utils/loss.py:147-181 (without its debug prints)."""
    differences = M[0] - M[1]
    nonzero_mask = torch.abs(differences) > eps
    n_nonzero = torch.sum(nonzero_mask)
    signs = torch.sign(differences)
    if n_nonzero > 0:
        x_normalized = torch.sum(signs) / n_nonzero
    else:
        x_normalized = torch.tensor(0.0, device=M.device)
    return (torch.abs(x_normalized) - 1.0) ** 2


class BatchLoss(nn.Module):
    """utils/loss.py:220-253: omic embeddings [B, 128] against the sampling grids [8 B, 2, 12, 12] (small tensors)."""

    def __init__(self, batch_size, world_size):
        super().__init__()
        self.batch_size = batch_size
        self.world_size = world_size

    def forward(self, omic, vgrid):
        N = self.batch_size * self.world_size
        if self.world_size > 1:
            omic = torch.cat(GatherLayer.apply(omic), dim=0)
            vgrid = torch.cat(GatherLayer.apply(vgrid), dim=0)
        omic = omic.view(N, -1)
        vgrid = vgrid.view(8, N, -1)                                    # Q18: 8 groups hard-coded
        similarity = _row_normalised(omic.mm(omic.t()))
        mean_vgrid_sim = _row_normalised(torch.bmm(vgrid, vgrid.transpose(1, 2))).mean(0)
        return (similarity - mean_vgrid_sim) ** 2 / N
