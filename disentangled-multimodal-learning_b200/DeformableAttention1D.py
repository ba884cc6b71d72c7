"""Drop-in mirror of the reference ``models/DeformableAttention1D.py`` (DeformCrossAttention1D).

Same class name, constructor keywords, forward signature and state_dict keys/shapes as the
reference (DeformableAttention1D.py:106-240); the compute runs
on the sm_100a kernels through ``ops.DeformCrossAttn1DFn``.  The submodules below only HOLD the
parameters under the reference's names - they are never called.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else d


def divisible_by(numer, denom):
    return (numer % denom) == 0


class CPB(nn.Module):
    """Parameter holder for the continuous position bias MLP (reference :60-102):
    mlp.0.0 = Linear(1, dim), mlp.1.0 = Linear(dim, dim), mlp.2 = Linear(dim, heads // offset_groups)."""

    def __init__(self, dim, *, heads, offset_groups, depth, log_distance=True):
        super().__init__()
        self.heads = heads
        self.offset_groups = offset_groups
        self.log_distance = log_distance
        self.mlp = nn.ModuleList([])
        self.mlp.append(nn.Sequential(nn.Linear(1, dim), nn.ReLU()))
        for _ in range(depth - 1):
            self.mlp.append(nn.Sequential(nn.Linear(dim, dim), nn.ReLU()))
        self.mlp.append(nn.Linear(dim, heads // offset_groups))


class DeformCrossAttention1D(nn.Module):
    def __init__(
        self,
        *,
        dim,
        dim_head=64,
        heads=8,
        dropout=0.,
        downsample_factor=4,
        offset_scale=None,
        offset_groups=4,
        offset_kernel_size=6,
        cpb_log_distance=True,
        group_queries=False,
        group_key_values=False,
    ):
        super().__init__()
        offset_scale = default(offset_scale, downsample_factor)
        assert offset_kernel_size >= downsample_factor, 'offset kernel size must be greater than or equal to the downsample factor'
        assert divisible_by(offset_kernel_size - downsample_factor, 2)
        offset_groups = default(offset_groups, heads)
        assert divisible_by(heads, offset_groups)

        inner_dim = dim_head * heads
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.offset_groups = offset_groups
        self.offset_scale = offset_scale
        self.offset_kernel_size = offset_kernel_size
        offset_dims = inner_dim // offset_groups
        self.downsample_factor = downsample_factor

        # indices 0 and 2 carry the parameters (keys to_offsets.0.{weight,bias}, to_offsets.2.weight)
        self.to_offsets = nn.Sequential(
            nn.Conv1d(offset_dims, offset_dims, offset_kernel_size, groups=offset_dims, stride=downsample_factor,
                      padding=(offset_kernel_size - downsample_factor) // 2),
            nn.GELU(),
            nn.Conv1d(offset_dims, 1, 1, bias=False),
            nn.Identity(),
            nn.Tanh(),
            nn.Identity(),
        )
        self.rel_pos_bias = CPB(dim // 4, offset_groups=offset_groups, heads=heads, depth=2, log_distance=cpb_log_distance)
        self.dropout = nn.Dropout(dropout)
        self.to_q = nn.Conv1d(dim, inner_dim, 1, groups=offset_groups if group_queries else 1, bias=False)
        self.to_k = nn.Conv1d(dim, inner_dim, 1, groups=offset_groups if group_key_values else 1, bias=False)
        self.to_v = nn.Conv1d(dim, inner_dim, 1, groups=offset_groups if group_key_values else 1, bias=False)
        self.to_out = nn.Conv1d(inner_dim, dim, 1)

        unsupported = []
        if group_queries or group_key_values:
            unsupported.append("grouped q/k/v projections")
        if not cpb_log_distance:
            unsupported.append("cpb_log_distance=False")
        if dim_head != 64 or inner_dim // offset_groups != 128 or dim != 128:
            unsupported.append(f"dim={dim}, dim_head={dim_head}, heads={heads}, offset_groups={offset_groups} "
                               "(kernels are built for dim=128, dim_head=64, 128 query channels per group - the "
                               "only configuration the reference instantiates, DeformCrossTransMIL.py:55-60)")
        if heads // offset_groups > 2:
            unsupported.append("more than 2 heads per offset group")
        self._unsupported = unsupported

    def prefetch_bias_table(self, n: int, device):
        """Start building the position-bias table for a sequence of n tokens now (it only depends on the CPB parameters): the
        next forward() on the same stream picks it up instead of building it right before the attention kernel."""
        mlp = self.rel_pos_bias.mlp
        ws = [t.detach().contiguous().float() for t in (mlp[0][0].weight.reshape(-1), mlp[0][0].bias, mlp[1][0].weight,
                                                        mlp[1][0].bias, mlp[2].weight, mlp[2].bias)]
        n_kv = ops.kv_length(n, self.offset_kernel_size, self.downsample_factor)
        if n_kv < 1:
            return
        self._prefetched = (n, ops.build_bias_table(ws, ws[0].shape[0], self.heads // self.offset_groups, n_kv,
                                                    float(self.offset_scale), device))

    def forward(self, x1, x2, return_vgrid=False, rows=None, _norm=None):
        """x1, x2: [b, dim, n] channel-first (reference layout).  Returns [b, dim, n] (+ vgrid [(b g), n_kv]).
        rows (extension, default None = the reference contract): compute the attention output only for the first
        `rows` query tokens -> [b, dim, rows]; offsets, keys and values still see every token, so those rows and
        every gradient are identical to the full call followed by a slice.
        _norm (extension used by DeformCrossTransLayer): an nn.LayerNorm; x1 / x2 are then the layer's UN-normalised token
        streams, the norm and the layer's residual x1 + attn are applied inside the fused function."""
        if self._unsupported:
            raise NotImplementedError("dml_b200 DeformCrossAttention1D: " + "; ".join(self._unsupported))
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("attention dropout > 0 is not implemented in the fused kernel "
                                      "(the reference never sets it for the 1-D layer)")
        mlp = self.rel_pos_bias.mlp
        pre = getattr(self, "_prefetched", None)
        self._prefetched = None
        cfg = (self.heads, self.dim_head, self.offset_groups, self.downsample_factor, self.offset_kernel_size,
               float(self.offset_scale), int(rows or 0), float(_norm.eps) if _norm is not None else 0.0,
               pre[1] if pre is not None and pre[0] == x1.shape[-1] else None)
        out_t, vgrid = ops.DeformCrossAttn1DFn.apply(
            x1.transpose(1, 2), x2.transpose(1, 2),
            self.to_q.weight, self.to_k.weight, self.to_v.weight, self.to_out.weight, self.to_out.bias,
            self.to_offsets[0].weight, self.to_offsets[0].bias, self.to_offsets[2].weight,
            mlp[0][0].weight, mlp[0][0].bias, mlp[1][0].weight, mlp[1][0].bias, mlp[2].weight, mlp[2].bias,
            _norm.weight if _norm is not None else None, _norm.bias if _norm is not None else None, cfg)
        out = out_t.transpose(1, 2)
        if return_vgrid:
            return out, vgrid
        return out
