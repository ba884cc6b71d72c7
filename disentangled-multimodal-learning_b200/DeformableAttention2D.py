"""Drop-in mirror of the reference ``DeformCrossAttention2D`` (models/DeformableAttention2D.py:162-342):
same constructor keywords, forward signature, return arity and ``state_dict`` keys, on the sm_100a kernels of
csrc/deform2d.cu / deform2d_bias.cu (``ops2d.DeformCrossAttn2DFn``) and the pair GEMM for ``to_out``.

The kernels are built for the one configuration the reference constructs (models/Modules.py:107-126, 178-197, 248-257 and
DeformCrossTransMIL.py:45-54): dim 128, 8 heads = 8 offset groups, dim_head 64, grouped projections; anything else raises.
"""
import math

import torch
from torch import nn

from . import ops, ops2d
from ._lib import DmlError


def default(val, d):
    return val if val is not None else d


class Scale(nn.Module):
    """models/DeformableAttention2D.py:111-117 (parameter-free; keeps the Sequential indices of the reference)."""

    def __init__(self, scale):
        super().__init__()
        self.scale = scale

    def forward(self, x):
        return x * self.scale


class CPB(nn.Module):
    """Parameter holder of the 2-input continuous position bias (models/DeformableAttention2D.py:121-142); evaluated by
    dml_da2_bias_fwd / dml_da2_bias_bwd."""

    def __init__(self, dim, *, heads, offset_groups, depth):
        super().__init__()
        self.heads = heads
        self.offset_groups = offset_groups
        self.mlp = nn.ModuleList([])
        self.mlp.append(nn.Sequential(nn.Linear(2, dim), nn.ReLU()))
        for _ in range(depth - 1):
            self.mlp.append(nn.Sequential(nn.Linear(dim, dim), nn.ReLU()))
        self.mlp.append(nn.Linear(dim, heads // offset_groups))


CPB2D = CPB


class DeformCrossAttention2D(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8, dropout=0., downsample_factor=4, offset_scale=4,
                 offset_groups=8, offset_kernel_size=6, group_queries=True, group_key_values=True):
        super().__init__()
        offset_scale = default(offset_scale, downsample_factor)
        assert offset_kernel_size >= downsample_factor, 'offset kernel size must be greater than or equal to the downsample factor'
        assert (offset_kernel_size - downsample_factor) % 2 == 0
        offset_groups = default(offset_groups, heads)
        assert heads % offset_groups == 0
        inner_dim = dim_head * heads
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.offset_groups = offset_groups
        offset_dims = inner_dim // offset_groups
        self.downsample_factor = downsample_factor
        self.offset_scale = offset_scale
        self.offset_kernel_size = offset_kernel_size
        self.to_offsets = nn.Sequential(
            nn.Conv2d(offset_dims, offset_dims, offset_kernel_size, groups=offset_dims, stride=downsample_factor,
                      padding=(offset_kernel_size - downsample_factor) // 2),
            nn.GELU(),
            nn.Conv2d(offset_dims, 2, 1, bias=False),
            nn.Tanh(),
            Scale(offset_scale),
        )
        self.rel_pos_bias = CPB(dim // 4, offset_groups=offset_groups, heads=heads, depth=2)
        self.dropout = nn.Dropout(dropout)
        self.to_q = nn.Conv2d(dim, inner_dim, 1, groups=offset_groups if group_queries else 1, bias=False)
        self.to_k = nn.Conv2d(dim, inner_dim, 1, groups=offset_groups if group_key_values else 1, bias=False)
        self.to_v = nn.Conv2d(dim, inner_dim, 1, groups=offset_groups if group_key_values else 1, bias=False)
        self.to_out = nn.Conv2d(inner_dim, dim, 1)
        self._supported = (dim == 128 and dim_head == 64 and heads == 8 and offset_groups == 8 and group_queries and group_key_values
                           and offset_kernel_size <= 6)

    def forward(self, x1, x2, return_vgrid=False):
        """x1, x2: [B, dim, n] with n a perfect square (:241-242).  Returns (out [B, dim, n], vgrid [(B G), 2, hk, wk]) when
        ``return_vgrid`` else (out, attn [B, heads, n, n_kv]) - the reference's teacher-mode return (:328-342)."""
        if not self._supported:
            raise DmlError("DeformCrossAttention2D kernels are built for dim 128, 8 heads = 8 offset groups, dim_head 64, grouped "
                           "projections (the only configuration the reference constructs)")
        B, dim, n = x1.shape
        side = int(math.isqrt(n))
        if side * side != n or x2.shape != x1.shape:
            raise DmlError("DeformCrossAttention2D views the sequence as a square grid: n must be a perfect square "
                           "(models/DeformableAttention2D.py:241-242)")
        mlp = self.rel_pos_bias.mlp
        cfg = (side, self.offset_kernel_size, self.downsample_factor, float(self.offset_scale), float(self.dropout.p), self.training)
        o, attn, vgrid = ops2d.DeformCrossAttn2DFn.apply(
            x1.transpose(1, 2), x2.transpose(1, 2), self.to_q.weight, self.to_k.weight, self.to_v.weight,
            self.to_offsets[0].weight, self.to_offsets[0].bias, self.to_offsets[2].weight,
            mlp[0][0].weight, mlp[0][0].bias, mlp[1][0].weight, mlp[1][0].bias, mlp[2].weight, mlp[2].bias, cfg)
        out = ops.linear_pg(o, self.to_out.weight.reshape(dim, -1), self.to_out.bias).transpose(1, 2)      # :322-326
        if return_vgrid:
            return out, vgrid
        return out, attn
