"""Drop-in mirror of ``nystrom_attention.NystromAttention`` as used by the reference
(vendored copy models/NystromAttention.py:20-157 == models/cmta_utils.py:147-281).

Same constructor, forward signature and state_dict keys (to_qkv.weight, to_out.0.{weight,bias},
res_conv.weight).  Landmark pooling, the three row softmaxes and the value-conv/merge/residual are
hand-written HBM kernels; every contraction (to_qkv, the three similarity products, the 24 products of
the pseudo-inverse recurrence, the aggregation products, to_out, and all their gradients) runs on the
hand-written tcgen05 / TMEM / TMA GEMM of csrc/gemm_tc.cu with fp16 hi+lo split operands (22 significant
bits, fp32 accumulation: the 6-step pinv recurrence does not survive 11-bit operands).
"""
from __future__ import annotations

from math import ceil

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .ops import mm_tc as mm_tf32   # every contraction on the tcgen05 split-fp16 GEMM (csrc/gemm_tc.cu): fp32-class accuracy


def moore_penrose_iter_pinv(x, iters=6):
    """NystromAttention.py:20-35 - the init scalar is a GLOBAL max over batch and heads (quirk T3)."""
    abs_x = torch.abs(x)
    col = abs_x.sum(dim=-1)
    row = abs_x.sum(dim=-2)
    z = x.transpose(-1, -2) / (torch.max(col) * torch.max(row))
    eye = torch.eye(x.shape[-1], device=x.device, dtype=x.dtype)[None]
    for _ in range(iters):
        xz = mm_tf32(x, z.contiguous())
        z = 0.25 * mm_tf32(z.contiguous(), 13 * eye - mm_tf32(xz, 15 * eye - mm_tf32(xz, 7 * eye - xz)))
    return z


class NystromAttention(nn.Module):
    def __init__(self, dim, dim_head=64, heads=8, num_landmarks=256, pinv_iterations=6, residual=True,
                 residual_conv_kernel=33, eps=1e-8, dropout=0.):
        super().__init__()
        self.eps = eps
        inner_dim = heads * dim_head
        self.num_landmarks = num_landmarks
        self.pinv_iterations = pinv_iterations
        self.heads = heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))
        self.residual = residual
        if residual:
            kernel_size = residual_conv_kernel
            padding = residual_conv_kernel // 2
            self.res_conv = nn.Conv2d(heads, heads, (kernel_size, 1), padding=(padding, 0), groups=heads, bias=False)

    def forward(self, x, mask=None, return_attn=False):
        if mask is not None:
            raise NotImplementedError("mask is never passed by any caller in the reference (SURVEY.md Q11)")
        b, n, _ = x.shape
        h, m, d = self.heads, self.num_landmarks, self.dim_head
        W = h * d
        rem = n % m
        pad = (m - rem) if rem > 0 else 0
        n_pad = n + pad
        l = ceil(n / m)

        # fused qkv projection into a buffer whose first `pad` rows are the zero front-padding (:79-90)
        qkv = mm_tf32(x.float(), self.to_qkv.weight.t())
        if pad > 0:
            qkv = F.pad(qkv, (0, 0, pad, 0), value=0.0)
        qkv = qkv.contiguous()
        q_s, k_s, v_s = qkv[..., :W], qkv[..., W:2 * W], qkv[..., 2 * W:]
        heads_first = lambda t: t.reshape(b, n_pad, h, d).transpose(1, 2)
        q = heads_first(q_s) * self.scale                                  # :98
        k, v = heads_first(k_s), heads_first(v_s)
        q_l = ops.LandmarkPoolFn.apply(q_s, l, h, d, self.scale / l)        # :102-118 (q already scaled)
        k_l = ops.LandmarkPoolFn.apply(k_s, l, h, d, 1.0 / l)

        attn1 = ops.SoftmaxRowsFn.apply(mm_tf32(q, k_l.transpose(-1, -2)))    # [b,h,n_pad,m]
        attn2 = ops.SoftmaxRowsFn.apply(mm_tf32(q_l, k_l.transpose(-1, -2)))   # [b,h,m,m]  (feeds the pinv)
        attn3 = ops.SoftmaxRowsFn.apply(mm_tf32(q_l, k.transpose(-1, -2)))    # [b,h,m,n_pad]
        attn2_inv = moore_penrose_iter_pinv(attn2, self.pinv_iterations)      # :138
        out = mm_tf32(mm_tf32(attn1, attn2_inv), mm_tf32(attn3, v))           # :140

        if self.residual:
            out = ops.ResConvMergeFn.apply(out, v_s, self.res_conv.weight)    # :144-149
        else:
            out = out.transpose(1, 2).reshape(b, n_pad, W)
        out = mm_tf32(out, self.to_out[0].weight.t()) + self.to_out[0].bias
        out = self.to_out[1](out)
        out = out[:, -n:]
        if return_attn:
            attn = attn1 @ attn2_inv @ attn3
            return out, attn
        return out
