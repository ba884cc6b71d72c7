"""Oracle restatement of DeformCrossAttention1D (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/models/DeformableAttention1D.py; every function cites
the lines it restates.  Parameters are passed as a dict keyed exactly like the
reference ``state_dict``:

    to_offsets.0.weight [C/G,1,ks]  to_offsets.0.bias [C/G]  to_offsets.2.weight [1,C/G,1]
    rel_pos_bias.mlp.0.0.{weight[32,1],bias[32]}  .1.0.{weight[32,32],bias[32]}
    rel_pos_bias.mlp.2.{weight[H/G,32],bias[H/G]}
    to_q/to_k/to_v.weight [C,dim,1]   to_out.{weight[dim,C,1],bias[dim]}

All functions are differentiable torch code, so gradients come from autograd.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

Params = Dict[str, torch.Tensor]


def kv_length(n: int, ksize: int = 6, stride: int = 4) -> int:
    """Output length of the strided depthwise conv (DeformableAttention1D.py:140):
    padding = (ksize - stride) // 2, n_kv = floor((n + 2p - ksize) / stride) + 1."""
    pad = (ksize - stride) // 2
    return (n + 2 * pad - ksize) // stride + 1


def normalize_grid(v: torch.Tensor) -> torch.Tensor:
    """DeformableAttention1D.py:45-48 - scale by the LAST-dim length of ``v``."""
    n = v.shape[-1]
    return 2.0 * v / max(n - 1, 1) - 1.0


def offsets_net(q_grouped: torch.Tensor, P: Params, stride: int, offset_scale: float) -> torch.Tensor:
    """to_offsets Sequential (DeformableAttention1D.py:139-146):
    depthwise Conv1d(k, stride, pad) + bias -> GELU(erf) -> Conv1d(C/G -> 1, no bias)
    -> tanh -> * offset_scale.  q_grouped: [(b g), C/G, n] -> [(b g), n_kv]."""
    w0, b0, w2 = P["to_offsets.0.weight"], P["to_offsets.0.bias"], P["to_offsets.2.weight"]
    ks = w0.shape[-1]
    y = F.conv1d(q_grouped, w0, b0, stride=stride, padding=(ks - stride) // 2, groups=w0.shape[0])
    y = F.gelu(y)
    y = F.conv1d(y, w2)[:, 0]
    return torch.tanh(y) * offset_scale


def grid_sample_1d_literal(feats: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """DeformableAttention1D.py:36-43, restated literally: the grid gets a trailing
    (value, 0) pair, so the learned coordinate is the x (size-1 W axis) coordinate
    and y = 0 samples the middle of the sequence (quirk T1 / Q1)."""
    g2 = torch.stack((grid, torch.zeros_like(grid)), dim=-1)[:, :, None, :]  # [B, n_kv, 1, 2]
    out = F.grid_sample(feats[..., None], g2, mode="bilinear", padding_mode="zeros", align_corners=False)
    return out[..., 0]


def centre_taps(n: int):
    """Integer artefact of the degenerate sample: y = 0, align_corners=False ->
    iy = ((0+1)*n - 1)/2.  Returns (idx0, idx1, w0, w1); idx1's weight is 0 for odd n."""
    iy = ((0.0 + 1.0) * n - 1.0) / 2.0
    i0 = int(math.floor(iy))
    w1 = iy - i0
    return i0, i0 + 1, 1.0 - w1, w1


def tent_weight(g: torch.Tensor) -> torch.Tensor:
    """x tap weight of F.grid_sample on a width-1 image (align_corners=False):
    ix = ((g+1)*1 - 1)/2, only pixel 0 exists -> weight 1-|ix| inside (-1,1), else 0.
    Written with the same fp32 operation order as the ATen kernel so that the
    closed form is bit-identical to grid_sample_1d_literal for odd n."""
    ix = ((g + 1.0) * 1.0 - 1.0) / 2.0
    x0 = torch.floor(ix)
    w_right = ix - x0            # weight of pixel x0+1
    w_left = (x0 + 1.0) - ix     # weight of pixel x0
    w = torch.where(x0 == 0, w_left, torch.where(x0 == -1, w_right, torch.zeros_like(ix)))
    return w


def grid_sample_1d_closed(feats: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """Closed form of grid_sample_1d_literal: centre token(s) x tent."""
    n = feats.shape[-1]
    i0, i1, w0, w1 = centre_taps(n)
    centre = feats[..., i0] * w0
    if w1 != 0.0:
        centre = centre + feats[..., i1] * w1
    return centre[:, :, None] * tent_weight(grid)[:, None, :]


def cpb_bias(seq_scaled: torch.Tensor, g: torch.Tensor, P: Params, groups: int) -> torch.Tensor:
    """CPB.forward (DeformableAttention1D.py:84-102).  seq_scaled [ni], g [(b G), n_kv]
    -> bias [b, H, ni, n_kv] with head index = group * (H/G) + mlp_output."""
    pos = seq_scaled[None, :, None, None] - g[:, None, :, None]
    t = torch.sign(pos) * torch.log(pos.abs() + 1)
    h = F.relu(F.linear(t, P["rel_pos_bias.mlp.0.0.weight"], P["rel_pos_bias.mlp.0.0.bias"]))
    h = F.relu(F.linear(h, P["rel_pos_bias.mlp.1.0.weight"], P["rel_pos_bias.mlp.1.0.bias"]))
    o = F.linear(h, P["rel_pos_bias.mlp.2.weight"], P["rel_pos_bias.mlp.2.bias"])  # [(bG), i, j, o]
    bg, ni, nj, no = o.shape
    b = bg // groups
    return o.reshape(b, groups, ni, nj, no).permute(0, 1, 4, 2, 3).reshape(b, groups * no, ni, nj)


def _attend_rows(q_rows, k, v, seq_rows, g, P, groups):
    """sim + CPB -> softmax -> attn @ v for a block of query rows
    (DeformableAttention1D.py:211-231).  q_rows [b,H,r,d] already scaled."""
    sim = torch.einsum("bhid,bhjd->bhij", q_rows, k)
    sim = sim + cpb_bias(seq_rows, g, P, groups)
    sim = sim - sim.amax(dim=-1, keepdim=True).detach()
    attn = sim.softmax(dim=-1)
    return torch.einsum("bhij,bhjd->bhid", attn, v)


def deform_cross_attention_1d(
    x1: torch.Tensor,
    x2: torch.Tensor,
    P: Params,
    *,
    heads: int = 8,
    dim_head: int = 64,
    offset_groups: int = 4,
    downsample_factor: int = 4,
    offset_scale: Optional[float] = None,
    row_block: Optional[int] = None,
    literal_gather: bool = False,
    return_aux: bool = False,
):
    """DeformCrossAttention1D.forward (DeformableAttention1D.py:156-240).
    x1, x2: [b, dim, n] channel-first.  ``row_block`` evaluates the attention in
    blocks of query rows with recompute-in-backward (identical maths; needed at
    n = 16385 where the shipped module materialises ~34 GB per CPB activation)."""
    offset_scale = downsample_factor if offset_scale is None else offset_scale
    b, _, n = x2.shape
    G, H = offset_groups, heads
    scale = dim_head ** -0.5

    q = F.conv1d(x1, P["to_q.weight"])                                   # :175
    C = q.shape[1]
    grouped_q = q.reshape(b * G, C // G, n)                               # :179-181
    offsets = offsets_net(grouped_q, P, downsample_factor, offset_scale)  # :182
    n_kv = offsets.shape[-1]
    vgrid = torch.arange(n_kv, device=x1.device) + offsets               # :186-187
    g = normalize_grid(vgrid)                                            # :188
    x2g = x2.reshape(b * G, x2.shape[1] // G, n)
    kv_feats = (grid_sample_1d_literal if literal_gather else grid_sample_1d_closed)(x2g, g)  # :190-193
    kv_feats = kv_feats.reshape(b, -1, n_kv)                              # :195
    k = F.conv1d(kv_feats, P["to_k.weight"])                              # :199
    v = F.conv1d(kv_feats, P["to_v.weight"])
    q = q * scale                                                         # :203
    split = lambda t: t.reshape(b, H, dim_head, t.shape[-1]).transpose(2, 3)  # :207
    qh, kh, vh = split(q), split(k), split(v)
    seq_scaled = normalize_grid(torch.arange(n, device=x1.device))        # :215-216
    seq_scaled = seq_scaled.to(g.dtype)
    if row_block is None or row_block >= n:
        out = _attend_rows(qh, kh, vh, seq_scaled, g, P, G)
    else:
        blocks = []
        for r0 in range(0, n, row_block):
            r1 = min(n, r0 + row_block)
            if torch.is_grad_enabled():
                blocks.append(checkpoint(_attend_rows, qh[:, :, r0:r1], kh, vh, seq_scaled[r0:r1], g, P, G,
                                         use_reentrant=False))
            else:
                blocks.append(_attend_rows(qh[:, :, r0:r1], kh, vh, seq_scaled[r0:r1], g, P, G))
        out = torch.cat(blocks, dim=2)
    out = out.transpose(2, 3).reshape(b, H * dim_head, n)                 # :232
    out = F.conv1d(out, P["to_out.weight"], P["to_out.bias"])             # :233
    if return_aux:
        return out, dict(vgrid=vgrid, g=g, offsets=offsets, n_kv=n_kv, tent=tent_weight(g),
                         centre=centre_taps(n))
    return out
