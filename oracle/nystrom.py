"""Oracle restatement of NystromAttention (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/models/NystromAttention.py:20-157 (bit-identical vendored
copy: models/cmta_utils.py:147-281).  The reference imports the class from the pip
package ``nystrom_attention`` (lucidrains/nystrom-attention, version unpinned - the
reference has no requirements file); that package is absent, the vendored file is
the published algorithm this module restates.

Parameter dict keys = reference state_dict keys:
    to_qkv.weight [3*h*d, dim]   to_out.0.weight [dim, h*d]   to_out.0.bias [dim]
    res_conv.weight [h, 1, K, 1]
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


def landmark_geometry(n: int, m: int):
    """Integer artefacts (NystromAttention.py:79-82,102): returns (front_pad, n_pad, l).
    The pad is PREPENDED (quirk Q8); l = ceil(n / m) uses the UNPADDED n."""
    rem = n % m
    pad = (m - rem) if rem > 0 else 0
    return pad, n + pad, math.ceil(n / m)


def moore_penrose_iter_pinv(x: torch.Tensor, iters: int = 6) -> torch.Tensor:
    """NystromAttention.py:20-35.  The init scalar is a GLOBAL max over batch and
    heads (quirk T3 / Q9)."""
    ax = x.abs()
    col = ax.sum(dim=-1)
    row = ax.sum(dim=-2)
    z = x.transpose(-1, -2) / (col.max() * row.max())
    eye = torch.eye(x.shape[-1], device=x.device, dtype=x.dtype)[None]
    for _ in range(iters):
        xz = x @ z
        z = 0.25 * z @ (13 * eye - (xz @ (15 * eye - (xz @ (7 * eye - xz)))))
    return z


def nystrom_attention(x: torch.Tensor, P: Params, *, heads: int = 8, dim_head: int = 64,
                      num_landmarks: int = 256, pinv_iterations: int = 6, residual: bool = True,
                      return_aux: bool = False):
    """NystromAttention.forward (NystromAttention.py:74-157), mask=None, eval-mode
    dropout (identity).  x [b, n, dim] -> [b, n, dim]."""
    b, n, _ = x.shape
    h, m = heads, num_landmarks
    pad, n_pad, l = landmark_geometry(n, m)
    if pad > 0:
        x = F.pad(x, (0, 0, pad, 0), value=0.0)                           # :82
    qkv = F.linear(x, P["to_qkv.weight"])                                # :89
    q, k, v = qkv.chunk(3, dim=-1)
    heads_first = lambda t: t.reshape(b, n_pad, h, dim_head).transpose(1, 2)  # :90
    q, k, v = heads_first(q), heads_first(k), heads_first(v)
    q = q * dim_head ** -0.5                                              # :98
    # landmarks: sum over l consecutive PADDED tokens, then divide by l  (:102-118)
    q_l = q.reshape(b, h, n_pad // l, l, dim_head).sum(dim=3) / l
    k_l = k.reshape(b, h, n_pad // l, l, dim_head).sum(dim=3) / l
    sim1 = q @ k_l.transpose(-1, -2)                                      # :123
    sim2 = q_l @ k_l.transpose(-1, -2)                                    # :124
    sim3 = q_l @ k.transpose(-1, -2)                                      # :125
    a1, a2, a3 = sim1.softmax(-1), sim2.softmax(-1), sim3.softmax(-1)     # :137
    a2_inv = moore_penrose_iter_pinv(a2, pinv_iterations)                 # :138
    out = (a1 @ a2_inv) @ (a3 @ v)                                        # :140
    if residual:
        w = P["res_conv.weight"]
        out = out + F.conv2d(v, w, padding=(w.shape[2] // 2, 0), groups=h)  # :144-145
    out = out.transpose(1, 2).reshape(b, n_pad, h * dim_head)             # :149
    out = F.linear(out, P["to_out.0.weight"], P["to_out.0.bias"])         # :150
    out = out[:, -n:]                                                     # :151
    if return_aux:
        return out, dict(pad=pad, n_pad=n_pad, l=l, q_landmarks=q_l, k_landmarks=k_l, attn2=a2,
                         attn2_inv=a2_inv)
    return out
