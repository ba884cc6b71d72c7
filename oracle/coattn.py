"""Oracle restatement of the raw-score MultiheadAttention used for MCAT / CMTA co-attention (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/models/MultiheadAttention.py:7-321 (``multi_head_attention_forward``; second copy
models/cmta_utils.py:667-) along the one branch its callers take (models/model.py:1007,1047,1168-1170,1229-1238):
packed ``in_proj_weight``, ``key is value``, no masks, no bias_k / zero_attn, dropout 0, ``need_weights`` and ``need_raw``.

Parameter dict keys = reference state_dict keys:
    in_proj_weight [3E, E]   in_proj_bias [3E]   out_proj.weight [E, E]   out_proj.bias [E]
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


def multihead_attention_raw(query: torch.Tensor, key: torch.Tensor, P: Params, num_heads: int = 1):
    """query [L, B, E], key (= value) [S, B, E] -> (attn_output [L, B, E], raw scores [B, heads, L, S])."""
    L, B, E = query.shape
    S = key.shape[0]
    hd = E // num_heads
    W, b = P["in_proj_weight"], P["in_proj_bias"]
    q = F.linear(query, W[:E], b[:E])                                  # :131-133
    k, v = F.linear(key, W[E:], b[E:]).chunk(2, dim=-1)                # :140-146
    q = q * (float(hd) ** -0.5)                                         # :190
    q = q.contiguous().view(L, B * num_heads, hd).transpose(0, 1)       # :234-238
    k = k.contiguous().view(S, B * num_heads, hd).transpose(0, 1)
    v = v.contiguous().view(S, B * num_heads, hd).transpose(0, 1)
    raw = torch.bmm(q, k.transpose(1, 2))                               # :266  [B h, L, S]
    attn = torch.softmax(raw, dim=-1)                                   # :288
    out = torch.bmm(attn, v)                                            # :290
    out = out.transpose(0, 1).contiguous().view(L, B, E)                # :292
    out = F.linear(out, P["out_proj.weight"], P["out_proj.bias"])       # :293
    return out, raw.view(B, num_heads, L, S)                            # :300-303
