"""Drop-in for the reference's ``models/MultiheadAttention.py`` (the nn.MultiheadAttention variant that returns the RAW,
pre-softmax scores; second copy at ``models/cmta_utils.py:667``) on the sm_100a co-attention kernels (csrc/coattn.cu).

Same constructor, parameter names / shapes (``in_proj_weight [3E, E]``, ``in_proj_bias [3E]``, ``out_proj.weight / bias``)
and ``forward(query, key, value, key_padding_mask=None, need_weights=True, need_raw=True, attn_mask=None)`` with
``[L, B, E]`` / ``[S, B, E]`` inputs -> ``(attn_output [L, B, E], raw scores [B, heads, L, S])``.

Served configuration = what MCAT / CMTA use (``models/model.py:1007,1047`` and ``:1168-1170,1229-1238``): one head,
``key is value``, no masks, no bias_k / zero_attn, no attention dropout, E = 256, and at most 8 tokens on one of the two
sides (the genomic side).  The long side is read exactly once per direction by one streaming kernel; the O(8 E^2) algebra
on the short side runs as ordinary (tiny) torch ops under autograd.  Anything else raises - there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
from torch.nn.init import constant_, xavier_normal_, xavier_uniform_
from torch.nn.modules.linear import NonDynamicallyQuantizableLinear
from torch.nn.parameter import Parameter

from . import _lib
from ._lib import call, ptr, stream

FEW_MAX = 8


def _rows_view(t: torch.Tensor):
    """[B, S, E] fp32 view with a unit inner stride (copy only if the inner dimension is strided or unaligned)."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.stride(2) != 1 or t.stride(0) % 4 or t.stride(1) % 4 or t.data_ptr() % 16:
        t = t.contiguous()
    return t


class _FewQueriesFn(torch.autograd.Function):
    """x [B, S, E] (any batch / row strides), qt [B, F, E], c [B, F] -> px [B, F, E], raw [B, F, S]."""

    @staticmethod
    def forward(ctx, x, qt, c):
        x = _rows_view(x)
        B, S, E = x.shape
        Fq = qt.shape[1]
        lib = _lib.load(check_device=True)
        qt, c = qt.contiguous().float(), c.contiguous().float()
        raw = torch.empty(B, Fq, S, device=x.device, dtype=torch.float32)
        px = torch.empty(B, Fq, E, device=x.device, dtype=torch.float32)
        lse = torch.empty(B, Fq, device=x.device, dtype=torch.float32)
        ws = torch.empty(lib.dml_coattn_fq_fwd_ws_floats(B, Fq, S, E), device=x.device, dtype=torch.float32)
        call("dml_coattn_fq_fwd", ptr(x), x.stride(0), x.stride(1), ptr(qt), ptr(c), B, Fq, S, E, ptr(raw), ptr(px), ptr(lse), ptr(ws),
             stream())
        ctx.save_for_backward(x, qt, raw, lse, px)
        return px, raw

    @staticmethod
    def backward(ctx, dpx, draw):
        x, qt, raw, lse, px = ctx.saved_tensors
        B, S, E = x.shape
        Fq = qt.shape[1]
        lib = _lib.load(check_device=True)
        dpx = dpx.contiguous().float()
        dsum = (dpx * px).sum(-1).contiguous()
        nchunk = lib.dml_coattn_chunks(B, S, 0)
        ws = torch.empty(B, nchunk, Fq, E + 1, device=x.device, dtype=torch.float32)
        dx = torch.empty(B, S, E, device=x.device, dtype=torch.float32)
        draw_p = None
        if draw is not None:
            draw = draw.contiguous().float()
            draw_p = ptr(draw)
        call("dml_coattn_fq_bwd", ptr(x), x.stride(0), x.stride(1), ptr(qt), ptr(raw), ptr(lse), ptr(dpx), ptr(dsum), draw_p, B, Fq, S, E,
             ptr(dx), ptr(ws), stream())
        red = ws.sum(1)
        return dx, red[..., :E], red[..., E]


class _FewKeysFn(torch.autograd.Function):
    """x [B, S, E], kt [B, F, E], c [B, F], vt [B, F, E], bo [E] -> out [B, S, E], raw [B, S, F]."""

    @staticmethod
    def forward(ctx, x, kt, c, vt, bo):
        x = _rows_view(x)
        B, S, E = x.shape
        Fk = kt.shape[1]
        _lib.load(check_device=True)
        kt, c, vt, bo = kt.contiguous().float(), c.contiguous().float(), vt.contiguous().float(), bo.contiguous().float()
        raw = torch.empty(B, S, Fk, device=x.device, dtype=torch.float32)
        out = torch.empty(B, S, E, device=x.device, dtype=torch.float32)
        call("dml_coattn_fk_fwd", ptr(x), x.stride(0), x.stride(1), ptr(kt), ptr(c), ptr(vt), ptr(bo), B, Fk, S, E, ptr(raw), ptr(out),
             stream())
        ctx.save_for_backward(x, kt, vt, raw)
        return out, raw

    @staticmethod
    def backward(ctx, dout, draw):
        x, kt, vt, raw = ctx.saved_tensors
        B, S, E = x.shape
        Fk = kt.shape[1]
        lib = _lib.load(check_device=True)
        dout = _rows_view(dout)
        nchunk = lib.dml_coattn_chunks(B, S, 0)
        per = (2 * Fk + 1) * E + Fk
        ws = torch.empty(B, nchunk, per, device=x.device, dtype=torch.float32)
        dx = torch.empty(B, S, E, device=x.device, dtype=torch.float32)
        draw_p = None
        if draw is not None:
            draw = draw.contiguous().float()
            draw_p = ptr(draw)
        call("dml_coattn_fk_bwd", ptr(x), x.stride(0), x.stride(1), ptr(dout), dout.stride(0), dout.stride(1), ptr(kt), ptr(vt), ptr(raw),
             draw_p, B, Fk, S, E, ptr(dx), ptr(ws), stream())
        red = ws.sum(1)
        dkt = red[:, :Fk * E].reshape(B, Fk, E)
        dvt = red[:, Fk * E:2 * Fk * E].reshape(B, Fk, E)
        dbo = red[:, 2 * Fk * E:(2 * Fk + 1) * E].sum(0)
        dc = red[:, (2 * Fk + 1) * E:]
        return dx, dkt, dc, dvt, dbo


class MultiheadAttention(nn.Module):
    """models/MultiheadAttention.py:332-488 (constructor, parameters and forward signature unchanged)."""

    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None, vdim=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.kdim = kdim if kdim is not None else embed_dim
        self.vdim = vdim if vdim is not None else embed_dim
        self._qkv_same_embed_dim = self.kdim == embed_dim and self.vdim == embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        if self._qkv_same_embed_dim is False:
            self.q_proj_weight = Parameter(torch.Tensor(embed_dim, embed_dim))
            self.k_proj_weight = Parameter(torch.Tensor(embed_dim, self.kdim))
            self.v_proj_weight = Parameter(torch.Tensor(embed_dim, self.vdim))
            self.register_parameter('in_proj_weight', None)
        else:
            self.in_proj_weight = Parameter(torch.empty(3 * embed_dim, embed_dim))
            self.register_parameter('q_proj_weight', None)
            self.register_parameter('k_proj_weight', None)
            self.register_parameter('v_proj_weight', None)
        if bias:
            self.in_proj_bias = Parameter(torch.empty(3 * embed_dim))
        else:
            self.register_parameter('in_proj_bias', None)
        self.out_proj = NonDynamicallyQuantizableLinear(embed_dim, embed_dim)
        if add_bias_kv:
            self.bias_k = Parameter(torch.empty(1, 1, embed_dim))
            self.bias_v = Parameter(torch.empty(1, 1, embed_dim))
        else:
            self.bias_k = self.bias_v = None
        self.add_zero_attn = add_zero_attn
        self._reset_parameters()

    def _reset_parameters(self):
        if self._qkv_same_embed_dim:
            xavier_uniform_(self.in_proj_weight)
        else:
            xavier_uniform_(self.q_proj_weight)
            xavier_uniform_(self.k_proj_weight)
            xavier_uniform_(self.v_proj_weight)
        if self.in_proj_bias is not None:
            constant_(self.in_proj_bias, 0.)
            constant_(self.out_proj.bias, 0.)
        if self.bias_k is not None:
            xavier_normal_(self.bias_k)
        if self.bias_v is not None:
            xavier_normal_(self.bias_v)

    def _weights(self):
        E = self.embed_dim
        if self._qkv_same_embed_dim:
            Wq, Wk, Wv = self.in_proj_weight[:E], self.in_proj_weight[E:2 * E], self.in_proj_weight[2 * E:]
        else:
            Wq, Wk, Wv = self.q_proj_weight, self.k_proj_weight, self.v_proj_weight
        if self.in_proj_bias is not None:
            bq, bk, bv = self.in_proj_bias[:E], self.in_proj_bias[E:2 * E], self.in_proj_bias[2 * E:]
        else:
            z = Wq.new_zeros(E)
            bq = bk = bv = z
        return Wq, Wk, Wv, bq, bk, bv

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, need_raw=True, attn_mask=None):
        E = self.embed_dim
        if key_padding_mask is not None or attn_mask is not None:
            raise NotImplementedError("dml_b200 MultiheadAttention: masks are not served (no caller of the reference passes one)")
        if self.num_heads != 1 or self.bias_k is not None or self.add_zero_attn or not self._qkv_same_embed_dim:
            raise NotImplementedError("dml_b200 MultiheadAttention serves num_heads=1 without bias_kv / zero_attn / kdim / vdim "
                                      "(models/model.py:1007,1168,1170)")
        if self.dropout > 0. and self.training:
            raise NotImplementedError("dml_b200 MultiheadAttention: attention dropout is not served (the reference builds it with 0.)")
        if key is not value:
            raise NotImplementedError("dml_b200 MultiheadAttention needs key is value (every reference call site)")
        if not query.is_cuda:
            raise _lib.DmlError("dml_b200 ops need CUDA tensors (there is no CPU fallback)")
        L, B, _ = query.shape
        S = key.shape[0]
        assert query.shape[2] == E and key.shape[2] == E and key.shape[1] == B
        scaling = float(self.head_dim) ** -0.5
        Wq, Wk, Wv, bq, bk, bv = self._weights()
        Wo, bo = self.out_proj.weight, self.out_proj.bias
        if bo is None:
            bo = Wo.new_zeros(E)
        if L <= FEW_MAX:
            # few queries over the long key side (MCAT coattn, CMTA G_in_P_Att): fold W_k / W_v into the query side
            q = (torch.nn.functional.linear(query, Wq, bq) * scaling).transpose(0, 1)          # [B, L, E]
            qt = q @ Wk                                                                         # (W_k^T q_l)
            c = q @ bk
            px, raw = _FewQueriesFn.apply(key.transpose(0, 1), qt, c)
            attn = torch.nn.functional.linear(px, Wv, bv)                                       # rows of P sum to one
            out = torch.nn.functional.linear(attn, Wo, bo).transpose(0, 1)                      # [L, B, E]
            raw = raw.view(B, 1, L, S)
        elif S <= FEW_MAX:
            # the long side asks, few keys answer (CMTA P_in_G_Att): fold W_q and out_proj into the key side
            k = torch.nn.functional.linear(key, Wk, bk).transpose(0, 1)                        # [B, S, E]
            v = torch.nn.functional.linear(value, Wv, bv).transpose(0, 1)
            kt = (k @ Wq) * scaling
            c = (k @ bq) * scaling
            vt = v @ Wo.t()
            out, raw = _FewKeysFn.apply(query.transpose(0, 1), kt, c, vt, bo)
            out = out.transpose(0, 1)                                                           # [L, B, E]
            raw = raw.view(B, 1, L, S)
        else:
            raise NotImplementedError(f"dml_b200 MultiheadAttention serves co-attention with <= {FEW_MAX} tokens on one side "
                                      f"(got L={L}, S={S})")
        if not need_weights:
            return out, None
        if need_raw:
            return out, raw
        return out, torch.softmax(raw, dim=-1).sum(dim=1) / self.num_heads
