"""torch.autograd.Function wrappers around the C ABI (include/dml_b200.h).

Host code here only allocates buffers, orders the launches on torch's current stream and
calls plain library GEMMs (torch.matmul -> cuBLAS) for the unfused projections; every
fused / hot operator is a hand-written sm_100a kernel reached through ``_lib.call``.
"""
from __future__ import annotations

import math
import os
from contextlib import contextmanager

import torch

from . import _lib
from ._lib import call, ptr, stream

BF16 = torch.bfloat16
F16 = torch.float16
F32 = torch.float32
CPB_GRAD_FLOATS = 1192  # DML_CPB_GRAD_FLOATS in include/dml_b200.h


@contextmanager
def tf32_matmul():
    """fp32 GEMMs of this path run on the tensor cores in TF32 (fp32 accumulate)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@contextmanager
def fp32_matmul():
    """Exact fp32 GEMMs (no TF32) for the few small products whose rounding is amplified downstream."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def mm_f32out(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16 x bf16 -> fp32 2-D GEMM (weight gradients: no bf16 rounding of the long-K sums)."""
    return torch.mm(a, b, out_dtype=F32)


def centre_taps(n: int):
    """Sequence taps of the degenerate grid sample (y = 0, align_corners=False), reference
    DeformableAttention1D.py:36-43: iy = ((0 + 1) * n - 1) / 2."""
    iy = ((0.0 + 1.0) * n - 1.0) / 2.0
    i0 = int(math.floor(iy))
    w1 = iy - i0
    return i0, min(i0 + 1, n - 1), 1.0 - w1, w1


def kv_length(n: int, ksize: int, stride: int) -> int:
    pad = (ksize - stride) // 2
    return (n + 2 * pad - ksize) // stride + 1


_SIDE_STREAMS = {}


def side_stream(dev) -> torch.cuda.Stream:
    """The auxiliary stream of the CURRENT stream (one per device and calling stream, so that the two towers of
    DeformPathomicNet, which run on two streams, do not serialise through a shared one): small kernels off the critical
    chain - the bias-table build in the forward, the weight gradients in the backward - run on it next to the chain."""
    idx = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=idx)
    return _SIDE_STREAMS[key]


def grad_scale(t: torch.Tensor) -> torch.Tensor:
    """Device-side power-of-two loss scale for an fp16 operand: float[2] = (s, 1/s) with
    4 < s * max|t| <= 8 (no host sync; s = 1 for an all-zero tensor)."""
    lo, hi = torch.aminmax(t.detach())          # one pass, no |t| temporary
    amax = torch.maximum(hi, -lo).float()
    s = torch.exp2(torch.floor(torch.log2(8.0 / amax.clamp_min(1e-30))).clamp(-60.0, 60.0))
    s = torch.where(amax > 0, s, torch.ones_like(s))
    return torch.stack([s, 1.0 / s]).contiguous()


_ONES = {}


def colsum(x2d: torch.Tensor) -> torch.Tensor:
    """Column sums of a tall [rows, C] matrix (bias gradients over 16 k tokens) as one GEMV with a cached ones vector:
    torch's column reduction takes 13-25 us on these shapes, the GEMV 4-5."""
    rows = x2d.shape[0]
    key = (x2d.device, rows)
    ones = _ONES.get(key)
    if ones is None:
        # the cached vector is read from several streams (towers, auxiliary streams): it must be complete before any of
        # them can see it, so the fill is synchronised once here (never inside a graph capture: shapes are first seen in
        # the eager warm-up passes; a shape first met while capturing gets an uncached, capture-local vector)
        ones = torch.ones(rows, device=x2d.device, dtype=F32)
        if x2d.is_cuda:
            if torch.cuda.is_current_stream_capturing():
                return torch.mv(x2d.t(), ones)
            torch.cuda.current_stream(x2d.device).synchronize()
        _ONES[key] = ones
    return torch.mv(x2d.t(), ones)


def wgrad_mm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a^T b for tall a [rows, M], b [rows, N] (weight gradients: the reduction runs over the tokens).  cuBLAS maps the
    single skinny GEMM to M N / (128 * 64) CTAs - 8 for a 512 x 128 gradient over 16 k tokens, 40-50 us; cut into 512-row
    chunks it is a batched GEMM over all SMs plus a small sum."""
    rows = a.shape[0]
    chunk = 512
    c = rows // chunk
    if c < 4:
        return a.t() @ b
    n0 = c * chunk
    out = torch.bmm(a[:n0].view(c, chunk, -1).transpose(1, 2), b[:n0].view(c, chunk, -1)).sum(0)
    if n0 < rows:
        out = out + a[n0:].t() @ b[n0:]
    return out


class AddRowBiasFn(torch.autograd.Function):
    """x [..., rows, C] + v [..., 1, C] (a per-bag vector broadcast over the tokens); the gradient of v is a column sum."""

    @staticmethod
    def forward(ctx, x, v):
        return x + v

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        B = g.shape[0] if g.dim() == 3 else 1
        gv = torch.stack([colsum(g.reshape(B, -1, g.shape[-1])[i]) for i in range(B)])[:, None, :] if g.dim() == 3 else colsum(g)[None]
        return g, gv


# largest dS^T scratch the backward may allocate per call (bytes); DML_B200_DS_WS_MAX_GB overrides, 0 disables it
DS_WS_MAX_BYTES = int(float(os.environ.get("DML_B200_DS_WS_MAX_GB", "6")) * (1 << 30))


def _proj_nt(x, W, out_shape=None):
    """x [B, rows, K] (tensor or an already split [1, B*rows, K] operand) times W[N, K]^T on the tcgen05 split GEMM."""
    if not isinstance(x, SplitOperand):
        out_shape = x.shape[:-1] + (W.shape[0],)
        x = SplitOperand(x.reshape(1, -1, x.shape[-1]), True)
    c = gemm_nt(x, SplitOperand(W[None], True), (1, x.rows, W.shape[0]))
    return c.reshape(out_shape)


def loss_scale_from_amax(amax_bits: torch.Tensor) -> torch.Tensor:
    """Device-side power-of-two loss scale from the bit pattern of max|t| (the `absmax` output of dml_pgemm): float[2] =
    (s, 1/s) with 4 < s * max|t| <= 8 (s = 1 for an all-zero tensor); no host synchronisation."""
    amax = amax_bits.view(F32).reshape(())
    s = torch.exp2(torch.floor(torch.log2(8.0 / amax.clamp_min(1e-30))).clamp(-60.0, 60.0))
    s = torch.where(amax > 0, s, torch.ones_like(s))
    return torch.stack([s, 1.0 / s]).contiguous()


class DeformCrossAttn1DFn(torch.autograd.Function):
    """Forward + backward of DeformCrossAttention1D (DeformableAttention1D.py:156-240) on token-major inputs x1t, x2t
    [B, n, dim] (fp32).  Returns (out [B, n, dim] fp32, vgrid [(B G), n_kv]).

    Every projection (to_q, to_k | to_v, to_out, the input gradients and all weight gradients) runs on the bf16-pair tcgen05
    GEMM (csrc/pgemm.cu: fp32-class, 16-bit operand pairs, fp32 accumulate) whose epilogue writes what the next kernel
    reads: q / k / v leave it as fp16 (the operand type of the fused attention kernels), to_out adds bias and residual, the
    dO product also returns max|dO| for the device-side loss scale.  The attention core keeps softmax statistics, the
    position bias, the output and all accumulators in fp32.

    ln_w / ln_b (extension used by DeformCrossTransLayer, None for the plain module call): x1t, x2t are then the
    UN-normalised streams of the layer (DeformCrossTransMIL.py:62-68); the shared LayerNorm is applied inside - to every
    row of x1t (written only as the to_q operand pair) and to the one or two centre rows of x2t that the degenerate gather
    reads - and the layer's residual x1t + attn is the to_out epilogue."""

    @staticmethod
    def forward(ctx, x1t, x2t, Wq, Wk, Wv, Wo, bo, w0, b0, w2, m_w1, m_b1, m_W2, m_b2, m_W3, m_b3, ln_w, ln_b, cfg):
        from .pairs import Pair, pgemm
        H, d, G, stride, ks, offset_scale, rows, ln_eps = cfg
        B, n, dim = x1t.shape
        C = H * d
        Cg = C // G
        nout = H // G
        hid = m_w1.shape[0]
        scale = d ** -0.5
        dev = x1t.device
        fused = ln_w is not None
        n_out = n if not rows else min(int(rows), n)      # leading query rows whose attention output is computed
        n_kv = kv_length(n, ks, stride)
        if n_kv < 1:
            raise _lib.DmlError(f"sequence of {n} tokens is too short for offset kernel {ks}/stride {stride}")
        st = stream()

        x1f = x1t.to(F32).contiguous()
        x2f = x2t.to(F32).contiguous()
        w0f, b0f, w2f = w0.reshape(Cg, ks).contiguous().float(), b0.contiguous().float(), w2.reshape(Cg).contiguous().float()
        mlp = [t.contiguous().float() for t in (m_w1.reshape(-1), m_b1, m_W2, m_b2, m_W3, m_b3)]
        Wq_p = Pair.from_f32(Wq.reshape(C, dim))
        Wkv_p = Pair.from_f32(torch.cat((Wk.reshape(C, dim), Wv.reshape(C, dim)), 0))
        Wo_p = Pair.from_f32(Wo.reshape(dim, C))

        # the bias table only depends on the MLP weights: build it on the side stream, next to the projections
        table = torch.empty(_lib.load().dml_cpb_table_bytes(), device=dev, dtype=torch.uint8)
        t_max = math.log1p(2.0 + 2.0 * float(offset_scale) / max(n_kv - 1, 1)) * 1.001 + 1e-3
        cur, side = torch.cuda.current_stream(), side_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            call("dml_cpb_table_build", *[ptr(t) for t in mlp], hid, nout, t_max, ptr(table), stream())

        i0, i1, wy0, wy1 = centre_taps(n)
        if fused:
            # shared LayerNorm: every row of x1 (as the to_q operand pair only), the centre rows of x2
            lnw, lnb = ln_w.contiguous().float(), ln_b.contiguous().float()
            x1p = Pair.empty((B, n, dim), dev)
            mean1 = torch.empty(B * n, device=dev, dtype=F32)
            rstd1 = torch.empty_like(mean1)
            call("dml_layernorm_fwd_pair", ptr(x1f), ptr(lnw), ptr(lnb), B * n, dim, float(ln_eps), None, ptr(x1p.planes),
                 x1p.planes.stride(0), ptr(mean1), ptr(rstd1), st)
            crow = [i0] if wy1 == 0.0 else [i0, i1]
            xc_in = x2f[:, crow].contiguous()                          # [B, k, dim]
            kc = len(crow)
            x2c = torch.empty_like(xc_in)
            mean2 = torch.empty(B * kc, device=dev, dtype=F32)
            rstd2 = torch.empty_like(mean2)
            call("dml_layernorm_fwd", ptr(xc_in), ptr(lnw), ptr(lnb), B * kc, dim, float(ln_eps), ptr(x2c), ptr(mean2), ptr(rstd2), st)
            gi0, gi1, gn = 0, kc - 1, kc
        else:
            lnw = mean1 = rstd1 = xc_in = mean2 = rstd2 = None
            x1p = Pair.from_f32(x1f)
            x2c, gi0, gi1, gn, crow = x2f, i0, i1, n, None

        # to_q (:175): fp32-class product, one fp16 rounding in the epilogue (q feeds the softmax exponent)
        q = torch.empty(B, n, C, device=dev, dtype=F16)
        pgemm(x1p, Wq_p.b1(), M=n, N=C, K=dim, batch=(B,), want_f32=False, half_out=q)
        vgrid = torch.empty(B * G, n_kv, device=dev, dtype=F32)
        g = torch.empty_like(vgrid)
        call("dml_offsets_fwd", ptr(q), ptr(w0f), ptr(b0f), ptr(w2f), B, n, C, G, ks, stride, float(offset_scale),
             ptr(vgrid), ptr(g), st)
        kv = torch.empty(B, n_kv, dim, device=dev, dtype=F32)
        call("dml_kv_gather_fwd", ptr(x2c), ptr(g), B, gn, dim, G, n_kv, gi0, gi1, wy0, wy1, ptr(kv), st)
        kv_p = Pair.from_f32(kv)
        kvh = torch.empty(B, n_kv, 2 * C, device=dev, dtype=F16)            # k | v (:199), one GEMM
        pgemm(kv_p, Wkv_p.b1(), M=n_kv, N=2 * C, K=dim, batch=(B,), want_f32=False, half_out=kvh)
        k, v = kvh[..., :C], kvh[..., C:]
        cur.wait_stream(side)                                             # bias table ready
        # offsets / keys / values always need every query position; the attention itself only the first n_out rows
        q_att = q if n_out == n else q[:, :n_out].contiguous()
        o = torch.empty(B, n_out, C, device=dev, dtype=F32)
        lse = torch.empty(B, H, n_out, device=dev, dtype=F32)
        call("dml_deform_attn_fwd_tc", ptr(q_att), ptr(k), ptr(v), ptr(g), ptr(table), B, H, d, n_out, n_kv, n, C, 2 * C, 2 * C, C,
             nout, scale, ptr(o), ptr(lse), st)
        o_p = Pair.from_f32(o)
        out, _ = pgemm(o_p, Wo_p.b1(), M=n_out, N=dim, K=C, batch=(B,), bias=bo.contiguous().float(),
                       resid=x1f[:, :n_out] if fused else None)          # to_out (:233) [+ the layer's residual]

        ctx.cfg = cfg
        ctx.taps = (i0, i1, wy0, wy1, gi0, gi1, gn, crow)
        ctx.pairs = (x1p, kv_p, o_p, Wq_p, Wkv_p, Wo_p)
        ctx.save_for_backward(x1f, x2c, q, q_att, kvh, g, table, o, lse, w0f, b0f, w2f, lnw, mean1, rstd1, xc_in, mean2, rstd2, *mlp)
        ctx.x2_shape = tuple(x2f.shape)
        return out, vgrid

    @staticmethod
    def backward(ctx, dout, dvgrid):
        from .pairs import Pair, pgemm
        (x1f, x2c, q, q_att, kvh, g, table, o, lse, w0f, b0f, w2f, lnw, mean1, rstd1, xc_in, mean2, rstd2, *mlp) = ctx.saved_tensors
        x1p, kv_p, o_p, Wq_p, Wkv_p, Wo_p = ctx.pairs
        H, d, G, stride, ks, offset_scale, _rows, ln_eps = ctx.cfg
        fused = lnw is not None
        n_out = o.shape[1]
        i0, i1, wy0, wy1, gi0, gi1, gn, crow = ctx.taps
        B, n, dim = x1f.shape
        C, Cg, nout, hid = H * d, (H * d) // G, H // G, mlp[0].shape[0]
        n_kv = kvh.shape[1]
        k, v = kvh[..., :C], kvh[..., C:]
        scale = d ** -0.5
        dev = x1f.device
        st = stream()

        dout = dout.contiguous().float()
        dout_p = Pair.from_f32(dout)
        # dO = dout Wo feeds dS = P (dP - D) directly: fp32-class product; its maximum (for the fp16 loss scale) is the epilogue's
        amax = torch.zeros(1, device=dev, dtype=torch.int32)
        d_o, _ = pgemm(dout_p, Wo_p.b1(), M=n_out, N=C, K=dim, batch=(B,), b_trans=True, absmax=amax)
        # weight gradients are leaves of this backward: they run on the auxiliary stream, next to the chain
        # dO -> attention backward -> offsets / gather backward -> input gradients, and join it at the end
        cur, side = torch.cuda.current_stream(), side_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            dWo = torch.empty(dim, C, device=dev, dtype=F32)
            _wgrad_pg(dout_p, o_p, dWo, M=dim, N=C, K=n_out, B=B)
            dbo = torch.empty(dim, device=dev, dtype=F32)
            call("dml_colsum", ptr(dout), B * n_out, dim, dim, ptr(dbo), stream())
        dscale = loss_scale_from_amax(amax)
        d_o16 = torch.empty(d_o.shape, device=dev, dtype=F16)
        torch.mul(d_o, dscale[0], out=d_o16)                               # scale and round in one pass

        dq_attn = torch.empty(B, n_out, C, device=dev, dtype=F32)
        dk = torch.empty(B, n_kv, C, device=dev, dtype=F32)
        dv = torch.empty_like(dk)
        dg = torch.empty(B * G, n_kv, device=dev, dtype=F32)
        segsum = torch.empty(_lib.load().dml_cpb_seg_max(), 4, device=dev, dtype=F32)
        dsum = torch.empty(B, H, n_out, device=dev, dtype=F32)
        # dS^T scratch (fp16, n x n_kv per head): lets dQ = dS K run as a streaming GEMM instead of recomputing P a second
        # time; bags too large for it fall back to the recomputing dQ kernel (same results)
        ws_bytes = _lib.load().dml_deform_attn_bwd_ws_bytes(B, H, n_out, n_kv)
        ds_ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8) if 0 < ws_bytes <= DS_WS_MAX_BYTES else None
        call("dml_deform_attn_bwd_tc", ptr(q_att), ptr(k), ptr(v), ptr(g), ptr(table), ptr(o), ptr(d_o16), ptr(lse), B, H, d,
             n_out, n_kv, n, C, 2 * C, 2 * C, C, nout, scale, ptr(dscale), ptr(dsum), ptr(dq_attn), ptr(dk), ptr(dv), ptr(dg),
             ptr(segsum), ptr(ds_ws) if ds_ws is not None else None, st)
        del ds_ws
        if n_out != n:                                 # the other query rows only receive the offset-path gradient
            full = torch.zeros(B, n, C, device=dev, dtype=F32)
            full[:, :n_out] = dq_attn
            dq_attn = full
        dk_p, dv_p = Pair.from_f32(dk), Pair.from_f32(dv)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            mlp_g = torch.empty(CPB_GRAD_FLOATS, device=dev, dtype=F32)
            call("dml_cpb_param_grad", *[ptr(t) for t in mlp], hid, nout, ptr(table), ptr(segsum), ptr(mlp_g), stream())
            dWk = torch.empty(C, dim, device=dev, dtype=F32)
            dWv = torch.empty(C, dim, device=dev, dtype=F32)
            _wgrad_pg(dk_p, kv_p, dWk, M=C, N=dim, K=n_kv, B=B)
            _wgrad_pg(dv_p, kv_p, dWv, M=C, N=dim, K=n_kv, B=B)
        Wk_p, Wv_p = Pair(Wkv_p.planes[:, :C]), Pair(Wkv_p.planes[:, C:])
        dkv, _ = pgemm(dk_p, Wk_p.b1(), M=n_kv, N=dim, K=C, batch=(B,), b_trans=True)           # [B, n_kv, dim]
        pgemm(dv_p, Wv_p.b1(), M=n_kv, N=dim, K=C, batch=(B,), b_trans=True, out=dkv, accumulate=True)
        dcentre = torch.empty(B, dim, device=dev, dtype=F32)
        call("dml_kv_gather_bwd", ptr(x2c), ptr(g), ptr(dkv), B, gn, dim, G, n_kv, gi0, gi1, wy0, wy1, ptr(dcentre),
             ptr(dg), st)
        dx2t = torch.zeros(ctx.x2_shape, device=dev, dtype=F32)
        dlw = dlb = None
        if fused:
            kc = len(crow)
            dxc = torch.stack([wy0 * dcentre] + ([wy1 * dcentre] if kc == 2 else []), 1).contiguous()      # [B, k, dim]
            dxc_in = torch.empty_like(dxc)
            dlw2 = torch.empty(dim, device=dev, dtype=F32)
            dlb2 = torch.empty_like(dlw2)
            call("dml_layernorm_bwd", ptr(dxc), ptr(xc_in), ptr(lnw), ptr(mean2), ptr(rstd2), B * kc, dim, ptr(dxc_in), ptr(dlw2),
                 ptr(dlb2), st)
            dx2t[:, crow] = dxc_in
        else:
            dx2t[:, i0] += wy0 * dcentre
            if wy1 != 0.0:
                dx2t[:, i1] += wy1 * dcentre

        d_off = dg * (2.0 / max(n_kv - 1, 1))                              # g = 2 vgrid / max(n_kv-1,1) - 1
        if dvgrid is not None:
            d_off = d_off + dvgrid
        d_off = d_off.contiguous()
        dy_ws = torch.empty(B * G, n_kv, Cg, device=dev, dtype=F32)
        wgrad = torch.empty(Cg * ks + 2 * Cg, device=dev, dtype=F32)
        dq_p = Pair.empty((B, n, C), dev)                                  # total query gradient, as the GEMM operand only
        call("dml_offsets_bwd_pair", ptr(q), ptr(w0f), ptr(b0f), ptr(w2f), ptr(d_off), ptr(dq_attn), scale, B, n, C, G, ks,
             stride, float(offset_scale), ptr(dy_ws), ptr(wgrad), None, ptr(dq_p.planes), dq_p.planes.stride(0), st)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            dWq = torch.empty(C, dim, device=dev, dtype=F32)
            _wgrad_pg(dq_p, x1p, dWq, M=C, N=dim, K=n, B=B)
        dx1t, _ = pgemm(dq_p, Wq_p.b1(), M=n, N=dim, K=C, batch=(B,), b_trans=True)
        if fused:
            dh = torch.empty_like(x1f)
            dlw = torch.empty(dim, device=dev, dtype=F32)
            dlb = torch.empty_like(dlw)
            call("dml_layernorm_bwd", ptr(dx1t), ptr(x1f), ptr(lnw), ptr(mean1), ptr(rstd1), B * n, dim, ptr(dh), ptr(dlw), ptr(dlb), st)
            dh[:, :n_out] += dout                                          # the layer's residual
            dx1t = dh
            dlw, dlb = dlw + dlw2, dlb + dlb2
        cur.wait_stream(side)
        for t in (dWo, dbo, mlp_g, dWk, dWv, dWq):                         # allocated on the auxiliary stream, consumed on this one
            t.record_stream(cur)
        for t in (dout, segsum, dout_p.planes, dk_p.planes, dv_p.planes, dq_p.planes):      # allocated here, read there
            t.record_stream(side)

        dw0 = wgrad[: Cg * ks].reshape(Cg, 1, ks)
        db0 = wgrad[Cg * ks: Cg * ks + Cg]
        dw2 = wgrad[Cg * ks + Cg:].reshape(1, Cg, 1)
        g_w1 = mlp_g[0:hid].reshape(hid, 1)
        g_b1 = mlp_g[32:32 + hid]
        g_W2 = mlp_g[64:64 + 1024].reshape(32, 32)[:hid, :hid]
        g_b2 = mlp_g[1088:1088 + hid]
        g_W3 = mlp_g[1120:1184].reshape(2, 32)[:nout, :hid]
        g_b3 = mlp_g[1184:1184 + nout]
        return (dx1t, dx2t, dWq.reshape(C, dim, 1), dWk.reshape(C, dim, 1), dWv.reshape(C, dim, 1),
                dWo.reshape(dim, C, 1), dbo, dw0, db0, dw2, g_w1, g_b1, g_W2.contiguous(), g_b2, g_W3.contiguous(),
                g_b3, dlw, dlb, None)


def _wgrad_pg(G_p, X_p, out, *, M, N, K, B):
    """out [M, N] = sum_b G_b^T X_b for token-major pairs G [B, K, M], X [B, K, N] (a weight gradient: the reduction runs over
    the tokens): split-K pair GEMM, every bag reducing into the one output."""
    from .pairs import pgemm
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    splits = max(1, min(((K + 63) // 64) // 8, (148 + tiles * B - 1) // (tiles * B)))
    out.zero_()
    pgemm(G_p, X_p, M=M, N=N, K=K, batch=(B,), a_trans=True, b_trans=True, out=out, splits=max(2, splits))
    return out


class LandmarkPoolFn(torch.autograd.Function):
    """mean over l consecutive (front-padded) tokens: x [B, n_pad, H*d] (may be a column slice of the
    fused qkv buffer) -> [B, H, n_pad/l, d] * mult  (NystromAttention.py:102-118)."""

    @staticmethod
    def forward(ctx, x, l, H, d, mult):
        B, n_pad, W = x.shape
        assert W == H * d and x.stride(2) == 1 and x.stride(0) == n_pad * x.stride(1)
        out = torch.empty(B, H, n_pad // l, d, device=x.device, dtype=F32)
        call("dml_landmark_pool_fwd", ptr(x), x.stride(1), 0, B, n_pad, l, H, d, float(mult), ptr(out), stream())
        ctx.meta = (B, n_pad, l, H, d, float(mult))
        return out

    @staticmethod
    def backward(ctx, dout):
        B, n_pad, l, H, d, mult = ctx.meta
        dout = dout.contiguous()
        dx = torch.empty(B, n_pad, H * d, device=dout.device, dtype=F32)
        call("dml_landmark_pool_bwd", ptr(dout), B, n_pad, l, H, d, mult, ptr(dx), stream())
        return dx, None, None, None, None


class SoftmaxRowsFn(torch.autograd.Function):
    """softmax over the last dim of a contiguous fp32 tensor (NystromAttention.py:137)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty_like(x)
        cols = x.shape[-1]
        call("dml_softmax_rows_fwd", ptr(x), ptr(y), x.numel() // cols, cols, stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(y)
        cols = y.shape[-1]
        call("dml_softmax_rows_bwd", ptr(y), ptr(dy), ptr(dx), y.numel() // cols, cols, stream())
        return dx


class ResConvMergeFn(torch.autograd.Function):
    """y[b,i,(h d)] = a[b,h,i,d] + depthwise conv_K(v)[b,i,(h d)]  (NystromAttention.py:144-149).
    v [B, n_pad, H*d] may be a column slice of the fused qkv buffer; w [H,1,K,1]."""

    @staticmethod
    def forward(ctx, a, v, w):
        B, H, n_pad, d = a.shape
        assert v.stride(2) == 1 and v.stride(0) == n_pad * v.stride(1)
        a = a.contiguous()
        wf = w.reshape(H, -1).contiguous().float()
        K = wf.shape[1]
        y = torch.empty(B, n_pad, H * d, device=a.device, dtype=F32)
        call("dml_res_conv_merge_fwd", ptr(a), ptr(v), v.stride(1), 0, ptr(wf), K, B, n_pad, H, d, ptr(y), stream())
        ctx.save_for_backward(v, wf)
        ctx.meta = (B, H, n_pad, d, K, tuple(w.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        v, wf = ctx.saved_tensors
        B, H, n_pad, d, K, wshape = ctx.meta
        dy = dy.contiguous()
        da = torch.empty(B, H, n_pad, d, device=dy.device, dtype=F32)
        dv = torch.empty(B, n_pad, H * d, device=dy.device, dtype=F32)
        dw = torch.empty(H, K, device=dy.device, dtype=F32)
        call("dml_res_conv_merge_bwd", ptr(dy), ptr(v), v.stride(1), 0, ptr(wf), K, B, n_pad, H, d, ptr(da), ptr(dv),
             ptr(dw), stream())
        return da, dv, dw.reshape(wshape)


class MatmulFn(torch.autograd.Function):
    """a @ b for fp32 operands, forward and backward (same batch dims); tf32=True takes the TF32
    tensor-core path (fp32 accumulate), tf32=False exact fp32."""

    @staticmethod
    def forward(ctx, a, b, tf32):
        ctx.save_for_backward(a, b)
        ctx.tf32 = tf32
        with (tf32_matmul() if tf32 else fp32_matmul()):
            return torch.matmul(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        with (tf32_matmul() if ctx.tf32 else fp32_matmul()):
            da = torch.matmul(g, b.transpose(-1, -2)) if ctx.needs_input_grad[0] else None
            db = torch.matmul(a.transpose(-1, -2), g) if ctx.needs_input_grad[1] else None
        if db is not None and db.dim() > b.dim():
            db = db.reshape(-1, *b.shape).sum(0)
        return da, db, None


def mm_tf32(a, b):
    return MatmulFn.apply(a, b, True)


def mm_fp32(a, b):
    return MatmulFn.apply(a, b, False)


def split_bf16(t: torch.Tensor):
    """fp32 -> (hi, lo) bf16 pair with hi + lo == t to 16 significant bits."""
    hi = t.to(BF16)
    return hi, (t - hi.float()).to(BF16)


class LinearBf16BagFn(torch.autograd.Function):
    """y = x @ W^T + b for a bf16 bag x [M, K] (the bag is GIVEN in bf16: no rounding is added to it) and fp32
    W [N, K]: the weight enters as a (hi, lo) bf16 pair and the incoming gradient likewise, so the product and
    the weight gradient carry fp32-class accuracy on the bf16 tensor-core path (fp32 accumulate / output)."""

    @staticmethod
    def forward(ctx, x, W, b):
        Wh, Wl = split_bf16(W)
        y = torch.mm(x, Wh.t(), out_dtype=F32)
        y += torch.mm(x, Wl.t(), out_dtype=F32)
        ctx.save_for_backward(x, W)
        return y + b

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dyh, dyl = split_bf16(dy)
        dW = torch.mm(dyh.t(), x, out_dtype=F32)
        dW += torch.mm(dyl.t(), x, out_dtype=F32)
        dx = None
        if ctx.needs_input_grad[0]:
            with tf32_matmul():
                dx = (dy @ W).to(x.dtype)
        return dx, dW, colsum(dy)


class LinearPgFn(torch.autograd.Function):
    """y = [relu](x @ W^T + b) on the bf16-pair tcgen05 GEMM (csrc/pgemm.cu).  x [rows, K] fp32 (split into a pair) or bf16
    (exact: one plane, no rounding added); W [N, K], b [N] fp32.  Bias and ReLU are the GEMM's epilogue; the weight
    gradient is a split-K GEMM over the rows, the bias gradient a column-sum kernel."""

    @staticmethod
    def forward(ctx, x, W, b, relu):
        from .pairs import Pair, pgemm
        rows, K = x.shape
        N = W.shape[0]
        xp = Pair.exact(x.contiguous()) if x.dtype == BF16 else Pair.from_f32(x)
        Wp = Pair.from_f32(W)
        y, _ = pgemm(xp, Wp, M=rows, N=N, K=K, bias=b.contiguous().float() if b is not None else None, relu=relu)
        ctx.relu, ctx.has_bias, ctx.xdtype = relu, b is not None, x.dtype
        ctx.pairs = (xp, Wp)
        ctx.save_for_backward(y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        from .pairs import Pair, pgemm
        xp, Wp = ctx.pairs
        (y,) = ctx.saved_tensors
        rows, K = xp.shape
        N = Wp.shape[0]
        dy = dy.contiguous().float()
        if ctx.relu:
            dy = torch.where(y > 0, dy, torch.zeros((), device=dy.device, dtype=dy.dtype))
        dyp = Pair.from_f32(dy)
        dW = torch.zeros(N, K, device=dy.device, dtype=F32)
        tiles = ((N + 127) // 128) * ((K + 127) // 128)
        splits = max(1, min(((rows + 63) // 64) // 8, (148 + tiles - 1) // tiles))
        if splits > 1:
            pgemm(dyp, xp, M=N, N=K, K=rows, a_trans=True, b_trans=True, out=dW, splits=splits)
        else:
            pgemm(dyp, xp, M=N, N=K, K=rows, a_trans=True, b_trans=True, out=dW)
        db = None
        if ctx.has_bias:
            db = torch.empty(N, device=dy.device, dtype=F32)
            call("dml_colsum", ptr(dy), rows, N, N, ptr(db), stream())
        dx = None
        if ctx.needs_input_grad[0]:
            dx, _ = pgemm(dyp, Wp, M=rows, N=K, K=N, b_trans=True)
            dx = dx.to(ctx.xdtype)
        return dx, dW, db, None


class FusionFn(torch.autograd.Function):
    """FusionNet (DeformCrossTransMIL.py:28-38,105,111) without the [B, N, 256] concat or the [B, N, 128] repeat of the omic
    vector: y[b] = path[b] @ Wp^T + vec[b], vec = omic @ Wo^T + bias a per-bag vector that enters as the GEMM's epilogue
    bias.  path [B, N, D] fp32, Wp [D_out, D], vec [B, D_out]."""

    @staticmethod
    def forward(ctx, path, Wp, vec):
        from .pairs import Pair, pgemm
        B, N, D = path.shape
        Do = Wp.shape[0]
        pp, Wpp = Pair.from_f32(path), Pair.from_f32(Wp)
        y, _ = pgemm(pp, Wpp.b1(), M=N, N=Do, K=D, batch=(B,), bias=vec.contiguous().float())
        ctx.pairs = (pp, Wpp)
        return y

    @staticmethod
    def backward(ctx, dy):
        from .pairs import Pair, pgemm
        pp, Wpp = ctx.pairs
        B, N, D = pp.shape
        Do = Wpp.shape[0]
        dy = dy.contiguous().float()
        dyp = Pair.from_f32(dy)
        dpath, _ = pgemm(dyp, Wpp.b1(), M=N, N=D, K=Do, batch=(B,), b_trans=True)
        dWp = torch.empty(Do, D, device=dy.device, dtype=F32)
        _wgrad_pg(dyp, pp, dWp, M=Do, N=D, K=N, B=B)
        dvec = torch.empty(B, Do, device=dy.device, dtype=F32)
        for b in range(B):
            call("dml_colsum", ptr(dy[b]), N, Do, Do, ptr(dvec[b]), stream())
        return dpath, dWp, dvec


def linear_pg(x: torch.Tensor, W: torch.Tensor, b, relu: bool = False) -> torch.Tensor:
    """[relu](x @ W^T + b) over the last dimension of x [..., K] (fp32 or bf16) -> fp32 [..., N]."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype not in (BF16, F32):
        x2 = x2.float()
    return LinearPgFn.apply(x2, W, b, relu).reshape(*lead, W.shape[0])


def fc1_bf16_bag(x: torch.Tensor, W: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """relu(x @ W^T + b) for a bf16 bag x [M, K] (DeformCrossTransMIL.py:100 on `path.float()`), fp32 output."""
    return LinearPgFn.apply(x, W, b, True)


class LayerNormFn(torch.autograd.Function):
    """LayerNorm over the last dim (128 / 256 / 512) of a contiguous fp32 tensor: one warp per row, statistics saved
    for the backward, weight / bias gradients reduced per CTA (DeformCrossTransLayer.norm, TransLayer.norm)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        x = x.contiguous().float()
        D = x.shape[-1]
        rows = x.numel() // D
        wf, bf = w.contiguous().float(), b.contiguous().float()
        y = torch.empty_like(x)
        mean = torch.empty(rows, device=x.device, dtype=F32)
        rstd = torch.empty_like(mean)
        call("dml_layernorm_fwd", ptr(x), ptr(wf), ptr(bf), rows, D, float(eps), ptr(y), ptr(mean), ptr(rstd), stream())
        ctx.save_for_backward(x, wf, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wf, mean, rstd = ctx.saved_tensors
        D = x.shape[-1]
        rows = x.numel() // D
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        dw = torch.empty(D, device=x.device, dtype=F32)
        db = torch.empty_like(dw)
        call("dml_layernorm_bwd", ptr(dy), ptr(x), ptr(wf), ptr(mean), ptr(rstd), rows, D, ptr(dx), ptr(dw), ptr(db), stream())
        return dx, dw, db, None


def layer_norm(x, norm: torch.nn.LayerNorm):
    """norm(x) through the row-LayerNorm kernel; `norm` only holds the parameters (reference state_dict keys)."""
    return LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)


# ---------------------------------------------------------------------------------------------------------------
# fp32-class batched GEMM on tcgen05 (csrc/gemm_tc.cu): used by NystromAttention
# ---------------------------------------------------------------------------------------------------------------
def _as_batched(t: torch.Tensor):
    """[..., R, C] -> (tensor whose (R, C) block is row-major with one uniform batch stride, batch, R, C, ld, batch_stride,
    transposed_view).  A transposed view of a row-major block is accepted as is (transposed_view = True: the memory
    holds [C, R]); anything else is made contiguous."""
    R, C = t.shape[-2], t.shape[-1]
    lead = t.shape[:-2]
    batch = 1
    for d in lead:
        batch *= d

    def uniform(tt):
        # the leading dims must collapse to one stride
        st, sh = list(tt.stride()[:-2]), list(tt.shape[:-2])
        dims = [(s_, n_) for s_, n_ in zip(st, sh) if n_ > 1]
        if not dims:
            return 0
        bs = dims[-1][0]
        expect = bs
        for s_, n_ in reversed(dims):
            if s_ != expect:
                return None
            expect = s_ * n_
        return bs

    if t.stride(-1) == 1 and t.stride(-2) >= C:
        bs = uniform(t)
        if bs is not None:
            return t, batch, R, C, t.stride(-2), bs, False
    if t.stride(-2) == 1 and t.stride(-1) >= R:
        bs = uniform(t)
        if bs is not None:
            return t, batch, C, R, t.stride(-1), bs, True       # memory is [C, R] row-major
    t = t.contiguous()
    return t, batch, R, C, C, R * C, False


class SplitOperand:
    """fp32 [..., R, C] -> (hi, lo) fp16 [batch, rows, ldo] with the K axis contiguous, plus its device-side scale.
    k_last=True: K is the last axis of `t` (rows = R); k_last=False: K is the second-to-last axis (rows = C, transposed
    while splitting)."""

    def __init__(self, t: torch.Tensor, k_last: bool):
        t = t.float()
        base, batch, R, C, ld, bs, tview = _as_batched(t)
        # memory block [R, C] (after undoing a transposed view); logical K-last?  XOR with the view flag
        transpose = (not k_last) != tview
        rows, K = (C, R) if transpose else (R, C)
        ldo = (K + 7) // 8 * 8
        dev = t.device
        self.hi = torch.empty(batch, rows, ldo, device=dev, dtype=F16)
        self.lo = torch.empty_like(self.hi)
        self.scale = torch.empty(2, device=dev, dtype=F32)
        ws = torch.empty(1, device=dev, dtype=torch.int32)
        call("dml_split_f16", ptr(base), bs, batch, R, C, ld, int(transpose), ldo, ptr(self.hi), ptr(self.lo),
             ptr(self.scale), ptr(ws), stream())
        self.batch, self.rows, self.K, self.ld = batch, rows, K, ldo


def gemm_nt(A: SplitOperand, Bm: SplitOperand, out_shape, alpha: float = 1.0) -> torch.Tensor:
    """C[b] = alpha * A[b] @ B[b]^T (fp32, contiguous `out_shape` = [..., M, N]); B may have batch 1 (shared)."""
    assert A.K == Bm.K and (Bm.batch == A.batch)
    c = torch.empty(out_shape, device=A.hi.device, dtype=F32)
    M, N = A.rows, Bm.rows
    call("dml_gemm_nt_split", ptr(A.hi), ptr(A.lo), ptr(Bm.hi), ptr(Bm.lo), ptr(A.scale), ptr(Bm.scale), float(alpha),
         A.batch, M, N, A.K, A.ld, Bm.ld, ptr(c), N, M * N, stream())
    return c


class MatmulTcFn(torch.autograd.Function):
    """a [..., M, K] @ b [..., K, N] (same leading dims, or b 2-D) on the tcgen05 split-fp16 GEMM, forward and backward."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        ctx.b2d = b.dim() == 2 and a.dim() > 2
        a3 = a.reshape(-1, a.shape[-1])[None] if ctx.b2d else a
        b3 = b[None] if ctx.b2d else b
        out = gemm_nt(SplitOperand(a3, True), SplitOperand(b3, False), a3.shape[:-1] + (b.shape[-1],))
        return out.reshape(a.shape[:-1] + (b.shape[-1],))

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g3 = g.reshape(-1, g.shape[-1])[None] if ctx.b2d else g
        a3 = a.reshape(-1, a.shape[-1])[None] if ctx.b2d else a
        b3 = b[None] if ctx.b2d else b
        da = db = None
        if ctx.needs_input_grad[0]:      # dA [M, K] = dC [M, N] . (B [K, N])^T
            da = gemm_nt(SplitOperand(g3, True), SplitOperand(b3, True), a3.shape).reshape(a.shape)
        if ctx.needs_input_grad[1]:      # dB [K, N] = A^T [K, M] . (dC^T [N, M])^T
            db = gemm_nt(SplitOperand(a3, False), SplitOperand(g3, False), b3.shape).reshape(b.shape)
        return da, db


def mm_tc(a, b):
    return MatmulTcFn.apply(a, b)
