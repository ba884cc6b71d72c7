"""torch.autograd.Function wrappers around the C ABI (include/dml_b200.h).

Host code here only allocates buffers and orders the launches on torch's current stream; every operator of the path,
the projections and weight gradients included (bf16-pair tcgen05 GEMM, ``pairs.pgemm``), is a hand-written sm_100a
kernel reached through ``_lib.call``.
"""
from __future__ import annotations

import math
import os
import threading

import torch

from . import _lib
from ._lib import call, ptr, stream

BF16 = torch.bfloat16
F16 = torch.float16
F32 = torch.float32
CPB_GRAD_FLOATS = 1192  # DML_CPB_GRAD_FLOATS in include/dml_b200.h


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
# Priority of the streams that carry latency-bound kernel chains (tower, side and graph-capture streams); branch_stream() keeps
# the default (lowest) one.  Measured on the B200 (same box, A/B): -1 costs the DeformPathomicNet step 0.13 ms (3.70 -> 3.84 ms:
# the weight-gradient kernels of the side streams then cut into the attention kernels) and changes TransMIL by < 1 %, so the
# default is 0 = all streams equal; DML_B200_CHAIN_PRIORITY=-1 is kept as a tuning knob.
CHAIN_PRIORITY = int(os.environ.get("DML_B200_CHAIN_PRIORITY", "0"))


def side_stream(dev) -> torch.cuda.Stream:
    """The auxiliary stream of the CURRENT stream (one per device and calling stream, so that the two towers of
    DeformPathomicNet, which run on two streams, do not serialise through a shared one): small kernels off the critical
    chain - the bias-table build in the forward, the weight gradients in the backward - run on it next to the chain."""
    idx = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=idx, priority=CHAIN_PRIORITY)
    return _SIDE_STREAMS[key]


def branch_stream(dev) -> torch.cuda.Stream:
    """A second auxiliary stream of the CURRENT stream, distinct from side_stream(), for BULK kernels that run next to a
    latency-bound chain: the token-sized kernels of NystromAttention next to the pseudo-inverse chain, the HBM-bound dQ GEMM of
    the deformable backward next to the dK / dV adjoint chain.  Always default (= lowest) priority; see CHAIN_PRIORITY for
    what raising the other streams above it does (the block scheduler hands out thread blocks of pending kernels in launch
    order within a priority, so a 500-CTA bulk kernel waiting for SMs holds back every small kernel launched after it)."""
    idx = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream, "branch")
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=idx)
    return _SIDE_STREAMS[key]


# largest dS^T scratch the backward may allocate per call (bytes); DML_B200_DS_WS_MAX_GB overrides, 0 disables it
DS_WS_MAX_BYTES = int(float(os.environ.get("DML_B200_DS_WS_MAX_GB", "6")) * (1 << 30))
DQ_OVERLAP = os.environ.get("DML_B200_DQ_OVERLAP", "1") != "0"      # dQ GEMM of the backward on its own stream (see backward)


def build_bias_table(mlp, hid: int, nout: int, n_kv: int, offset_scale: float, dev):
    """Launch dml_cpb_table_build for the CPB MLP parameters `mlp` (contiguous fp32: w1, b1, W2, b2, W3, b3) on the auxiliary
    stream of the current stream; returns (table, event recorded after the build)."""
    table = torch.empty(_lib.load().dml_cpb_table_bytes(), device=dev, dtype=torch.uint8)
    t_max = math.log1p(2.0 + 2.0 * float(offset_scale) / max(n_kv - 1, 1)) * 1.001 + 1e-3
    cur, side = torch.cuda.current_stream(), side_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        call("dml_cpb_table_build", *[ptr(t) for t in mlp], hid, nout, t_max, ptr(table), stream())
        ev = side.record_event()
    return table, ev


# How many attention-forward launches of the same shape share the GPU (set by the caller that knows: DeformPathomicNet runs its two
# towers on two streams).  Host-side planning state only; the C entry point takes the resulting block count explicitly.
_CONCURRENT_ATTN = threading.local()


class concurrent_attention_launches:
    """with concurrent_attention_launches(2): ... - the deformable-attention forwards issued inside are planned as `count`
    same-shaped launches sharing the SMs (see fwd_half_blocks)."""

    def __init__(self, count: int):
        self.count = max(1, int(count))

    def __enter__(self):
        self.prev = getattr(_CONCURRENT_ATTN, "count", 1)
        _CONCURRENT_ATTN.count = self.count
        return self

    def __exit__(self, *a):
        _CONCURRENT_ATTN.count = self.prev
        return False


def plan_half_blocks(n: int, B: int, G: int, concurrent: int, nsm: int) -> int:
    """`half_blocks` of dml_deform_attn_fwd_tc_split.  The forward kernel's CTAs take 256 queries (two 128-row softmax groups) of
    one (bag, offset group) and cost the same; `concurrent` launches of U such CTAs each fill floor(c U / nsm) whole waves and
    spill the rest into one more.  When the spill, cut into 128-query CTAs (measured at ~0.8 of a 256-query CTA's time, not 0.5:
    one group alone cannot alternate with another on the tensor pipe), fits into ONE wave together with the trailing one-group
    CTAs, the step ends that much earlier; otherwise nothing is split.  n = 16 385, G = 4, two towers, 148 SMs: 512 CTAs = 3 waves +
    68 -> 9 blocks per (bag, group) become 18 short CTAs (measured A/B on one box: 3.680 -> 3.631 ms per step; 8 or 10 blocks are
    worse than none)."""
    rg = -(-n // 128)
    nf = rg // 2
    per = G * B
    tot = concurrent * nf * per
    waves = tot // nsm
    spill = tot - waves * nsm
    if waves == 0 or spill == 0 or nf == 0:
        return 0
    t = min(-(-spill // (concurrent * per)), nf)
    halves = concurrent * (2 * t + (rg - 2 * nf)) * per
    slack = waves * nsm - concurrent * (nf - t) * per
    return t if slack >= 0 and halves <= nsm + slack else 0


_FWD_HALF_ENV = os.environ.get("DML_B200_FWD_HALF_BLOCKS")      # tuning aid: a fixed count instead of the plan ("0" = never split)


def fwd_half_blocks(n: int, B: int, G: int, dev) -> int:
    if _FWD_HALF_ENV is not None:
        return max(0, int(_FWD_HALF_ENV))
    c = getattr(_CONCURRENT_ATTN, "count", 1)
    if c <= 1:
        return 0
    return plan_half_blocks(n, B, G, c, torch.cuda.get_device_properties(dev).multi_processor_count)


def loss_scale_from_amax(amax_bits: torch.Tensor) -> torch.Tensor:
    """Device-side power-of-two loss scale from the bit pattern of max|t| (the `absmax` output of dml_pgemm): float[2] =
    (s, 1/s) with 4 < s * max|t| <= 8 (s = 1 for an all-zero tensor); no host synchronisation.  One launch on the GPU (it sits
    on the chain between the forward and the backward of the step); the torch expression below is the CPU form of the same."""
    if amax_bits.is_cuda:
        out = torch.empty(2, device=amax_bits.device, dtype=F32)
        call("dml_loss_scale_from_amax", ptr(amax_bits), ptr(out), stream())
        return out
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
        H, d, G, stride, ks, offset_scale, rows, ln_eps = cfg[:8]
        prefetched = cfg[8] if len(cfg) > 8 else None
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

        # the bias table only depends on the MLP weights: it is built on the side stream, next to the projections - or was
        # started even earlier by the caller (prefetch_bias_table, at the top of DeformCrossTransMIL.forward)
        cur, side = torch.cuda.current_stream(), side_stream(dev)
        if prefetched is not None:
            table, table_ready = prefetched
        else:
            table, table_ready = build_bias_table(mlp, hid, nout, n_kv, offset_scale, dev)
        i0, i1, wy0, wy1 = centre_taps(n)
        if fused:
            # shared LayerNorm: every row of x1 (as the to_q operand pair only), the centre rows of x2
            lnw, lnb = ln_w.contiguous().float(), ln_b.contiguous().float()
            x1p = Pair.empty((B, n, dim), dev)
            mean1 = torch.empty(B * n, device=dev, dtype=F32)
            rstd1 = torch.empty_like(mean1)
            call("dml_layernorm_fwd_pair", ptr(x1f), ptr(lnw), ptr(lnb), B * n, dim, float(ln_eps), None, ptr(x1p.planes),
                 x1p.planes.stride(0), ptr(mean1), ptr(rstd1), st)
            kc = 1 if wy1 == 0.0 else 2                                # the centre rows i0 (, i0 + 1) are adjacent
            crow = (i0, i0 + kc)
            xc_in = x2f[:, crow[0]:crow[1]].contiguous()               # [B, k, dim]
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
        cur.wait_event(table_ready)                                       # bias table ready
        table.record_stream(cur)
        # offsets / keys / values always need every query position; the attention itself only the first n_out rows
        q_att = q if n_out == n else q[:, :n_out].contiguous()
        o = torch.empty(B, n_out, C, device=dev, dtype=F32)
        lse = torch.empty(B, H, n_out, device=dev, dtype=F32)
        call("dml_deform_attn_fwd_tc_split", ptr(q_att), ptr(k), ptr(v), ptr(g), ptr(table), B, H, d, n_out, n_kv, n, C, 2 * C, 2 * C, C,
             nout, scale, ptr(o), ptr(lse), fwd_half_blocks(n_out, B, G, dev), st)
        o_p = Pair.from_f32(o)
        out, _ = pgemm(o_p, Wo_p.b1(), M=n_out, N=dim, K=C, batch=(B,), bias=bo.contiguous().float(),
                       resid=x1f[:, :n_out] if fused else None)          # to_out (:233) [+ the layer's residual]

        ctx.cfg = cfg[:8]
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
        if d_o.numel() % 8 == 0:
            call("dml_scale_to_half", ptr(d_o), ptr(dscale), d_o.numel(), ptr(d_o16), st)      # scale and round in one pass
        else:
            torch.mul(d_o, dscale[0], out=d_o16)

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
        # With the workspace the HBM-bound dQ GEMM is issued apart, on its own stream: everything that only needs dK / dV / dg
        # (the key / value projection adjoints, the gather and offset-network backward - a latency-bound chain of ~150 us at
        # the end of the step) runs next to it and joins before dq is combined with the offset-path gradient.
        split_dq = ds_ws is not None and DQ_OVERLAP
        call("dml_deform_attn_bwd_tc", ptr(q_att), ptr(k), ptr(v), ptr(g), ptr(table), ptr(o), ptr(d_o16), ptr(lse), B, H, d,
             n_out, n_kv, n, C, 2 * C, 2 * C, C, nout, scale, ptr(dscale), ptr(dsum), None if split_dq else ptr(dq_attn), ptr(dk),
             ptr(dv), ptr(dg), ptr(segsum), ptr(ds_ws) if ds_ws is not None else None, st, kernels=2 if split_dq else 3)
        dqs = None
        if split_dq:
            dqs = branch_stream(dev)
            dqs.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(dqs):
                call("dml_deform_attn_dq_from_ds", ptr(ds_ws), ptr(k), ptr(dscale), B, H, d, n_out, n_kv, 2 * C, ptr(dq_attn), stream())
        else:
            del ds_ws
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
        # the gradient of x2 (only its centre rows are non-zero) is a leaf of this backward: a handful of microsecond-sized kernels
        # that leave the critical chain (gather backward -> offset-network backward -> dq) for the auxiliary stream
        dlw = dlb = dlw2 = dlb2 = None
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            dx2t = torch.zeros(ctx.x2_shape, device=dev, dtype=F32)
            if fused:
                kc = crow[1] - crow[0]
                dxc = torch.stack([wy0 * dcentre] + ([wy1 * dcentre] if kc == 2 else []), 1).contiguous()      # [B, k, dim]
                dxc_in = torch.empty_like(dxc)
                dlw2 = torch.empty(dim, device=dev, dtype=F32)
                dlb2 = torch.empty_like(dlw2)
                call("dml_layernorm_bwd", ptr(dxc), ptr(xc_in), ptr(lnw), ptr(mean2), ptr(rstd2), B * kc, dim, ptr(dxc_in), ptr(dlw2),
                     ptr(dlb2), stream())
                dx2t[:, crow[0]:crow[1]] = dxc_in
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
        if dqs is not None:                            # offset-network backward first, the combination once dq has arrived
            call("dml_offsets_bwd_pair", ptr(q), ptr(w0f), ptr(b0f), ptr(w2f), ptr(d_off), None, scale, B, n, C, G, ks,
                 stride, float(offset_scale), ptr(dy_ws), ptr(wgrad), None, None, 0, st, kernels=1)
            cur.wait_stream(dqs)
            del ds_ws                                  # allocated on this stream, last read on the other one: released after the join
        if n_out != n:                                 # the other query rows only receive the offset-path gradient
            full = torch.zeros(B, n, C, device=dev, dtype=F32)
            full[:, :n_out] = dq_attn
            dq_attn = full
        call("dml_offsets_bwd_pair", ptr(q), ptr(w0f), ptr(b0f), ptr(w2f), None if dqs is not None else ptr(d_off), ptr(dq_attn),
             scale, B, n, C, G, ks, stride, float(offset_scale), ptr(dy_ws), ptr(wgrad), None, ptr(dq_p.planes),
             dq_p.planes.stride(0), st, kernels=1 if dqs is not None else 2)
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
        cur.wait_stream(side)
        if fused:
            dlw, dlb = dlw + dlw2, dlb + dlb2
        for t in (dWo, dbo, mlp_g, dWk, dWv, dWq, dx2t, dlw2, dlb2):       # allocated on the auxiliary stream, consumed on this one
            if t is not None:
                t.record_stream(cur)
        for t in (dout, segsum, dout_p.planes, dk_p.planes, dv_p.planes, dq_p.planes, dcentre):      # allocated here, read there
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
        if ctx.relu and N % 8 == 0:
            dym = torch.empty_like(dy)                  # the incoming gradient belongs to autograd: the masked copy is ours
            dyp = Pair.empty((rows, N), dy.device)
            call("dml_relu_mask_pair", ptr(dy), ptr(dym), ptr(y), rows, N, N, N, ptr(dyp.planes), N, dyp.planes.stride(0), stream())
            dy = dym
        else:
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


class MaxNetFn(torch.autograd.Function):
    """The omic MLP of MaxNet (models/model.py:173-218): 4 x (Linear -> ELU -> AlphaDropout) -> ReLU as one kernel per
    direction (csrc/maxnet.cu).  x [B, in]; params = W0, b0, ..., W3, b3; p = AlphaDropout probability (0 in eval mode)."""

    @staticmethod
    def forward(ctx, x, p, *params):
        import ctypes as C
        x = x.contiguous().float()
        B = x.shape[0]
        Ws = [w.contiguous().float() for w in params[0::2]]
        bs = [b.contiguous().float() for b in params[1::2]]
        dims = [x.shape[1]] + [w.shape[0] for w in Ws]
        dev = x.device
        na, nh = sum(dims[1:]), sum(dims[:-1])
        u = torch.rand(B, na, device=dev, dtype=F32) if p > 0.0 else None
        act = torch.empty(B, na, device=dev, dtype=F32)
        hsave = torch.empty(B, nh, device=dev, dtype=F32)
        feat = torch.empty(B, dims[-1], device=dev, dtype=F32)
        Wp = (C.c_void_p * 4)(*[w.data_ptr() for w in Ws])
        bp = (C.c_void_p * 4)(*[b.data_ptr() for b in bs])
        dm = (C.c_int * 5)(*dims)
        call("dml_maxnet_fwd", ptr(x), Wp, bp, dm, B, ptr(u) if u is not None else None, float(p), ptr(act), ptr(hsave),
             ptr(feat), stream())
        ctx.p, ctx.dims = float(p), dims
        ctx.save_for_backward(u, act, hsave, feat, *Ws, *bs)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        import ctypes as C
        u, act, hsave, feat, *wb = ctx.saved_tensors
        Ws, bs = wb[:4], wb[4:]
        dims = ctx.dims
        B = feat.shape[0]
        dfeat = dfeat.contiguous().float()
        sizes = [dims[l + 1] * dims[l] + dims[l + 1] for l in range(4)]
        dparams = torch.empty(sum(sizes), device=feat.device, dtype=F32)
        dx = torch.empty(B, dims[0], device=feat.device, dtype=F32) if ctx.needs_input_grad[0] else None
        Wp = (C.c_void_p * 4)(*[w.data_ptr() for w in Ws])
        bp = (C.c_void_p * 4)(*[b.data_ptr() for b in bs])
        dm = (C.c_int * 5)(*dims)
        call("dml_maxnet_bwd", ptr(dfeat), Wp, bp, dm, B, ptr(u) if u is not None else None, ctx.p, ptr(act), ptr(hsave), ptr(feat),
             ptr(dparams), ptr(dx) if dx is not None else None, stream())
        grads, off = [], 0
        for l in range(4):
            nw = dims[l + 1] * dims[l]
            grads.append(dparams[off: off + nw].view(dims[l + 1], dims[l]))
            grads.append(dparams[off + nw: off + nw + dims[l + 1]])
            off += sizes[l]
        return (dx, None, *grads)


class TowerHeadFn(torch.autograd.Function):
    """(encoded, logits) = heads on the cls row of the layer output (DeformCrossTransMIL.py:128-151): h = LayerNorm(x[:, 0]),
    logits = _fc2(h), encoded = multimodal_projection(h) - one kernel per direction (csrc/heads.cu).  x [B, n, D]."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, W2, b2, Wp, bp, eps):
        B, n, D = x.shape
        dev = x.device
        x0 = x[:, 0].contiguous().float()
        ws = [t.contiguous().float() for t in (ln_w, ln_b, W2, b2, Wp, bp)]
        nc, De = W2.shape[0], Wp.shape[0]
        hn = torch.empty(B, D, device=dev, dtype=F32)
        stats = torch.empty(B, 2, device=dev, dtype=F32)
        logits = torch.empty(B, nc, device=dev, dtype=F32)
        enc = torch.empty(B, De, device=dev, dtype=F32)
        call("dml_tower_head_fwd", ptr(x0), D, B, D, ptr(ws[0]), ptr(ws[1]), float(eps), ptr(ws[2]), ptr(ws[3]), nc, ptr(ws[4]),
             ptr(ws[5]), De, ptr(hn), ptr(stats), ptr(logits), ptr(enc), stream())
        ctx.shape = (B, n, D, nc, De)
        ctx.save_for_backward(x0, hn, stats, ws[0], ws[2], ws[4])
        return enc, logits

    @staticmethod
    def backward(ctx, denc, dlogits):
        x0, hn, stats, lw, W2, Wp = ctx.saved_tensors
        B, n, D, nc, De = ctx.shape
        dev = x0.device
        dx = torch.zeros(B, n, D, device=dev, dtype=F32)           # only the cls row receives a gradient
        dparams = torch.empty(2 * D + nc * D + nc + De * D + De, device=dev, dtype=F32)
        denc = denc.contiguous().float() if denc is not None else None
        dlogits = dlogits.contiguous().float() if dlogits is not None else None
        call("dml_tower_head_bwd", ptr(x0), D, B, D, ptr(lw), ptr(W2), nc, ptr(Wp), De, ptr(hn), ptr(stats),
             ptr(dlogits) if dlogits is not None else None, ptr(denc) if denc is not None else None, ptr(dparams), ptr(dx),
             n * D, stream())
        o = 0
        dlw, o = dparams[o: o + D], o + D
        dlb, o = dparams[o: o + D], o + D
        dW2, o = dparams[o: o + nc * D].view(nc, D), o + nc * D
        db2, o = dparams[o: o + nc], o + nc
        dWp, o = dparams[o: o + De * D].view(De, D), o + De * D
        dbp = dparams[o: o + De]
        return dx, dlw, dlb, dW2, db2, dWp, dbp, None


class Linear3Fn(torch.autograd.Function):
    """The three classifiers of DeformPathomicNet (models/model.py:535-558) as one kernel per direction: yc = act(Wc cat(a, b) +
    bc), ya = act(Wa a + ba), yb = act(Wb b + bb); act = sigmoid for the survival task."""

    @staticmethod
    def forward(ctx, a, b, Wc, bc, Wa, ba, Wb, bb, sigmoid):
        a, b = a.contiguous().float(), b.contiguous().float()
        ws = [t.contiguous().float() for t in (Wc, bc, Wa, ba, Wb, bb)]
        B, Da, Db, nc = a.shape[0], a.shape[1], b.shape[1], Wc.shape[0]
        ys = [torch.empty(B, nc, device=a.device, dtype=F32) for _ in range(3)]
        call("dml_linear3_fwd", ptr(a), ptr(b), B, Da, Db, *[ptr(t) for t in ws], nc, int(bool(sigmoid)), *[ptr(y) for y in ys], stream())
        ctx.meta = (B, Da, Db, nc, int(bool(sigmoid)))
        ctx.save_for_backward(a, b, ws[0], ws[2], ws[4], *ys)
        return tuple(ys)

    @staticmethod
    def backward(ctx, gyc, gya, gyb):
        a, b, Wc, Wa, Wb, yc, ya, yb = ctx.saved_tensors
        B, Da, Db, nc, sig = ctx.meta
        dev = a.device
        gs = [g.contiguous().float() if g is not None else None for g in (gyc, gya, gyb)]
        dparams = torch.empty(nc * (Da + Db) + nc + nc * Da + nc + nc * Db + nc, device=dev, dtype=F32)
        da = torch.empty(B, Da, device=dev, dtype=F32)
        db = torch.empty(B, Db, device=dev, dtype=F32)
        call("dml_linear3_bwd", ptr(a), ptr(b), B, Da, Db, ptr(Wc), ptr(Wa), ptr(Wb), nc, sig, ptr(yc), ptr(ya), ptr(yb),
             *[ptr(g) if g is not None else None for g in gs], ptr(dparams), ptr(da), ptr(db), stream())
        o = 0
        dWc, o = dparams[o: o + nc * (Da + Db)].view(nc, Da + Db), o + nc * (Da + Db)
        dbc, o = dparams[o: o + nc], o + nc
        dWa, o = dparams[o: o + nc * Da].view(nc, Da), o + nc * Da
        dba, o = dparams[o: o + nc], o + nc
        dWb, o = dparams[o: o + nc * Db].view(nc, Db), o + nc * Db
        dbb = dparams[o: o + nc]
        return da, db, dWc, dbc, dWa, dba, dWb, dbb, None


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
