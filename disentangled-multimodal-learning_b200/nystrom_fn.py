"""Forward and backward of NystromAttention (reference: models/NystromAttention.py:20-157 == models/cmta_utils.py:147-281) as
ONE autograd function over the sm_100a kernels: every contraction on the bf16-pair tcgen05 GEMM (`dml_pgemm`), landmark
pooling / long-row softmax / value convolution on the pair HBM kernels (csrc/nystrom_pair.cu).

What is fused where (per layer; B bags, n tokens, n_pad = front-padded length, H heads of width d, m landmarks):
  to_qkv        one GEMM whose A map starts `pad` rows before the tokens (TMA zero fill = the front padding, :79-85, no padded
                copy); the epilogue scales the q columns (:98) and writes the qkv PAIR (no fp32 copy, no split pass)
  sim1, sim2    GEMM + row softmax in the epilogue (:123-124,137): attn1 / attn2 leave the kernel as pairs
  sim3          GEMM -> long-row softmax kernel (:125,137); attn3 @ v as a split-K GEMM over the tokens
  pinv          6 x 4 GEMMs (:31-33) with the polynomial terms `c I - (.)` in the epilogues, operands stay pairs;
                xz (7 I - xz) is evaluated as 7 xz - xz xz (one product, the 7 xz term enters as the epilogue's residual)
  aggregation   attn1 @ (attn2_inv @ (attn3 @ v)) (:140 re-associated: the [n, m] x [m, m] product is never formed), written
                straight into the head-merged layout; + res_conv(v) (:144-145) in the kernel that writes the to_out operand
  to_out        GEMM + bias on the last n rows only (:150-151)
The backward is the hand-derived adjoint of exactly these steps (softmax backward of sim1 in the GEMM epilogue, the pinv
recurrence reversed with 8 GEMMs per step from the saved iterates, every weight gradient a split-K GEMM over the tokens).
"""
from __future__ import annotations

from math import ceil

import os

import torch

from . import _lib
from ._lib import call, ptr, stream
from .pairs import Pair, chain, pgemm

F32 = torch.float32


def _ksplits(K: int, tiles: int) -> int:
    """Split-K factor for a token-reduction product with `tiles` output tiles: fill ~148 SMs, at least 8 k-blocks per part."""
    kb = (K + 63) // 64
    return max(1, min(kb // 8, (148 + tiles - 1) // max(tiles, 1)))


def _mm_tokens(A, Bm, out, *, M, N, K, batch, reduce_into_2d=False, **kw):
    """out (zeroed here) += A^T-form product over the token axis, split-K."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    nb = 1
    for b_ in batch:
        nb *= b_
    s = _ksplits(K, tiles * (1 if reduce_into_2d else nb))
    if s == 1 and not reduce_into_2d:
        pgemm(A, Bm, M=M, N=N, K=K, batch=batch, out=out, **kw)
    else:
        out.zero_()
        pgemm(A, Bm, M=M, N=N, K=K, batch=batch, out=out, splits=max(s, 2), **kw)
    return out


PINV_TWO_STREAMS = os.environ.get("DML_B200_PINV_TWO_STREAMS", "1") != "0"


def _tls_chain():
    from . import pairs
    return pairs._tls


def CHAIN_DEFAULT_ON():
    from . import pairs
    return pairs.CHAIN_DEFAULT


def _pinv_side(dev):
    """The auxiliary stream for the independent products of a pseudo-inverse step (None: everything on the current stream -
    the chained cooperative launches are one stream by construction)."""
    from .ops import side_stream
    if not PINV_TWO_STREAMS or getattr(_tls_chain(), "chain", None) is not None or CHAIN_DEFAULT_ON():
        return None
    return side_stream(dev)


NY_BRANCH = os.environ.get("DML_B200_NY_BRANCH", "1") != "0"


def _branch(dev):
    from .ops import branch_stream
    return branch_stream(dev) if NY_BRANCH else None


class _on:
    """Run the block on `s` (a no-op for None); ordering is the caller's business (events / wait_stream)."""

    def __init__(self, s):
        self.s = s

    def __enter__(self):
        if self.s is not None:
            self.ctx = torch.cuda.stream(self.s)
            self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        if self.s is not None:
            self.ctx.__exit__(*a)
        return False


def pinv_forward(x_pair: Pair, x_f32: torch.Tensor, iters: int, B: int, H: int, m: int, save: bool):
    """moore_penrose_iter_pinv (NystromAttention.py:20-35).  Returns (z pair, saved iterates).

    One step z' = 1/4 z (13 I - P (15 I - P (7 I - P))), P = x z, is evaluated as
        P = x z;   t3 = 15 I - (7 P - P P)  ||  Bz = z P;   z' = 13/4 z - 1/4 Bz t3
    (z (13 I - P t3) = 13 z - (z P) t3): the same four products, but three deep instead of four - the two middle ones are
    independent and run side by side on two streams (each occupies 64 of the 148 SMs).  Each product costs ~10 us of launch /
    load / epilogue latency whatever its size, so the depth of the chain is what the pseudo-inverse costs."""
    bt = (B, H)
    lib = _lib.load(check_device=True)
    x_f32 = x_f32.contiguous()
    dev = x_f32.device
    shape = (B, H, m, m)
    sums = torch.empty(lib.dml_ny_pinv_init_sums_floats(B * H, m), device=dev, dtype=F32)
    z = Pair.empty(shape, dev)
    # z0 = x^T / (max row abs-sum * max column abs-sum), both maxima GLOBAL over batch and heads (quirk T3): two launches
    call("dml_ny_pinv_init_fwd", ptr(x_f32), B * H, m, ptr(sums), ptr(z.planes), z.planes.stride(0), stream())
    z_f = z.float()                                          # the 13/4 z term enters the last product's epilogue in fp32
    saved = [sums]
    with chain():        # DML_B200_PGEMM_CHAIN=1: the dependent products leave as chained cooperative launches instead
        side = _pinv_side(dev)
        cur = torch.cuda.current_stream()
        for _ in range(iters):
            xz_f, xz = pgemm(x_pair, z, M=m, N=m, K=m, b_trans=True, batch=bt, want_pair=True)
            t3, zp = Pair.empty(shape, dev), Pair.empty(shape, dev)          # both allocated on the calling stream
            if side is not None:
                side.wait_stream(cur)
            with _on(side):
                pgemm(z, xz, M=m, N=m, K=m, b_trans=True, batch=bt, want_f32=False, pair_out=zp)                  # z P
            # t3 = 15 I - xz (7 I - xz) = 15 I - (7 xz - xz xz)
            pgemm(xz, xz, M=m, N=m, K=m, b_trans=True, batch=bt, alpha=-1.0, resid=xz_f, resid_scale=7.0, diag=15.0,
                  want_f32=False, pair_out=t3)
            if side is not None:
                cur.wait_stream(side)
            z_f_new, z_new = pgemm(zp, t3, M=m, N=m, K=m, b_trans=True, batch=bt, alpha=-0.25, resid=z_f, resid_scale=3.25,
                                   want_pair=True)
            if save:
                saved.append((z, xz, t3, zp))
            z, z_f = z_new, z_f_new
    return z, saved


def pinv_backward(x_pair: Pair, x_f32: torch.Tensor, saved, G_f: torch.Tensor, G: Pair, B: int, H: int, m: int):
    """Adjoint of pinv_forward: G = d z_final (fp32 + pair) -> d x (fp32 [B, H, m, m]).

    Per step (z' = 13/4 z - 1/4 Bz t3, Bz = z P, t3 = 15 I - 7 P + P P, P = x z), eight products four deep on two streams:
        main   a  dB  = -1/4 G t3^T        d  dP  = z^T dB              f  dP += P^T dt3 + dPe      h  G' = dz + x^T dP
        side   b  dt3 = -1/4 Bz^T G        e  dPe = dt3 P^T - 7 dt3     c  dz  = 13/4 G + dB P^T    g  dx += dP z^T
    (c waits for a, f for e, g for f, h for c).  Every buffer is allocated on the calling stream."""
    bt = (B, H)
    dx = None
    sums, saved = saved[0], saved[1:]
    dev = x_f32.device
    shape = (B, H, m, m)
    with chain():
        side = _pinv_side(dev)
        cur = torch.cuda.current_stream()

        def mark(s):
            if side is None:
                return None
            ev = torch.cuda.Event()
            ev.record(s)
            return ev

        def after(s, ev):
            if ev is not None:
                s.wait_event(ev)

        for (z, P, t3, zp) in reversed(saved):
            first = dx is None
            dB, dt3, dPp = Pair.empty(shape, dev), Pair.empty(shape, dev), Pair.empty(shape, dev)
            dt3_f, dPe, dz_f, dP = (torch.empty(shape, device=dev, dtype=F32) for _ in range(4))
            if first:
                dx = torch.empty(shape, device=dev, dtype=F32)
            if side is not None:
                side.wait_stream(cur)
            with _on(side):
                pgemm(zp, G, M=m, N=m, K=m, a_trans=True, b_trans=True, batch=bt, alpha=-0.25, out=dt3_f, pair_out=dt3)      # b
                pgemm(dt3, P, M=m, N=m, K=m, batch=bt, resid=dt3_f, resid_scale=-7.0, out=dPe)                              # e
            ev_e = mark(side)
            pgemm(G, t3, M=m, N=m, K=m, batch=bt, alpha=-0.25, want_f32=False, pair_out=dB)                                 # a
            ev_a = mark(cur)
            pgemm(z, dB, M=m, N=m, K=m, a_trans=True, b_trans=True, batch=bt, out=dP)                                       # d
            with _on(side):
                after(side, ev_a)
                pgemm(dB, P, M=m, N=m, K=m, batch=bt, resid=G_f, resid_scale=3.25, out=dz_f)                                # c
            ev_c = mark(side)
            after(cur, ev_e)
            pgemm(P, dt3, M=m, N=m, K=m, a_trans=True, b_trans=True, batch=bt, resid=dPe, out=dP, accumulate=True,
                  pair_out=dPp)                                                                                             # f
            ev_f = mark(cur)
            with _on(side):
                after(side, ev_f)
                pgemm(dPp, z, M=m, N=m, K=m, batch=bt, out=dx, accumulate=not first)                                        # g
            after(cur, ev_c)
            G_f, G = pgemm(x_pair, dPp, M=m, N=m, K=m, a_trans=True, b_trans=True, batch=bt, out=dz_f, accumulate=True,
                           want_pair=True)                                                                                  # h
            if side is not None:
                cur.wait_stream(side)          # the step's operands may be released / overwritten from here on
    # z0 = x^T / (max row-sum * max column-sum): the scalar couples all bags and heads; two launches, the chain's dx added in
    lib = _lib.load(check_device=True)
    x_f32 = x_f32.contiguous()
    G_f = G_f.contiguous()
    part = torch.empty(lib.dml_ny_pinv_init_part_floats(B * H, m), device=x_f32.device, dtype=F32)
    out = torch.empty_like(x_f32)
    call("dml_ny_pinv_init_bwd", ptr(G_f), ptr(x_f32), ptr(sums), None if dx is None else ptr(dx.contiguous()), B * H, m, ptr(part),
         ptr(out), stream())
    return out


class NystromAttnFn(torch.autograd.Function):
    """y [B, n, dim] = to_out(aggregate(softmax kernels of to_qkv(x))) with the layer's optional LayerNorm in front.

    x: fp32 [B, n, dim].  ln_w / ln_b: LayerNorm parameters or None (TransLayer fuses its norm: the normalised rows only ever
    exist as the GEMM's operand pair).  Wqkv [3 H d, dim], Wout [dim, H d], bout [dim], wconv [H, 1, K, 1] or None."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, Wqkv, Wout, bout, wconv, cfg):
        H, d, m, iters, ln_eps = cfg
        B, n, dim = x.shape
        W = H * d
        dev = x.device
        scale = d ** -0.5
        rem = n % m
        pad = (m - rem) if rem > 0 else 0
        n_pad = n + pad
        l = ceil(n / m)
        if m > 256 or (m % 8) or (d % 8) or (dim % 8):
            raise _lib.DmlError(f"NystromAttention: num_landmarks={m}, dim_head={d}, dim={dim} outside the kernels' range "
                                "(landmarks <= 256, multiples of 8)")
        st = stream()
        x = x.contiguous().float()
        if ln_w is not None:
            xn = Pair.empty((B, n, dim), dev)
            mean = torch.empty(B * n, device=dev, dtype=F32)
            rstd = torch.empty_like(mean)
            call("dml_layernorm_fwd_pair", ptr(x), ptr(ln_w), ptr(ln_b), B * n, dim, float(ln_eps), None, ptr(xn.planes),
                 xn.planes.stride(0), ptr(mean), ptr(rstd), st)
        else:
            xn, mean, rstd = Pair.from_f32(x), None, None
        Wqkv_p, Wout_p = Pair.from_f32(Wqkv), Pair.from_f32(Wout)

        # to_qkv into the front-padded pair buffer; q columns scaled (:79-98)
        qkv = Pair.empty((B, n_pad, 3 * W), dev)
        pgemm(xn, Wqkv_p.b1(), M=n_pad, N=3 * W, K=dim, batch=(B,), a_row_offset=-pad, ncol_split=W, alpha2=scale,
              want_f32=False, pair_out=qkv)
        q_h, k_h, v_h = (qkv.head_slices(i, 3, H, d) for i in range(3))
        lm = Pair.empty((2, B, H, m, d), dev)                                       # q_l, k_l (:102-118)
        call("dml_ny_landmark_pool", ptr(qkv.planes), qkv.planes.stride(0), 3 * W, B, n_pad, l, H, d, 1.0 / l, 1.0 / l,
             ptr(lm.planes), lm.planes.stride(0), st)
        q_l, k_l = Pair(lm.planes[:, 0]), Pair(lm.planes[:, 1])
        bt = (B, H)
        attn2_f, attn2 = pgemm(q_l, k_l, M=m, N=m, K=d, batch=bt, softmax=1, want_pair=True)                    # :124,137
        # Two independent branches from here: the pseudo-inverse of attn2 - a chain of 18 small dependent products, ~10 us of
        # latency each, on 64-128 SMs - and the token-sized kernels (attn1, sim3 -> attn3, attn3 @ v).  The second branch runs
        # on an auxiliary stream next to the chain; its buffers are allocated here, on the calling stream, and live until the
        # streams have joined.
        attn1, attn3 = Pair.empty((B, H, n_pad, m), dev), Pair.empty((B, H, m, n_pad), dev)
        sim3 = torch.empty(B, H, m, n_pad, device=dev, dtype=F32)
        T_f = torch.empty(B, H, m, d, device=dev, dtype=F32)
        cur = torch.cuda.current_stream()
        br = _branch(dev)
        if br is not None:
            br.wait_stream(cur)
        with _on(br):
            pgemm(q_h, k_l, M=n_pad, N=m, K=d, batch=bt, softmax=1, want_f32=False, pair_out=attn1)             # :123,137
            pgemm(q_l, k_h, M=m, N=n_pad, K=d, batch=bt, out=sim3)                                              # :125
            call("dml_ny_softmax_rows_fwd", ptr(sim3), B * H * m, n_pad, ptr(attn3.planes), attn3.planes.stride(0), stream())
            _mm_tokens(attn3, v_h, T_f, M=m, N=d, K=n_pad, batch=bt, b_trans=True)                              # attn3 @ v
        z, saved = pinv_forward(attn2, attn2_f, iters, B, H, m, save=True)                                      # :138
        if br is not None:
            cur.wait_stream(br)
        del sim3
        T = Pair.from_f32(T_f)
        _, Wm = pgemm(z, T, M=m, N=d, K=m, b_trans=True, batch=bt, want_f32=False, want_pair=True)              # attn2_inv @ (.)
        agg = torch.empty(B, n_pad, W, device=dev, dtype=F32)
        pgemm(attn1, Wm, M=n_pad, N=d, K=m, b_trans=True, batch=bt, out=agg.view(B, n_pad, H, d).permute(0, 2, 1, 3))   # :140
        if wconv is not None:
            wf = wconv.reshape(H, -1).contiguous().float()
            om = Pair.empty((B, n_pad, W), dev)
            call("dml_ny_res_conv_fwd", ptr(agg), ptr(qkv.planes), qkv.planes.stride(0), 3 * W, 2 * W, ptr(wf), wf.shape[1], B,
                 n_pad, H, d, ptr(om.planes), om.planes.stride(0), st)                                          # :144-145
        else:
            wf, om = None, Pair.from_f32(agg)
        del agg
        y, _ = pgemm(om, Wout_p.b1(), M=n, N=dim, K=W, batch=(B,), a_row_offset=pad, bias=bout.contiguous().float())   # :150-151
        ctx.cfg = (H, d, m, iters, B, n, dim, pad, n_pad, l, scale)
        ctx.has_ln, ctx.has_conv = ln_w is not None, wconv is not None
        ctx.wshape = tuple(wconv.shape) if wconv is not None else None
        ctx.saved_pairs = (xn, Wqkv_p, Wout_p, qkv, lm, attn1, attn2, attn3, z, T, Wm, om, saved)
        ctx.save_for_backward(x, ln_w, mean, rstd, attn2_f, wf)
        return y

    @staticmethod
    def backward(ctx, dy):
        H, d, m, iters, B, n, dim, pad, n_pad, l, scale = ctx.cfg
        x, ln_w, mean, rstd, attn2_f, wf = ctx.saved_tensors
        xn, Wqkv_p, Wout_p, qkv, lm, attn1, attn2, attn3, z, T, Wm, om, saved = ctx.saved_pairs
        W = H * d
        dev = dy.device
        st = stream()
        bt = (B, H)
        q_h, k_h, v_h = (qkv.head_slices(i, 3, H, d) for i in range(3))
        q_l, k_l = Pair(lm.planes[:, 0]), Pair(lm.planes[:, 1])
        dy = dy.contiguous().float()
        dyp = Pair.from_f32(dy)
        dbout = torch.empty(dim, device=dev, dtype=F32)
        call("dml_colsum", ptr(dy), B * n, dim, dim, ptr(dbout), st)
        # to_out: y = om[pad:] Wout^T + b
        dWout = torch.empty(dim, W, device=dev, dtype=F32)
        _mm_tokens(dyp, om, dWout, M=dim, N=W, K=n, batch=(B,), reduce_into_2d=True, a_trans=True, b_trans=True, b_k_offset=pad)
        dOm_f, dOm = pgemm(dyp, Wout_p.b1(), M=n_pad, N=W, K=dim, batch=(B,), a_row_offset=-pad, b_trans=True, want_pair=True)
        acc = torch.empty(B, n_pad, 3 * W, device=dev, dtype=F32)       # gradients of (scaled q, k, v), every element written below
        acc_h = [acc.view(B, n_pad, 3, H, d)[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
        if ctx.has_conv:
            dwconv = torch.empty(H, wf.shape[1], device=dev, dtype=F32)
            call("dml_ny_res_conv_bwd", ptr(dOm_f), ptr(qkv.planes), qkv.planes.stride(0), 3 * W, 2 * W, ptr(wf), wf.shape[1], B,
                 n_pad, H, d, ptr(acc), 3 * W, 2 * W, ptr(dwconv), st)
        else:
            dwconv = None
            acc_h[2].zero_()
        del dOm_f
        dO_h = Pair(dOm.planes.view(2, B, n_pad, H, d).permute(0, 1, 3, 2, 4))
        # O = attn1 Wm
        _, dS1 = pgemm(dO_h, Wm, M=n_pad, N=m, K=d, batch=bt, softmax=2, aux=attn1, want_f32=False, want_pair=True)
        dWm_f = torch.empty(B, H, m, d, device=dev, dtype=F32)
        _mm_tokens(attn1, dO_h, dWm_f, M=m, N=d, K=n_pad, batch=bt, a_trans=True, b_trans=True)
        dWm = Pair.from_f32(dWm_f)
        # sim1 = q k_l^T
        pgemm(dS1, k_l, M=n_pad, N=d, K=m, b_trans=True, batch=bt, out=acc_h[0])
        dl = torch.empty(2, B, H, m, d, device=dev, dtype=F32)          # d q_l, d k_l
        _mm_tokens(dS1, q_h, dl[1], M=m, N=d, K=n_pad, batch=bt, a_trans=True, b_trans=True)
        del dS1
        # Wm = z T
        dZ_f, dZ = pgemm(dWm, T, M=m, N=m, K=d, batch=bt, want_pair=True)
        _, dT = pgemm(z, dWm, M=m, N=d, K=m, a_trans=True, b_trans=True, batch=bt, want_f32=False, want_pair=True)
        # T = attn3 v ; attn3 = softmax(q_l k^T)
        # the token-sized kernels below do not need the pseudo-inverse adjoint (a chain of 24 small dependent products): they run
        # on the auxiliary stream next to it; buffers allocated here, on the calling stream, and released after the join
        dA3 = torch.empty(B, H, m, n_pad, device=dev, dtype=F32)
        dS3 = Pair.empty((B, H, m, n_pad), dev)
        cur = torch.cuda.current_stream()
        br = _branch(dev)
        if br is not None:
            br.wait_stream(cur)
        with _on(br):
            pgemm(dT, v_h, M=m, N=n_pad, K=d, batch=bt, out=dA3)
            call("dml_ny_softmax_rows_bwd", ptr(attn3.planes), attn3.planes.stride(0), ptr(dA3), B * H * m, n_pad, ptr(dS3.planes),
                 dS3.planes.stride(0), stream())
            pgemm(attn3, dT, M=n_pad, N=d, K=m, a_trans=True, b_trans=True, batch=bt, out=acc_h[2], accumulate=True)
            _mm_tokens(dS3, k_h, dl[0], M=m, N=d, K=n_pad, batch=bt, b_trans=True)
            pgemm(dS3, q_l, M=n_pad, N=d, K=m, a_trans=True, b_trans=True, batch=bt, out=acc_h[1])
        # pinv and sim2 = softmax(q_l k_l^T)
        dA2 = pinv_backward(attn2, attn2_f, saved, dZ_f, dZ, B, H, m)
        if br is not None:
            cur.wait_stream(br)
        del dA3, dS3
        dS2_f = attn2_f * (dA2 - (dA2 * attn2_f).sum(-1, keepdim=True))
        dS2 = Pair.from_f32(dS2_f)
        pgemm(dS2, k_l, M=m, N=d, K=m, b_trans=True, batch=bt, out=dl[0], accumulate=True)
        pgemm(dS2, q_l, M=m, N=d, K=m, a_trans=True, b_trans=True, batch=bt, out=dl[1], accumulate=True)
        # d(qkv) as a pair, then to_qkv
        dqkv = Pair.empty((B, n_pad, 3 * W), dev)
        call("dml_ny_dqkv_finalize", ptr(acc), ptr(dl), B, n_pad, l, H, d, float(scale), ptr(dqkv.planes), dqkv.planes.stride(0), st)
        del acc
        dxn, _ = pgemm(dqkv, Wqkv_p.b1(), M=n, N=dim, K=3 * W, batch=(B,), a_row_offset=pad, b_trans=True)
        dWqkv = torch.empty(3 * W, dim, device=dev, dtype=F32)
        _mm_tokens(dqkv, xn, dWqkv, M=3 * W, N=dim, K=n, batch=(B,), reduce_into_2d=True, a_trans=True, a_k_offset=pad, b_trans=True)
        if ctx.has_ln:
            dx = torch.empty_like(x)
            dlw = torch.empty(dim, device=dev, dtype=F32)
            dlb = torch.empty_like(dlw)
            call("dml_layernorm_bwd", ptr(dxn), ptr(x), ptr(ln_w), ptr(mean), ptr(rstd), B * n, dim, ptr(dx), ptr(dlw), ptr(dlb), st)
        else:
            dx, dlw, dlb = dxn, None, None
        return (dx, dlw, dlb, dWqkv, dWout, dbout, dwconv.reshape(ctx.wshape) if dwconv is not None else None, None)


class PPEGFn(torch.autograd.Function):
    """PPEG (models/mil.py:192-206) as one 7x7 depthwise stencil: x [B, 1 + side^2, C], wsum [C, 49], bsum [C]."""

    @staticmethod
    def forward(ctx, x, wsum, bsum, side):
        x = x.contiguous().float()
        B, _, C = x.shape
        wsum, bsum = wsum.contiguous().float(), bsum.contiguous().float()
        y = torch.empty_like(x)
        call("dml_ppeg_stencil", ptr(x), ptr(wsum), ptr(bsum), B, side, C, 0, ptr(y), stream())
        ctx.save_for_backward(x, wsum, bsum)
        ctx.side = side
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wsum, bsum = ctx.saved_tensors
        B, _, C = x.shape
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        call("dml_ppeg_stencil", ptr(dy), ptr(wsum), ptr(bsum), B, ctx.side, C, 1, ptr(dx), stream())
        dw = torch.empty(C, 49, device=x.device, dtype=F32)
        db = torch.empty(C, device=x.device, dtype=F32)
        call("dml_ppeg_wgrad", ptr(x), ptr(dy), B, ctx.side, C, ptr(dw), ptr(db), stream())
        return dx, dw, db, None
