"""torch.autograd.Function wrappers of the DeformCrossAttention2D / ClusterMergeNet kernels (csrc/deform2d.cu,
deform2d_bias.cu, cluster.cu).  Host code allocates buffers and orders launches on torch's current stream."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream

F32 = torch.float32
BIAS_GRAD_FLOATS = 1192     # DML_DA2_BIAS_GRAD_FLOATS
OFF_GRAD_FLOATS = 2496


def kv_side(side: int, ks: int, stride: int) -> int:
    return int(_lib.load().dml_da2_kv_side(side, ks, stride))


def _gproj(x, W, rows):
    y = torch.empty(rows, 512, device=x.device, dtype=F32)
    call("dml_da2_gproj_fwd", ptr(x), ptr(W), rows, ptr(y), stream())
    return y


def _gproj_bwd(dy, x, W, rows, dx, accumulate):
    lib = _lib.load()
    parts = torch.empty(lib.dml_da2_gproj_parts(rows), 8192, device=dy.device, dtype=F32)
    dW = torch.empty(512, 16, device=dy.device, dtype=F32)
    call("dml_da2_gproj_bwd", ptr(dy), ptr(x), ptr(W), rows, int(accumulate), ptr(dx) if dx is not None else None, ptr(parts), ptr(dW),
         stream())
    return dW


class DeformCrossAttn2DFn(torch.autograd.Function):
    """DeformCrossAttention2D.forward up to (not including) to_out (DeformableAttention2D.py:241-321) on token-major fp32
    inputs x1t, x2t [B, n, 128], n = side^2.  Returns (o [B, n, 512], attn [B, 8, n, m], vgrid [(B 8), 2, hk, hk]).
    Gradients may arrive at all three outputs (the teacher / student losses read attn and vgrid, utils/loss.py)."""

    @staticmethod
    def forward(ctx, x1t, x2t, Wq, Wk, Wv, wdw, bdw, w2, m_W1, m_b1, m_W2, m_b2, m_W3, m_b3, cfg):
        side, ks, stride, offset_scale, drop_p, training = cfg
        B, n, dim = x1t.shape
        if dim != 128 or side * side != n:
            raise _lib.DmlError("DeformCrossAttention2D kernels: dim 128 and a square token grid (DeformableAttention2D.py:241)")
        dev = x1t.device
        st = stream()
        hk = kv_side(side, ks, stride)
        if hk < 1:
            raise _lib.DmlError(f"a {side} x {side} grid is too small for offset kernel {ks} / stride {stride}")
        m = hk * hk
        x1f, x2f = x1t.contiguous().float(), x2t.contiguous().float()
        Wqf, Wkf, Wvf = (t.reshape(512, 16).contiguous().float() for t in (Wq, Wk, Wv))
        offw = [wdw.reshape(64, ks * ks).contiguous().float(), bdw.contiguous().float(), w2.reshape(2, 64).contiguous().float()]
        mlp = [t.contiguous().float() for t in (m_W1, m_b1, m_W2, m_b2, m_W3.reshape(-1), m_b3)]
        q = _gproj(x1f, Wqf, B * n)
        vgrid = torch.empty(B * 8, 2, hk, hk, device=dev, dtype=F32)
        vs = torch.empty(B * 8, m, 2, device=dev, dtype=F32)
        call("dml_da2_offsets_fwd", ptr(q), *[ptr(t) for t in offw], B, side, ks, stride, float(offset_scale), ptr(vgrid), ptr(vs), st)
        kvf = torch.empty(B, m, 128, device=dev, dtype=F32)
        call("dml_da2_gather_fwd", ptr(x2f), ptr(vs), B, side, m, ptr(kvf), st)
        k = _gproj(kvf, Wkf, B * m)
        v = _gproj(kvf, Wvf, B * m)
        attn = torch.empty(B, 8, n, m, device=dev, dtype=F32)
        call("dml_da2_bias_fwd", ptr(vs), *[ptr(t) for t in mlp], B, side, m, ptr(attn), st)
        keep, keep_scale = None, 1.0
        if training and drop_p > 0.0:
            keep = (torch.rand(B, 8, n, m, device=dev) >= drop_p).to(torch.uint8)
            keep_scale = 1.0 / (1.0 - drop_p)
        o = torch.empty(B, n, 512, device=dev, dtype=F32)
        ws = torch.empty(_lib.load().dml_da2_attn_ws_bytes(B, n, m, 0), device=dev, dtype=torch.uint8)
        call("dml_da2_attn_fwd", ptr(q), ptr(k), ptr(v), ptr(attn), ptr(keep) if keep is not None else None, keep_scale, B, n, m,
             64 ** -0.5, ptr(ws), ptr(o), st)
        ctx.set_materialize_grads(False)
        ctx.cfg = (side, ks, stride, float(offset_scale), keep_scale, hk)
        ctx.keep = keep
        ctx.save_for_backward(x1f, x2f, Wqf, Wkf, Wvf, *offw, *mlp, q, vs, kvf, k, v, attn)
        return o, attn, vgrid

    @staticmethod
    def backward(ctx, do, dattn, dvgrid):
        (x1f, x2f, Wqf, Wkf, Wvf, wdw, bdw, w2, W1, b1, W2, b2, W3, b3, q, vs, kvf, k, v, attn) = ctx.saved_tensors
        side_len, ks, stride, offset_scale, keep_scale, hk = ctx.cfg
        keep = ctx.keep
        lib = _lib.load()
        B, n, _ = x1f.shape
        m = hk * hk
        dev = x1f.device
        st = stream()
        do = torch.zeros(B, n, 512, device=dev, dtype=F32) if do is None else do.contiguous().float()
        dA = dattn.contiguous().float() if dattn is not None else None
        ds = torch.empty_like(attn)
        dq = torch.empty(B, n, 512, device=dev, dtype=F32)
        dkv = torch.empty(2, B, m, 512, device=dev, dtype=F32)
        parts = torch.empty(lib.dml_da2_cols_chunks(B, n, m), 2, B, m, 512, device=dev, dtype=F32)
        ws = torch.empty(lib.dml_da2_attn_ws_bytes(B, n, m, 1), device=dev, dtype=torch.uint8)
        call("dml_da2_attn_bwd", ptr(q), ptr(k), ptr(v), ptr(attn), ptr(do), ptr(dA) if dA is not None else None,
             ptr(keep) if keep is not None else None, keep_scale, B, n, m, 64 ** -0.5, ptr(ws), ptr(ds), ptr(dq), ptr(parts), ptr(dkv), st)
        # keys / values -> gathered features: small launches that only need dk / dv, on the auxiliary stream next to the
        # position-bias backward (buffers are allocated on the main stream first)
        from .ops import side_stream
        dkvf = torch.empty(B, m, 128, device=dev, dtype=F32)
        dx2 = torch.zeros(B, n, 128, device=dev, dtype=F32)
        kparts = [torch.empty(lib.dml_da2_gproj_parts(B * m), 8192, device=dev, dtype=F32) for _ in range(2)]
        dWk, dWv = torch.empty(512, 16, device=dev, dtype=F32), torch.empty(512, 16, device=dev, dtype=F32)
        cur, side = torch.cuda.current_stream(), side_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            s2 = stream()
            call("dml_da2_gproj_bwd", ptr(dkv[0]), ptr(kvf), ptr(Wkf), B * m, 0, ptr(dkvf), ptr(kparts[0]), ptr(dWk), s2)
            call("dml_da2_gproj_bwd", ptr(dkv[1]), ptr(kvf), ptr(Wvf), B * m, 1, ptr(dkvf), ptr(kparts[1]), ptr(dWv), s2)
        # position-bias MLP: parameter gradients and the gradient at the sampling positions
        dvs = torch.zeros(B * 8, m, 2, device=dev, dtype=F32)
        bparts = torch.empty(lib.dml_da2_bias_bwd_parts(B, side_len), BIAS_GRAD_FLOATS, device=dev, dtype=F32)
        bg = torch.empty(BIAS_GRAD_FLOATS, device=dev, dtype=F32)
        call("dml_da2_bias_bwd", ptr(vs), ptr(W1), ptr(b1), ptr(W2), ptr(b2), ptr(W3), ptr(ds), B, side_len, m, ptr(bparts), ptr(bg), ptr(dvs), st)
        cur.wait_stream(side)
        call("dml_da2_gather_bwd", ptr(dkvf), ptr(x2f), ptr(vs), B, side_len, m, ptr(dx2), ptr(dvs), st)
        # offset net (adds its share to dq), then the query projection
        dconv = torch.empty(B * 8, m, 64, device=dev, dtype=F32)
        oparts = torch.empty(lib.dml_da2_offsets_parts(B, side_len, ks, stride), OFF_GRAD_FLOATS, device=dev, dtype=F32)
        og = torch.empty(OFF_GRAD_FLOATS, device=dev, dtype=F32)
        dvg = dvgrid.contiguous().float() if dvgrid is not None else None
        call("dml_da2_offsets_bwd", ptr(q), ptr(wdw), ptr(bdw), ptr(w2), ptr(dvs), ptr(dvg) if dvg is not None else None, B, side_len, ks,
             stride, offset_scale, ptr(dconv), ptr(oparts), ptr(og), ptr(dq), st)
        dx1 = torch.empty(B, n, 128, device=dev, dtype=F32)
        dWq = _gproj_bwd(dq, x1f, Wqf, B * n, dx1, False)
        taps = ks * ks
        dwdw, dbdw, dw2 = og[:64 * taps].view(64, 1, ks, ks), og[64 * taps:64 * taps + 64], og[64 * taps + 64:64 * taps + 192].view(2, 64, 1, 1)
        dW1, db1, dW2, db2 = bg[0:64].view(32, 2), bg[64:96], bg[96:1120].view(32, 32), bg[1120:1152]
        dW3, db3 = bg[1152:1184].view(1, 32), bg[1184:1185]
        return (dx1, dx2, dWq.view(512, 16, 1, 1), dWk.view(512, 16, 1, 1), dWv.view(512, 16, 1, 1), dwdw, dbdw, dw2,
                dW1, db1, dW2, db2, dW3, db3, None)


# ---------------------------------------------------------------------------------------------------------------------
# ClusterMergeNet
# ---------------------------------------------------------------------------------------------------------------------
def dpc_knn(x: torch.Tensor, cluster_num: int, noise: torch.Tensor):
    """cluster_dpc_knn (models/ClusterMergeNet.py:68-128; k = 5, no token mask) on x [B, N, 128] fp32: returns
    (idx_cluster [B, N] int64, index_down [B, cluster_num] int64).  The N x N distance matrix is never formed."""
    B, N, C = x.shape
    x = x.detach().contiguous().float()
    st = stream()
    density = torch.empty(B, N, device=x.device, dtype=F32)
    rowmax2 = torch.empty_like(density)
    planes = torch.empty(2, B * N, C, device=x.device, dtype=torch.float16)       # fp16 hi / lo parts of the scaled tokens (22 bits)
    norms = torch.empty(B * N, device=x.device, dtype=F32)
    inv_s2 = torch.empty(1, device=x.device, dtype=F32)
    amax = x.abs().amax().reshape(1)
    call("dml_dpc_split", ptr(x), ptr(amax), B * N, C, ptr(planes), ptr(norms), ptr(inv_s2), st)
    call("dml_dpc_density", ptr(planes), ptr(norms), ptr(inv_s2), ptr(noise.contiguous().float()), B, N, C, ptr(density), ptr(rowmax2), st)
    dist_max = (rowmax2.max(dim=1)[0].sqrt() / (C ** 0.5)).contiguous()
    parent = torch.empty_like(density)
    call("dml_dpc_parent", ptr(planes), ptr(norms), ptr(inv_s2), ptr(density), ptr(dist_max), B, N, C, ptr(parent), st)
    score = parent * density                                                        # :117
    index_down = torch.topk(score, k=cluster_num, dim=-1)[1].contiguous()           # :118
    idx = torch.empty(B, N, device=x.device, dtype=torch.int64)
    call("dml_dpc_assign", ptr(x), ptr(index_down), B, N, C, cluster_num, ptr(idx), st)
    idx.scatter_(1, index_down, torch.arange(cluster_num, device=x.device).expand(B, cluster_num))    # :126-128
    return idx, index_down


class MergeTokensFn(torch.autograd.Function):
    """merge_tokens (models/ClusterMergeNet.py:133-166): weighted mean of the tokens of each cluster.  x [B, N, 128],
    w [B, N] -> (merged [B, K, 128], all_w [B, K])."""

    @staticmethod
    def forward(ctx, x, w, idx, K):
        B, N, C = x.shape
        x, w = x.contiguous().float(), w.contiguous().float()
        merged = torch.empty(B, K, C, device=x.device, dtype=F32)
        all_w = torch.empty(B, K, device=x.device, dtype=F32)
        ws = torch.empty(_lib.load().dml_merge_ws_floats(B, N, K), device=x.device, dtype=F32)
        call("dml_merge_fwd", ptr(x), ptr(w), ptr(idx), B, N, C, K, ptr(ws), ptr(merged), ptr(all_w), stream())
        ctx.save_for_backward(x, w, idx, merged, all_w)
        ctx.mark_non_differentiable(all_w)
        return merged, all_w

    @staticmethod
    def backward(ctx, dmerged, _):
        x, w, idx, merged, all_w = ctx.saved_tensors
        B, N, C = x.shape
        K = merged.shape[1]
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        call("dml_merge_bwd", ptr(dmerged.contiguous().float()), ptr(x), ptr(w), ptr(idx), ptr(merged), ptr(all_w), B, N, C, K, ptr(dx),
             ptr(dw), stream())
        return dx, dw, None, None
