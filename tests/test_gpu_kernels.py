"""Kernel-level parity on the GPU, through the C ABI, against plain torch fp32 maths on the same
(bf16-rounded where the kernel stores bf16) inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

from dml_b200 import _lib, ops, synth
from dml_b200._lib import call, call_test, ptr, stream
from oracle import deform1d as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"


def mlp_params(seed, hid=32, nout=2, gain=2.0):
    shp = {"w1": (hid, 1), "b1": (hid,), "W2": (hid, hid), "b2": (hid,), "W3": (nout, hid), "b3": (nout,)}
    P = synth.fill_like(shp, seed, gain)
    P["w1"] = synth.uniform((hid, 1), seed, "w1x", 2.0)        # wide slopes -> many breakpoints inside [-T, T]
    P["b1"] = synth.uniform((hid,), seed, "b1x", 1.0)
    return {k: v.to(DEV) for k, v in P.items()}


def loss_scale(t):
    """(s, 1/s) with 4 < s max|t| <= 8, from the maximum as dml_pgemm's absmax epilogue would deliver it."""
    bits = t.abs().max().reshape(1).view(torch.int32)
    return ops.loss_scale_from_amax(bits)


def dense_mlp(t, P):
    h = F.relu(t[:, None] * P["w1"][:, 0] + P["b1"])
    h = F.relu(h @ P["W2"].t() + P["b2"])
    return h @ P["W3"].t() + P["b3"]


def build_table(P, t_max, hid=32, nout=2):
    table = torch.empty(_lib.load().dml_cpb_table_bytes(), device=DEV, dtype=torch.uint8)
    args = [P[k].reshape(-1).contiguous() if k == "w1" else P[k].contiguous() for k in ("w1", "b1", "W2", "b2", "W3", "b3")]
    call("dml_cpb_table_build", *[ptr(a) for a in args], hid, nout, t_max, ptr(table), stream())
    return table, args


@pytest.mark.parametrize("seed,hid,nout", [(1, 32, 2), (2, 32, 2), (3, 16, 1), (4, 32, 1)])
def test_cpb_table_matches_dense_mlp(seed, hid, nout):
    P = mlp_params(seed, hid, nout)
    T = 1.2
    table, _ = build_table(P, T, hid, nout)
    hdr = table[:16].view(torch.int32)
    nseg, kmax = int(hdr[0]), int(hdr[1])
    assert 1 <= nseg <= 1089 and 0 <= kmax <= 2048   # hdr[1] = cells holding >= 2 breakpoints (slow-path cells)
    t = torch.linspace(-T * 0.999, T * 0.999, 200001, device=DEV)
    out = torch.empty(t.numel(), 2, device=DEV)
    seg = torch.empty(t.numel(), device=DEV, dtype=torch.int32)
    call("dml_cpb_eval", ptr(table), ptr(t), t.numel(), ptr(out), ptr(seg), stream())
    ref = dense_mlp(t.double(), {k: v.double() for k, v in P.items()})
    err = (out[:, :nout].double() - ref).abs().max().item()
    assert err < 2e-5 * max(1.0, ref.abs().max().item()), (err, nseg, kmax)
    assert int(seg.min()) == 0 and int(seg.max()) == nseg - 1
    assert bool((seg[1:] >= seg[:-1]).all())          # segments are ordered in t


def attn_reference(q, k, v, g, P, H_, nout, scale, n):
    """fp32 torch maths of DeformableAttention1D.py:203-231 on [B,n,H*64] / [B,n_kv,H*64] inputs."""
    B, n_kv = k.shape[0], k.shape[1]
    d = 64
    qh = q.float().reshape(B, n, H_, d).transpose(1, 2)
    kh = k.float().reshape(B, n_kv, H_, d).transpose(1, 2)
    vh = v.float().reshape(B, n_kv, H_, d).transpose(1, 2)
    sim = torch.einsum("bhid,bhjd->bhij", qh, kh) * scale
    seq = O.normalize_grid(torch.arange(n, device=q.device)).float()
    Pd = {"rel_pos_bias.mlp.0.0.weight": P["w1"], "rel_pos_bias.mlp.0.0.bias": P["b1"],
          "rel_pos_bias.mlp.1.0.weight": P["W2"], "rel_pos_bias.mlp.1.0.bias": P["b2"],
          "rel_pos_bias.mlp.2.weight": P["W3"], "rel_pos_bias.mlp.2.bias": P["b3"]}
    sim = sim + O.cpb_bias(seq, g, Pd, H_ // nout)
    attn = sim.softmax(-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, vh)
    return out.transpose(1, 2).reshape(B, n, H_ * d)


@pytest.mark.parametrize("B,n,n_kv", [(1, 193, 48), (2, 130, 64), (1, 517, 129), (1, 64, 1), (2, 700, 200), (1, 2049, 512)])
def test_deform_attn_fwd_tcgen05_matches_torch(B, n, n_kv):
    """The tcgen05/TMEM/TMA forward against exact fp32 torch maths on the same fp16 inputs, and against the
    mma.sync forward it replaces (same contract)."""
    Hh, d, nout = 8, 64, 2
    G, C = Hh // nout, Hh * d
    seed = 300 + n
    q = (synth.normal((B, n, C), seed, "q") * 0.7).to(DEV).to(torch.float16)
    k = (synth.normal((B, n_kv, C), seed, "k") * 0.7).to(DEV).to(torch.float16)
    v = synth.normal((B, n_kv, C), seed, "v").to(DEV).to(torch.float16)
    vgrid = torch.arange(n_kv, device=DEV)[None] + synth.uniform((B * G, n_kv), seed, "off", 2.0).to(DEV)
    g = O.normalize_grid(vgrid).contiguous()
    P = mlp_params(seed)
    t_max = math.log1p(2.0 + 4.0 / max(n_kv - 1, 1)) * 1.001 + 1e-3
    table, _ = build_table(P, t_max)
    scale = d ** -0.5
    o = torch.full((B, n, C), float("nan"), device=DEV)
    lse = torch.full((B, Hh, n), float("nan"), device=DEV)
    call("dml_deform_attn_fwd_tc", ptr(q), ptr(k), ptr(v), ptr(g), ptr(table), B, Hh, d, n, n_kv, n, C, C, C, C, nout,
         scale, ptr(o), ptr(lse), stream())
    ref = attn_reference(q.float(), k.float(), v.float(), g, P, Hh, nout, scale, n)
    H.assert_close(o, ref, 1e-3, "attention output (tcgen05)")
    o2 = torch.empty_like(o)
    lse2 = torch.empty_like(lse)
    call_test("dml_deform_attn_fwd", ptr(q), ptr(k), ptr(v), ptr(g), ptr(table), B, Hh, d, n, n_kv, C, C, C, C, nout, scale,
         ptr(o2), ptr(lse2), stream())
    H.assert_close(lse, lse2, 1e-5, "log-sum-exp (tcgen05 vs mma.sync)")
    # the same launch with trailing 256-query blocks cut into 128-query CTAs (dml_deform_attn_fwd_tc_split): identical results
    for hb in (1, 3, 1000):
        o3 = torch.full_like(o, float("nan"))
        lse3 = torch.full_like(lse, float("nan"))
        call("dml_deform_attn_fwd_tc_split", ptr(q), ptr(k), ptr(v), ptr(g), ptr(table), B, Hh, d, n, n_kv, n, C, C, C, C, nout,
             scale, ptr(o3), ptr(lse3), hb, stream())
        assert torch.equal(o3, o) and torch.equal(lse3, lse), f"half_blocks = {hb}"


@pytest.mark.parametrize("B,n,n_kv", [(1, 300, 70), (2, 1000, 257)])
def test_dq_from_ds_workspace_matches_matmul(B, n, n_kv):
    """The streaming dQ GEMM on its own: dq = (1/s) dS K from a dS^T workspace in the backward's layout."""
    Hh, d = 8, 64
    C = Hh * d
    n_pad, n_kv_pad = -(-n // 32) * 32, -(-n_kv // 128) * 128
    seed = 900 + n
    ds = synth.normal((B * Hh, n_kv, n), seed, "ds").to(DEV)
    ws = torch.zeros(B * Hh, n_kv_pad, n_pad, device=DEV, dtype=torch.float16)
    ws[:, :n_kv, :n] = ds.to(torch.float16)
    k = synth.normal((B, n_kv, C), seed, "k").to(DEV).to(torch.float16)
    dscale = torch.tensor([4.0, 0.25], device=DEV)
    dq = torch.full((B, n, C), float("nan"), device=DEV)
    call("dml_deform_attn_dq_from_ds", ptr(ws), ptr(k), ptr(dscale), B, Hh, d, n, n_kv, C, ptr(dq), stream())
    dsf = ws[:, :n_kv, :n].float().reshape(B, Hh, n_kv, n)
    kf = k.float().reshape(B, n_kv, Hh, d).permute(0, 2, 1, 3)
    ref = 0.25 * torch.einsum("bhji,bhjd->bihd", dsf, kf).reshape(B, n, C)
    H.assert_close(dq, ref, 1e-5, "dq from dS^T")


def test_deform_attn_fwd_tcgen05_is_deterministic_with_large_scores():
    """No atomics in the forward: repeated launches must agree bit for bit (a race between the two threads that share a
    query row, or between the softmax warps and the MMA warps, would show up here).  Large |q|, |k| force the softmax
    reference to be raised - and the O rows rescaled in TMEM - many times."""
    B, n, n_kv, Hh, d, nout = 1, 3001, 750, 8, 64, 2
    G, C = Hh // nout, Hh * d
    seed = 77
    q = (synth.normal((B, n, C), seed, "q") * 3.0).to(DEV).to(torch.float16)
    k = (synth.normal((B, n_kv, C), seed, "k") * 3.0).to(DEV).to(torch.float16)
    v = synth.normal((B, n_kv, C), seed, "v").to(DEV).to(torch.float16)
    vgrid = torch.arange(n_kv, device=DEV)[None] + synth.uniform((B * G, n_kv), seed, "off", 2.0).to(DEV)
    g = O.normalize_grid(vgrid).contiguous()
    P = mlp_params(seed)
    table, _ = build_table(P, math.log1p(2.0 + 4.0 / (n_kv - 1)) * 1.001 + 1e-3)
    scale = d ** -0.5
    outs = []
    for _ in range(6):
        o = torch.full((B, n, C), float("nan"), device=DEV)
        lse = torch.full((B, Hh, n), float("nan"), device=DEV)
        call("dml_deform_attn_fwd_tc", ptr(q), ptr(k), ptr(v), ptr(g), ptr(table), B, Hh, d, n, n_kv, n, C, C, C, C, nout,
             scale, ptr(o), ptr(lse), stream())
        outs.append((o, lse))
    for o, lse in outs[1:]:
        assert torch.equal(o, outs[0][0]) and torch.equal(lse, outs[0][1])
    ref = attn_reference(q.float(), k.float(), v.float(), g, P, Hh, nout, scale, n)
    H.assert_close(outs[0][0], ref, 2e-3, "attention output, peaked softmax")


# tc_ws: dS^T workspace + streaming dQ GEMM; tc_general: the dK/dV kernel's path for tables with too many segments
@pytest.mark.parametrize("impl", ["mma", "tc", "tc_ws", "tc_general"])
# (1, 3100, 200): few key blocks, many query tiles -> the dK/dV kernel splits the query range across CTAs (reductions)
@pytest.mark.parametrize("B,n,n_kv", [(1, 193, 48), (2, 130, 64), (1, 517, 129), (1, 64, 1), (1, 1100, 300), (1, 3100, 200),
                                      (2, 3200, 129)])      # batch 2 + work list + a mostly padded second key block
def test_deform_attn_fwd_bwd_matches_torch(B, n, n_kv, impl):
    Hh, d, nout = 8, 64, 2
    G, C = Hh // nout, Hh * d
    seed = 100 + n
    q = (synth.normal((B, n, C), seed, "q") * 0.7).to(DEV).to(torch.float16)
    k = (synth.normal((B, n_kv, C), seed, "k") * 0.7).to(DEV).to(torch.float16)
    v = synth.normal((B, n_kv, C), seed, "v").to(DEV).to(torch.float16)
    vgrid = torch.arange(n_kv, device=DEV)[None] + synth.uniform((B * G, n_kv), seed, "off", 2.0).to(DEV)
    g = O.normalize_grid(vgrid).contiguous()
    P = mlp_params(seed)
    t_max = math.log1p(2.0 + 4.0 / max(n_kv - 1, 1)) * 1.001 + 1e-3
    table, margs = build_table(P, t_max)
    scale = d ** -0.5
    o = torch.empty(B, n, C, device=DEV, dtype=torch.float32)
    lse = torch.empty(B, Hh, n, device=DEV)
    call_test("dml_deform_attn_fwd", ptr(q), ptr(k), ptr(v), ptr(g), ptr(table), B, Hh, d, n, n_kv, C, C, C, C, nout, scale,
         ptr(o), ptr(lse), stream())
    qf, kf, vf = (t.float().requires_grad_() for t in (q, k, v))
    gf = g.clone().requires_grad_()
    Pf = {kk: vv.clone().requires_grad_() for kk, vv in P.items()}
    ref = attn_reference(qf, kf, vf, gf, Pf, Hh, nout, scale, n)
    H.assert_close(o, ref, 2e-3, "attention output (fp16 P in the PV MMA)")

    r = (synth.normal((B, n, C), seed, "r") * 1e-3).to(DEV)            # small upstream gradient: exercises the fp16 loss scale
    dscale = loss_scale(r)
    r16 = (r * dscale[0]).to(torch.float16)
    r = r16.float() * dscale[1]
    loss = (ref * r).sum()
    grads = torch.autograd.grad(loss, [qf, kf, vf, gf] + [Pf[x] for x in ("w1", "b1", "W2", "b2", "W3", "b3")])
    dq = torch.empty(B, n, C, device=DEV)
    dk = torch.empty(B, n_kv, C, device=DEV)
    dv = torch.empty_like(dk)
    dg = torch.empty(B * G, n_kv, device=DEV)
    segsum = torch.empty(_lib.load().dml_cpb_seg_max(), 4, device=DEV)
    dsum = torch.empty(B, Hh, n, device=DEV)
    if impl == "mma":
        call_test("dml_deform_attn_bwd", ptr(q), ptr(k), ptr(v), ptr(g), ptr(table), ptr(o), ptr(r16), ptr(lse), B, Hh, d, n, n_kv,
             C, C, C, C, nout, scale, ptr(dscale), ptr(dsum), ptr(dq), ptr(dk), ptr(dv), ptr(dg), ptr(segsum), stream())
    else:
        ws = None
        bwd = call
        if impl == "tc_general":          # the knob lives in the test-only build of the same source (libdml_b200_test.so)
            _lib.load_test().dml_debug_set_seg_limit(3)
            bwd = call_test
        if impl == "tc_ws":
            nbytes = _lib.load().dml_deform_attn_bwd_ws_bytes(B, Hh, n, n_kv)
            assert nbytes == B * Hh * (-(-n_kv // 128) * 128) * (-(-n // 32) * 32) * 2
            ws = torch.full((nbytes // 2,), float("nan"), device=DEV, dtype=torch.float16)   # every element the GEMM reads must be written
        bwd("dml_deform_attn_bwd_tc", ptr(q), ptr(k), ptr(v), ptr(g), ptr(table), ptr(o), ptr(r16), ptr(lse), B, Hh, d, n,
            n_kv, n, C, C, C, C, nout, scale, ptr(dscale), ptr(dsum), ptr(dq), ptr(dk), ptr(dv), ptr(dg), ptr(segsum),
            ptr(ws) if ws is not None else None, stream())
        if impl == "tc_general":
            _lib.load_test().dml_debug_set_seg_limit(0)
    mg = torch.empty(ops.CPB_GRAD_FLOATS, device=DEV)
    call("dml_cpb_param_grad", *[ptr(a) for a in margs], 32, nout, ptr(table), ptr(segsum), ptr(mg), stream())
    tol = 3e-3     # fp16 P / dS operands in the MMAs; compared against exact fp32 maths
    zero = 1e-6 if n_kv == 1 else 0.0   # one key: softmax == 1, dS == 0 -> dq, dk are exactly zero in the reference
    H.assert_close(dq * scale, grads[0], tol, "dq", atol=zero)
    H.assert_close(dk, grads[1], tol, "dk", atol=zero)
    H.assert_close(dv, grads[2], tol, "dv")
    if n_kv > 1:
        H.assert_close(dg, grads[3], tol, "dg")
    H.assert_close(mg[0:32], grads[4][:, 0], tol, "d mlp.w1", atol=zero)
    H.assert_close(mg[32:64], grads[5], tol, "d mlp.b1", atol=zero)
    H.assert_close(mg[64:1088].reshape(32, 32), grads[6], tol, "d mlp.W2", atol=zero)
    H.assert_close(mg[1088:1120], grads[7], tol, "d mlp.b2", atol=zero)
    H.assert_close(mg[1120:1184].reshape(2, 32), grads[8], tol, "d mlp.W3", atol=zero)
    assert float(mg[1184:1186].abs().max()) < 1e-2 * max(1e-3, float(grads[8].abs().max()))   # softmax shift invariance


@pytest.mark.parametrize("B,n", [(2, 193), (1, 128), (1, 1030)])
def test_offsets_and_gather_match_oracle(B, n):
    G, C, dim, ks, stride, osc = 4, 512, 128, 6, 4, 2.0
    seed = 200 + n
    P = {k: v.to(DEV) for k, v in synth.fill_like(H.deform_shapes(), seed, gain=2.0).items()}
    q = synth.normal((B, n, C), seed, "q").to(DEV).to(torch.float16)
    n_kv = O.kv_length(n, ks, stride)
    vgrid = torch.empty(B * G, n_kv, device=DEV)
    g = torch.empty_like(vgrid)
    w0 = P["to_offsets.0.weight"].reshape(128, ks).contiguous()
    b0 = P["to_offsets.0.bias"].contiguous()
    w2 = P["to_offsets.2.weight"].reshape(128).contiguous()
    call("dml_offsets_fwd", ptr(q), ptr(w0), ptr(b0), ptr(w2), B, n, C, G, ks, stride, osc, ptr(vgrid), ptr(g), stream())
    qf = q.float().requires_grad_()
    Pf = {k: v.clone().requires_grad_() for k, v in P.items()}
    gq = qf.transpose(1, 2).reshape(B * G, C // G, n)
    off = O.offsets_net(gq, Pf, stride, osc)
    vg_ref = torch.arange(n_kv, device=DEV) + off
    assert n_kv == _lib.load().dml_offsets_kv_len(n, ks, stride)
    H.assert_close(vgrid, vg_ref, 1e-5, "vgrid")
    assert float((g - O.normalize_grid(vg_ref)).abs().max()) < 2e-6

    # gather forward / backward against the closed form AND the literal grid_sample
    x2 = synth.normal((B, n, dim), seed, "x2").to(DEV)
    kv = torch.empty(B, n_kv, dim, device=DEV, dtype=torch.float32)
    i0, i1, wy0, wy1 = ops.centre_taps(n)
    assert (i0, wy0, wy1) == (O.centre_taps(n)[0], O.centre_taps(n)[2], O.centre_taps(n)[3])
    call("dml_kv_gather_fwd", ptr(x2), ptr(g), B, n, dim, G, n_kv, i0, i1, wy0, wy1, ptr(kv), stream())
    x2f = x2.clone().requires_grad_()
    gf = g.clone().requires_grad_()
    lit = O.grid_sample_1d_literal(x2f.transpose(1, 2).reshape(B * G, dim // G, n), gf).reshape(B, dim, n_kv).transpose(1, 2)
    H.assert_close(kv, lit, 1e-6, "kv_feats")
    assert torch.equal(kv, lit) or n % 2 == 0     # the closed form of the degenerate sample is bit-exact for odd n
    dkv = synth.normal((B, n_kv, dim), seed, "dkv").to(DEV)
    gx2, gg = torch.autograd.grad((lit * dkv).sum(), (x2f, gf))
    dcentre = torch.empty(B, dim, device=DEV)
    dg = torch.zeros(B * G, n_kv, device=DEV)
    call("dml_kv_gather_bwd", ptr(x2), ptr(g), ptr(dkv), B, n, dim, G, n_kv, i0, i1, wy0, wy1, ptr(dcentre), ptr(dg), stream())
    dx2 = torch.zeros_like(x2)
    dx2[:, i0] += wy0 * dcentre
    if wy1:
        dx2[:, i1] += wy1 * dcentre
    H.assert_close(dx2, gx2, 1e-5, "d x2")
    H.assert_close(dg, gg, 1e-5, "d g (gather)")

    # offsets backward
    d_off = synth.normal((B * G, n_kv), seed, "d_off").to(DEV)
    dq_attn = synth.normal((B, n, C), seed, "dq_attn").to(DEV)
    grads = torch.autograd.grad((off * d_off).sum(), [qf, Pf["to_offsets.0.weight"], Pf["to_offsets.0.bias"], Pf["to_offsets.2.weight"]])
    dy_ws = torch.empty(B * G, n_kv, 128, device=DEV)
    wgrad = torch.empty(128 * ks + 256, device=DEV)
    dq = torch.empty(B, n, C, device=DEV, dtype=torch.float32)
    call("dml_offsets_bwd", ptr(q), ptr(w0), ptr(b0), ptr(w2), ptr(d_off), ptr(dq_attn), 0.125, B, n, C, G, ks, stride, osc,
         ptr(dy_ws), ptr(wgrad), ptr(dq), stream())
    H.assert_close(dq, grads[0] + 0.125 * dq_attn, 1e-5, "dq total")
    H.assert_close(wgrad[:128 * ks].reshape(128, 1, ks), grads[1], 1e-4, "d to_offsets.0.weight")
    H.assert_close(wgrad[128 * ks:128 * ks + 128], grads[2], 1e-4, "d to_offsets.0.bias")
    H.assert_close(wgrad[128 * ks + 128:].reshape(1, 128, 1), grads[3], 1e-4, "d to_offsets.2.weight")


@pytest.mark.parametrize("rows,D", [(5, 128), (16385, 128), (1000, 256), (6085, 512)])
def test_layernorm_rows(rows, D):
    x = (synth.normal((rows, D), 21, "x") * 2 + 0.5).to(DEV).requires_grad_()
    ln = torch.nn.LayerNorm(D).to(DEV)
    with torch.no_grad():
        ln.weight.copy_(1.0 + 0.2 * synth.uniform((D,), 21, "w").to(DEV))
        ln.bias.copy_(synth.uniform((D,), 21, "b", 0.3).to(DEV))
    y = ops.layer_norm(x, ln)
    ref = ln(x)
    H.assert_close(y, ref, 2e-6, "layernorm")
    r = synth.normal((rows, D), 22, "r").to(DEV)
    g = torch.autograd.grad((y * r).sum(), (x, ln.weight, ln.bias))
    gr = torch.autograd.grad((ref * r).sum(), (x, ln.weight, ln.bias))
    H.assert_close(g[0], gr[0], 1e-5, "d x")
    H.assert_close(g[1], gr[1], 2e-5, "d weight")
    H.assert_close(g[2], gr[2], 2e-5, "d bias")


@pytest.mark.parametrize("B,din", [(1, 59), (3, 361), (2, 431)])
def test_maxnet_fused_kernels_match_torch(B, din):
    """MaxNet (models/model.py:173-218) as one kernel per direction: eval mode against the torch modules (fp64), training mode
    against the same formula evaluated in torch with the kernel's own uniform numbers (AlphaDropout, torch's constants)."""
    from dml_b200.model import MaxNet
    torch.manual_seed(0)
    net = MaxNet(input_dim=din, omic_dim=128, dropout_rate=0.25, label_dim=4).to(DEV)
    x = synth.normal((B, din), 13, "xo").to(DEV).requires_grad_()
    r = synth.normal((B, 128), 13, "ro").to(DEV)
    net.eval()
    feat = net(x_omic=x)[0]
    g = torch.autograd.grad((feat * r).sum(), [x] + [p for p in net.encoder.parameters()])
    ref_net = MaxNet(input_dim=din, omic_dim=128, dropout_rate=0.25, label_dim=4).double().to(DEV)
    ref_net.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    ref_net.eval()
    xd = x.detach().double().requires_grad_()
    ref = ref_net.relu(ref_net.encoder(xd))
    gr = torch.autograd.grad((ref * r.double()).sum(), [xd] + [p for p in ref_net.encoder.parameters()])
    H.assert_close(feat, ref, 1e-5, "MaxNet features")
    for a, b in zip(g, gr):
        H.assert_close(a, b, 2e-5, "MaxNet gradient")
    # training mode: reproduce the kernel's AlphaDropout from its saved uniform numbers
    net.train()
    torch.manual_seed(7)
    feat = net(x_omic=x)[0]
    g = torch.autograd.grad((feat * r).sum(), [x] + [p for p in net.encoder.parameters()])
    torch.manual_seed(7)
    u = torch.rand(B, 64 + 48 + 32 + 128, device=DEV).double()
    p, alpha = 0.25, 1.7580993408473766
    a_ = 1.0 / math.sqrt((alpha * alpha * p + 1) * (1 - p))
    h, off = xd, 0
    for blk in ref_net.encoder:
        y = F.elu(blk[0](h))
        uu = u[:, off: off + y.shape[1]]
        h = torch.where(uu < p, torch.full_like(y, alpha * a_ * (p - 1)), a_ * y + alpha * a_ * p)
        off += y.shape[1]
    ref = F.relu(h)
    gr = torch.autograd.grad((ref * r.double()).sum(), [xd] + [p_ for p_ in ref_net.encoder.parameters()])
    H.assert_close(feat, ref, 1e-5, "MaxNet features (training)")
    for a, b in zip(g, gr):
        H.assert_close(a, b, 2e-5, "MaxNet gradient (training)")


@pytest.mark.parametrize("B,n", [(1, 37), (3, 5)])
def test_tower_head_kernels_match_torch(B, n):
    """norm(h)[:, 0] -> _fc2, multimodal_projection (DeformCrossTransMIL.py:128-151) as one kernel per direction."""
    D, nc, De = 128, 4, 128
    h = synth.normal((B, n, D), 17, "h").to(DEV).requires_grad_()
    P = {k: v.to(DEV).requires_grad_() for k, v in synth.fill_like(
        {"norm.weight": (D,), "norm.bias": (D,), "fc2.weight": (nc, D), "fc2.bias": (nc,), "proj.weight": (De, D), "proj.bias": (De,)}, 17).items()}
    enc, logits = ops.TowerHeadFn.apply(h, P["norm.weight"], P["norm.bias"], P["fc2.weight"], P["fc2.bias"], P["proj.weight"],
                                        P["proj.bias"], 1e-5)
    r1, r2 = synth.normal((B, De), 17, "r1").to(DEV), synth.normal((B, nc), 17, "r2").to(DEV)
    names = list(P)
    g = torch.autograd.grad((enc * r1).sum() + (logits * r2).sum(), [h] + [P[k] for k in names])
    hd = h.detach().double().requires_grad_()
    Pd = {k: v.detach().double().requires_grad_() for k, v in P.items()}
    hn = F.layer_norm(hd, (D,), Pd["norm.weight"], Pd["norm.bias"], 1e-5)[:, 0]
    enc_r, log_r = F.linear(hn, Pd["proj.weight"], Pd["proj.bias"]), F.linear(hn, Pd["fc2.weight"], Pd["fc2.bias"])
    gr = torch.autograd.grad((enc_r * r1.double()).sum() + (log_r * r2.double()).sum(), [hd] + [Pd[k] for k in names])
    H.assert_close(enc, enc_r, 1e-5, "encoded")
    H.assert_close(logits, log_r, 1e-5, "logits")
    for nm, a, b in zip(["h"] + names, g, gr):
        H.assert_close(a, b, 2e-5, "tower head grad " + nm)


@pytest.mark.parametrize("sigmoid", [False, True])
def test_three_classifier_head_matches_torch(sigmoid):
    B, Da, Db, nc = 3, 128, 128, 4
    a = synth.normal((B, Da), 19, "a").to(DEV).requires_grad_()
    b = synth.normal((B, Db), 19, "b").to(DEV).requires_grad_()
    P = {k: v.to(DEV).requires_grad_() for k, v in synth.fill_like(
        {"c.weight": (nc, Da + Db), "c.bias": (nc,), "a.weight": (nc, Da), "a.bias": (nc,), "b.weight": (nc, Db), "b.bias": (nc,)}, 19).items()}
    ys = ops.Linear3Fn.apply(a, b, P["c.weight"], P["c.bias"], P["a.weight"], P["a.bias"], P["b.weight"], P["b.bias"], sigmoid)
    rs = [synth.normal((B, nc), 19, f"r{i}").to(DEV) for i in range(3)]
    names = list(P)
    g = torch.autograd.grad(sum((y * r).sum() for y, r in zip(ys, rs)), [a, b] + [P[k] for k in names], retain_graph=True)
    ad, bd = a.detach().double().requires_grad_(), b.detach().double().requires_grad_()
    Pd = {k: v.detach().double().requires_grad_() for k, v in P.items()}
    act = torch.sigmoid if sigmoid else (lambda t: t)
    yr = [act(F.linear(torch.cat((ad, bd), 1), Pd["c.weight"], Pd["c.bias"])), act(F.linear(ad, Pd["a.weight"], Pd["a.bias"])),
          act(F.linear(bd, Pd["b.weight"], Pd["b.bias"]))]
    gr = torch.autograd.grad(sum((y * r.double()).sum() for y, r in zip(yr, rs)), [ad, bd] + [Pd[k] for k in names])
    for y, r_ in zip(ys, yr):
        H.assert_close(y, r_, 1e-5, "classifier outputs")
    for nm, x, y in zip(["a", "b"] + names, g, gr):
        H.assert_close(x, y, 2e-5, "classifier grad " + nm)
    # only the fused head's output drives the loss in training (train_test.py:833-853): the other two gradients are None
    (ga,) = torch.autograd.grad(ys[0].sum(), [a], allow_unused=True)
    assert ga is not None


@pytest.mark.gpu
def test_loss_scale_kernel_matches_the_host_expression():
    """dml_loss_scale_from_amax (one thread, exponent arithmetic) against the torch expression it replaces on the step's chain."""
    vals = [0.0, 1e-42, 3e-39, 1e-30, 3e-7, 0.02, 0.5, 1.0, 1.0000001, 2.0, 3.999, 4.0, 7.9, 8.0, 8.5, 900.0, 2.0 ** 40, 3e30, 3e38]
    for a in vals:
        bits = torch.tensor([a], dtype=torch.float32).view(torch.int32)
        ref = ops.loss_scale_from_amax(bits)                       # CPU: the torch form
        got = ops.loss_scale_from_amax(bits.to(DEV)).cpu()         # GPU: the kernel
        assert torch.equal(ref, got), (a, ref.tolist(), got.tolist())
        av = float(torch.tensor(a, dtype=torch.float32))
        if 1e-17 < av < 1e17:
            assert 4.0 < float(got[0]) * av <= 8.0
