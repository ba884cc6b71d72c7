"""DeformCrossAttention2D and ClusterMergeNet on the kernels of csrc/deform2d.cu, deform2d_bias.cu and
cluster.cu: every kernel through the C ABI against the oracle's torch maths on the same device, the modules against the
reference's own goldens (tests/golden/deform2d_*.npz, clustermerge_*.npz) and against the oracle at the reference's bag size."""
import math

import pytest
import torch
import torch.nn.functional as F

from dml_b200 import _lib, ops2d, synth
from dml_b200._lib import call, ptr
from dml_b200.ClusterMergeNet import ClusterMergeNet
from dml_b200.DeformableAttention2D import DeformCrossAttention2D
from oracle import deform2d as O2
from oracle.golden_cases import CLUSTER_CASES, DEFORM2D_CASES, thin
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3          # north_star: fp32-class path
# The six parameter gradients of the position-bias MLP are cancellation-dominated sums over ReLU-masked terms (rows of dS sum to
# zero): the REFERENCE's own fp32 result is 1e-4 .. 9e-4 (max-norm) away from its fp64 result on the golden cases (measured with
# the oracle), and the kernels sit 6e-5 .. 6e-4 from fp64 (test_position_bias_mlp_fwd_bwd) - two fp32-class results
# can therefore differ by more than 1e-3 from EACH OTHER.  Against the fp32 goldens these six tensors are held to 3e-3.
TOL_MLP = 3e-3


def tol_of(name):
    return TOL_MLP if "rel_pos_bias.mlp" in name else TOL


def st():
    return torch.cuda.current_stream().cuda_stream


def params(seed, gain=2.0):
    return {k: v.to(DEV) for k, v in synth.fill_like(H.attn2d_shapes(""), seed, gain=gain).items()}


# ---------------------------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows", [1, 77, 5000])
def test_grouped_projection_fwd_bwd(rows):
    x = synth.normal((rows, 128), 1, "x").to(DEV)
    W = synth.uniform((512, 16), 1, "W", 0.25).to(DEV)
    dy = synth.normal((rows, 512), 1, "dy").to(DEV)
    y = torch.empty(rows, 512, device=DEV)
    call("dml_da2_gproj_fwd", ptr(x), ptr(W), rows, ptr(y), st())
    xr = x.clone().requires_grad_()
    Wr = W.clone().requires_grad_()
    ref = F.conv2d(xr.t().reshape(1, 128, rows, 1), Wr.reshape(512, 16, 1, 1), groups=8).reshape(512, rows).t()
    H.assert_close(y, ref, 1e-5, "y")
    gx, gW = torch.autograd.grad((ref * dy).sum(), (xr, Wr))
    parts = torch.empty(_lib.load().dml_da2_gproj_parts(rows), 8192, device=DEV)
    dx = torch.full((rows, 128), 7.0, device=DEV)
    dW = torch.empty(512, 16, device=DEV)
    call("dml_da2_gproj_bwd", ptr(dy), ptr(x), ptr(W), rows, 0, ptr(dx), ptr(parts), ptr(dW), st())
    H.assert_close(dx, gx, 1e-5, "dx")
    H.assert_close(dW, gW, 1e-5, "dW")
    call("dml_da2_gproj_bwd", ptr(dy), ptr(x), ptr(W), rows, 1, ptr(dx), ptr(parts), None, st())
    H.assert_close(dx, 2 * gx, 1e-5, "dx accumulated")


@pytest.mark.parametrize("B,side", [(2, 20), (1, 50), (1, 23), (1, 6), (1, 101)])
def test_offsets_fwd_bwd(B, side):
    P = params(3)
    n = side * side
    hk = O2.kv_side(side)
    m = hk * hk
    q = synth.normal((B, n, 512), 3, "q").to(DEV)
    assert _lib.load().dml_da2_kv_side(side, 6, 4) == hk
    wdw, bdw, w2 = P["to_offsets.0.weight"].reshape(64, 36).contiguous(), P["to_offsets.0.bias"], P["to_offsets.2.weight"].reshape(2, 64).contiguous()
    vgrid = torch.empty(B * 8, 2, hk, hk, device=DEV)
    vs = torch.empty(B * 8, m, 2, device=DEV)
    call("dml_da2_offsets_fwd", ptr(q), ptr(wdw), ptr(bdw), ptr(w2), B, side, 6, 4, 4.0, ptr(vgrid), ptr(vs), st())
    # oracle on the same device
    Pr = {k: v.clone().requires_grad_() for k, v in P.items()}
    qr = q.clone().requires_grad_()
    qg = qr.transpose(1, 2).reshape(B * 8, 64, side, side)
    off = O2.offsets_net(qg, Pr, 4, 4.0)
    vg = O2.xy_grid(hk, hk).to(DEV) + off
    vx, vy = O2.normalize_xy(vg[:, 0], vg[:, 1], hk, hk)
    vsr = torch.stack((vx, vy), -1).reshape(B * 8, m, 2)
    H.assert_close(vgrid, vg, 1e-5, "vgrid")
    H.assert_close(vs, vsr, 1e-5, "vs")
    dvs = synth.normal((B * 8, m, 2), 3, "dvs").to(DEV)
    dvg = synth.normal((B * 8, 2, hk, hk), 3, "dvg").to(DEV)
    names = ["to_offsets.0.weight", "to_offsets.0.bias", "to_offsets.2.weight"]
    gs = torch.autograd.grad((vsr * dvs).sum() + (vg * dvg).sum(), [qr] + [Pr[k] for k in names])
    dconv = torch.empty(B * 8, m, 64, device=DEV)
    parts = torch.empty(_lib.load().dml_da2_offsets_parts(B, side, 6, 4), 2496, device=DEV)
    grads = torch.empty(2496, device=DEV)
    dq = torch.zeros(B, n, 512, device=DEV)
    call("dml_da2_offsets_bwd", ptr(q), ptr(wdw), ptr(bdw), ptr(w2), ptr(dvs), ptr(dvg), B, side, 6, 4, 4.0, ptr(dconv), ptr(parts),
         ptr(grads), ptr(dq), st())
    H.assert_close(dq, gs[0], 1e-4, "dq")
    H.assert_close(grads[:2304].view(64, 1, 6, 6), gs[1], 1e-4, "d depthwise weight")
    H.assert_close(grads[2304:2368], gs[2], 1e-4, "d depthwise bias")
    H.assert_close(grads[2368:2496].view(2, 64, 1, 1), gs[3], 1e-4, "d pointwise weight")


@pytest.mark.parametrize("B,side,m", [(2, 20, 25), (1, 50, 144), (1, 9, 4), (1, 101, 625)])
def test_bilinear_gather_fwd_bwd(B, side, m):
    n = side * side
    x2 = synth.normal((B, n, 128), 5, "x2").to(DEV)
    # positions inside, on the border and outside the image
    vs = (synth.uniform((B * 8, m, 2), 5, "vs", 1.25)).to(DEV)
    kvf = torch.empty(B, m, 128, device=DEV)
    call("dml_da2_gather_fwd", ptr(x2), ptr(vs), B, side, m, ptr(kvf), st())
    xr, vr = x2.clone().requires_grad_(), vs.clone().requires_grad_()
    img = xr.transpose(1, 2).reshape(B * 8, 16, side, side)
    ref = F.grid_sample(img, vr.reshape(B * 8, 1, m, 2), mode="bilinear", padding_mode="zeros", align_corners=False)   # [(B 8), 16, 1, m]
    ref = ref.reshape(B, 128, m).transpose(1, 2)
    H.assert_close(kvf, ref, 1e-5, "kvf")
    d = synth.normal((B, m, 128), 5, "d").to(DEV)
    gx, gv = torch.autograd.grad((ref * d).sum(), (xr, vr))
    dx2 = torch.zeros(B, n, 128, device=DEV)
    dvs = torch.ones(B * 8, m, 2, device=DEV)
    call("dml_da2_gather_bwd", ptr(d), ptr(x2), ptr(vs), B, side, m, ptr(dx2), ptr(dvs), st())
    H.assert_close(dx2, gx, 1e-5, "dx2")
    H.assert_close(dvs - 1.0, gv, 1e-4, "dvs")


@pytest.mark.parametrize("B,side,m", [(2, 20, 25), (1, 50, 144), (1, 7, 16), (1, 33, 49), (1, 101, 625)])
def test_position_bias_mlp_fwd_bwd(B, side, m):
    """The tensor-core MLP (bf16-pair mma.sync) against the dense fp64 MLP: bias, the six parameter gradients and d vs."""
    P = params(7)
    n = side * side
    vs = synth.uniform((B * 8, m, 2), 7, "vs", 1.1).to(DEV)
    names = ["rel_pos_bias.mlp.0.0.weight", "rel_pos_bias.mlp.0.0.bias", "rel_pos_bias.mlp.1.0.weight", "rel_pos_bias.mlp.1.0.bias",
             "rel_pos_bias.mlp.2.weight", "rel_pos_bias.mlp.2.bias"]
    mlp = [P[k].contiguous() for k in names]
    bias = torch.empty(B, 8, n, m, device=DEV)
    call("dml_da2_bias_fwd", ptr(vs), *[ptr(t) for t in mlp], B, side, m, ptr(bias), st())
    Pd = {k: P[k].double().requires_grad_() for k in names}
    vd = vs.double().requires_grad_()
    g = O2.xy_grid(side, side).to(DEV)
    gx, gy = O2.normalize_xy(g[0], g[1], side, side)
    gq = torch.stack((gx, gy), -1).reshape(1, n, 1, 2).double()
    ref = O2.bias_mlp(gq - vd.reshape(B * 8, 1, m, 2), Pd)[..., 0].reshape(B, 8, n, m)
    H.assert_close(bias, ref.float(), 2e-5, "bias")
    ds = (synth.normal((B, 8, n, m), 7, "ds") * 0.01).to(DEV)
    ds = ds - ds.mean(dim=-1, keepdim=True)              # rows of dS sum to zero (softmax), the cancellation the real path has
    gs = torch.autograd.grad((ref * ds.double()).sum(), [Pd[k] for k in names] + [vd])
    parts = torch.empty(_lib.load().dml_da2_bias_bwd_parts(B, side), 1192, device=DEV)
    grads = torch.empty(1192, device=DEV)
    dvs = torch.zeros(B * 8, m, 2, device=DEV)
    call("dml_da2_bias_bwd", ptr(vs), *[ptr(t) for t in mlp[:5]], ptr(ds), B, side, m, ptr(parts), ptr(grads), ptr(dvs), st())
    # every parameter gradient is a cancellation-dominated sum over ReLU-masked terms (rows of dS sum to zero): torch's own fp32
    # evaluation of the same MLP is 1e-4 .. 2e-3 away from fp64 on them, growing with the number of pairs.  The bar: 1e-3 against
    # fp64, or three times the fp32-torch error where that is larger (both printed)
    P32 = {k: P[k].clone().requires_grad_() for k in names}
    v32 = vs.clone().requires_grad_()
    ref32 = O2.bias_mlp(gq.float() - v32.reshape(B * 8, 1, m, 2), P32)[..., 0].reshape(B, 8, n, m)
    g32 = torch.autograd.grad((ref32 * ds).sum(), [P32[k] for k in names] + [v32])
    got = {"dW1": grads[0:64].view(32, 2), "db1": grads[64:96], "dW2": grads[96:1120].view(32, 32), "db2": grads[1120:1152],
           "dW3": grads[1152:1184].view(1, 32), "dvs": dvs}
    keys = ["dW1", "db1", "dW2", "db2", "dW3", "db3", "dvs"]
    ref_g, ref_32 = dict(zip(keys, gs)), dict(zip(keys, g32))
    errs = {k: max(H.rel_l2(v, ref_g[k].float()), H.max_rel(v, ref_g[k].float())) for k, v in got.items()}
    e32 = {k: max(H.rel_l2(ref_32[k], ref_g[k].float()), H.max_rel(ref_32[k], ref_g[k].float())) for k in got}
    print("bias-MLP gradient errors vs fp64 (kernel, torch fp32):", {k: (f"{errs[k]:.1e}", f"{e32[k]:.1e}") for k in got})
    assert abs(float(grads[1184]) - float(gs[5])) <= 1e-4 * float(ds.abs().sum())
    bad = {k: (errs[k], e32[k]) for k in got if errs[k] > max(TOL, 3.0 * e32[k])}
    assert not bad, bad


@pytest.mark.parametrize("B,n,m,drop", [(2, 400, 25, False), (1, 2500, 144, True), (1, 37, 70, False), (1, 10201, 625, False)])
def test_attention_rows_and_columns(B, n, m, drop):
    q = synth.normal((B, n, 512), 9, "q").to(DEV)
    k = synth.normal((B, m, 512), 9, "k").to(DEV)
    v = synth.normal((B, m, 512), 9, "v").to(DEV)
    bias = synth.normal((B, 8, n, m), 9, "bias").to(DEV)
    keep = (synth.uniform((B, 8, n, m), 9, "keep", 0.5) > -0.4).to(torch.uint8).to(DEV) if drop else None
    ks = 1.0 / 0.9 if drop else 1.0
    attn = bias.clone()
    o = torch.empty(B, n, 512, device=DEV)
    scale = 64 ** -0.5
    ws = torch.empty(_lib.load().dml_da2_attn_ws_bytes(B, n, m, 1), device=DEV, dtype=torch.uint8)
    call("dml_da2_attn_fwd", ptr(q), ptr(k), ptr(v), ptr(attn), ptr(keep) if drop else None, ks, B, n, m, scale, ptr(ws), ptr(o), st())
    qr, kr, vr, br = (t.clone().requires_grad_() for t in (q, k, v, bias))
    hd = lambda t: t.reshape(B, -1, 8, 64).transpose(1, 2)
    sim = hd(qr) @ hd(kr).transpose(2, 3) * scale + br
    a = sim.softmax(-1)
    ad = a * keep.float() * ks if drop else a
    oref = (ad @ hd(vr)).transpose(1, 2).reshape(B, n, 512)
    H.assert_close(attn, a, 1e-5, "attn")
    H.assert_close(o, oref, 1e-5, "o")
    do = synth.normal((B, n, 512), 9, "do").to(DEV)
    dA = synth.normal((B, 8, n, m), 9, "dA").to(DEV)
    gq, gk, gv, gb = torch.autograd.grad((oref * do).sum() + (a * dA).sum(), (qr, kr, vr, br))
    ds = torch.empty_like(attn)
    dq = torch.empty(B, n, 512, device=DEV)
    dkv = torch.empty(2, B, m, 512, device=DEV)
    parts = torch.empty(_lib.load().dml_da2_cols_chunks(B, n, m), 2, B, m, 512, device=DEV)
    call("dml_da2_attn_bwd", ptr(q), ptr(k), ptr(v), ptr(attn), ptr(do), ptr(dA), ptr(keep) if drop else None, ks, B, n, m, scale,
         ptr(ws), ptr(ds), ptr(dq), ptr(parts), ptr(dkv), st())
    H.assert_close(ds, gb, 1e-4, "dS")
    H.assert_close(dq, gq, 1e-4, "dq")
    H.assert_close(dkv[0], gk, 1e-4, "dk")
    H.assert_close(dkv[1], gv, 1e-4, "dv")


# ---------------------------------------------------------------------------------------------------------------------
# module
# ---------------------------------------------------------------------------------------------------------------------
def _module(seed):
    mod = DeformCrossAttention2D(dim=128, dim_head=64, heads=8, dropout=0.1, downsample_factor=4, offset_scale=4, offset_groups=8,
                                 offset_kernel_size=6)
    mod.load_state_dict(synth.fill_like(H.attn2d_shapes(""), seed, gain=2.0), strict=True)
    return mod.to(DEV)


@pytest.mark.parametrize("c", DEFORM2D_CASES, ids=lambda c: c["name"])
def test_module_matches_reference_goldens(c):
    G = H.golden(c["name"])
    mod = _module(c["seed"]).eval()
    n = c["side"] ** 2
    x1 = synth.normal((c["b"], 128, n), c["seed"], "x1").to(DEV).requires_grad_()
    x2 = synth.normal((c["b"], 128, n), c["seed"], "x2").to(DEV).requires_grad_()
    r = synth.normal((c["b"], 128, n), c["seed"], "r").to(DEV)
    out, attn = mod(x1, x2)
    _, vgrid = mod(x1, x2, return_vgrid=True)
    assert out.shape == (c["b"], 128, n) and attn.shape[:3] == (c["b"], 8, n)
    r2 = synth.normal(tuple(attn.shape), c["seed"], "r2").to(DEV)
    H.assert_close(vgrid.cpu(), G["vgrid"], 1e-5, "vgrid")
    H.assert_close(thin(out.cpu()), G["out"], TOL, "out")
    H.assert_close(thin(attn.cpu()), G["attn"], TOL, "attn")
    ((out * r).sum() + (attn * r2).sum()).backward()
    H.assert_close(thin(x1.grad.cpu()), G["gx1"], TOL, "gx1")
    H.assert_close(thin(x2.grad.cpu()), G["gx2"], TOL, "gx2")
    for k, p in mod.named_parameters():
        H.assert_close(thin(p.grad.cpu()), G["grad." + k], tol_of(k), "grad " + k, atol=1e-3 if k.endswith("mlp.2.bias") else 0.0)


@pytest.mark.parametrize("B,side,train", [(4, 50, False), (2, 50, True), (1, 64, False), (8, 12, True), (3, 8, False), (1, 101, False), (2, 6, False),
                                          (16, 10, False)])
def test_module_matches_oracle_at_bag_size(B, side, train):
    """The teacher's batch (config: batch_size 4, 2 500 patches, 144 keys), training mode with the attention dropout (same keep
    mask through the oracle), gradients arriving at out, attn and vgrid."""
    seed = 100 + side + B
    mod = _module(seed).train(train)
    n = side * side
    m = O2.kv_side(side) ** 2
    x1 = synth.normal((B, 128, n), seed, "x1").to(DEV).requires_grad_()
    x2 = synth.normal((B, 128, n), seed, "x2").to(DEV).requires_grad_()
    r = synth.normal((B, 128, n), seed, "r").to(DEV)
    r2 = synth.normal((B, 8, n, m), seed, "r2").to(DEV)
    r3 = synth.normal((B * 8, 2, int(math.isqrt(m)), int(math.isqrt(m))), seed, "r3").to(DEV)
    torch.manual_seed(seed)
    out, attn = mod(x1, x2)
    torch.manual_seed(seed)
    keep = (torch.rand(B, 8, n, m, device=DEV) >= 0.1) if train else None
    torch.manual_seed(seed)
    out_v, vgrid = mod(x1, x2, return_vgrid=True)
    assert torch.equal(out_v, out)
    loss = (out * r).sum() + (attn * r2).sum() + (vgrid * r3).sum()
    gs = torch.autograd.grad(loss, [x1, x2] + list(mod.parameters()))
    # The oracle runs twice on the same device, in fp32 (the reference's arithmetic) and in fp64.  Every tensor must match ONE of
    # them within tolerance: the position-bias MLP gradients are ill-conditioned sums where two fp32 results differ from each other
    # by more than either differs from fp64 (tol_of), while the bilinear sampling path is DISCONTINUOUS in the sampling position
    # (floor): a key within one ulp of a grid line legitimately takes the other cell in fp64 (seen at side = 101: the fp32 oracle
    # and the kernels agree to 1e-5 and both sit 8e-3 from fp64 on d x1 / to_q / to_offsets).
    names = [k for k, _ in mod.named_parameters()]

    def oracle(dt):
        P = {k: v.detach().to(dt).requires_grad_() for k, v in mod.state_dict().items()}
        y1, y2 = x1.detach().to(dt).requires_grad_(), x2.detach().to(dt).requires_grad_()
        oo, oa, ov = O2.deform_cross_attention_2d(y1, y2, P, drop_keep=keep, drop_p=0.1)
        rs = torch.autograd.grad((oo * r.to(dt)).sum() + (oa * r2.to(dt)).sum() + (ov * r3.to(dt)).sum(), [y1, y2] + [P[k] for k in names])
        return [t.float() for t in (oo, oa, ov)] + [t.float() for t in rs]

    ref32, ref64 = oracle(torch.float32), oracle(torch.float64)
    labels = ["out", "attn", "vgrid", "gx1", "gx2"] + names
    ours = [out, attn, vgrid] + list(gs)
    report, bad = {}, {}
    for name, a, b32, b64 in zip(labels, ours, ref32, ref64):
        e32, e64 = max(H.rel_l2(a, b32), H.max_rel(a, b32)), max(H.rel_l2(a, b64), H.max_rel(a, b64))
        report[name] = (f"{e32:.1e}", f"{e64:.1e}")
        tol = 1e-5 if name == "vgrid" else tol_of(name)
        if min(e32, e64) > tol and not (name.endswith("mlp.2.bias") and float((a - b64).abs().max()) <= 1e-3):
            bad[name] = (e32, e64)
    print("module vs oracle (fp32, fp64):", report)
    assert not bad, bad


def test_large_bag_stress_100k_patches():
    """BASELINE configs[3]: one slide of 316 x 316 = 99 856 patches, 79 x 79 = 6 241 sampled keys (a 20 GB attention map).
    Size-independent checks: probability rows sum to one, and 48 random query rows of the map and of the output equal the oracle
    evaluated for those rows only; the backward runs and is finite."""
    side, seed = 316, 123
    n = side * side
    mod = _module(seed).eval()
    g = torch.Generator(device=DEV).manual_seed(seed)
    x1 = torch.randn(1, 128, n, device=DEV, generator=g).requires_grad_()
    x2 = torch.randn(1, 128, n, device=DEV, generator=g).requires_grad_()
    out, attn = mod(x1, x2)
    assert attn.shape == (1, 8, n, 6241) and out.shape == (1, 128, n)
    dev = float((attn.sum(-1) - 1.0).abs().max())
    assert dev < 1e-4, dev
    rows = torch.randint(0, n, (48,), device=DEV, generator=g)
    P = {k: v.detach() for k, v in mod.state_dict().items()}
    with torch.no_grad():
        oo, oa, _ = O2.deform_cross_attention_2d(x1.detach(), x2.detach(), P, rows=rows)
    H.assert_close(attn[:, :, rows], oa, TOL, "attn rows")
    H.assert_close(out[:, :, rows], oo, TOL, "out rows")
    out.sum().backward()
    assert bool(torch.isfinite(x1.grad).all()) and bool(torch.isfinite(x2.grad).all())
    assert all(bool(torch.isfinite(p.grad).all()) for p in mod.parameters())
    assert float(x1.grad.abs().max()) > 0 and float(x2.grad.abs().max()) > 0


def test_cluster_merge_100k_tokens():
    """ClusterMergeNet at N = 99 856 (path_cluster_num 0.0008 -> 80 clusters): structural properties of the clustering that hold
    at any size - every index in range, every cluster non-empty, merged rows = the weighted mean of their members."""
    N, seed = 99856, 77
    mod = ClusterMergeNet(sample_ratio=0.0008, dim_out=128)
    mod.load_state_dict(synth.fill_like(H.cluster_shapes(), seed), strict=True)
    mod = mod.to(DEV)
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(1, N, 128, device=DEV, generator=g).requires_grad_()
    tok = dict(x=x, token_num=N, idx_token=torch.arange(N, device=DEV)[None], agg_weight=x.new_ones(1, N, 1))
    down, full = mod(tok)
    K = math.ceil(N * 0.0008)
    idx = down["idx_token"]
    assert down["x"].shape == (1, K, 128) and int(idx.min()) == 0 and int(idx.max()) == K - 1
    counts = torch.bincount(idx[0], minlength=K)
    assert int(counts.min()) >= 1 and int(counts.sum()) == N
    w = full["token_score"].exp()[0, :, 0]
    xn = full["x"][0]
    ref = torch.zeros(K, 128, device=DEV).index_add_(0, idx[0], xn * w[:, None]) / (torch.zeros(K, device=DEV).index_add_(0, idx[0], w) + 1e-6)[:, None]
    H.assert_close(down["x"][0], ref, 1e-4, "merged")
    down["x"].sum().backward()
    assert bool(torch.isfinite(x.grad).all())


# ---------------------------------------------------------------------------------------------------------------------
# ClusterMergeNet
# ---------------------------------------------------------------------------------------------------------------------
def _cluster_module(c):
    mod = ClusterMergeNet(sample_ratio=c["ratio"], dim_out=128)
    mod.load_state_dict(synth.fill_like(H.cluster_shapes(), c["seed"]), strict=True)
    noise = (synth.uniform((c["B"], c["N"]), c["seed"], "noise", 0.5) + 0.5).to(DEV)
    mod.noise_fn = lambda B, N, dev: noise
    return mod.to(DEV)


@pytest.mark.parametrize("c", CLUSTER_CASES, ids=lambda c: c["name"])
def test_cluster_merge_matches_reference_goldens(c):
    G = H.golden(c["name"])
    mod = _cluster_module(c)
    x = synth.normal((c["B"], c["N"], 128), c["seed"], "x").to(DEV).requires_grad_()
    tok = dict(x=x, token_num=c["N"], idx_token=torch.arange(c["N"], device=DEV)[None].repeat(c["B"], 1),
               agg_weight=x.new_ones(c["B"], c["N"], 1))
    down, full = mod(tok)
    assert down["token_num"] == G["merged"].shape[1] and full["token_score"].shape == (c["B"], c["N"], 1)
    assert torch.equal(down["idx_token"].cpu(), G["idx_cluster"]), "cluster indices must be bit-exact"
    H.assert_close(down["x"].cpu(), G["merged"], TOL, "merged")
    r = synth.normal(tuple(down["x"].shape), c["seed"], "r").to(DEV)
    (down["x"] * r).sum().backward()
    H.assert_close(thin(x.grad.cpu()), G["gx"], TOL, "gx")
    for k, p in mod.named_parameters():
        # d/d(score.bias) is mathematically ~0 (a common factor e^b of all token weights cancels in the weighted mean, up to the
        # 1e-6 of ClusterMergeNet.py:159): rounding noise only
        H.assert_close(p.grad.cpu(), G["grad." + k], TOL, "grad " + k, atol=1e-4 if k == "score.bias" else 0.0)


@pytest.mark.parametrize("B,N,K", [(2, 1000, 13), (1, 4099, 40)])
def test_dpc_knn_matches_oracle(B, N, K):
    """Clustered tokens (well separated densities), N not a multiple of the 32-token tile: indices bit-exact."""
    centres = synth.normal((B, 24, 128), 11, "centres")
    pick = torch.from_numpy(synth._rng(11, "pick").integers(0, 24, size=(B, N)))
    x = (torch.gather(centres, 1, pick[..., None].expand(B, N, 128)) + 0.35 * synth.normal((B, N, 128), 11, "jit")).to(DEV)
    noise = (synth.uniform((B, N), 11, "noise", 0.5) + 0.5).to(DEV)
    idx, down = ops2d.dpc_knn(x, K, noise)
    ridx, rdown = O2.dpc_knn(x.cpu(), K, noise.cpu())
    assert torch.equal(down.cpu(), rdown)
    assert torch.equal(idx.cpu(), ridx)
