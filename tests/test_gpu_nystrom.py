"""NystromAttention / TransLayer / PPEG on the pair kernels: kernel-level checks against plain torch fp32/fp64 maths, the fused
module against the oracle at sizes with front padding, and CUDA-graph capture of a whole TransMIL training step."""
import pytest
import torch
import torch.nn.functional as F

from dml_b200 import synth
from dml_b200._lib import call, ptr, stream
from dml_b200.mil import PPEG, TransLayer
from dml_b200.model import Args, define_net
from dml_b200.NystromAttention import NystromAttention
from dml_b200.pairs import Pair
from oracle import nystrom as ON
from oracle import towers
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3          # north_star: fp32-class path


def test_landmark_pool_pair_is_exact_segment_mean():
    B, Hh, d, n_pad, l = 2, 8, 16, 192, 3
    W = Hh * d
    qkv = synth.normal((B, n_pad, 3 * W), 3, "qkv").to(DEV)
    qkv[:, :5] = 0.0                                             # front padding rows count in the mean (quirk Q8)
    P = Pair.from_f32(qkv)
    m = n_pad // l
    out = Pair.empty((2, B, Hh, m, d), DEV)
    call("dml_ny_landmark_pool", ptr(P.planes), P.planes.stride(0), 3 * W, B, n_pad, l, Hh, d, 0.5 / l, 1.0 / l, ptr(out.planes),
         out.planes.stride(0), stream())
    src = P.float().double()
    for which, mult in ((0, 0.5 / l), (1, 1.0 / l)):
        ref = src[..., which * W:(which + 1) * W].reshape(B, m, l, Hh, d).sum(2).permute(0, 2, 1, 3) * mult
        H.assert_close(out.float()[which], ref, 1e-5, f"landmarks {which}")


def test_long_row_softmax_pair_forward_backward():
    rows, cols = 37, 16640
    x = (synth.normal((rows, cols), 4, "x") * 3).to(DEV)
    y = Pair.empty((rows, cols), DEV)
    call("dml_ny_softmax_rows_fwd", ptr(x), rows, cols, ptr(y.planes), y.planes.stride(0), stream())
    ref = x.double().softmax(-1)
    H.assert_close(y.float(), ref, 2e-5, "softmax")
    dy = synth.normal((rows, cols), 4, "dy").to(DEV)
    dx = Pair.empty((rows, cols), DEV)
    call("dml_ny_softmax_rows_bwd", ptr(y.planes), y.planes.stride(0), ptr(dy), rows, cols, ptr(dx.planes), dx.planes.stride(0), stream())
    yy = y.float().double()
    H.assert_close(dx.float(), yy * (dy.double() - (dy.double() * yy).sum(-1, keepdim=True)), 2e-5, "softmax backward")


@pytest.mark.parametrize("side,C", [(13, 512), (128, 512), (9, 96)])
def test_ppeg_stencil_matches_the_three_depthwise_convolutions(side, C):
    B = 2
    mod = PPEG(dim=C).to(DEV)
    with torch.no_grad():
        for p in mod.parameters():
            p.copy_(synth.uniform(tuple(p.shape), 9, "p" + str(p.numel()), 0.2).to(DEV))
    x = synth.normal((B, 1 + side * side, C), 9, "x").to(DEV).requires_grad_()
    r = synth.normal((B, 1 + side * side, C), 9, "r").to(DEV)
    y = mod(x, side, side)
    g = torch.autograd.grad((y * r).sum(), [x] + list(mod.parameters()))
    xd = x.detach().double().requires_grad_()
    P = {k: v.detach().double().requires_grad_() for k, v in mod.state_dict().items()}
    ref = towers.ppeg(xd, P, side, side)
    gr = torch.autograd.grad((ref * r.double()).sum(), [xd] + [P[k] for k, _ in mod.named_parameters()])
    H.assert_close(y, ref, 1e-5, "PPEG")
    for nm, a, b in zip(["x"] + [k for k, _ in mod.named_parameters()], g, gr):
        H.assert_close(a, b, 2e-5, "PPEG grad " + nm)


@pytest.mark.parametrize("b,n,dim,dh,m", [(1, 1000, 512, 64, 256), (2, 700, 256, 32, 128), (1, 512, 128, 16, 64)])
def test_nystrom_attention_against_oracle_with_front_padding(b, n, dim, dh, m):
    seed = 17
    mod = NystromAttention(dim=dim, dim_head=dh, heads=8, num_landmarks=m, pinv_iterations=6, residual=True, dropout=0.1)
    mod.load_state_dict(synth.fill_like(H.nystrom_shapes(dim, dh), seed, 2.0), strict=True)
    mod.to(DEV).eval()
    x = synth.normal((b, n, dim), seed, "x").to(DEV).requires_grad_()
    r = synth.normal((b, n, dim), seed, "r").to(DEV)
    out = mod(x)
    names = [k for k, _ in mod.named_parameters()]
    g = torch.autograd.grad((out * r).sum(), [x] + list(mod.parameters()))
    P = {k: v.detach().double().requires_grad_() for k, v in mod.state_dict().items()}
    xd = x.detach().double().requires_grad_()
    ref, aux = ON.nystrom_attention(xd, P, heads=8, dim_head=dh, num_landmarks=m, return_aux=True)
    assert (aux["pad"], aux["n_pad"], aux["l"]) == ON.landmark_geometry(n, m)
    gr = torch.autograd.grad((ref * r.double()).sum(), [xd] + [P[k] for k in names])
    H.assert_close(out, ref, TOL, "out")
    for nm, a, bb in zip(["x"] + names, g, gr):
        H.assert_close(a, bb, TOL, "grad " + nm)


def test_translayer_fuses_its_layernorm():
    seed, n = 23, 777
    mod = TransLayer(dim=512)
    sd = {"norm.weight": 1.0 + 0.1 * synth.uniform((512,), seed, "nw"), "norm.bias": synth.uniform((512,), seed, "nb", 0.1)}
    sd.update({"attn." + k: v for k, v in synth.fill_like(H.nystrom_shapes(512, 64), seed, 1.5).items()})
    mod.load_state_dict(sd, strict=True)
    mod.to(DEV).eval()
    x = synth.normal((1, n, 512), seed, "x").to(DEV).requires_grad_()
    r = synth.normal((1, n, 512), seed, "r").to(DEV)
    out = mod(x)
    names = [k for k, _ in mod.named_parameters()]
    g = torch.autograd.grad((out * r).sum(), [x] + list(mod.parameters()))
    P = {k: v.detach().double().requires_grad_() for k, v in mod.state_dict().items()}
    xd = x.detach().double().requires_grad_()
    ref = towers.trans_layer(xd, P)
    gr = torch.autograd.grad((ref * r.double()).sum(), [xd] + [P[k] for k in names])
    H.assert_close(out, ref, TOL, "out")
    for nm, a, bb in zip(["x"] + names, g, gr):
        H.assert_close(a, bb, TOL, "grad " + nm)


def test_transmil_training_step_is_graph_capturable_and_matches_eager():
    """bench.py replays the TransMIL step as one CUDA graph: capture must succeed (no host synchronisation, no legacy-stream
    work anywhere in the path) and the replayed losses must equal the eager step's (eval mode: dropout off)."""
    from dml_b200.graph import GraphedTrainStep
    N = 900
    net = define_net(Args(mode="path", label_dim=3))
    net.load_state_dict(synth.fill_like(H.transmil_shapes(), 42), strict=True)
    net.to(DEV).eval()
    bag = synth.synthetic_bag(N, 5)
    inp = {"x": bag["x_path"].to(torch.bfloat16).to(DEV), "label": bag["label_grade"].to(DEV)}
    w = torch.tensor([1.47, 1.51, 1.0], device=DEV)
    loss_fn = lambda out, b: F.cross_entropy(out[1], b["label"], weight=w)   # noqa: E731
    eager = float(loss_fn(net(inp["x"]), inp))
    step = GraphedTrainStep(net, loss_fn, inp, optimizer=None, model_keys=("x",), warmup=2)
    l1 = float(step(inp))
    l2 = float(step(inp))
    assert abs(l1 - eager) <= 1e-5 * max(1.0, abs(eager)) and l1 == l2
    # and in train mode (to_out dropout live) the capture must still go through
    net.train()
    step = GraphedTrainStep(net, loss_fn, inp, optimizer=None, model_keys=("x",), warmup=2)
    assert torch.isfinite(step(inp)).all()


@pytest.mark.parametrize("NB,m", [(8, 256), (16, 64), (3, 40)])
def test_pinv_initial_iterate_and_its_adjoint(NB, m):
    """z0 = x^T / (max row abs-sum * max column abs-sum) with GLOBAL maxima (NystromAttention.py:20-27, quirk T3), and the full
    gradient through the transpose, the scale and both arg-max rows / columns, against torch autograd in fp64."""
    lib = __import__("dml_b200")._lib.load(check_device=True)
    x = synth.normal((NB, m, m), 17, "x").to(DEV)
    G = synth.normal((NB, m, m), 17, "g").to(DEV)
    add = synth.normal((NB, m, m), 17, "add").to(DEV)
    sums = torch.empty(lib.dml_ny_pinv_init_sums_floats(NB, m), device=DEV)
    z = Pair.empty((NB, m, m), DEV)
    call("dml_ny_pinv_init_fwd", ptr(x), NB, m, ptr(sums), ptr(z.planes), z.planes.stride(0), stream())
    xd = x.double().requires_grad_()
    ax = xd.abs()
    z_ref = xd.transpose(-1, -2) / (ax.sum(-1).max() * ax.sum(-2).max())
    H.assert_close(z.float(), z_ref, 1e-5, "z0")
    (g_ref,) = torch.autograd.grad(z_ref, xd, G.double())
    part = torch.empty(lib.dml_ny_pinv_init_part_floats(NB, m), device=DEV)
    dx = torch.empty_like(x)
    for addend in (None, add):
        call("dml_ny_pinv_init_bwd", ptr(G), ptr(x), ptr(sums), None if addend is None else ptr(addend), NB, m, ptr(part), ptr(dx),
             stream())
        H.assert_close(dx, g_ref + (0 if addend is None else addend.double()), 1e-5, "d x")
