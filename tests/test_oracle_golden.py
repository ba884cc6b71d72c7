"""Pin the oracle restatement against outputs of the reference itself (tests/golden/*.npz,
made by oracle/make_goldens.py from /root/reference).  fp32 on CPU: tolerance 2e-5."""
import pytest
import torch

from dml_b200 import synth
from oracle import coattn, deform1d, losses, nystrom, towers
from oracle.golden_cases import (COATTN_CASES, DEFORM_CASES, LOSS_CASES, NYSTROM_CASES, PATHOMIC_CASES, TOWER_CASES, TRANSMIL_CASES,
                                 loss_inputs, thin)
from tests import helpers as H

TOL = 2e-5


def _check_param_grads(P, loss, G, tol=TOL, skip=()):
    names = [k for k in P if P[k].requires_grad]
    gs = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    seen = 0
    for k, g in zip(names, gs):
        key = "grad." + k
        if g is None:
            assert key not in G, f"oracle gives no grad for {k} but the reference does"
            continue
        if key not in G:
            assert float(g.abs().max()) == 0.0 or k in skip, f"reference has no grad for {k}"
            continue
        # d/d(mlp.2.bias) is mathematically 0 (softmax shift invariance): rounding noise only
        H.assert_close(thin(g), G[key], tol, key, atol=1e-4 if k.endswith("rel_pos_bias.mlp.2.bias") else 0.0)
        seen += 1
    assert seen > 0


@pytest.mark.parametrize("c", DEFORM_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("row_block", [None, 64])
def test_deform1d_matches_reference(c, row_block):
    G = H.golden(c["name"])
    P = H.leafify(synth.fill_like(H.deform_shapes(), c["seed"], gain=2.0))
    x1 = synth.normal((c["b"], 128, c["n"]), c["seed"], "x1").requires_grad_()
    x2 = synth.normal((c["b"], 128, c["n"]), c["seed"], "x2").requires_grad_()
    r = synth.normal((c["b"], 128, c["n"]), c["seed"], "r")
    out, aux = deform1d.deform_cross_attention_1d(x1, x2, P, offset_scale=2, row_block=row_block, return_aux=True)
    assert aux["n_kv"] == deform1d.kv_length(c["n"]) == G["vgrid"].shape[-1]
    H.assert_close(aux["vgrid"], G["vgrid"], 1e-6, "vgrid")
    H.assert_close(thin(out), G["out"], TOL, "out")
    loss = (out * r).sum()
    gx1, gx2 = torch.autograd.grad(loss, (x1, x2), retain_graph=True)
    H.assert_close(thin(gx1), G["gx1"], TOL, "gx1")
    H.assert_close(thin(gx2), G["gx2"], TOL, "gx2")
    _check_param_grads(P, loss, G)


@pytest.mark.parametrize("n", [5, 33, 193, 1025])
def test_degenerate_gather_closed_form_is_bit_exact_for_odd_n(n):
    """grid_sample_1d == centre token x tent weight, torch.equal for odd n."""
    feats = synth.normal((8, 32, n), 3, "feats")
    nkv = deform1d.kv_length(n)
    grid = deform1d.normalize_grid(torch.arange(nkv) + synth.uniform((8, nkv), 3, "off", 2.0))
    a = deform1d.grid_sample_1d_literal(feats, grid)
    b = deform1d.grid_sample_1d_closed(feats, grid)
    assert torch.equal(a, b)
    i0, i1, w0, w1 = deform1d.centre_taps(n)
    assert i0 == (n - 1) // 2 and w1 == 0.0 and w0 == 1.0


@pytest.mark.parametrize("n", [4, 128, 1000])
def test_degenerate_gather_even_n_close(n):
    feats = synth.normal((4, 32, n), 4, "feats")
    nkv = deform1d.kv_length(n)
    grid = deform1d.normalize_grid(torch.arange(nkv) + synth.uniform((4, nkv), 4, "off", 2.0))
    a = deform1d.grid_sample_1d_literal(feats, grid)
    b = deform1d.grid_sample_1d_closed(feats, grid)
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    assert deform1d.centre_taps(n)[2:] == (0.5, 0.5)


def test_kv_length_and_landmark_geometry_integers():
    """Integer artefacts quoted in SURVEY.md section 8 header."""
    assert deform1d.kv_length(16385) == 4096
    assert deform1d.kv_length(6) == 1 and deform1d.kv_length(4) == 1 and deform1d.kv_length(2049) == 512
    assert nystrom.landmark_geometry(16385, 256) == (255, 16640, 65)
    assert nystrom.landmark_geometry(6085, 256) == (59, 6144, 24)
    assert nystrom.landmark_geometry(512, 256) == (0, 512, 2)
    assert towers.square_side(6000) == 78 and towers.square_side(16384) == 128


@pytest.mark.parametrize("c", NYSTROM_CASES, ids=lambda c: c["name"])
def test_nystrom_matches_reference(c):
    G = H.golden(c["name"])
    P = H.leafify(synth.fill_like(H.nystrom_shapes(c["dim"], c["dim_head"]), c["seed"], gain=2.0))
    x = synth.normal((c["b"], c["n"], c["dim"]), c["seed"], "x").requires_grad_()
    r = synth.normal((c["b"], c["n"], c["dim"]), c["seed"], "r")
    out = nystrom.nystrom_attention(x, P, heads=8, dim_head=c["dim_head"], num_landmarks=c["m"])
    H.assert_close(thin(out), G["out"], TOL, "out")
    loss = (out * r).sum()
    (gx,) = torch.autograd.grad(loss, (x,), retain_graph=True)
    H.assert_close(thin(gx), G["gx"], TOL, "gx")
    _check_param_grads(P, loss, G)


@pytest.mark.parametrize("c", TOWER_CASES, ids=lambda c: c["name"])
def test_deform_cross_trans_mil_matches_reference(c):
    G = H.golden(c["name"])
    P = H.leafify(synth.fill_like(H.dctmil_shapes(), c["seed"]))
    path = synth.synthetic_bag(c["N"], c["seed"], c["B"])["x_path"].requires_grad_()
    omic = synth.normal((c["B"], 128), c["seed"], "omic").requires_grad_()
    enc, logits = towers.deform_cross_trans_mil(path, omic, P)
    H.assert_close(enc, G["encoded"], TOL, "encoded")
    H.assert_close(logits, G["logits"], TOL, "logits")
    loss = (enc * synth.normal(enc.shape, c["seed"], "r_enc")).sum() + (logits * synth.normal(logits.shape, c["seed"], "r_log")).sum()
    gpath, gomic = torch.autograd.grad(loss, (path, omic), retain_graph=True)
    H.assert_close(thin(gpath[0]), G["gpath"], TOL, "gpath")
    H.assert_close(gomic, G["gomic"], TOL, "gomic")
    _check_param_grads(P, loss, G)


@pytest.mark.parametrize("c", TRANSMIL_CASES, ids=lambda c: c["name"])
def test_transmil_matches_reference(c):
    G = H.golden(c["name"])
    P = H.leafify(synth.fill_like(H.transmil_shapes(), c["seed"]))
    x = synth.synthetic_bag(c["N"], c["seed"], c["B"])["x_path"].requires_grad_()
    enc, logits = towers.trans_mil(x, P)
    H.assert_close(enc, G["encoded"], 5e-5, "encoded")
    H.assert_close(logits, G["logits"], 5e-5, "logits")
    loss = (enc * synth.normal(enc.shape, c["seed"], "r_enc")).sum() + (logits * synth.normal(logits.shape, c["seed"], "r_log")).sum()
    (gx,) = torch.autograd.grad(loss, (x,), retain_graph=True)
    H.assert_close(thin(gx[0]), G["gx"], 5e-5, "gx")
    _check_param_grads(P, loss, G, tol=5e-5)


@pytest.mark.parametrize("c", PATHOMIC_CASES, ids=lambda c: c["name"])
def test_deform_pathomic_net_matches_reference(c):
    G = H.golden(c["name"])
    P = H.leafify(synth.fill_like(H.pathomic_shapes(), c["seed"]))
    bag = synth.synthetic_bag(c["N"], c["seed"], c["B"])
    feats, vt, vi, logits = towers.deform_pathomic_net(bag["x_path"], bag["x_omic_tumor"], bag["x_omic_immune"], P,
                                                       task_type=c["task"])
    H.assert_close(feats, G["features"], TOL, "features")
    H.assert_close(logits[0], G["hazard_tumor"], TOL, "hazard_tumor")
    H.assert_close(logits[1], G["hazard_immune"], TOL, "hazard_immune")
    H.assert_close(logits[2], G["hazard"], TOL, "hazard")
    label = bag["label_diag"] if c["task"] == "diag2021" else bag["label_surv"]
    loss = towers.bag_loss(logits, label, c["task"], bag["censor"])
    H.assert_close(loss, G["loss"], TOL, "loss")
    _check_param_grads(P, loss, G)


@pytest.mark.parametrize("c", COATTN_CASES, ids=lambda c: c["name"])
def test_raw_score_multihead_attention_matches_reference(c):
    """models/MultiheadAttention.py as MCAT / CMTA call it (1 head, E = 256), both directions; a gradient reaches the raw scores."""
    G = H.golden(c["name"])
    P = H.leafify(synth.fill_like(H.mha_shapes(256), c["seed"], gain=2.0))
    q = synth.normal((c["L"], c["B"], 256), c["seed"], "query").requires_grad_()
    kv = synth.normal((c["S"], c["B"], 256), c["seed"], "key").requires_grad_()
    r = synth.normal((c["L"], c["B"], 256), c["seed"], "r")
    r2 = synth.normal((c["B"], 1, c["L"], c["S"]), c["seed"], "r2", scale=0.1)
    out, raw = coattn.multihead_attention_raw(q, kv, P)
    H.assert_close(thin(out), G["out"], TOL, "out")
    H.assert_close(thin(raw), G["raw"], TOL, "raw scores")
    loss = (out * r).sum() + (raw * r2).sum()
    gq, gkv = torch.autograd.grad(loss, (q, kv), retain_graph=True)
    H.assert_close(thin(gq), G["gq"], TOL, "d query")
    H.assert_close(thin(gkv), G["gkv"], TOL, "d key/value")
    _check_param_grads(P, loss, G)


@pytest.mark.parametrize("c", LOSS_CASES, ids=lambda c: c["name"])
def test_batch_losses_match_reference(c):
    """utils/loss.py PathBatchLoss / OmicDomainScaleLoss / BatchLoss at world_size 1: values and input gradients."""
    G = H.golden(c["name"])
    x = {k: v.requires_grad_() for k, v in loss_inputs(c).items()}
    pb = losses.path_batch_loss(x["a1_10"], x["a1_20"])
    od = losses.omic_domain_scale_loss(x["a1_10"], x["a1_20"], x["a2_10"], x["a2_20"])
    bl = losses.batch_loss(x["omic"], x["vgrid"])
    H.assert_close(pb, G["path_batch"], TOL, "PathBatchLoss")
    H.assert_close(od, G["omic_domain"], TOL, "OmicDomainScaleLoss")
    H.assert_close(bl, G["batch"], TOL, "BatchLoss")
    g = torch.autograd.grad(pb.sum(), (x["a1_10"], x["a1_20"]), retain_graph=True)
    H.assert_close(thin(g[0]), G["pb.g10"], TOL, "d PathBatchLoss / d att10")
    H.assert_close(thin(g[1]), G["pb.g20"], TOL, "d PathBatchLoss / d att20")
    g = torch.autograd.grad(od, (x["a1_10"], x["a1_20"], x["a2_10"], x["a2_20"]), retain_graph=True)
    for k, v in zip(("a1_10", "a1_20", "a2_10", "a2_20"), g):
        H.assert_close(thin(v), G["od.g_" + k], 5 * TOL, "d OmicDomainScaleLoss / d " + k)
    g = torch.autograd.grad(bl.sum(), (x["omic"], x["vgrid"]))
    H.assert_close(thin(g[0]), G["bl.g_omic"], TOL, "d BatchLoss / d omic")
    H.assert_close(thin(g[1]), G["bl.g_vgrid"], TOL, "d BatchLoss / d vgrid")
