"""Pin the oracle restatement of DeformCrossAttention2D / ClusterMergeNet against outputs of the reference
itself (tests/golden/deform2d_*.npz, clustermerge_*.npz made by oracle/make_goldens.py).  fp32 on CPU: tolerance 2e-5;
cluster indices bit-exact."""
import pytest
import torch

from dml_b200 import synth
from oracle import deform2d
from oracle.golden_cases import CLUSTER_CASES, DEFORM2D_CASES, thin
from tests import helpers as H

TOL = 2e-5


def deform2d_inputs(c):
    n = c["side"] ** 2
    P = H.leafify(synth.fill_like(H.attn2d_shapes(""), c["seed"], gain=2.0))
    x1 = synth.normal((c["b"], 128, n), c["seed"], "x1").requires_grad_()
    x2 = synth.normal((c["b"], 128, n), c["seed"], "x2").requires_grad_()
    r = synth.normal((c["b"], 128, n), c["seed"], "r")
    m = deform2d.kv_side(c["side"]) ** 2
    r2 = synth.normal((c["b"], 8, n, m), c["seed"], "r2")
    return P, x1, x2, r, r2


@pytest.mark.parametrize("c", DEFORM2D_CASES, ids=lambda c: c["name"])
def test_deform2d_matches_reference(c):
    G = H.golden(c["name"])
    P, x1, x2, r, r2 = deform2d_inputs(c)
    with torch.backends.mkldnn.flags(enabled=False):
        out, attn, vgrid = deform2d.deform_cross_attention_2d(x1, x2, P)
        assert vgrid.shape == G["vgrid"].shape
        H.assert_close(vgrid, G["vgrid"], 1e-6, "vgrid")
        H.assert_close(thin(out), G["out"], TOL, "out")
        H.assert_close(thin(attn), G["attn"], TOL, "attn")
        loss = (out * r).sum() + (attn * r2).sum()
        names = list(P)
        gs = torch.autograd.grad(loss, [x1, x2] + [P[k] for k in names], allow_unused=True)
    H.assert_close(thin(gs[0]), G["gx1"], TOL, "gx1")
    H.assert_close(thin(gs[1]), G["gx2"], TOL, "gx2")
    for k, g in zip(names, gs[2:]):
        # d/d(mlp.2.bias) is mathematically 0 (softmax shift invariance): rounding noise only
        H.assert_close(thin(g), G["grad." + k], TOL, k, atol=1e-3 if k.endswith("mlp.2.bias") else 0.0)


def test_row_subset_oracle_equals_full_rows():
    """The rows= variant used for the 100k-token spot checks is the full computation restricted to those query rows."""
    c = DEFORM2D_CASES[0]
    P, x1, x2, _, _ = deform2d_inputs(c)
    rows = torch.tensor([0, 7, 199, 200, 399])
    with torch.no_grad():
        out, attn, vg = deform2d.deform_cross_attention_2d(x1, x2, P)
        o2, a2, v2 = deform2d.deform_cross_attention_2d(x1, x2, P, rows=rows)
    assert torch.equal(vg, v2)
    H.assert_close(o2, out[:, :, rows], 1e-6, "out rows")
    H.assert_close(a2, attn[:, :, rows], 1e-6, "attn rows")


def test_kv_side_integers():
    assert deform2d.kv_side(50) == 12 and deform2d.kv_side(316) == 79 and deform2d.kv_side(23) == 5 and deform2d.kv_side(6) == 1


def cluster_inputs(c):
    P = H.leafify(synth.fill_like(H.cluster_shapes(), c["seed"]))
    x = synth.normal((c["B"], c["N"], 128), c["seed"], "x").requires_grad_()
    noise = synth.uniform((c["B"], c["N"]), c["seed"], "noise", 0.5) + 0.5
    return P, x, noise


@pytest.mark.parametrize("c", CLUSTER_CASES, ids=lambda c: c["name"])
def test_cluster_merge_matches_reference(c):
    G = H.golden(c["name"])
    P, x, noise = cluster_inputs(c)
    merged, idx, _, _ = deform2d.cluster_merge_net(x, P, c["ratio"], noise)
    assert torch.equal(idx, G["idx_cluster"])
    H.assert_close(merged, G["merged"], TOL, "merged")
    r = synth.normal(tuple(merged.shape), c["seed"], "r")
    names = list(P)
    gs = torch.autograd.grad((merged * r).sum(), [x] + [P[k] for k in names])
    H.assert_close(thin(gs[0]), G["gx"], TOL, "gx")
    for k, g in zip(names, gs[1:]):
        H.assert_close(g, G["grad." + k], TOL, k)
