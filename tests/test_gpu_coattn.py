"""Raw-score co-attention (MCAT / CMTA, models/MultiheadAttention.py) on the streaming kernels of csrc/coattn.cu: the module
against the reference's own goldens and against the oracle at bag sizes, both directions, with a gradient on the raw scores."""
import pytest
import torch

from dml_b200 import synth
from dml_b200.MultiheadAttention import MultiheadAttention
from oracle import coattn as OC
from oracle.golden_cases import COATTN_CASES, thin
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3          # north_star: fp32 path


def _module(seed):
    mod = MultiheadAttention(embed_dim=256, num_heads=1)
    mod.load_state_dict(synth.fill_like(H.mha_shapes(256), seed, gain=2.0), strict=True)
    return mod.to(DEV)


def _inputs(c):
    q = synth.normal((c["L"], c["B"], 256), c["seed"], "query").to(DEV).requires_grad_()
    kv = synth.normal((c["S"], c["B"], 256), c["seed"], "key").to(DEV).requires_grad_()
    r = synth.normal((c["L"], c["B"], 256), c["seed"], "r").to(DEV)
    r2 = synth.normal((c["B"], 1, c["L"], c["S"]), c["seed"], "r2", scale=0.1).to(DEV)
    return q, kv, r, r2


@pytest.mark.parametrize("c", COATTN_CASES, ids=lambda c: c["name"])
def test_module_matches_reference_goldens(c):
    G = H.golden(c["name"])
    mod = _module(c["seed"])
    q, kv, r, r2 = _inputs(c)
    out, raw = mod(q, kv, kv)
    assert out.shape == (c["L"], c["B"], 256) and raw.shape == (c["B"], 1, c["L"], c["S"])
    H.assert_close(thin(out.cpu()), G["out"], TOL, "out")
    H.assert_close(thin(raw.cpu()), G["raw"], TOL, "raw scores")
    loss = (out * r).sum() + (raw * r2).sum()
    loss.backward()
    H.assert_close(thin(q.grad.cpu()), G["gq"], TOL, "d query")
    H.assert_close(thin(kv.grad.cpu()), G["gkv"], TOL, "d key/value")
    for k, p in mod.named_parameters():
        H.assert_close(thin(p.grad.cpu()), G["grad." + k], TOL, "grad " + k)


@pytest.mark.parametrize("L,S,B", [(4, 2500, 8), (4, 16384, 2), (1, 1000, 1), (8, 777, 3), (3, 130, 2), (2500, 4, 8), (16384, 6, 1),
                                   (999, 8, 2), (513, 1, 2), (130, 5, 1)])
def test_module_matches_oracle_at_bag_sizes(L, S, B):
    """config_others.yaml:60 (B = 8, N = 2 500) and the 16k bag, every short-side count the dispatcher instantiates, ragged
    chunk tails; the bag side arrives as the transposed view the reference builds (model.py:1041: wsi_net(x).transpose(0, 1))."""
    seed = 70 + L % 7 + S % 5
    mod = _module(seed)
    few_q = L <= 8
    n_long = S if few_q else L
    bag = synth.normal((B, n_long, 256), seed, "bag").to(DEV).requires_grad_()
    few = synth.normal(((L if few_q else S), B, 256), seed, "few").to(DEV).requires_grad_()
    long_side = bag.transpose(0, 1)                       # [n_long, B, E] view over [B, n_long, E]
    q, kv = (few, long_side) if few_q else (long_side, few)
    out, raw = mod(q, kv, kv)
    r = synth.normal((L, B, 256), seed, "r").to(DEV)
    r2 = synth.normal((B, 1, L, S), seed, "r2", scale=0.1).to(DEV)
    ((out * r).sum() + (raw * r2).sum()).backward()

    P = {k: v.detach().double().requires_grad_() for k, v in mod.state_dict().items()}
    bag_d, few_d = bag.detach().double().requires_grad_(), few.detach().double().requires_grad_()
    qd, kvd = (few_d, bag_d.transpose(0, 1)) if few_q else (bag_d.transpose(0, 1), few_d)
    ro, rr = OC.multihead_attention_raw(qd, kvd, P)
    ((ro * r.double()).sum() + (rr * r2.double()).sum()).backward()
    H.assert_close(out, ro, TOL, "out")
    H.assert_close(raw, rr, TOL, "raw scores")
    H.assert_close(bag.grad, bag_d.grad, TOL, "d bag")
    H.assert_close(few.grad, few_d.grad, TOL, "d few")
    for k, p in mod.named_parameters():
        H.assert_close(p.grad, P[k].grad, TOL, "grad " + k)


def test_need_raw_false_returns_head_averaged_softmax():
    mod = _module(5)
    q = synth.normal((4, 2, 256), 5, "q").to(DEV)
    kv = synth.normal((300, 2, 256), 5, "kv").to(DEV)
    out, w = mod(q, kv, kv, need_raw=False)
    _, raw = mod(q, kv, kv)
    assert w.shape == (2, 4, 300)
    H.assert_close(w, torch.softmax(raw[:, 0], -1), 1e-6, "weights")
    assert mod(q, kv, kv, need_weights=False)[1] is None


def test_unserved_configurations_raise():
    mod = _module(6)
    a = synth.normal((20, 1, 256), 6, "a").to(DEV)
    with pytest.raises(NotImplementedError):
        mod(a, a, a)                                          # 20 x 20: no short side
    with pytest.raises(NotImplementedError):
        mod(a[:4], a, a, attn_mask=torch.zeros(4, 20, device=DEV))
    with pytest.raises(NotImplementedError):
        MultiheadAttention(256, 8).to(DEV)(a[:4], a, a)
