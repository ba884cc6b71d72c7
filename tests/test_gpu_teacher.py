"""The teacher / student callers of the 2-D operator (mirror dml_b200.Modules of models/Modules.py) on the GPU against goldens
generated from the reference's own TeacherNet / StudentNet (tests/golden/teacher_*.npz, student_*.npz): logits, hazards,
features, attention maps and every parameter gradient."""
from types import SimpleNamespace

import pytest
import torch

from dml_b200 import Modules, synth
from oracle.golden_cases import TEACHER_CASES, thin
from oracle.make_goldens import teacher_inputs
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3
TOL_MLP = 3e-3      # position-bias MLP gradients: see tests/test_gpu_deform2d.py


def _build(c):
    args = SimpleNamespace(path_dim=128, label_dim=4, attn_dim=2, path_cluster_num=0.0008)
    mod = (Modules.TeacherNet if c["kind"] == "teacher" else Modules.StudentNet)(args).eval()
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    mod.load_state_dict(synth.fill_like(shapes, c["seed"]), strict=True)
    return mod.to(DEV)


@pytest.mark.parametrize("c", TEACHER_CASES, ids=lambda c: c["name"])
def test_teacher_student_match_reference_goldens(c):
    G = H.golden(c["name"])
    mod = _build(c)
    bag, omic, noise = teacher_inputs(c)
    bag = bag.to(DEV).requires_grad_()
    omic = [o.to(DEV) for o in omic]
    if c["kind"] == "student":
        nz = noise.to(DEV)
        mod.encoder.cluster_merge.noise_fn = lambda B, N, dev: nz
    out = mod(bag, omic)
    logits = out[0]
    H.assert_close(logits.cpu(), G["logits"], TOL, "logits")
    H.assert_close(out[1].cpu(), G["hazards"], TOL, "hazards")
    H.assert_close(out[3].cpu(), G["risk"], TOL, "risk")
    if c["kind"] == "teacher":
        atts = out[6:8]
        H.assert_close(out[4].cpu(), G["feature1"], TOL, "feature1")
        H.assert_close(out[5].cpu(), G["feature2"], TOL, "feature2")
        H.assert_close(thin(out[6].cpu()), G["att1"], TOL, "att1")
        H.assert_close(thin(out[7].cpu()), G["att2"], TOL, "att2")
    else:
        atts = out[5:6]
        H.assert_close(out[4].cpu(), G["feature"], TOL, "feature")
        H.assert_close(thin(out[5].cpu()), G["att"], TOL, "att")
    loss = (logits * synth.normal(tuple(logits.shape), c["seed"], "r_log").to(DEV)).sum()
    for i, a in enumerate(atts):
        loss = loss + (a * synth.normal(tuple(a.shape), c["seed"], f"r_att{i}").to(DEV)).sum() * 0.01
    loss.backward()
    H.assert_close(thin(bag.grad[0].cpu()), G["gbag"], TOL, "d bag")
    seen = 0
    for k, p in mod.named_parameters():
        key = "grad." + k
        if p.grad is None:
            assert key not in G, k
            continue
        tol = TOL_MLP if "rel_pos_bias.mlp" in k else TOL
        # mathematically-zero gradients (softmax shift invariance; common factor of the cluster weights): rounding noise only
        atol = 1e-3 if k.endswith("rel_pos_bias.mlp.2.bias") else 1e-5 if k.endswith("cluster_merge.score.bias") else 0.0
        H.assert_close(thin(p.grad.cpu()), G[key], tol, key, atol=atol)
        seen += 1
    assert seen >= 20
