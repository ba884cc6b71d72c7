"""Shared test helpers: golden loading, seeded inputs, error norms."""
import os

import numpy as np
import torch

from dml_b200 import synth
from oracle.golden_cases import thin

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name + ".npz")).items()}


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    """max|a-b| / max|b|  (the norm SURVEY.md H4 fixes for parity)."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, tol, what="", atol=0.0, max_factor=1.0):
    """rel-L2 <= tol and max-abs/max-abs <= max_factor * tol, or every |a-b| <= atol (for gradients
    that are mathematically zero, e.g. the CPB output bias under softmax shift invariance)."""
    a = thin(a) if a.shape != b.shape else a
    assert a.shape == b.shape, (what, a.shape, b.shape)
    a, b = a.detach(), b.detach()
    if atol > 0 and float((a.double() - b.double()).abs().max()) <= atol:
        return
    e1, e2 = rel_l2(a, b), max_rel(a, b)
    assert e1 <= tol and e2 <= tol * max_factor, f"{what}: rel_l2={e1:.3e} max_rel={e2:.3e} tol={tol:.1e}"


def deform_shapes(dim=128, heads=8, dim_head=64, groups=4, ks=6):
    C = heads * dim_head
    return {
        "to_offsets.0.weight": (C // groups, 1, ks), "to_offsets.0.bias": (C // groups,),
        "to_offsets.2.weight": (1, C // groups, 1),
        "rel_pos_bias.mlp.0.0.weight": (dim // 4, 1), "rel_pos_bias.mlp.0.0.bias": (dim // 4,),
        "rel_pos_bias.mlp.1.0.weight": (dim // 4, dim // 4), "rel_pos_bias.mlp.1.0.bias": (dim // 4,),
        "rel_pos_bias.mlp.2.weight": (heads // groups, dim // 4), "rel_pos_bias.mlp.2.bias": (heads // groups,),
        "to_q.weight": (C, dim, 1), "to_k.weight": (C, dim, 1), "to_v.weight": (C, dim, 1),
        "to_out.weight": (dim, C, 1), "to_out.bias": (dim,),
    }


def nystrom_shapes(dim, dim_head, heads=8, ks=33):
    inner = heads * dim_head
    return {"to_qkv.weight": (3 * inner, dim), "to_out.0.weight": (dim, inner), "to_out.0.bias": (dim,),
            "res_conv.weight": (heads, 1, ks, 1)}


def attn2d_shapes(p):
    """Unused-but-present DeformCrossAttention2D parameters."""
    return {
        p + "to_offsets.0.weight": (64, 1, 6, 6), p + "to_offsets.0.bias": (64,), p + "to_offsets.2.weight": (2, 64, 1, 1),
        p + "rel_pos_bias.mlp.0.0.weight": (32, 2), p + "rel_pos_bias.mlp.0.0.bias": (32,),
        p + "rel_pos_bias.mlp.1.0.weight": (32, 32), p + "rel_pos_bias.mlp.1.0.bias": (32,),
        p + "rel_pos_bias.mlp.2.weight": (1, 32), p + "rel_pos_bias.mlp.2.bias": (1,),
        p + "to_q.weight": (512, 16, 1, 1), p + "to_k.weight": (512, 16, 1, 1), p + "to_v.weight": (512, 16, 1, 1),
        p + "to_out.weight": (128, 512, 1, 1), p + "to_out.bias": (128,),
    }


def dctmil_shapes(n_classes=4, prefix=""):
    s = {"cls_token": (1, 1, 128), "fusion_layer.fusion_layer.weight": (128, 256), "fusion_layer.fusion_layer.bias": (128,),
         "_fc1.0.weight": (128, 1024), "_fc1.0.bias": (128,), "layer3.norm.weight": (128,), "layer3.norm.bias": (128,),
         "norm.weight": (128,), "norm.bias": (128,), "_fc2.weight": (n_classes, 128), "_fc2.bias": (n_classes,),
         "pooler.dense.weight": (128, 128), "pooler.dense.bias": (128,),
         "multimodal_projection.weight": (128, 128), "multimodal_projection.bias": (128,)}
    s.update(attn2d_shapes("layer3.attn2d."))
    s.update({"layer3.attn1d." + k: v for k, v in deform_shapes().items()})
    return {prefix + k: v for k, v in s.items()}


def transmil_shapes(label_dim=3, path_dim=128):
    s = {"cls_token": (1, 1, 512), "_fc1.0.weight": (512, 1024), "_fc1.0.bias": (512,),
         "pos_layer.proj.weight": (512, 1, 7, 7), "pos_layer.proj.bias": (512,),
         "pos_layer.proj1.weight": (512, 1, 5, 5), "pos_layer.proj1.bias": (512,),
         "pos_layer.proj2.weight": (512, 1, 3, 3), "pos_layer.proj2.bias": (512,),
         "norm.weight": (512,), "norm.bias": (512,), "_fc2.weight": (label_dim, 512), "_fc2.bias": (label_dim,),
         "multimodal_projection.weight": (path_dim, 512), "multimodal_projection.bias": (path_dim,)}
    for l in ("layer1", "layer2"):
        s[l + ".norm.weight"] = (512,)
        s[l + ".norm.bias"] = (512,)
        s.update({l + ".attn." + k: v for k, v in nystrom_shapes(512, 64).items()})
    return s


def maxnet_shapes(input_dim, label_dim=4, prefix=""):
    hid = [input_dim, 64, 48, 32, 128]
    s = {"output_range": (1,), "output_shift": (1,), "classifier.0.weight": (label_dim, 128), "classifier.0.bias": (label_dim,)}
    for i in range(4):
        s[f"encoder.{i}.0.weight"] = (hid[i + 1], hid[i])
        s[f"encoder.{i}.0.bias"] = (hid[i + 1],)
    return {prefix + k: v for k, v in s.items()}


def pathomic_shapes(label_dim=4):
    s = {"output_range": (1,), "output_shift": (1,), "classifier.weight": (label_dim, 256), "classifier.bias": (label_dim,),
         "classifier_tumor.0.weight": (label_dim, 128), "classifier_tumor.0.bias": (label_dim,),
         "classifier_immune.0.weight": (label_dim, 128), "classifier_immune.0.bias": (label_dim,)}
    s.update(maxnet_shapes(59, label_dim, "omic_net_tumor."))
    s.update(maxnet_shapes(361, label_dim, "omic_net_immune."))
    s.update(dctmil_shapes(label_dim, "pathomic_net_tumor."))
    s.update(dctmil_shapes(label_dim, "pathomic_net_immune."))
    return s


def mha_shapes(E=256):
    """state_dict of models/MultiheadAttention.py:MultiheadAttention(embed_dim=E) (packed in-projection)."""
    return {"in_proj_weight": (3 * E, E), "in_proj_bias": (3 * E,), "out_proj.weight": (E, E), "out_proj.bias": (E,)}


def leafify(P):
    return {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in P.items()}


def cluster_shapes(C=128):
    """state_dict of models/ClusterMergeNet.py:ClusterMergeNet(dim_out=C)."""
    return {"norm.weight": (C,), "norm.bias": (C,), "score.weight": (1, C), "score.bias": (1,)}
