"""Oracle restatement of the callers of the attention ops (TEST INFRASTRUCTURE ONLY).

DeformCrossTransMIL   /root/reference/models/DeformCrossTransMIL.py:28-161
TransLayer/PPEG/TransMIL  /root/reference/models/mil.py:171-259
MaxNet / DeformPathomicNet  /root/reference/models/model.py:173-218, 471-568
Losses used by the bench:  train_test.py:791-853 (weighted CE), utils/utils.py:245-261 (NLL hazard)

Functional style: ``P`` is a flat dict keyed like the reference ``state_dict``;
``sub(P, prefix)`` selects a sub-module.  Everything runs in eval mode (dropout =
identity).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .deform1d import deform_cross_attention_1d
from .nystrom import nystrom_attention

Params = Dict[str, torch.Tensor]


def sub(P: Params, prefix: str) -> Params:
    prefix = prefix + "."
    return {k[len(prefix):]: v for k, v in P.items() if k.startswith(prefix)}


def layer_norm(x, P, name):
    return F.layer_norm(x, (x.shape[-1],), P[name + ".weight"], P[name + ".bias"], 1e-5)


def deform_cross_trans_mil(path: torch.Tensor, omic: torch.Tensor, P: Params, *,
                           row_block: Optional[int] = None, return_hidden: bool = False):
    """DeformCrossTransMIL.forward with args.attn_dim == 1, return_vgrid False
    (DeformCrossTransMIL.py:97-161).  path [B,N,1024], omic [B,128]."""
    path = F.relu(F.linear(path.float(), P["_fc1.0.weight"], P["_fc1.0.bias"]))            # :100
    omic_rep = omic.float()[:, None, :].expand(-1, path.shape[1], -1)                      # :105
    h = F.linear(torch.cat((path, omic_rep), dim=-1),                                      # :35-37,111 cat(path, omic)
                 P["fusion_layer.fusion_layer.weight"], P["fusion_layer.fusion_layer.bias"])
    B = h.shape[0]
    cls = P["cls_token"].expand(B, -1, -1)
    h = torch.cat((cls, h), dim=1)                                                         # :119
    path = torch.cat((cls, path), dim=1)                                                   # :122
    # DeformCrossTransLayer (:62-68): ONE LayerNorm shared by both streams (Q4)
    x1 = layer_norm(h, P, "layer3.norm").transpose(1, 2)
    x2 = layer_norm(path, P, "layer3.norm").transpose(1, 2)
    a = deform_cross_attention_1d(x1, x2, sub(P, "layer3.attn1d"), heads=8, dim_head=64, offset_groups=4,
                                  downsample_factor=4, offset_scale=2, row_block=row_block)
    h = h + a.transpose(1, 2)
    hidden = h
    h = layer_norm(h, P, "norm")[:, 0]                                                     # :128
    logits = F.linear(h, P["_fc2.weight"], P["_fc2.bias"])                                 # :132
    encoded = F.linear(h, P["multimodal_projection.weight"], P["multimodal_projection.bias"])  # :151
    if return_hidden:
        return encoded, logits, hidden
    return encoded, logits


def trans_layer(x, P, dim=512):
    """TransLayer (mil.py:171-189): x + Nystrom(LN(x)); dim_head=dim/8, m=dim/2."""
    y = nystrom_attention(layer_norm(x, P, "norm"), sub(P, "attn"), heads=8, dim_head=dim // 8,
                          num_landmarks=dim // 2, pinv_iterations=6, residual=True)
    return x + y


def ppeg(x, P, H, W):
    """PPEG (mil.py:192-206): 7x7 + 5x5 + 3x3 depthwise convs + identity on the square grid."""
    B, _, C = x.shape
    cls, feat = x[:, :1], x[:, 1:]
    f = feat.transpose(1, 2).reshape(B, C, H, W)
    y = (F.conv2d(f, P["proj.weight"], P["proj.bias"], padding=3, groups=C) + f
         + F.conv2d(f, P["proj1.weight"], P["proj1.bias"], padding=2, groups=C)
         + F.conv2d(f, P["proj2.weight"], P["proj2.bias"], padding=1, groups=C))
    return torch.cat((cls, y.flatten(2).transpose(1, 2)), dim=1)


def square_side(N: int) -> int:
    """mil.py:233 - ceil(sqrt(N)) evaluated the way numpy does (float64)."""
    return int(math.ceil(math.sqrt(N)))


def trans_mil(x: torch.Tensor, P: Params):
    """TransMIL.forward (mil.py:225-259).  x [B,N,1024] -> (encoded, logits)."""
    h = F.relu(F.linear(x.float(), P["_fc1.0.weight"], P["_fc1.0.bias"]))
    N = h.shape[1]
    side = square_side(N)
    add = side * side - N
    h = torch.cat((h, h[:, :add]), dim=1)                                                  # :235 wrap-pad (Q12)
    h = torch.cat((P["cls_token"].expand(h.shape[0], -1, -1), h), dim=1)
    h = trans_layer(h, sub(P, "layer1"))
    h = ppeg(h, sub(P, "pos_layer"), side, side)
    h = trans_layer(h, sub(P, "layer2"))
    h = layer_norm(h, P, "norm")[:, 0]
    logits = F.linear(h, P["_fc2.weight"], P["_fc2.bias"])
    encoded = F.linear(h, P["multimodal_projection.weight"], P["multimodal_projection.bias"])
    return encoded, logits


def max_net(x, P):
    """MaxNet (model.py:173-218), eval mode: 4x(Linear+ELU) -> ReLU; classifier head unused here."""
    for i in range(4):
        x = F.elu(F.linear(x, P[f"encoder.{i}.0.weight"], P[f"encoder.{i}.0.bias"]))
    return F.relu(x)


def deform_pathomic_net(x_path, x_omic_tumor, x_omic_immune, P: Params, *, task_type="diag2021",
                        row_block: Optional[int] = None):
    """DeformPathomicNet.forward (model.py:511-568), fusion_type == 'concat'.
    Returns (features, vec_tumor, vec_immune, [hazard_tumor, hazard_immune, hazard])."""
    ot = max_net(x_omic_tumor, sub(P, "omic_net_tumor"))
    vt, _ = deform_cross_trans_mil(x_path, ot, sub(P, "pathomic_net_tumor"), row_block=row_block)
    oi = max_net(x_omic_immune, sub(P, "omic_net_immune"))
    vi, _ = deform_cross_trans_mil(x_path, oi, sub(P, "pathomic_net_immune"), row_block=row_block)
    feats = torch.cat((vt, vi), dim=1)
    hz = F.linear(feats, P["classifier.weight"], P["classifier.bias"])
    ht = F.linear(vt, P["classifier_tumor.0.weight"], P["classifier_tumor.0.bias"])
    hi = F.linear(vi, P["classifier_immune.0.weight"], P["classifier_immune.0.bias"])
    if task_type == "survival":
        hz, ht, hi = torch.sigmoid(hz), torch.sigmoid(ht), torch.sigmoid(hi)               # :555-558
    return feats, vt, vi, [ht, hi, hz]


DIAG2021_CE_WEIGHTS = (1.0, 4.15, 2.93, 2.43)     # train_test.py:790
GRADE_CE_WEIGHTS = (1.47, 1.51, 1.0)              # train_test.py:791


def nll_surv_loss(hazards, Y, c, alpha=0.0, eps=1e-7):
    """utils/utils.py:245-261 with S = cumprod(1 - hazards) (train_test.py:826)."""
    B = len(Y)
    Y = Y.view(B, 1)
    c = c.view(B, 1).float()
    S = torch.cumprod(1 - hazards, dim=1)
    S_pad = torch.cat([torch.ones_like(c), S], 1)
    unc = -(1 - c) * (torch.log(torch.gather(S_pad, 1, Y).clamp(min=eps))
                      + torch.log(torch.gather(hazards, 1, Y).clamp(min=eps)))
    cen = -c * torch.log(torch.gather(S_pad, 1, Y + 1).clamp(min=eps))
    return ((1 - alpha) * (cen + unc) + alpha * unc).mean()


def bag_loss(logits, label, task_type="diag2021", censor=None):
    """Loss on the fused head only (train_test.py:833-853: loss = loss3 on logits[2])."""
    hz = logits[2]
    if task_type == "diag2021":
        return F.cross_entropy(hz, label, weight=torch.tensor(DIAG2021_CE_WEIGHTS, device=hz.device))
    if task_type == "grade":
        return F.cross_entropy(hz, label, weight=torch.tensor(GRADE_CE_WEIGHTS, device=hz.device))
    if task_type == "survival":
        return nll_surv_loss(hz, label, censor, alpha=0.0)
    raise ValueError(task_type)
