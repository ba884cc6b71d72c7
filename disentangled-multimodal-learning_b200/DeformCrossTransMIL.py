"""Drop-in mirror of the reference ``models/DeformCrossTransMIL.py`` (FusionNet, DeformCrossTransLayer,
DeformCrossTransMIL, Pooler) with the same constructor/forward signatures and state_dict keys
(SURVEY.md section 8b / appendix A).  attn_dim == 1 is the supported (and the only working) branch.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .DeformableAttention1D import DeformCrossAttention1D
from .DeformableAttention2D import DeformCrossAttention2D


class FusionNet(nn.Module):
    """Linear(2*feature_dim -> feature_dim) on cat(gene_features, image_features)
    (reference :28-38).  Evaluated as two half-GEMMs so that the [B, N, 2*dim] concat (and the
    [B, N, dim] repeat of the per-bag omic vector, :105) are never materialised."""

    def __init__(self, feature_dim=128):
        super().__init__()
        self.feature_dim = feature_dim
        self.fusion_layer = nn.Linear(feature_dim * 2, feature_dim)

    def forward(self, gene_features, image_features):
        W, b = self.fusion_layer.weight, self.fusion_layer.bias
        fd = self.feature_dim
        if image_features.dim() == gene_features.dim() - 1:
            # per-bag vector [B, dim], broadcast over the tokens: the epilogue bias of the path-half GEMM
            vec = F.linear(image_features, W[:, fd:]) + b
            return ops.FusionFn.apply(gene_features.float(), W[:, :fd], vec)
        return ops.linear_pg(torch.cat((gene_features, image_features), dim=-1), W, b)      # token-level second input


class DeformCrossTransLayer(nn.Module):
    def __init__(self, norm_layer=nn.LayerNorm, dim=128):
        super().__init__()
        self.norm = norm_layer(dim)
        self.attn2d = DeformCrossAttention2D(dim=128, dim_head=64, heads=8, dropout=0.1, downsample_factor=4,
                                             offset_scale=4, offset_groups=8, offset_kernel_size=6)
        self.attn1d = DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6)

    def forward(self, x1, x2, attn_dim, return_vgrid, rows=None):
        if attn_dim == 1:
            # one LayerNorm shared by both streams (reference :44,66) and the residual (:67), both inside the fused function
            return self.attn1d(x1.transpose(1, 2), x2.transpose(1, 2), rows=rows, _norm=self.norm).transpose(1, 2)
        raise NotImplementedError("attn_dim == 2 is broken in the reference as shipped (SURVEY.md Q6) and "
                                  "DeformCrossAttention2D is not built yet (row N1)")


class Pooler(nn.Module):
    def __init__(self, hidden_size):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(torch.mean(hidden_states, dim=1)))


class DeformCrossTransMIL(nn.Module):
    def __init__(self, args, n_classes=4):
        super().__init__()
        self.fusion_layer = FusionNet(feature_dim=128)
        self._fc1 = nn.Sequential(nn.Linear(1024, args.path_dim), nn.ReLU())
        self.cls_token = nn.Parameter(torch.randn(1, 1, args.path_dim))
        self.args = args
        self.n_classes = n_classes
        self.layer3 = DeformCrossTransLayer(dim=args.path_dim)
        self.norm = nn.LayerNorm(args.path_dim)
        self._fc2 = nn.Linear(args.path_dim, self.n_classes)
        self.pooler = Pooler(args.path_dim)
        self.multimodal_projection = nn.Linear(args.path_dim, self.args.path_dim)

    def forward(self, path, omic):
        if getattr(self.args, "attn_dim", 1) != 1:
            raise NotImplementedError("only attn_dim == 1 is supported (SURVEY.md Q6)")
        if getattr(self.args, "return_vgrid", False):
            raise NotImplementedError("return_vgrid with attn_dim == 1 raises in the reference (SURVEY.md Q6)")
        if path.is_cuda:      # the bias table depends on the CPB weights only: start it before fc1, off the critical chain
            self.layer3.attn1d.prefetch_bias_table(path.shape[1] + 1, path.device)
        fc1 = self._fc1[0]
        # fc1 + ReLU (:100) on the pair GEMM, bias and ReLU in its epilogue; a bf16 bag enters as it is (one exact plane)
        path = ops.linear_pg(path, fc1.weight, fc1.bias, relu=True)
        ready = getattr(omic, "_dml_ready", None)      # omic vector produced on another stream (model._omic_ahead)
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready)
            omic.record_stream(torch.cuda.current_stream())
        h = self.fusion_layer(path, omic.float())
        B = h.shape[0]
        cls_tokens = self.cls_token.expand(B, -1, -1).to(h.device)
        h = torch.cat((cls_tokens, h), dim=1)
        path = torch.cat((cls_tokens, path), dim=1)
        # Only the cls row of the layer output is consumed (reference :128).  args.cls_row_only = True
        # (an extension, off by default) asks the attention for that row alone: logits and every gradient are unchanged
        # (the offsets, keys and values still see every token), the n x n_kv attention work drops to 1 x n_kv.
        rows = 1 if getattr(self.args, "cls_row_only", False) else None
        h = self.layer3(h, path, 1, False, rows=rows)
        # norm(h)[:, 0] -> _fc2, multimodal_projection (:128-151): LayerNorm is per token, so only the cls row is normalised;
        # the row's norm and both heads are one kernel per direction
        if h.is_cuda and self.norm.elementwise_affine:
            encoded, logits = ops.TowerHeadFn.apply(h, self.norm.weight, self.norm.bias, self._fc2.weight, self._fc2.bias,
                                                    self.multimodal_projection.weight, self.multimodal_projection.bias,
                                                    self.norm.eps)
        else:
            h = self.norm(h[:, 0])
            logits = self._fc2(h)
            encoded = self.multimodal_projection(h)
        return encoded, logits, None
