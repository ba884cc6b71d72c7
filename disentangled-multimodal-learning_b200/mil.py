"""Drop-in mirror of the Nystrom users in the reference ``models/mil.py``: TransLayer (:171-189),
PPEG (:192-206), TransMIL (:209-259).  Same names, signatures and state_dict keys."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .NystromAttention import NystromAttention


class TransLayer(nn.Module):
    def __init__(self, norm_layer=nn.LayerNorm, dim=512):
        super().__init__()
        self.norm = norm_layer(dim)
        self.attn = NystromAttention(dim=dim, dim_head=dim // 8, heads=8, num_landmarks=dim // 2, pinv_iterations=6,
                                     residual=True, dropout=0.1)

    def forward(self, x):
        return x + self.attn(ops.layer_norm(x, self.norm))


class PPEG(nn.Module):
    def __init__(self, dim=512):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, 7, 1, 7 // 2, groups=dim)
        self.proj1 = nn.Conv2d(dim, dim, 5, 1, 5 // 2, groups=dim)
        self.proj2 = nn.Conv2d(dim, dim, 3, 1, 3 // 2, groups=dim)

    def forward(self, x, H, W):
        B, _, C = x.shape
        cls_token, feat_token = x[:, 0], x[:, 1:]
        cnn_feat = feat_token.transpose(1, 2).view(B, C, H, W)
        x = self.proj(cnn_feat) + cnn_feat + self.proj1(cnn_feat) + self.proj2(cnn_feat)
        x = x.flatten(2).transpose(1, 2)
        return torch.cat((cls_token.unsqueeze(1), x), dim=1)


class TransMIL(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.pos_layer = PPEG(dim=512)
        self._fc1 = nn.Sequential(nn.Linear(1024, 512), nn.ReLU())
        self.cls_token = nn.Parameter(torch.randn(1, 1, 512))
        self.n_classes = self.args.label_dim
        self.layer1 = TransLayer(dim=512)
        self.layer2 = TransLayer(dim=512)
        self.norm = nn.LayerNorm(512)
        self._fc2 = nn.Linear(512, self.n_classes)
        self.multimodal_projection = nn.Linear(512, self.args.path_dim)

    def forward(self, x):
        fc1 = self._fc1[0]
        h = F.relu(ops.mm_tc(x.float(), fc1.weight.t()) + fc1.bias)
        N = h.shape[1]
        side = int(np.ceil(np.sqrt(N)))
        add_length = side * side - N
        h = torch.cat([h, h[:, :add_length, :]], dim=1)          # wrap-pad to a square (quirk Q12)
        B = h.shape[0]
        cls_tokens = self.cls_token.expand(B, -1, -1).to(h.device)
        h = torch.cat((cls_tokens, h), dim=1)
        h = self.layer1(h)
        h = self.pos_layer(h, side, side)
        h = self.layer2(h)
        h = self.norm(h[:, 0])
        logits = self._fc2(h)
        encoded = self.multimodal_projection(h)
        return encoded, logits, None
