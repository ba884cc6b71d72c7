"""Drop-in mirror of the Nystrom users in the reference ``models/mil.py``: TransLayer (:171-189),
PPEG (:192-206), TransMIL (:209-259).  Same names, signatures and state_dict keys."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .NystromAttention import NystromAttention
from .nystrom_fn import PPEGFn


class TransLayer(nn.Module):
    def __init__(self, norm_layer=nn.LayerNorm, dim=512):
        super().__init__()
        self.norm = norm_layer(dim)
        self.attn = NystromAttention(dim=dim, dim_head=dim // 8, heads=8, num_landmarks=dim // 2, pinv_iterations=6,
                                     residual=True, dropout=0.1)

    def forward(self, x):
        return x + self.attn(x, _norm=self.norm)      # LayerNorm fused into the attention function (mil.py:186)


class PPEG(nn.Module):
    def __init__(self, dim=512):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, 7, 1, 7 // 2, groups=dim)
        self.proj1 = nn.Conv2d(dim, dim, 5, 1, 5 // 2, groups=dim)
        self.proj2 = nn.Conv2d(dim, dim, 3, 1, 3 // 2, groups=dim)

    def forward(self, x, H, W):
        """x [B, 1 + H W, C]: cls token + the H x W grid (H == W: TransMIL wrap-pads the bag to a square).  The three
        depthwise convolutions and the identity are one 7x7 stencil kernel (weights summed here: exact up to fp32
        reassociation); autograd splits the stencil's weight gradient back onto proj / proj1 / proj2."""
        assert H == W, "PPEG runs on the square grid TransMIL builds (mil.py:232-235)"
        C = x.shape[-1]
        wsum = (self.proj.weight.reshape(C, 7, 7) + F.pad(self.proj1.weight.reshape(C, 5, 5), (1, 1, 1, 1))
                + F.pad(self.proj2.weight.reshape(C, 3, 3), (2, 2, 2, 2))).reshape(C, 49)
        bsum = self.proj.bias + self.proj1.bias + self.proj2.bias
        return PPEGFn.apply(x, wsum, bsum, H)


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
        h = ops.linear_pg(x, fc1.weight, fc1.bias, relu=True)            # mil.py:229 (bf16 bags enter as they are)
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
