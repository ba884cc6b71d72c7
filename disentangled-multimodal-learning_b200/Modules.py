"""Mirror of the teacher / student callers of the 2-D operator (reference models/Modules.py:66-458) so that the GPU box, which
has no /root/reference, can build them: same class names, constructor arguments, forward signatures, return arity and
``state_dict`` keys.  Where /root/reference exists the UNMODIFIED models/Modules.py runs on the shadowed operators instead
(tests/test_integration_reference.py).

What runs on this repository's kernels: ``_fc1`` (pair GEMM with bias + ReLU epilogue), every LayerNorm, both
``DeformCrossAttention2D`` and ``ClusterMergeNet``.  The single-key ``nn.MultiheadAttention`` of ``TransFusionLayer``
(one key / value token per bag: the softmax is identically 1), the poolers and the classifier are the reference's own tiny
torch modules - callers, kept as they are (SURVEY.md 8a "caller, kept as-is").
"""
import torch
from torch import nn

from . import ops
from .ClusterMergeNet import ClusterMergeNet
from .DeformableAttention2D import DeformCrossAttention2D


class Pooler(nn.Module):
    """models/Modules.py:460-498: mean over the tokens -> Linear -> tanh."""

    def __init__(self, hidden_size):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(torch.mean(hidden_states, dim=1)))


class FusionNet(nn.Module):
    """models/Modules.py:66-76: Linear over the concatenation of two token streams (one pair GEMM over [x1 | x2])."""

    def __init__(self, feature_dim=128):
        super().__init__()
        self.fusion_layer = nn.Linear(feature_dim * 2, feature_dim)

    def forward(self, feature1, feature2):
        return ops.linear_pg(torch.cat((feature1, feature2), dim=-1), self.fusion_layer.weight, self.fusion_layer.bias)


class TransFusionLayer(nn.Module):
    """models/Modules.py:78-99."""

    def __init__(self, norm_layer=nn.LayerNorm, dim=128):
        super().__init__()
        self.norm = norm_layer(dim)
        self.multihead_attn = nn.MultiheadAttention(embed_dim=128, num_heads=8, dropout=0.1)
        self.pooler = Pooler(dim)

    def forward(self, x1, x2):
        x, w = self.multihead_attn(ops.layer_norm(x1, self.norm), ops.layer_norm(x2, self.norm), ops.layer_norm(x2, self.norm),
                                   attn_mask=None)
        x = x1 + x
        x = self.pooler(ops.layer_norm(x.transpose(0, 1), self.norm))
        return x.unsqueeze(dim=1), w


def _attn2d():
    return DeformCrossAttention2D(dim=128, dim_head=64, heads=8, dropout=0.1, downsample_factor=4, offset_scale=4, offset_groups=8,
                                  offset_kernel_size=6)


class TeacherEncoder(nn.Module):
    """models/Modules.py:171-240 (UniTeacherEncoder :100-169 is the same module fed with [path, path])."""

    def __init__(self, args, norm_layer=nn.LayerNorm, dim=128):
        super().__init__()
        self.norm = norm_layer(dim)
        self.args = args
        self.attn2d_omic1 = _attn2d()
        self.attn2d_omic2 = _attn2d()
        self.fusion_layer = FusionNet(feature_dim=128)
        self.transfusion_layer1 = TransFusionLayer(dim=128)
        self.transfusion_layer2 = TransFusionLayer(dim=128)

    def forward(self, x1, x2, attn_dim, return_vgrid=False):
        n2 = ops.layer_norm(x2, self.norm).transpose(1, 2)
        x_omic1, attn_omic1 = self.attn2d_omic1(ops.layer_norm(x1[0], self.norm).transpose(1, 2), n2, return_vgrid=False)
        x_omic2, attn_omic2 = self.attn2d_omic2(ops.layer_norm(x1[1], self.norm).transpose(1, 2), n2, return_vgrid=False)
        x_out1 = x1[0] + x_omic1.transpose(1, 2)
        x_out2 = x1[1] + x_omic2.transpose(1, 2)
        x = self.fusion_layer(x_out1, x_out2)
        query = ops.layer_norm(x, self.norm).transpose(0, 1)                                     # [L, B, D]
        kv1 = ops.layer_norm(x1[0][:, 0, :].unsqueeze(dim=1), self.norm).transpose(0, 1)
        kv2 = ops.layer_norm(x1[1][:, 0, :].unsqueeze(dim=1), self.norm).transpose(0, 1)
        x_fusion1, _ = self.transfusion_layer1(query, kv1)
        x_fusion2, _ = self.transfusion_layer2(query, kv2)
        return x_fusion1, x_fusion2, attn_omic1, attn_omic2


UniTeacherEncoder = TeacherEncoder


class StudentEncoder(nn.Module):
    """models/Modules.py:242-309."""

    def __init__(self, args, norm_layer=nn.LayerNorm, dim=128):
        super().__init__()
        self.norm = norm_layer(dim)
        self.args = args
        self.attn2d = _attn2d()
        self.cluster_merge = ClusterMergeNet(sample_ratio=self.args.path_cluster_num, dim_out=dim)

    def forward(self, x1, x2, attn_dim, return_vgrid=False):
        x, attn_path = self.attn2d(ops.layer_norm(x1, self.norm).transpose(1, 2), ops.layer_norm(x2, self.norm).transpose(1, 2),
                                   return_vgrid=False)
        x = x1 + x.transpose(1, 2)
        B, N, _ = x.shape
        token_dict = {'x': x, 'token_num': N, 'idx_token': torch.arange(N, device=x.device)[None, :].repeat(B, 1),
                      'agg_weight': x.new_ones(B, N, 1)}
        token_dict, _ = self.cluster_merge(token_dict)
        return token_dict['x'], attn_path


def _survival_head(logits):
    hazards = torch.sigmoid(logits)
    S = torch.cumprod(1 - hazards, dim=1)
    return hazards, S, -torch.sum(S, dim=1)


class TeacherNet(nn.Module):
    """models/Modules.py:357-397: the omic vectors are broadcast over the patches and act as the query stream."""
    uni = False

    def __init__(self, args):
        super().__init__()
        self._fc1 = nn.Sequential(nn.Linear(1024, args.path_dim), nn.ReLU())
        self.args = args
        self.encoder = TeacherEncoder(args=self.args, dim=args.path_dim)
        self.norm = nn.LayerNorm(args.path_dim)
        self.pooler1 = Pooler(args.path_dim)
        self.pooler2 = Pooler(args.path_dim)
        self.classifier = nn.Linear(args.path_dim * 2, args.label_dim)

    def forward(self, path, omic_list=None):
        path = ops.linear_pg(path if path.dtype == torch.bfloat16 else path.float(), self._fc1[0].weight, self._fc1[0].bias, relu=True)
        if self.uni:
            streams = [path, path]
        else:
            streams = [o.float().unsqueeze(1).repeat(1, path.shape[1], 1) for o in omic_list[:2]]
        feature1, feature2, att_omic1, att_omic2 = self.encoder(streams, path, self.args.attn_dim)
        feature1 = self.pooler1(ops.layer_norm(feature1, self.norm))
        feature2 = self.pooler2(ops.layer_norm(feature2, self.norm))
        logits = self.classifier(torch.cat((feature1, feature2), dim=-1))
        hazards, S, risk = _survival_head(logits)
        return logits, hazards, S, risk, feature1, feature2, att_omic1, att_omic2


class UniTeacherNet(TeacherNet):
    """models/Modules.py:312-354: the teacher with the patch stream on both query sides."""
    uni = True


class StudentNet(nn.Module):
    """models/Modules.py:429-458: path-only branch, two merged cluster tokens concatenated for the classifier."""

    def __init__(self, args):
        super().__init__()
        self._fc1 = nn.Sequential(nn.Linear(1024, args.path_dim), nn.ReLU())
        self.args = args
        self.encoder = StudentEncoder(args=self.args, dim=args.path_dim)
        self.norm = nn.LayerNorm(args.path_dim)
        self.pooler1 = Pooler(args.path_dim)
        self.classifier = nn.Linear(args.path_dim * 2, args.label_dim)

    def forward(self, path, omic_list=None):
        path = ops.linear_pg(path if path.dtype == torch.bfloat16 else path.float(), self._fc1[0].weight, self._fc1[0].bias, relu=True)
        feature, att = self.encoder(path, path, self.args.attn_dim)
        feature = torch.cat((feature[:, 0, :], feature[:, 1, :]), dim=-1)
        logits = self.classifier(feature)
        hazards, S, risk = _survival_head(logits)
        return logits, hazards, S, risk, feature, att
