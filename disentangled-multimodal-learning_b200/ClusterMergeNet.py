"""Drop-in mirror of the reference ``ClusterMergeNet`` (models/ClusterMergeNet.py:68-207): DPC-KNN token
clustering and the weighted merge on the kernels of csrc/cluster.cu (the N x N distance matrix is never materialised).
Same constructor, ``forward(token_dict) -> (down_dict, token_dict)`` contract and ``state_dict`` keys (norm.*, score.*)."""
import math

import torch
from torch import nn

from . import ops, ops2d


def cluster_dpc_knn(token_dict, cluster_num, k=5, token_mask=None, noise=None):
    """models/ClusterMergeNet.py:68-128.  ``noise`` replaces the ``torch.rand`` of :103 (tests pass a seeded tensor)."""
    if k != 5 or token_mask is not None:
        raise NotImplementedError("cluster_dpc_knn kernels: k = 5 and no token mask (the reference never passes another)")
    x = token_dict['x']
    if noise is None:
        noise = torch.rand(x.shape[:2], device=x.device, dtype=torch.float32)
    idx_cluster, _ = ops2d.dpc_knn(x, cluster_num, noise)
    return idx_cluster, cluster_num


def merge_tokens(token_dict, idx_cluster, cluster_num, token_weight=None):
    """models/ClusterMergeNet.py:133-179."""
    x = token_dict['x']
    idx_token = token_dict['idx_token']
    agg_weight = token_dict['agg_weight']
    B, N, C = x.shape
    if token_weight is None:
        token_weight = x.new_ones(B, N, 1)
    merged, all_w = ops2d.MergeTokensFn.apply(x, token_weight.reshape(B, N), idx_cluster, cluster_num)
    norm_weight = token_weight / torch.gather(all_w, 1, idx_cluster)[..., None]                 # :160
    idx_token_new = torch.gather(idx_cluster, 1, idx_token)                                     # :168
    weight_t = torch.gather(norm_weight, 1, idx_token[..., None])                               # :169
    return {'x': merged, 'token_num': cluster_num, 'idx_token': idx_token_new, 'agg_weight': agg_weight * weight_t}


class ClusterMergeNet(nn.Module):
    def __init__(self, sample_ratio, dim_out):
        super().__init__()
        self.sample_ratio = sample_ratio
        self.dim_out = dim_out
        self.norm = nn.LayerNorm(self.dim_out)
        self.score = nn.Linear(self.dim_out, 1)
        self.noise_fn = None          # tests: callable (B, N, device) -> the tie-breaking noise of ClusterMergeNet.py:103

    def forward(self, token_dict):
        token_dict = token_dict.copy()
        x = ops.layer_norm(token_dict['x'], self.norm)                                          # :193
        token_score = torch.nn.functional.linear(x, self.score.weight, self.score.bias)        # :194 ([B, N, 1]: a matvec)
        token_weight = token_score.exp()
        token_dict['x'] = x
        B, N, C = x.shape
        token_dict['token_score'] = token_score
        cluster_num = max(math.ceil(N * self.sample_ratio), 1)
        noise = self.noise_fn(B, N, x.device) if self.noise_fn is not None else None
        idx_cluster, cluster_num = cluster_dpc_knn(token_dict, cluster_num, k=5, noise=noise)
        down_dict = merge_tokens(token_dict, idx_cluster, cluster_num, token_weight)
        return down_dict, token_dict
