"""Oracle restatement of DeformCrossAttention2D and ClusterMergeNet (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/models/DeformableAttention2D.py:89-342 and /root/reference/models/ClusterMergeNet.py:68-207;
every function cites the lines it restates.  Parameters are a dict keyed like the reference ``state_dict``:

    to_offsets.0.{weight [C/G,1,ks,ks], bias [C/G]}   to_offsets.2.weight [2,C/G,1,1]
    rel_pos_bias.mlp.0.0.{weight [hid,2], bias [hid]}  .1.0.{weight [hid,hid], bias [hid]}  .2.{weight [H/G,hid], bias [H/G]}
    to_q.weight [C, dim/G, 1, 1] (grouped)  to_k / to_v.weight [C, dim/G, 1, 1]  to_out.{weight [dim,C,1,1], bias [dim]}

Differentiable torch code: gradients come from autograd.  Pinned against the reference by tests/golden/deform2d_*.npz and
clustermerge_*.npz (oracle/make_goldens.py).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


def kv_side(side: int, ksize: int = 6, stride: int = 4) -> int:
    """Output side of the strided depthwise Conv2d (DeformableAttention2D.py:209)."""
    pad = (ksize - stride) // 2
    return (side + 2 * pad - ksize) // stride + 1


def xy_grid(rows: int, cols: int) -> torch.Tensor:
    """create_grid_like (DeformableAttention2D.py:89-99): [2, rows, cols], channel 0 = column index x, channel 1 = row index y."""
    ys, xs = torch.meshgrid(torch.arange(rows, dtype=torch.float32), torch.arange(cols, dtype=torch.float32), indexing="ij")
    return torch.stack((xs, ys), 0)


def normalize_xy(x: torch.Tensor, y: torch.Tensor, rows: int, cols: int):
    """normalize_grid (DeformableAttention2D.py:101-109): the FIRST channel (x) is divided by rows - 1 and the second (y) by
    cols - 1, exactly as the reference does (it names them grid_h / grid_w; identical for the square grids it is used on)."""
    return 2.0 * x / max(rows - 1, 1) - 1.0, 2.0 * y / max(cols - 1, 1) - 1.0


def offsets_net(q_grouped: torch.Tensor, P: Params, stride: int, offset_scale: float) -> torch.Tensor:
    """to_offsets (DeformableAttention2D.py:208-214): depthwise Conv2d(ks, stride, pad) + bias -> GELU(erf) -> Conv2d(C/G -> 2,
    1x1, no bias) -> tanh -> * offset_scale.  [(b g), C/G, h, w] -> [(b g), 2, hk, wk]."""
    w0, b0, w2 = P["to_offsets.0.weight"], P["to_offsets.0.bias"], P["to_offsets.2.weight"]
    ks = w0.shape[-1]
    y = F.conv2d(q_grouped, w0, b0, stride=stride, padding=(ks - stride) // 2, groups=w0.shape[0])
    return torch.tanh(F.conv2d(F.gelu(y), w2)) * offset_scale


def bias_mlp(pos: torch.Tensor, P: Params) -> torch.Tensor:
    """CPB.forward (DeformableAttention2D.py:144-158) on relative positions [..., 2]: signed log, then 2 -> hid -> hid -> H/G."""
    t = torch.sign(pos) * torch.log(pos.abs() + 1)
    t = F.relu(F.linear(t, P["rel_pos_bias.mlp.0.0.weight"], P["rel_pos_bias.mlp.0.0.bias"]))
    t = F.relu(F.linear(t, P["rel_pos_bias.mlp.1.0.weight"], P["rel_pos_bias.mlp.1.0.bias"]))
    return F.linear(t, P["rel_pos_bias.mlp.2.weight"], P["rel_pos_bias.mlp.2.bias"])


def deform_cross_attention_2d(x1: torch.Tensor, x2: torch.Tensor, P: Params, *, heads: int = 8, groups: int = 8,
                              stride: int = 4, offset_scale: float = 4.0, drop_keep: Optional[torch.Tensor] = None,
                              drop_p: float = 0.0, rows: Optional[torch.Tensor] = None):
    """DeformCrossAttention2D.forward (DeformableAttention2D.py:224-342).  x1, x2: [B, dim, n] with n a perfect square.
    Returns (out [B, dim, n], attn [B, heads, n, n_kv], vgrid [(B G), 2, hk, wk]).  ``drop_keep`` (bool [B, heads, n, n_kv]) is the
    dropout keep-mask applied to attn before the aggregation (:316, training mode); None = eval.
    ``rows`` (long [r], oracle-only): evaluate the attention for these query tokens only - out [B, dim, r], attn [B, heads, r,
    n_kv] - so that 100k-token bags (whose full map is 20 GB per activation) can be spot-checked; same maths row for row."""
    B, dim, n = x1.shape
    side = int(math.isqrt(n))
    assert side * side == n, "the reference views the sequence as a square grid (:241-242)"
    G, H = groups, heads
    x1 = x1.reshape(B, dim, side, side)
    x2 = x2.reshape(B, dim, side, side)
    q = F.conv2d(x1, P["to_q.weight"], groups=G)                                            # :248 (grouped 1x1)
    C = q.shape[1]
    grp = lambda t: t.reshape(B * G, t.shape[1] // G, *t.shape[2:])                         # :253
    offsets = offsets_net(grp(q), P, stride, offset_scale)                                  # :257
    hk, wk = offsets.shape[-2:]
    vgrid = xy_grid(hk, wk).to(offsets) + offsets                                           # :263-266
    vx, vy = normalize_xy(vgrid[:, 0], vgrid[:, 1], hk, wk)                                  # :270
    vs = torch.stack((vx, vy), -1)                                                          # [(B G), hk, wk, 2]
    kv = F.grid_sample(grp(x2), vs, mode="bilinear", padding_mode="zeros", align_corners=False)   # :274-277
    kv = kv.reshape(B, dim, hk, wk)                                                         # :280
    k = F.conv2d(kv, P["to_k.weight"], groups=G)
    v = F.conv2d(kv, P["to_v.weight"], groups=G)                                            # :285
    d = C // H
    q = q * d ** -0.5                                                                       # :290
    heads_of = lambda t: t.reshape(B, H, d, -1).transpose(2, 3)                             # :294  [B, H, tokens, d]
    qh, kh, vh = heads_of(q), heads_of(k), heads_of(v)
    g = xy_grid(side, side).to(x1)
    gx, gy = normalize_xy(g[0], g[1], side, side)                                           # :302-303
    gq = torch.stack((gx, gy), -1).reshape(1, n, 1, 2)
    if rows is not None:
        qh, gq, n = qh[:, :, rows], gq[:, rows], int(rows.numel())
    sim = qh @ kh.transpose(2, 3)                                                           # :298
    pos = gq - vs.reshape(B * G, 1, hk * wk, 2)                                             # :150
    bias = bias_mlp(pos, P)                                                                 # [(B G), n, n_kv, H/G]
    bias = bias.reshape(B, G, n, hk * wk, H // G).permute(0, 1, 4, 2, 3).reshape(B, H, n, hk * wk)   # :156
    sim = sim + bias
    sim = sim - sim.amax(dim=-1, keepdim=True).detach()                                     # :309
    attn = sim.softmax(dim=-1)                                                              # :313
    a = attn if drop_keep is None else attn * drop_keep.to(attn) / (1.0 - drop_p)           # :316
    out = a @ vh                                                                            # :320
    out = out.transpose(2, 3).reshape(B, C, n, 1)                                           # :321 (a 1x1 conv follows: the grid shape is immaterial)
    out = F.conv2d(out, P["to_out.weight"], P["to_out.bias"])                               # :322
    return out.reshape(B, dim, n), attn, vgrid


# ---------------------------------------------------------------------------------------------------------------------
# ClusterMergeNet (models/ClusterMergeNet.py)
# ---------------------------------------------------------------------------------------------------------------------
def dpc_knn(x: torch.Tensor, cluster_num: int, noise: torch.Tensor, k: int = 5):
    """cluster_dpc_knn (ClusterMergeNet.py:68-128) without the token mask (no caller passes one).  ``noise`` [B, N] stands for
    the ``torch.rand(...)`` of :103 (x 1e-6 is applied here).  Returns (idx_cluster [B, N] long, index_down [B, cluster_num])."""
    with torch.no_grad():
        B, N, C = x.shape
        dist = torch.cdist(x, x) / (C ** 0.5)                                               # :88
        near, _ = torch.topk(dist, k=k, dim=-1, largest=False)                              # :98
        density = (-(near ** 2).mean(dim=-1)).exp() + noise * 1e-6                          # :100-104
        higher = (density[:, None, :] > density[:, :, None]).to(x.dtype)                    # :111-112
        dist_max = dist.flatten(1).max(dim=-1)[0][:, None, None]                            # :113
        parent_dist, _ = (dist * higher + dist_max * (1 - higher)).min(dim=-1)              # :114
        score = parent_dist * density                                                       # :117
        _, index_down = torch.topk(score, k=cluster_num, dim=-1)                            # :118
        to_centres = torch.gather(dist, 1, index_down[:, :, None].expand(B, cluster_num, N))   # :121 index_points
        idx_cluster = to_centres.argmin(dim=1)                                              # :123
        rows = torch.arange(B)[:, None].expand(B, cluster_num)
        idx_cluster[rows.reshape(-1), index_down.reshape(-1)] = torch.arange(cluster_num).repeat(B)   # :126-128
    return idx_cluster, index_down


def merge_tokens(x: torch.Tensor, idx_cluster: torch.Tensor, cluster_num: int, token_weight: torch.Tensor) -> torch.Tensor:
    """merge_tokens (ClusterMergeNet.py:133-179), the part that feeds the model: weighted mean of the tokens of a cluster.
    x [B, N, C], token_weight [B, N, 1] -> [B, cluster_num, C]."""
    B, N, C = x.shape
    idx = (idx_cluster + torch.arange(B)[:, None] * cluster_num).reshape(B * N)
    all_w = token_weight.new_zeros(B * cluster_num, 1).index_add(0, idx, token_weight.reshape(B * N, 1)) + 1e-6   # :156-159
    norm_w = token_weight / all_w[idx].reshape(B, N, 1)                                     # :160
    merged = x.new_zeros(B * cluster_num, C).index_add(0, idx, (x * norm_w).reshape(B * N, C))   # :163-166
    return merged.reshape(B, cluster_num, C)


def cluster_merge_net(x: torch.Tensor, P: Params, sample_ratio: float, noise: torch.Tensor):
    """ClusterMergeNet.forward (ClusterMergeNet.py:191-207): LayerNorm -> score Linear -> exp weight -> DPC-KNN on the
    normalised tokens -> weighted merge.  Returns (merged [B, cluster_num, C], idx_cluster, normalised x, token_score)."""
    C = x.shape[-1]
    xn = F.layer_norm(x, (C,), P["norm.weight"], P["norm.bias"])
    score = F.linear(xn, P["score.weight"], P["score.bias"])
    weight = score.exp()
    cluster_num = max(math.ceil(x.shape[1] * sample_ratio), 1)
    idx_cluster, _ = dpc_knn(xn, cluster_num, noise)
    return merge_tokens(xn, idx_cluster, cluster_num, weight), idx_cluster, xn, score
