"""Parameter-compatible placeholder of the reference ``DeformCrossAttention2D``
(models/DeformableAttention2D.py:162-342).

``DeformCrossTransLayer`` always constructs this module and its parameters are part of every
``DeformPathomicNet`` checkpoint, so the keys and shapes must exist
even though the ``attn_dim == 1`` hot path never runs it (and the reference's own ``attn_dim == 2``
branch crashes as shipped, quirk Q6).  The 2-D operator itself is SURVEY.md section 8(f) row N1
("next") and is not built yet: ``forward`` raises.
"""
from torch import nn


def default(val, d):
    return val if val is not None else d


class CPB2D(nn.Module):
    def __init__(self, dim, *, heads, offset_groups, depth):
        super().__init__()
        self.heads = heads
        self.offset_groups = offset_groups
        self.mlp = nn.ModuleList([])
        self.mlp.append(nn.Sequential(nn.Linear(2, dim), nn.ReLU()))
        for _ in range(depth - 1):
            self.mlp.append(nn.Sequential(nn.Linear(dim, dim), nn.ReLU()))
        self.mlp.append(nn.Linear(dim, heads // offset_groups))


class DeformCrossAttention2D(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8, dropout=0., downsample_factor=4, offset_scale=4,
                 offset_groups=8, offset_kernel_size=6, group_queries=True, group_key_values=True):
        super().__init__()
        offset_scale = default(offset_scale, downsample_factor)
        assert offset_kernel_size >= downsample_factor
        assert (offset_kernel_size - downsample_factor) % 2 == 0
        offset_groups = default(offset_groups, heads)
        assert heads % offset_groups == 0
        inner_dim = dim_head * heads
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.offset_groups = offset_groups
        offset_dims = inner_dim // offset_groups
        self.downsample_factor = downsample_factor
        self.to_offsets = nn.Sequential(
            nn.Conv2d(offset_dims, offset_dims, offset_kernel_size, groups=offset_dims, stride=downsample_factor,
                      padding=(offset_kernel_size - downsample_factor) // 2),
            nn.GELU(),
            nn.Conv2d(offset_dims, 2, 1, bias=False),
            nn.Tanh(),
            nn.Identity(),
        )
        self.rel_pos_bias = CPB2D(dim // 4, offset_groups=offset_groups, heads=heads, depth=2)
        self.dropout = nn.Dropout(dropout)
        self.to_q = nn.Conv2d(dim, inner_dim, 1, groups=offset_groups if group_queries else 1, bias=False)
        self.to_k = nn.Conv2d(dim, inner_dim, 1, groups=offset_groups if group_key_values else 1, bias=False)
        self.to_v = nn.Conv2d(dim, inner_dim, 1, groups=offset_groups if group_key_values else 1, bias=False)
        self.to_out = nn.Conv2d(inner_dim, dim, 1)

    def forward(self, x1, x2, return_vgrid=False):
        raise NotImplementedError("DeformCrossAttention2D (SURVEY.md 8(f) N1) has no sm_100a kernel yet; "
                                  "use attn_dim == 1")
