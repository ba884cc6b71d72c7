"""Host-side count (no GPU) of the sweep variants the dK/dV kernel would take on the bench table: segments of the position-bias
MLP from a dense scan, keys at their un-displaced positions, 6 random key blocks x all 32-query tiles."""
import sys, math
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from dml_b200 import synth
from tests import helpers as H
sd = synth.fill_like(H.deform_shapes(), 42)
w1 = sd["rel_pos_bias.mlp.0.0.weight"].double().reshape(-1); b1 = sd["rel_pos_bias.mlp.0.0.bias"].double()
W2 = sd["rel_pos_bias.mlp.1.0.weight"].double(); b2 = sd["rel_pos_bias.mlp.1.0.bias"].double()
n, n_kv = 16385, 4096
T = math.log2(1 + 2.0 + 2.0 * 2 / (n_kv - 1)) * 1.001
xs = torch.linspace(-T, T, 4_000_001, dtype=torch.float64)
# the reference feeds sign(p) * log(|p| + 1) (natural log); the kernel works in log2 units: x_nat = x * ln2
t = xs * math.log(2.0)
h1 = t[:, None] * w1[None] + b1[None]
a1 = h1 > 0
h2 = torch.relu(h1) @ W2.T + b2
a2 = h2 > 0
pat = torch.cat([a1, a2], 1)
chg = (pat[1:] != pat[:-1]).any(1)
bp = xs[1:][chg].numpy()
print("segments", len(bp) + 1, "T", T)
w = np.diff(bp)
print("segment width quantiles (x units):", np.quantile(w, [0.1, 0.5, 0.9]))
# positions: s_i = 2 i / (n - 1) - 1, g_j ~ centre of the key's window
i = np.arange(n); s = 2.0 * i / (n - 1) - 1.0
j = np.arange(n_kv); g = 2.0 * (4.0 * j + 1.0) / (n - 1) - 1.0
def xof(p): return np.sign(p) * np.log2(np.abs(p) + 1.0)
ntile = (n + 31) // 32
modes = np.zeros(4, dtype=np.int64)
per_tile_heavy = 0
rng = np.random.default_rng(0)
kb_list = rng.choice(n_kv // 128, 6, replace=False)
tile_tot = 0
for kb in kb_list:
    gj = g[kb * 128:(kb + 1) * 128]                          # 128 keys = 4 warps of 32 lanes
    for tt in range(ntile):
        heavy = False
        for half in range(2):
            q0 = tt * 32 + half * 16
            if q0 >= n: continue
            q1 = min(q0 + 15, n - 1)
            xf = xof(s[q0] - gj); xl = xof(s[q1] - gj)
            c = np.searchsorted(bp, xl, side="right") - np.searchsorted(bp, xf, side="right")   # boundaries crossed per lane
            for wq in range(4):
                cw = c[wq * 32:(wq + 1) * 32]
                m = 3 if cw.max() == 0 else (0 if cw.max() <= 1 else (1 if cw.max() <= 3 else 2))
                modes[m] += 1
                heavy |= m in (1, 2)
        per_tile_heavy += heavy
        tile_tot += 1
tot = modes.sum()
print("warp-sweeps by mode  3 (no boundary) / 0 (<=1) / 1 (2-3) / 2 (>3):", (modes[[3, 0, 1, 2]] / tot).round(4))
print("tiles with at least one warp in mode 1 or 2:", round(per_tile_heavy / tile_tot, 4))
