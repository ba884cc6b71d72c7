"""Time DeformCrossAttention2D per C-ABI entry point and as a whole: CUDA events around every call of an
eager forward + backward of the module (eval mode, gradients arriving at out and attn), then events around REPS whole
forward + backward passes.  `python scripts/prof_deform2d.py [B] [side] [reps]`; prints one JSON object.

Work per pass (B bags, n = side^2 queries, m keys, 8 heads): pairs = 8 B n m; bias MLP dense maths 2 * 1120 FLOP per pair
forward (2 x that backward + the recompute); the attention map itself is an OUTPUT (4 B per pair written forward; dS written and
P, dS, dA read backward)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import _lib, synth
from dml_b200.DeformableAttention2D import DeformCrossAttention2D
from tests import helpers as H

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
side = int(sys.argv[2]) if len(sys.argv) > 2 else 50
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = "cuda"
mod = DeformCrossAttention2D(dim=128, dim_head=64, heads=8, dropout=0.1, downsample_factor=4, offset_scale=4, offset_groups=8,
                             offset_kernel_size=6)
mod.load_state_dict(synth.fill_like(H.attn2d_shapes(""), 5, gain=2.0), strict=True)
mod = mod.to(dev).eval()
n = side * side
g = torch.Generator(device=dev).manual_seed(1)
x1 = torch.randn(B, 128, n, device=dev, generator=g).requires_grad_()
x2 = torch.randn(B, 128, n, device=dev, generator=g).requires_grad_()


def fwd_bwd():
    out, attn = mod(x1, x2)
    r = out.sum() + (attn * attn).sum()
    r.backward()
    return attn.shape[-1]


for _ in range(3):
    m = fwd_bwd()
torch.cuda.synchronize()
events = []


def hook(name, phase):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    events.append((name, ev))


_lib._timing_hook = hook
fwd_bwd()
torch.cuda.synchronize()
_lib._timing_hook = None
kt = {}
for i in range(0, len(events), 2):
    kt.setdefault(events[i][0], []).append(events[i][1].elapsed_time(events[i + 1][1]))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    fwd_bwd()
b.record()
torch.cuda.synchronize()
pairs = 8 * B * n * m
kms = {k: [round(x, 4) for x in v] for k, v in kt.items()}
bf, bb = sum(kt["dml_da2_bias_fwd"]), sum(kt["dml_da2_bias_bwd"])
print(json.dumps({"B": B, "side": side, "n": n, "m": m, "pairs": pairs, "ms_fwd_bwd": a.elapsed_time(b) / reps,
                  "entry_points_ms": kms,
                  "bias_fwd_dense_tflops": pairs * 2240 / (bf * 1e-3) / 1e12, "bias_bwd_dense_tflops": pairs * 2 * 2240 / (bb * 1e-3) / 1e12,
                  "bias_fwd_issued_mma_tflops": pairs * 3 * 2048 / (bf * 1e-3) / 1e12,
                  "bias_bwd_issued_mma_tflops": pairs * 9 * 2048 / (bb * 1e-3) / 1e12}))
