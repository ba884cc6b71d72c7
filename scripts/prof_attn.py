"""Profiling driver: one DeformCrossAttention1D fwd+bwd at the north-star size (n=16385, n_kv=4096),
weights from synth (same as bench.py).  Used under ncu; prints CUDA-event times per entry point."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import synth, _lib
from dml_b200.DeformableAttention1D import DeformCrossAttention1D
from tests import helpers as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16385
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda"
mod = DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6)
mod.load_state_dict(synth.fill_like(H.deform_shapes(), 42), strict=True)
mod.to(dev)
x1 = synth.normal((1, 128, n), 1, "x1").to(dev).requires_grad_()
x2 = synth.normal((1, 128, n), 1, "x2").to(dev).requires_grad_()
events = []
def hook(name, phase):
    ev = torch.cuda.Event(enable_timing=True); ev.record(); events.append((name, phase, ev))
for r in range(reps):
    if r == reps - 1:
        _lib._timing_hook = hook
    out = mod(x1, x2)
    out.square().sum().backward()
torch.cuda.synchronize()
res = {}
for i in range(0, len(events), 2):
    res[events[i][0]] = events[i][2].elapsed_time(events[i + 1][2])
print(json.dumps(res))
