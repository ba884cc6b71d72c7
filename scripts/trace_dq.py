import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from dml_b200 import synth, _lib
from dml_b200.DeformableAttention1D import DeformCrossAttention1D
from tests import helpers as H
n = 16385; dev = "cuda"
mod = DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6)
mod.load_state_dict(synth.fill_like(H.deform_shapes(), 42), strict=True); mod.to(dev)
x1 = synth.normal((1, 128, n), 1, "x1").to(dev).requires_grad_(); x2 = synth.normal((1, 128, n), 1, "x2").to(dev).requires_grad_()
nt = 4096 // 32
buf = torch.zeros(nt * 2 * 4, dtype=torch.int64, device=dev)
for r in range(2):
    if r == 1: _lib.load().dml_debug_set_trace(buf.data_ptr())
    mod(x1, x2).square().sum().backward()
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(nt, 2, 4).astype(np.int64)
t0 = t[4:120]
for g in range(2):
    s_ready, sweep_done, p_seen, issued = (t0[:, g, k] for k in range(4))
    E = sweep_done - s_ready; wake = p_seen - sweep_done; iss = issued - p_seen
    H = s_ready[1:] - sweep_done[:-1]; period = s_ready[1:] - s_ready[:-1]
    print(f"group {g}: EW sweep {E.mean():.0f}  arrive->MMA warp sees P {wake.mean():.0f}  MMA issue {iss.mean():.0f}  sweep done->next S ready {H.mean():.0f}  period {period.mean():.0f}")
print("phase offset g1-g0 of S ready:", (t0[:,1,0]-t0[:,0,0]).mean())
