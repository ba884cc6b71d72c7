"""In-kernel clock64() trace of the dK/dV kernel (CTA 0): where a query tile's time goes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from dml_b200 import synth, _lib
from dml_b200.DeformableAttention1D import DeformCrossAttention1D
from tests import helpers as H
n = 16385; dev = "cuda"
mod = DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6)
mod.load_state_dict(synth.fill_like(H.deform_shapes(), 42), strict=True); mod.to(dev)
x1 = synth.normal((1, 128, n), 1, "x1").to(dev).requires_grad_(); x2 = synth.normal((1, 128, n), 1, "x2").to(dev).requires_grad_()
nt = (n + 31) // 32
buf = torch.zeros(nt * 8, dtype=torch.int64, device=dev)
for r in range(2):
    if r == 1: _lib.load().dml_debug_set_trace(buf.data_ptr())
    mod(x1, x2).square().sum().backward()
torch.cuda.synchronize()
_lib.load().dml_debug_set_trace(None)
t = buf.cpu().numpy().reshape(nt, 8).astype(np.int64)
np.save("gpurun_out/trace_dkv.npy", t)
for lo, hi in ((8, 100), (200, 300), (330, 430)):
    w = t[lo:hi]
    print(f"tiles {lo}-{hi}: period {np.diff(w[:, 1]).mean():.0f}"
          f" | warp0: sweep {(w[:, 2] - w[:, 1]).mean():.0f} of which segment hand-over {(w[:, 2] - w[:, 0]).mean():.0f}"
          f" | MMA: warp0 sweep done -> P seen {(w[:, 5] - w[:, 2]).mean():.0f}, issue dV/dK {(w[:, 6] - w[:, 5]).mean():.0f},"
          f" S(t+1) issued - S(t) seen by warp0 {(w[:, 7] - w[:, 1]).mean():.0f}")
