"""Dump the bias table and g of the synthetic north-star attention (for offline analysis of the table paths)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from dml_b200 import synth, ops
from dml_b200.DeformableAttention1D import DeformCrossAttention1D
from tests import helpers as H
n = 16385; dev = "cuda"
mod = DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6)
mod.load_state_dict(synth.fill_like(H.deform_shapes(), 42), strict=True); mod.to(dev)
x1 = synth.normal((1, 128, n), 1, "x1").to(dev).requires_grad_(); x2 = synth.normal((1, 128, n), 1, "x2").to(dev).requires_grad_()
saved = {}
orig = ops.DeformCrossAttn1DFn.backward
def hook(ctx, *a):
    t = ctx.saved_tensors
    saved["g"] = t[7].detach().cpu().numpy(); saved["table"] = t[8].detach().cpu().numpy()
    return orig(ctx, *a)
ops.DeformCrossAttn1DFn.backward = staticmethod(hook)
mod(x1, x2).square().sum().backward()
torch.cuda.synchronize()
np.savez("gpurun_out/table_dump.npz", **saved)
print("ok", saved["g"].shape, saved["table"].shape)
