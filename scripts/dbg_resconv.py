import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from dml_b200 import ops, synth
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
for (B, n_pad, Hh, d, K) in [(1, 256, 8, 16, 33), (1, 256, 8, 64, 33), (1, 512, 8, 64, 33), (1, 320, 8, 16, 33)]:
    W = Hh * d
    qkv = synth.normal((B, n_pad, 3 * W), 11, "qkv").to(DEV)
    a = synth.normal((B, Hh, n_pad, d), 11, "a").to(DEV).requires_grad_()
    w = synth.uniform((Hh, 1, K, 1), 11, "w", 0.3).to(DEV).requires_grad_()
    v = qkv[..., 2 * W:].detach().requires_grad_()
    y = ops.ResConvMergeFn.apply(a, v, w)
    vh = v.reshape(B, n_pad, Hh, d).transpose(1, 2)
    ref = (a + F.conv2d(vh, w, padding=(K // 2, 0), groups=Hh)).transpose(1, 2).reshape(B, n_pad, W)
    r = synth.normal((B, n_pad, W), 12, "r").to(DEV)
    ga, gv, gw = torch.autograd.grad((y * r).sum(), (a, v, w))
    ra, rv, rw = torch.autograd.grad((ref * r).sum(), (a, v, w))
    e = lambda x, y: float((x - y).abs().max() / y.abs().max())
    print((B, n_pad, Hh, d, K), "y", e(y, ref), "da", e(ga, ra), "dv", e(gv, rv), "dw", e(gw, rw))
    if e(gw, rw) > 1e-4:
        print(((gw - rw).abs().reshape(Hh, K) / rw.abs().max()).amax(1))
        print(((gw - rw).abs().reshape(Hh, K) / rw.abs().max()).amax(0))
