import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import synth
from dml_b200.NystromAttention import NystromAttention
from oracle import nystrom as ON
from oracle.golden_cases import NYSTROM_CASES
from tests import helpers as H
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
for c in NYSTROM_CASES[1:]:
    G = H.golden(c["name"])
    mod = NystromAttention(dim=c["dim"], dim_head=c["dim_head"], heads=8, num_landmarks=c["m"], pinv_iterations=6, residual=True, dropout=0.1)
    sd = synth.fill_like(H.nystrom_shapes(c["dim"], c["dim_head"]), c["seed"], 2.0)
    mod.load_state_dict(sd, strict=True); mod = mod.to(DEV).eval()
    x = synth.normal((c["b"], c["n"], c["dim"]), c["seed"], "x").to(DEV).requires_grad_()
    r = synth.normal((c["b"], c["n"], c["dim"]), c["seed"], "r").to(DEV)
    out = mod(x)
    names = [k for k, _ in mod.named_parameters()]
    gs = torch.autograd.grad((out * r).sum(), [x] + [p for _, p in mod.named_parameters()])
    for dt in (torch.float32, torch.float64):
        P = {k: v.detach().to(dt).clone().requires_grad_() for k, v in mod.state_dict().items()}
        xo = x.detach().to(dt).requires_grad_()
        ref = ON.nystrom_attention(xo, P, heads=8, dim_head=c["dim_head"], num_landmarks=c["m"])
        rg = torch.autograd.grad((ref * r.to(dt)).sum(), [xo] + [P[k] for k in names])
        e = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
        print(c["name"], dt, "out", e(out, ref), *[(nm, "%.2e" % e(a, b)) for nm, a, b in zip(["x"] + names, gs, rg)])
        if dt == torch.float64:
            print("  golden vs fp64 oracle: res_conv", e(G["grad.res_conv.weight"].to(DEV), rg[-1]), "max|ref|", float(rg[-1].abs().max()), "mean|ref|", float(rg[-1].abs().mean()))
