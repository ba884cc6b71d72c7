"""Profiling driver: eager training steps of the bench workload (DeformPathomicNet, one 16 384-patch bf16 bag,
diag2021 CE, AdamW).  Used under ncu for the per-launch time list of the whole step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import synth
from dml_b200.model import Args, bag_loss, define_net

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
net = define_net(Args(task_type="diag2021"))
net.load_state_dict(synth.fill_like({k: tuple(v.shape) for k, v in net.state_dict().items()}, 42), strict=True)
net.to(dev).train()
opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=2e-4, weight_decay=0.01, fused=True)
b = synth.synthetic_bag(N, seed=1000)
bag = {"x_path": b["x_path"].to(torch.bfloat16).to(dev), "x_omic_tumor": b["x_omic_tumor"].to(dev),
       "x_omic_immune": b["x_omic_immune"].to(dev), "label": b["label_diag"].to(dev)}
for s in range(steps):
    if s == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("last_step")
    out = net(x_path=bag["x_path"], x_omic_tumor=bag["x_omic_tumor"], x_omic_immune=bag["x_omic_immune"])
    loss = bag_loss(out[3], bag["label"], "diag2021")
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
