"""Profiling driver: eager training steps of TransMIL (2 x NystromAttention + PPEG) on one N-patch bf16 bag, grade CE.
Only the LAST step lies between cudaProfilerStart/Stop (run under `ncu --profile-from-start off` for the launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from dml_b200 import synth
from dml_b200.model import Args, define_net

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
net = define_net(Args(mode="path", label_dim=3))
net.load_state_dict(synth.fill_like({k: tuple(v.shape) for k, v in net.state_dict().items()}, 42), strict=True)
net.to(dev).train()
b = synth.synthetic_bag(N, seed=1000)
x, label = b["x_path"].to(torch.bfloat16).to(dev), b["label_grade"].to(dev)
w = torch.tensor([1.47, 1.51, 1.0], device=dev)
for s in range(steps):
    last = s == steps - 1
    if last:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    loss = F.cross_entropy(net(x)[1], label, weight=w)
    net.zero_grad(set_to_none=True)
    loss.backward()
    if last:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
print("loss", float(loss))
