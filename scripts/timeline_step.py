"""Kernel timeline of the graph-replayed training step (torch.profiler / CUPTI): where the GPU is not running an
attention kernel, what is it running, and how much of the step is idle."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from dml_b200 import synth
from dml_b200.model import Args, bag_loss, define_net
from dml_b200.graph import GraphedTrainStep

N = 16384
dev = torch.device("cuda", 0)
net = define_net(Args(task_type="diag2021"))
net.load_state_dict(synth.fill_like({k: tuple(v.shape) for k, v in net.state_dict().items()}, 42), strict=True)
net.to(dev).train()
params = [p for p in net.parameters() if p.requires_grad]
opt = torch.optim.AdamW(params, lr=2e-4, weight_decay=0.01, fused=True)
b = synth.synthetic_bag(N, seed=1000)
bag = {"x_path": b["x_path"].to(torch.bfloat16).to(dev), "x_omic_tumor": b["x_omic_tumor"].to(dev),
       "x_omic_immune": b["x_omic_immune"].to(dev), "label": b["label_diag"].to(dev)}
keys = ("x_path", "x_omic_tumor", "x_omic_immune")
step = GraphedTrainStep(net, lambda out, bb: bag_loss(out[3], bb["label"], "diag2021"), bag, model_keys=keys,
                        flat_optimizer=lambda ps: torch.optim.AdamW(ps, lr=2e-4, weight_decay=0.01, fused=True))
for _ in range(5):
    step(bag)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step(bag)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
rows = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev))
json.dump(rows, open("gpurun_out/timeline_step.json", "w"))
print("kernel events", len(rows))
