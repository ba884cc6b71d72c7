"""Kernel timeline of one graph-replayed DeformPathomicNet training step (the bench.py step) from torch.profiler (CUPTI):
per kernel stream / start / duration, the gaps in which no attention kernel runs, and what runs in them."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from dml_b200 import synth
from dml_b200.model import Args, bag_loss, define_net
from dml_b200.graph import GraphedTrainStep

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
TASK = "diag2021"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
net = define_net(Args(task_type=TASK))
net.load_state_dict(synth.fill_like({k: tuple(v.shape) for k, v in net.state_dict().items()}, 42), strict=True)
net.to(dev).train()
b = synth.synthetic_bag(N, seed=1000)
bag = {"x_path": b["x_path"].to(torch.bfloat16).to(dev), "x_omic_tumor": b["x_omic_tumor"].to(dev),
       "x_omic_immune": b["x_omic_immune"].to(dev), "label": b["label_diag"].to(dev)}
keys = ("x_path", "x_omic_tumor", "x_omic_immune")
gstep = GraphedTrainStep(net, lambda out, bb: bag_loss(out[3], bb["label"], TASK), bag,
                         flat_optimizer=lambda ps: torch.optim.AdamW(ps, lr=2e-4, weight_decay=0.01, fused=True),
                         model_keys=keys, warmup=3)
for _ in range(5):
    gstep(bag)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        gstep(bag)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# split into steps at the first attention forward of each replay
rows = [(e.name, e.time_range.start, e.time_range.end, getattr(e, "device_resource_id", getattr(e, "stream", -1))) for e in ev]
print("cuda events:", len(rows))
# the middle replay: between the 2nd and 3rd occurrences of the first kernel name pattern 'Memcpy' / first kernel of a replay
starts = [i for i, r in enumerate(rows) if "deform_attn_fwd" in r[0]]
print("attention forwards:", len(starts))
if len(starts) >= 6:
    # each replay has two attention forwards; take replay 2: from the kernel after replay 1's last kernel
    fw = starts[2]
    # walk back to the start of this replay: the biggest idle gap before fw
    lo = fw
    while lo > 0 and rows[lo][1] - max(r[2] for r in rows[max(0, lo - 40):lo]) < 20.0 and fw - lo < 400:
        lo -= 1
    hi = starts[4]
    while hi > lo and rows[hi][1] - max(r[2] for r in rows[max(0, hi - 40):hi]) < 20.0 and starts[4] - hi < 400:
        hi -= 1
    seg = rows[lo:hi]
    t0 = seg[0][1]
    print(f"replay: {len(seg)} kernels, {seg[-1][2] - t0:.1f} us from first start to last end")
    out = []
    for name, s, e, st in seg:
        short = name.split("(")[0].replace("dml::", "").replace("void ", "")[:48]
        out.append((round(s - t0, 1), round(e - s, 1), st, short))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/timeline_step.json", "w") as f:
        json.dump(out, f)
    big = [o for o in out if o[1] > 100.0]
    print("kernels > 100 us:", big)
    # time not covered by any kernel > 100 us, and the kernels that run there
    cover = sorted((o[0], o[0] + o[1]) for o in big)
    gaps, cur = [], 0.0
    for s, e in cover:
        if s > cur:
            gaps.append((cur, s))
        cur = max(cur, e)
    gaps.append((cur, out[-1][0] + out[-1][1]))
    for g0, g1 in gaps:
        inside = [o for o in out if o[0] >= g0 - 0.5 and o[0] < g1 and o[1] <= 100.0]
        busy = sum(o[1] for o in inside)
        print(f"gap {g0:8.1f} .. {g1:8.1f} ({g1 - g0:6.1f} us): {len(inside)} kernels, {busy:.1f} us of kernel time")
        for o in inside[:60]:
            print("     ", o)
