"""Time ClusterMergeNet (DPC-KNN + merge, csrc/cluster.cu) per entry point: `python scripts/prof_cluster.py [B] [N] [ratio]`.
Distance work: B N^2 x 128 squared differences for the density pass and again for the parent pass (the reference forms the
N x N matrix with cdist and reads it five more times)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import _lib, synth
from dml_b200.ClusterMergeNet import ClusterMergeNet
from tests import helpers as H

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 99856
ratio = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0008
dev = "cuda"
mod = ClusterMergeNet(sample_ratio=ratio, dim_out=128)
mod.load_state_dict(synth.fill_like(H.cluster_shapes(), 3), strict=True)
mod = mod.to(dev)
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(B, N, 128, device=dev, generator=g).requires_grad_()


def fwd_bwd():
    tok = dict(x=x, token_num=N, idx_token=torch.arange(N, device=dev)[None].repeat(B, 1), agg_weight=x.new_ones(B, N, 1))
    down, _ = mod(tok)
    down["x"].sum().backward()
    return down["x"].shape[1]


K = fwd_bwd()
torch.cuda.synchronize()
events = []


def hook(name, phase):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    events.append((name, ev))


_lib._timing_hook = hook
fwd_bwd()
torch.cuda.synchronize()
_lib._timing_hook = None
kt = {}
for i in range(0, len(events), 2):
    kt[events[i][0]] = kt.get(events[i][0], 0.0) + events[i][1].elapsed_time(events[i + 1][1])
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    fwd_bwd()
b.record()
torch.cuda.synchronize()
flop = 3.0 * B * N * N * 128
print(json.dumps({"B": B, "N": N, "clusters": K, "ms_fwd_bwd": a.elapsed_time(b) / 3, "entry_points_ms": {k: round(v, 3) for k, v in kt.items()},
                  "density_pass_tflops_fp32": flop / (kt["dml_dpc_density"] * 1e-3) / 1e12,
                  "parent_pass_tflops_fp32": flop / (kt["dml_dpc_parent"] * 1e-3) / 1e12,
                  "reference_distance_matrix_gb": B * N * N * 4 / 1e9}))
