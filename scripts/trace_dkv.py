"""Per-tile clock64() stamps of CTA 0 of the dK/dV kernel (build the test library with DML_B200_TRACE=1 first:
`DML_B200_TRACE=1 python disentangled-multimodal-learning_b200/build.py --force`).  Stamp k of tile t (trace[t * 8 + k]):
0 warp 0 end of sweep, 1 / 2 warp 0 tile start (barriers passed) / end, 3 / 4 the same for warp 3, 5 MMA warp: P of the
tile arrived, 6 MMA warp: dV / dK MMAs issued, 7 MMA warp: S of the next tile issued (start of the wait for P)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import synth, _lib
from dml_b200.DeformableAttention1D import DeformCrossAttention1D
from tests import helpers as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16385
dev = "cuda"
mod = DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6)
mod.load_state_dict(synth.fill_like(H.deform_shapes(), 42), strict=True)
mod.to(dev)
x1 = synth.normal((1, 128, n), 1, "x1").to(dev).requires_grad_()
x2 = synth.normal((1, 128, n), 1, "x2").to(dev).requires_grad_()
ntiles = (n + 31) // 32
trace = torch.zeros(8 * ntiles, dtype=torch.int64, device=dev)
orig = _lib.call
def routed(name, *args, **kw):
    if name == "dml_deform_attn_bwd_tc":
        return _lib.call_test(name, *args)
    return orig(name, *args, **kw)
import dml_b200.ops as ops
for m in (ops, _lib):
    if getattr(m, "call", None) is orig:
        m.call = routed
for r in range(3):
    if r == 2:
        _lib.load_test().dml_debug_set_trace(trace.data_ptr())
    out = mod(x1, x2)
    out.square().sum().backward()
torch.cuda.synchronize()
_lib.load_test().dml_debug_set_trace(None)
tr = trace.cpu().view(ntiles, 8)
used = int((tr[:, 1] != 0).sum())
tr = tr[:used].double()
t0 = tr[0, 1]
def stat(x): return {"mean": round(float(x.mean()), 1), "p50": round(float(x.median()), 1), "p90": round(float(x.quantile(0.9)), 1), "max": round(float(x.max()), 1)}
res = {
    "tiles": used,
    "cycles_per_tile": round(float((tr[-1, 2] - tr[0, 1]) / used), 1),
    "w0_busy": stat(tr[:, 2] - tr[:, 1]), "w3_busy": stat(tr[:, 4] - tr[:, 3]),
    "w0_gap_to_next_start": stat(tr[1:, 1] - tr[:-1, 2]), "w3_gap_to_next_start": stat(tr[1:, 3] - tr[:-1, 4]),
    "mma_wait_P": stat(tr[:, 5] - tr[:, 7]), "mma_issue": stat(tr[:, 6] - tr[:, 5]),
    "w0_end_to_mma_P": stat(tr[:, 5] - tr[:, 2]), "w3_end_to_mma_P": stat(tr[:, 5] - tr[:, 4]),
    "mma_S_next_issued_to_w0_next_start": stat(tr[1:, 1] - tr[:-1, 7]),
}
print(json.dumps(res, indent=1))
mid = used // 2
print("sample rows (relative cycles):")
for t in range(mid, mid + 6):
    print([int(v - tr[mid, 1]) for v in tr[t]])
per = (tr[1:, 1] - tr[:-1, 1])
print("period (start to start) per 16 tiles:", [int(per[i:i + 16].mean()) for i in range(0, len(per), 16)])
print("w0 busy per 16 tiles:", [int((tr[i:i + 16, 2] - tr[i:i + 16, 1]).mean()) for i in range(0, used, 16)])
print("w0 sweep (start..sweep end) per 16 tiles:", [int((tr[i:i + 16, 0] - tr[i:i + 16, 1]).mean()) for i in range(0, used, 16)])
