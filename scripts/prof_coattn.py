"""Achieved HBM bandwidth of the four co-attention streaming kernels (csrc/coattn.cu) at MCAT / CMTA sizes: CUDA events around
each C-ABI call, input sets larger than L2 rotated between repetitions.  Algorithmic bytes per launch:
  fq_fwd  B S E 4 (x read once) + B F S 4 (raw written)          fq_bwd  2 B S E 4 (x read, dx written) + B F S 4 (raw read)
  fk_fwd  2 B S E 4 (x read, out written) + B S F 4              fk_bwd  3 B S E 4 (x, dout read, dx written) + B S F 4
Prints one JSON object; `python scripts/prof_coattn.py [B] [S] [F]`."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import _lib
from dml_b200._lib import call, ptr, stream

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
F = int(sys.argv[3]) if len(sys.argv) > 3 else 4
E, dev = 256, "cuda"
lib = _lib.load(check_device=True)
peak = 6454.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
nset = max(2, int(300e6 // (B * S * E * 4)) + 1)          # rotate > 126 MB of inputs
g = torch.Generator(device=dev).manual_seed(1)
xs = [torch.randn(B, S, E, device=dev, generator=g) for _ in range(nset)]
gs = [torch.randn(B, S, E, device=dev, generator=g) for _ in range(nset)]
few = torch.randn(B, F, E, device=dev, generator=g) * 0.05
few2 = torch.randn(B, F, E, device=dev, generator=g)
c = torch.randn(B, F, device=dev, generator=g)
bo = torch.randn(E, device=dev, generator=g)
raw_q = torch.empty(B, F, S, device=dev); px = torch.empty(B, F, E, device=dev); lse = torch.empty(B, F, device=dev)
raw_k = torch.empty(B, S, F, device=dev); out = torch.empty(B, S, E, device=dev); dx = torch.empty(B, S, E, device=dev)
ws1 = torch.empty(lib.dml_coattn_fq_fwd_ws_floats(B, F, S, E), device=dev)
ws2 = torch.empty(lib.dml_coattn_fq_bwd_ws_floats(B, F, S, E), device=dev)
ws3 = torch.empty(lib.dml_coattn_fk_bwd_ws_floats(B, F, S, E), device=dev)
dsum = torch.zeros(B, F, device=dev)
sb, sr = S * E, E


def run(name, i):
    x, gg = xs[i % nset], gs[i % nset]
    if name == "fq_fwd":
        call("dml_coattn_fq_fwd", ptr(x), sb, sr, ptr(few), ptr(c), B, F, S, E, ptr(raw_q), ptr(px), ptr(lse), ptr(ws1), stream())
    elif name == "fq_bwd":
        call("dml_coattn_fq_bwd", ptr(x), sb, sr, ptr(few), ptr(raw_q), ptr(lse), ptr(few2), ptr(dsum), None, B, F, S, E, ptr(dx), ptr(ws2),
             stream())
    elif name == "fk_fwd":
        call("dml_coattn_fk_fwd", ptr(x), sb, sr, ptr(few), ptr(c), ptr(few2), ptr(bo), B, F, S, E, ptr(raw_k), ptr(out), stream())
    else:
        call("dml_coattn_fk_bwd", ptr(x), sb, sr, ptr(gg), sb, sr, ptr(few), ptr(few2), ptr(raw_k), None, B, F, S, E, ptr(dx), ptr(ws3),
             stream())


bytes_ = {"fq_fwd": B * S * E * 4 + B * F * S * 4, "fq_bwd": 2 * B * S * E * 4 + B * F * S * 4,
          "fk_fwd": 2 * B * S * E * 4 + B * S * F * 4, "fk_bwd": 3 * B * S * E * 4 + B * S * F * 4}
res = {"B": B, "S": S, "F": F, "E": E, "hbm_peak_gbs": peak, "input_sets": nset}
for name in ("fq_fwd", "fq_bwd", "fk_fwd", "fk_bwd"):
    for i in range(3):
        run(name, i)
    torch.cuda.synchronize()
    reps, inner = 10, 10          # events around `inner` back-to-back launches: the host call overhead stays off the clock
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        for j in range(inner):
            run(name, i * inner + j)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) / inner for i in range(reps))
    med = ts[reps // 2]
    res[name] = {"ms_median": round(med, 4), "ms_min": round(ts[0], 4), "algorithmic_mb": round(bytes_[name] / 1e6, 1),
                 "gbs": round(bytes_[name] / med / 1e6, 1), "frac_of_hbm_peak": round(bytes_[name] / med / 1e6 / peak, 3)}
print(json.dumps(res))
