"""Achieved HBM bandwidth of the similarity kernels of the batch losses (csrc/gram.cu) at the reference's sizes: attention maps
[N, 8, 2500, 144] fp32 (teacher mode, utils/loss.py:42-52; N = batch x world).  Algorithmic bytes: forward = the maps read once
(N G K 4, twice that when A and B differ); adjoint = N G K 4 read + nloc G K 4 written.  `python scripts/prof_gram.py [N] [nloc]`."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import _lib
from dml_b200._lib import call, ptr, stream
from dml_b200.loss import _row_table

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nloc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
G, K, dev = 8, 2500 * 144, "cuda"
lib = _lib.load(check_device=True)
peak = 6454.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
g = torch.Generator(device=dev).manual_seed(1)
A = torch.rand(N, G, K, device=dev, generator=g)
B = torch.rand(N, G, K, device=dev, generator=g)
ta, tb = _row_table(A, N, G * K), _row_table(B, N, G * K)
nsplit = lib.dml_gram_splits(G, K, 0)
part = torch.empty(G, nsplit, N, N, device=dev)
W = torch.rand(G, nloc, N, device=dev, generator=g)
out = torch.empty(nloc, G, K, device=dev)
mb = N * G * K * 4 / 1e6


def timed(fn, reps=5, inner=4):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        for _ in range(inner):
            fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return sorted(ev[i].elapsed_time(ev[i + 1]) / inner for i in range(reps))[reps // 2]


res = {"N": N, "G": G, "K": K, "nloc": nloc, "hbm_peak_gbs": peak, "map_mb": round(mb, 1), "k_splits": nsplit}
for name, fn, byts in (
    ("gram_same", lambda: call("dml_gram_fwd", ptr(ta), K, ptr(ta), K, G, N, K, ptr(part), stream()), mb),
    ("gram_cross", lambda: call("dml_gram_fwd", ptr(ta), K, ptr(tb), K, G, N, K, ptr(part), stream()), 2 * mb),
    ("rows_mix", lambda: call("dml_rows_mix", ptr(W), ptr(tb), K, G, nloc, N, K, ptr(out), K, G * K, stream()), mb + nloc * G * K * 4 / 1e6),
):
    ms = timed(fn)
    res[name] = {"ms": round(ms, 4), "algorithmic_mb": round(byts, 1), "gbs": round(byts / ms, 1), "frac_of_hbm_peak": round(byts / ms / peak, 3)}
print(json.dumps(res))
