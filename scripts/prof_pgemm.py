"""Representative dml_pgemm launches of a NystromAttention layer at the 16k bag (for ncu --set full)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200.pairs import Pair, pgemm

dev = "cuda"
torch.manual_seed(0)
H, d, m, n_pad, dim = 8, 64, 256, 16640, 512
def P(*shape): return Pair.from_f32(torch.randn(*shape, device=dev))
q, kl, ql, k = P(1, H, n_pad, d), P(1, H, m, d), P(1, H, m, d), P(1, H, n_pad, d)
x, z = P(1, H, m, m), P(1, H, m, m)
xn, Wqkv = P(1, n_pad, dim), P(3 * H * d, dim)
xz_f = torch.randn(1, H, m, m, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ev = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
for r in range(reps):
    ev[0].record()
    pgemm(q, kl, M=n_pad, N=m, K=d, batch=(1, H), softmax=1, want_f32=False, want_pair=True)            # sim1 + softmax
    ev[1].record()
    pgemm(ql, k, M=m, N=n_pad, K=d, batch=(1, H))                                                        # sim3
    ev[2].record()
    pgemm(x, z, M=m, N=m, K=m, b_trans=True, batch=(1, H), alpha=-1.0, resid=xz_f, resid_scale=7.0, diag=15.0, want_f32=False, want_pair=True)
    ev[3].record()
    pgemm(xn, Wqkv.b1(), M=n_pad, N=3 * H * d, K=dim, batch=(1,), ncol_split=512, alpha2=0.125, want_f32=False, want_pair=True)
    ev[4].record()
torch.cuda.synchronize()
print("ms: sim1+softmax %.3f  sim3 %.3f  pinv product %.3f  to_qkv %.3f" % tuple(ev[i].elapsed_time(ev[i + 1]) for i in range(4)))
