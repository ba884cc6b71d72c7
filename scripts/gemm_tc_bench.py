"""Throughput of the tcgen05 split-fp16 GEMM (csrc/gemm_tc.cu) at the Nystrom-layer shapes of TransMIL @ N = 16 384
(SURVEY.md 2b K11-K17), CUDA events, against torch's exact-fp32 matmul on the same shapes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import ops, synth
dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
def t(fn, reps=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
shapes = {"to_qkv [16640,512]x[512,1536]": ((), 16640, 1536, 512), "sim1 q.k_l^T [8][16640,64]x[64,256]": ((8,), 16640, 256, 64),
          "pinv product [8][256,256]x[256,256]": ((8,), 256, 256, 256), "attn3.v [8][256,16640]x[16640,64]": ((8,), 256, 64, 16640),
          "attn1.(Z W) [8][16640,256]x[256,64]": ((8,), 16640, 64, 256)}
res = {}
for name, (batch, M, N, K) in shapes.items():
    a = synth.normal(batch + (M, K), 1, "a").to(dev); b = synth.normal(batch + (K, N), 1, "b").to(dev)
    A, Bm = ops.SplitOperand(a, True), ops.SplitOperand(b, False)
    nb = 1
    for d in batch: nb *= d
    flop = 2.0 * nb * M * N * K
    ms_gemm = t(lambda: ops.gemm_nt(A, Bm, batch + (M, N)))
    ms_all = t(lambda: ops.mm_tc(a, b))
    ms_ref = t(lambda: a @ b)
    res[name] = {"gemm_kernel_ms": round(ms_gemm, 4), "algorithmic_TFLOPs": round(flop / ms_gemm / 1e9, 1),
                 "issued_mma_TFLOPs (3 MMAs per k-step)": round(3 * flop / ms_gemm / 1e9, 1),
                 "with_split_kernels_ms": round(ms_all, 4), "torch_fp32_matmul_ms": round(ms_ref, 4)}
print(json.dumps(res, indent=1))
