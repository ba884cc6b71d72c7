"""Achieved HBM GB/s of the bandwidth-bound kernels at the north-star sizes, CUDA events on the
launching stream, inputs larger than L2 rotated (8 distinct copies) or L2 flushed.  Prints one JSON object."""
import json, os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import synth, ops
from dml_b200._lib import call, ptr, stream

dev = "cuda"
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
def report(name, ms, nbytes):
    res[name] = {"ms": round(ms, 4), "MB": round(nbytes / 1e6, 2), "GB/s": round(nbytes / ms / 1e6, 1), "frac_of_measured_peak": round(nbytes / ms / 1e6 / PEAK, 3)}

# ---- config A: n = 16385, n_kv = 4096, C = 512, G = 4 ----
B, n, C, G, ks, stride, dim = 1, 16385, 512, 4, 6, 4, 128
n_kv = (n + 2 - ks) // stride + 1
q = synth.normal((B, n, C), 1, "q").to(dev).half()
w0 = synth.uniform((128, ks), 1, "w0", 0.4).to(dev); b0 = synth.uniform((128,), 1, "b0", 0.1).to(dev); w2 = synth.uniform((128,), 1, "w2", 0.1).to(dev)
vgrid = torch.empty(B * G, n_kv, device=dev); g = torch.empty_like(vgrid)
report("offsets_fwd (reads q fp16 once)", timeit(lambda: call("dml_offsets_fwd", ptr(q), ptr(w0), ptr(b0), ptr(w2), B, n, C, G, ks, stride, 2.0, ptr(vgrid), ptr(g), stream())),
       q.numel() * 2 * 1.5 + vgrid.numel() * 8)          # 6-tap / stride-4 window: each q row is read 1.5 times from L1/L2, once from HBM -> count the algorithmic 1x + halo
x2 = synth.normal((B, n, dim), 1, "x2").to(dev)
kv = torch.empty(B, n_kv, dim, device=dev)
i0, i1, wy0, wy1 = ops.centre_taps(n)
report("kv_gather_fwd (writes kv fp32)", timeit(lambda: call("dml_kv_gather_fwd", ptr(x2), ptr(g), B, n, dim, G, n_kv, i0, i1, wy0, wy1, ptr(kv), stream())), kv.numel() * 4 + g.numel() * 4)
d_off = synth.normal((B * G, n_kv), 2, "d").to(dev); dq_attn = synth.normal((B, n, C), 2, "dqa").to(dev)
dy_ws = torch.empty(B * G, n_kv, 128, device=dev); wgrad = torch.empty(128 * ks + 256, device=dev); dq = torch.empty(B, n, C, device=dev)
report("offsets_bwd (q fp16 in, dq_attn fp32 in, dq fp32 out)", timeit(lambda: call("dml_offsets_bwd", ptr(q), ptr(w0), ptr(b0), ptr(w2), ptr(d_off), ptr(dq_attn), 0.125, B, n, C, G, ks, stride, 2.0, ptr(dy_ws), ptr(wgrad), ptr(dq), stream())),
       q.numel() * 2 + dq_attn.numel() * 4 + dq.numel() * 4 + 2 * dy_ws.numel() * 4)
# ---- LayerNorm [16385, 128] ----
x = synth.normal((n, 128), 3, "x").to(dev); ln = torch.nn.LayerNorm(128).to(dev)
y = torch.empty_like(x); mean = torch.empty(n, device=dev); rstd = torch.empty(n, device=dev)
report("layernorm_fwd [16385,128]", timeit(lambda: call("dml_layernorm_fwd", ptr(x), ptr(ln.weight), ptr(ln.bias), n, 128, 1e-5, ptr(y), ptr(mean), ptr(rstd), stream())), x.numel() * 8)
dyv = synth.normal((n, 128), 4, "dy").to(dev); dx = torch.empty_like(x); dw = torch.empty(128, device=dev); db = torch.empty(128, device=dev)
report("layernorm_bwd [16385,128]", timeit(lambda: call("dml_layernorm_bwd", ptr(dyv), ptr(x), ptr(ln.weight), ptr(mean), ptr(rstd), n, 128, ptr(dx), ptr(dw), ptr(db), stream())), x.numel() * 12)
# ---- Nystrom layer @16k: n_pad = 16640, H = 8, d = 64, l = 65 ----
n_pad, Hh, d, l = 16640, 8, 64, 65
qkv = synth.normal((1, n_pad, 3 * Hh * d), 5, "qkv").to(dev)
out = torch.empty(1, Hh, n_pad // l, d, device=dev)
report("landmark_pool_fwd (q columns of the fused qkv buffer)", timeit(lambda: call("dml_landmark_pool_fwd", ptr(qkv), 3 * Hh * d, 0, 1, n_pad, l, Hh, d, 1.0 / l, ptr(out), stream())), n_pad * Hh * d * 4)
a = synth.normal((1, Hh, n_pad, d), 6, "a").to(dev); wc = synth.uniform((Hh, 33), 6, "w", 0.2).to(dev); yv = torch.empty(1, n_pad, Hh * d, device=dev)
vs = qkv[..., 2 * Hh * d:]
report("res_conv_merge_fwd (a + conv33(v))", timeit(lambda: call("dml_res_conv_merge_fwd", ptr(a), ptr(vs), 3 * Hh * d, 0, ptr(wc), 33, 1, n_pad, Hh, d, ptr(yv), stream())), n_pad * Hh * d * 4 * 3)
s1 = synth.normal((Hh, n_pad, 256), 7, "s").to(dev); y1 = torch.empty_like(s1)
report("softmax_rows_fwd [8*16640, 256]", timeit(lambda: call("dml_softmax_rows_fwd", ptr(s1), ptr(y1), Hh * n_pad, 256, stream())), s1.numel() * 8)
s3 = synth.normal((Hh, 256, n_pad), 8, "s3").to(dev); y3 = torch.empty_like(s3)
report("softmax_rows_fwd [8*256, 16640]", timeit(lambda: call("dml_softmax_rows_fwd", ptr(s3), ptr(y3), Hh * 256, n_pad, stream())), s3.numel() * 8)
print(json.dumps({"hbm_peak_gbs_measured": PEAK, "kernels": res}, indent=1))
