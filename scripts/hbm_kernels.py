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
# ---- Nystrom layer @16k on pair storage: n_pad = 16640, H = 8, d = 64, l = 65, m = 256 ----
from dml_b200.pairs import Pair
n_pad, Hh, d, l, m = 16640, 8, 64, 65, 256
W = Hh * d
xin = synth.normal((16385, 512), 4, "xin").to(dev); lnw = torch.ones(512, device=dev); lnb = torch.zeros(512, device=dev)
xnp = Pair.empty((16385, 512), dev); mean5 = torch.empty(16385, device=dev); rstd5 = torch.empty_like(mean5)
report("layernorm_fwd_pair [16385,512] (fp32 in, operand pair out)", timeit(lambda: call("dml_layernorm_fwd_pair", ptr(xin), ptr(lnw), ptr(lnb), 16385, 512, 1e-5, None, ptr(xnp.planes), xnp.planes.stride(0), ptr(mean5), ptr(rstd5), stream())), xin.numel() * 8)
qkv = Pair.from_f32(synth.normal((1, n_pad, 3 * W), 5, "qkv").to(dev))
lm = Pair.empty((2, 1, Hh, m, d), dev)
report("ny_landmark_pool (q and k of the qkv pair)", timeit(lambda: call("dml_ny_landmark_pool", ptr(qkv.planes), qkv.planes.stride(0), 3 * W, 1, n_pad, l, Hh, d, 1.0 / l, 1.0 / l, ptr(lm.planes), lm.planes.stride(0), stream())), 2 * n_pad * W * 4)
a = synth.normal((1, n_pad, W), 6, "a").to(dev); wc = synth.uniform((Hh, 33), 6, "w", 0.2).to(dev); om = Pair.empty((1, n_pad, W), dev)
report("ny_res_conv_fwd (a + conv33(v) -> to_out operand pair)", timeit(lambda: call("dml_ny_res_conv_fwd", ptr(a), ptr(qkv.planes), qkv.planes.stride(0), 3 * W, 2 * W, ptr(wc), 33, 1, n_pad, Hh, d, ptr(om.planes), om.planes.stride(0), stream())), n_pad * W * 4 * 3)
acc = torch.empty(1, n_pad, 3 * W, device=dev); dwc = torch.empty(Hh, 33, device=dev)
report("ny_res_conv_bwd (dy fp32, v pair -> dv fp32, dw)", timeit(lambda: call("dml_ny_res_conv_bwd", ptr(a), ptr(qkv.planes), qkv.planes.stride(0), 3 * W, 2 * W, ptr(wc), 33, 1, n_pad, Hh, d, ptr(acc), 3 * W, 2 * W, ptr(dwc), stream())), n_pad * W * 4 * 3)
s3 = synth.normal((Hh * m, n_pad), 8, "s3").to(dev); y3 = Pair.empty((Hh * m, n_pad), dev); d3 = Pair.empty((Hh * m, n_pad), dev)
report("ny_softmax_rows_fwd [8*256, 16640] (fp32 in, pair out)", timeit(lambda: call("dml_ny_softmax_rows_fwd", ptr(s3), Hh * m, n_pad, ptr(y3.planes), y3.planes.stride(0), stream())), s3.numel() * 8)
report("ny_softmax_rows_bwd [8*256, 16640] (pair + fp32 in, pair out)", timeit(lambda: call("dml_ny_softmax_rows_bwd", ptr(y3.planes), y3.planes.stride(0), ptr(s3), Hh * m, n_pad, ptr(d3.planes), d3.planes.stride(0), stream())), s3.numel() * 12)
dl = synth.normal((2, 1, Hh, m, d), 9, "dl").to(dev); dq = Pair.empty((1, n_pad, 3 * W), dev)
report("ny_dqkv_finalize (fp32 in, pair out)", timeit(lambda: call("dml_ny_dqkv_finalize", ptr(acc), ptr(dl), 1, n_pad, l, Hh, d, 0.125, ptr(dq.planes), dq.planes.stride(0), stream())), acc.numel() * 8)
xp = synth.normal((1, 1 + 128 * 128, 512), 10, "xp").to(dev); yp = torch.empty_like(xp)
ws = synth.uniform((512, 49), 10, "ws", 0.1).to(dev); bs = torch.zeros(512, device=dev); dws = torch.empty(512, 49, device=dev); dbs = torch.empty(512, device=dev)
report("ppeg_stencil [128x128 grid, 512 ch] (7x7 depthwise, fp32)", timeit(lambda: call("dml_ppeg_stencil", ptr(xp), ptr(ws), ptr(bs), 1, 128, 512, 0, ptr(yp), stream())), xp.numel() * 8)
report("ppeg_wgrad [128x128 grid, 512 ch]", timeit(lambda: call("dml_ppeg_wgrad", ptr(xp), ptr(yp), 1, 128, 512, ptr(dws), ptr(dbs), stream())), xp.numel() * 8)
big = synth.normal((16385, 512), 11, "big").to(dev); bp = Pair.empty((16385, 512), dev)
report("pair_from_f32 [16385,512]", timeit(lambda: call("dml_pair_from_f32", ptr(big), 16385, 512, 512, 1.0, ptr(bp.planes), 512, bp.planes.stride(0), stream())), big.numel() * 8)
cs = torch.empty(512, device=dev)
report("colsum [16385,512]", timeit(lambda: call("dml_colsum", ptr(big), 16385, 512, 512, ptr(cs), stream())), big.numel() * 4)
print(json.dumps({"hbm_peak_gbs_measured": PEAK, "kernels": res}, indent=1))
