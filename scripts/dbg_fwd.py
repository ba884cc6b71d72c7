import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import synth, _lib
from dml_b200._lib import call, ptr, stream
from oracle import deform1d as O
from tests.test_gpu_kernels import mlp_params, build_table, attn_reference
DEV = "cuda"
for (B, n, n_kv) in [(1, 193, 48), (1, 517, 129)]:
    Hh, d, nout = 8, 64, 2
    G, C = Hh // nout, Hh * d
    seed = 300 + n
    q = (synth.normal((B, n, C), seed, "q") * 0.7).to(DEV).half()
    k = (synth.normal((B, n_kv, C), seed, "k") * 0.7).to(DEV).half()
    v = synth.normal((B, n_kv, C), seed, "v").to(DEV).half()
    vgrid = torch.arange(n_kv, device=DEV)[None] + synth.uniform((B * G, n_kv), seed, "off", 2.0).to(DEV)
    g = O.normalize_grid(vgrid).contiguous()
    P = mlp_params(seed)
    table, _ = build_table(P, math.log1p(2.0 + 4.0 / max(n_kv - 1, 1)) * 1.001 + 1e-3)
    ref = attn_reference(q.float(), k.float(), v.float(), g, P, Hh, nout, d ** -0.5, n)
    for rep in range(3):
        o = torch.full((B, n, C), float("nan"), device=DEV); lse = torch.full((B, Hh, n), float("nan"), device=DEV)
        call("dml_deform_attn_fwd_tc", ptr(q), ptr(k), ptr(v), ptr(g), ptr(table), B, Hh, d, n, n_kv, n, C, C, C, C, nout, d ** -0.5, ptr(o), ptr(lse), stream())
        err = (o - ref).abs().reshape(B, n, Hh, d).amax(-1)[0]      # [n, H]
        bad = (err > 1e-2 * ref.abs().max()).nonzero()
        print((B, n, n_kv), "rep", rep, "max err", float(err.max()), "bad rows/heads:", bad.shape[0], bad[:12].tolist())
