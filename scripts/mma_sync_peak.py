"""Measured issue rate of the legacy tensor path (mma.sync m16n8k16 bf16) on this GPU: the denominator for the position-bias
MLP kernels.  Test-only library; prints one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dml_b200 import _lib

out = torch.zeros(4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
res = {}
for ctas_per_sm in (1, 2, 4):
    for chains in (4, 8, 16):
        ctas, iters = 148 * ctas_per_sm, 20000
        _lib.call_test("dml_test_mma_sync_peak", ctas, chains, 100, out.data_ptr(), st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.call_test("dml_test_mma_sync_peak", ctas, chains, iters, out.data_ptr(), st)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        res[f"{ctas_per_sm}cta_x8warps_{chains}chains"] = round(ctas * 8 * chains * iters * 2 * 4096 / (ms * 1e-3) / 1e12, 1)
print(json.dumps({"mma_sync_m16n8k16_bf16_tflops": res, "best": max(res.values())}))
