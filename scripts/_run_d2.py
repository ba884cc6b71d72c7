import sys, json
sys.path.insert(0, '/root/repo')
import torch, bench
r = bench.bench_deform2d("cuda", cpu=False)
print(json.dumps({k: r[k] for k in ("value", "ms_per_step", "config")}))
