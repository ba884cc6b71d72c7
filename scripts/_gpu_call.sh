python -m pytest tests/test_gpu_nystrom.py tests/test_gpu_modules.py -m gpu -q -x -k "nystrom or transmil or pinv or Nystrom or TransMIL or res_conv" 2>&1 | tail -3
for n in 16384 6000; do python bench.py --workload transmil --n-patches $n --no-cpu-baseline > gpurun_out/r2_bench_transmil_rc_$n.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_transmil_rc_$n.json').read().strip().splitlines()[-1]); print('rc $n', d['ms_per_step'], {k:round(v,3) for k,v in list(d['kernel_ms_per_step'].items())[:8]})"; done
