python -m pytest tests/test_gpu_pgemm.py tests/test_gpu_nystrom.py -m gpu -q -x 2>&1 | tail -2
python bench.py --workload transmil --no-cpu-baseline > gpurun_out/r2_bench_transmil_narrow.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_transmil_narrow.json').read().strip().splitlines()[-1]); print('narrow', d['ms_per_step'], d['kernel_ms_per_step']['dml_pgemm'])"
DML_B200_PGEMM_NARROW=0 python bench.py --workload transmil --no-cpu-baseline > gpurun_out/r2_bench_transmil_wide.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_transmil_wide.json').read().strip().splitlines()[-1]); print('wide', d['ms_per_step'], d['kernel_ms_per_step']['dml_pgemm'])"
python bench.py --workload transmil --n-patches 6000 --no-cpu-baseline > gpurun_out/r2_bench_transmil_narrow_6000.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_transmil_narrow_6000.json').read().strip().splitlines()[-1]); print('narrow 6000', d['ms_per_step'], d['kernel_ms_per_step']['dml_pgemm'])"
