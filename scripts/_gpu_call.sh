set -x
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py -m gpu -q -x > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2_pytest_gpu2.log
python scripts/prof_attn.py 16385 3 > gpurun_out/r2_prof_attn_events2.json 2>&1; cat gpurun_out/r2_prof_attn_events2.json
python bench.py --no-cpu-baseline --no-transmil --no-cls-row-only > gpurun_out/r2_bench_mode3.json 2> gpurun_out/r2_bench_mode3.err; echo "exit $?"; cut -c1-400 gpurun_out/r2_bench_mode3.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deform_attn_fwd -s 1 -c 1 -o gpurun_out/r2_attn_fwd_full -f python scripts/prof_attn.py 16385 2 > gpurun_out/ncu_attn_fwd_full.log 2>&1; echo "exit $?"
