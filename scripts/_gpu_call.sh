python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $?"; tail -2 gpurun_out/r2_bench_final.err
