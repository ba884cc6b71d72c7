python -m pytest tests/test_gpu_losses.py -m gpu -q -x > gpurun_out/r2_pytest_losses.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/r2_pytest_losses.log
for n in 8 32 64; do python scripts/prof_gram.py $n 4 > gpurun_out/r2_gram_bw_n$n.json 2>&1; cat gpurun_out/r2_gram_bw_n$n.json; done
