python -m pytest tests/test_gpu_trainer.py -m gpu -q -x > gpurun_out/r2_pytest_trainer.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/r2_pytest_trainer.log | cut -c1-400
