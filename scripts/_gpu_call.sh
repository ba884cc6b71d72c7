python -m pytest tests/test_gpu_coattn.py -m gpu -q -x > gpurun_out/r2_pytest_coattn.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_coattn.log
python scripts/prof_coattn.py 8 16384 4 > gpurun_out/r2_coattn_bw_b8_s16384_f4.json 2>&1; cat gpurun_out/r2_coattn_bw_b8_s16384_f4.json
python scripts/prof_coattn.py 8 2500 4 > gpurun_out/r2_coattn_bw_b8_s2500_f4.json 2>&1; cat gpurun_out/r2_coattn_bw_b8_s2500_f4.json
python scripts/prof_coattn.py 1 16384 6 > gpurun_out/r2_coattn_bw_b1_s16384_f6.json 2>&1; cat gpurun_out/r2_coattn_bw_b1_s16384_f6.json
