mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_separate_gpu.py tests/test_train_gpu.py -m gpu -q --tb=short 2>&1 > gpurun_out/r2_pytest_sep.log
grep -E "^E  |^FAILED|passed|failed" gpurun_out/r2_pytest_sep.log | cut -c1-300 | head -60
