mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tf32_gpu.py tests/test_continuous_gpu.py -m gpu -q --tb=short 2>&1 > gpurun_out/r2_pytest_x.log
grep -E "^E  |^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_x.log | cut -c1-400 | head -30
