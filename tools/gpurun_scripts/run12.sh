mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/r2_pytest_full.log
grep -E "^E  |^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_full.log | cut -c1-300 | head -60
