mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tf32_gpu.py tests/test_train_gpu.py tests/test_phase2_gpu.py -m gpu -q --tb=short 2>&1 > gpurun_out/r2_pytest_tf32.log
grep -E "^E  |^FAILED|passed|failed" gpurun_out/r2_pytest_tf32.log | cut -c1-250 | head -40
timeout 300 python tools/tf32_bench.py 2>&1 | grep -v -i warn > gpurun_out/r2_tf32_bench2.txt
grep -E "layer|cfg2" gpurun_out/r2_tf32_bench2.txt | cut -c1-175
