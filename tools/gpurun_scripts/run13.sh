mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_phase2_gpu.py tests/test_train_gpu.py tests/test_tf32_gpu.py -m gpu -q --tb=short 2>&1 | grep -E "^E  |^FAILED|passed|failed" | cut -c1-200
PROFILE_DTYPE=f32 timeout 300 python tools/profile_update.py cfg2 2 2>&1 | grep -v -i warn | head -8 | cut -c1-160
MLB_TF32_ROUND=1 timeout 300 python tools/tf32_bench.py 2>&1 | grep cfg2 | cut -c1-200
