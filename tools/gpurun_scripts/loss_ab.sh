mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_phase2_gpu.py tests/test_twohot_gpu.py tests/test_hlgauss_gpu.py tests/test_continuous_gpu.py tests/test_golden_gpu.py tests/test_train_gpu.py tests/test_tc_gpu.py tests/test_update_shapes_gpu.py tests/test_select_gpu.py -m gpu -q --tb=short 2>&1 | grep -E "^E  |^FAILED|^ERROR|passed|failed" | cut -c1-300 | head -20
for h in 1 0 1 0; do MLB_LOSS_PERSIST=$h timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 60 --warmup 20 2>/dev/null | python -c "
import sys, json
b=json.loads(sys.stdin.readline()); print('loss_persist=$h', b['ms_per_step'], b['value'], b['e2e']['value'])"; done
