mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_phase2_gpu.py tests/test_twohot_gpu.py tests/test_hlgauss_gpu.py tests/test_golden_gpu.py -m gpu -q 2>&1 | tail -3
MLB_PDL=0 timeout 300 python tools/profile_update.py cfg2 2 2>&1 | grep -v -i warn | head -9
timeout 300 python - <<'PY' 2>&1 | grep -v Warn | tail -2 | cut -c1-200
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import torch, bench_configs as b
b.run('cfg2 PPO MLP 3x256, 8192x32', 8192, 32, 1, 256, 3, 4, 4, torch.bfloat16, steps=30, warm=10)
PY
