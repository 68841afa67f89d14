mkdir -p gpurun_out
MLB_TC_EPI=1 timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_update_shapes_gpu.py tests/test_train_gpu.py -m gpu -q 2>&1 | tail -3
for e in 0 1 0 1; do
MLB_TC_EPI=$e timeout 300 python - <<'PY' 2>&1 | grep -v Warn | tail -1 | cut -c1-160
import sys, os; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import torch, bench_configs as b
b.run('cfg2 epi=' + os.environ['MLB_TC_EPI'], 8192, 32, 1, 256, 3, 4, 4, torch.bfloat16, steps=40, warm=10)
PY
done
