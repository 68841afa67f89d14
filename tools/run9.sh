mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_recurrent_gpu.py -m gpu -q -x -s 2>&1 | grep "PARITY\|passed\|failed\|Error\|error" | tail -12
timeout 600 python - <<'PY' 2>&1 | grep -v Warn | tail -4
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import torch, bench_configs as b
b.run('cfg4 LSTM-256 actor-critic, 16384x128, 4 BPTT chunks, value-norm EMA', 16384, 128, 4, 256, 2, 4, 2, torch.bfloat16, rnn=256, normalize_values=True, steps=3, warm=3)
PY
