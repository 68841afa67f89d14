mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q --tb=short 2>&1 | grep -E "^E  |^FAILED|passed|failed|skipped" | cut -c1-250 | tee gpurun_out/r2_pytest_dp_2gpu_final.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_update_check.py 2>&1 | grep "DP_UPDATE_OK\|AssertionError\|Error" | head -3 | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | grep "^{" > gpurun_out/r2_bench_2gpu_final.json
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2_bench_2gpu_final.json').readline())
print(b['n_gpus'], b['ms_per_step'], b['value'], b['e2e']['value'], b.get('dp_check'), b['config'].get('permutation'), json.dumps(b.get('cfg3'))[:300])
PY
