mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/r2_bench_2gpu.json'))
    print({k:b[k] for k in ('value','ms_per_step','n_gpus','dp_check')}, b['e2e'], b['cfg3'], b['config'])
except Exception as e:
    print('no json', e)
PY
