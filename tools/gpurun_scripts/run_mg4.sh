mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1200 python -m pytest tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2_pytest_dp_4gpu.log
for n in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err; echo "bench$n rc=$?"
  tail -c 600 gpurun_out/r2_bench_${n}gpu.err
  python - <<PY
import json
try:
    b=json.loads([l for l in open('gpurun_out/r2_bench_${n}gpu.json') if l.startswith('{')][-1])
    print({k:b[k] for k in ('value','ms_per_step','n_gpus','dp_check')}, b['e2e']['value'], b['cfg3']['value'], b['cfg3']['ms_per_step'], b['config']['permutation'], b['config']['fused_allreduce'])
except Exception as e:
    print('no json', e)
PY
done
