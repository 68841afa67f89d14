timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_update_check.py 2>&1 | grep "DP_UPDATE_OK\|AssertionError\|Error" | head -3 | cut -c1-700
for v in "MLB_X=1" "MLB_DP_ASSIGN=slice"; do
  echo "== $v"
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "
import sys, json
b=json.loads(sys.stdin.readline()); print(b['ms_per_step'], b['value'], b['e2e']['ms_per_step'], b['config']['permutation'])"
done
