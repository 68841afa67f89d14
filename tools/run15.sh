mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/r2_pytest_full.log
grep -E "^E  |^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_full.log | cut -c1-300 | head -30
for h in 1 0 1 0; do MLB_FUSE_POST_STEP=$h timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 60 --warmup 20 2>/dev/null | python -c "
import sys, json
b=json.loads(sys.stdin.readline()); print('fuse_post_step=$h', b['ms_per_step'], b['value'], b['e2e']['value'])"; done
