mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/r2_pytest_full.log
grep -E "^E  |^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_full.log | cut -c1-300 | head -30
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 30 --warmup 10 2>/dev/null | python -c "
import sys, json
b=json.loads(sys.stdin.readline()); print(b['ms_per_step'], b['value'], b['e2e']['value'])"
